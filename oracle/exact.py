"""O1 -- the exact oracle (TEST INFRASTRUCTURE; see oracle/__init__.py).

NumPy restatement of the reference forward path with exact integer accumulators and ONE
fixed fp32 epilogue order.  The CUDA kernels are required to match
these functions bit-for-bit wherever the inputs of a layer are integer levels.

Reference lines restated here:
  quantize / quantized_tanh   layers/quantized_ops.py:49-66, 87-100
  round_through (tf.round)    layers/quantized_ops.py:8-14   (round-half-to-even)
  binary_tanh / binarize      layers/binary_ops.py:16-24, 37-64
  _ternarize / ternarize      layers/ternary_ops.py:15-41
  QuantizedConv2D.call        layers/quantized_layers.py:164-194 (scaling identity dropped: a value identity)
  QuantizedDense.call         layers/quantized_layers.py:79-88
  BinaryConv2D/Dense.call     layers/binary_layers.py:160-187, 78-85
  TernaryConv2D/Dense.call    layers/ternary_layers.py:156-174, 77-84
  graphs                      models/vgg.py:5-44, models/resnet.py:15-147 (via oracle.netspec)
"""
from __future__ import annotations

import numpy as np

from . import netspec

F32 = np.float32
BIN_THRESHOLD = F32(2.0 ** -24)   # binary_tanh(y) = +1  <=>  y > 2^-24 in fp32


# --------------------------------------------------------------------------- quantisers
def quantize_levels(w, nb):
    """Integer level k of quantize(W, nb): value = k / 2^(nb-1)  (quantized_ops.py:58-64)."""
    m = F32(2 ** (nb - 1))
    k = np.rint(np.asarray(w, F32) * m)                 # fp32 multiply by a power of two is exact
    return np.clip(k, -float(m), float(m) - 1).astype(np.int32)


def binarize_bits(w, H=1.0):
    """1 where binarize(W,H) == +H  (binary_ops.py:23-24,49,62)."""
    x = np.asarray(w, F32) / F32(H)
    hs = np.clip(F32(0.5) * x + F32(0.5), F32(0), F32(1))
    return (np.rint(hs) > 0.5).astype(np.uint8)          # == (x > 2^-24)


def binarize_levels(w, H=1.0):
    return binarize_bits(w, H).astype(np.int32) * 2 - 1


def mean_abs_fixed_order(a):
    """mean(|a|) in a FIXED order shared with the CUDA packer: 1024 strided partial sums in
    float64 (element i goes to lane i % 1024, accumulated in increasing i), combined by a
    halving tree (lane t += lane t + n/2), divided by the count in float64, rounded to fp32."""
    a = np.abs(np.asarray(a, F32).reshape(-1)).astype(np.float64)
    n = a.size
    lanes = 1024
    pad = (-n) % lanes
    if pad:
        a = np.concatenate([a, np.zeros(pad, np.float64)])
    rows = a.reshape(-1, lanes)
    part = np.zeros(lanes, np.float64)
    for r in range(rows.shape[0]):
        part = part + rows[r]
    w = lanes
    while w > 1:
        w //= 2
        part = part[:w] + part[w:2 * w]
    return F32(part[0] / np.float64(n))


def ternarize_levels(w, H=1.0):
    """_ternarize (ternary_ops.py:22-28): cutoff = 0.7*mean|W/H|; +1 if W>c; -1 if W<=-c; else 0."""
    x = np.asarray(w, F32) / F32(H)
    cutoff = F32(0.7) * mean_abs_fixed_order(x)
    t = np.where(x > cutoff, 1, np.where(x <= -cutoff, -1, 0))
    return t.astype(np.int32)


def weight_levels(kernel, wkind, nb, H=1.0):
    """-> (integer levels, weight scale as python float)."""
    if wkind == "quantized":
        return quantize_levels(kernel, nb), 1.0 / float(2 ** (nb - 1))
    if wkind == "binary":
        return binarize_levels(kernel, H), float(H)
    if wkind == "ternary":
        return ternarize_levels(kernel, H), float(H)
    if wkind == "float":                                   # no quantiser (keras Conv2D / Dense): the values themselves
        return np.asarray(kernel, np.float64), 1.0
    raise ValueError(wkind)


def act_quant_levels(z, abits):
    """quantized_tanh on fp32 values -> integer levels (quantized_ops.py:95-98)."""
    m = F32(2 ** (abits - 1))
    q = np.rint(np.asarray(z, F32) * m)
    return np.clip(q, -float(m), float(m) - 1).astype(np.int32)


def act_binary_levels(z):
    return np.where(np.asarray(z, F32) > BIN_THRESHOLD, 1, -1).astype(np.int32)


def leaky(z, alpha=netspec.LEAKY_ALPHA):
    z = np.asarray(z, F32)
    return np.where(z > 0, z, F32(alpha) * z).astype(F32)


# --------------------------------------------------------------------------- bit packing
def pack_bits_lastdim(levels_pm1):
    """+-1 levels [..., C] -> uint32 words [..., ceil(C/32)], bit (c%32) of word c//32 set
    iff level == +1.  Channels beyond C are 0 bits (they are masked by the weight side)."""
    lv = np.asarray(levels_pm1)
    C = lv.shape[-1]
    words = (C + 31) // 32
    bits = (lv > 0).astype(np.uint64)
    pad = words * 32 - C
    if pad:
        bits = np.concatenate([bits, np.zeros(lv.shape[:-1] + (pad,), np.uint64)], axis=-1)
    bits = bits.reshape(lv.shape[:-1] + (words, 32))
    sh = np.arange(32, dtype=np.uint64)
    return (bits << sh).sum(axis=-1).astype(np.uint32)


# --------------------------------------------------------------------------- geometry
def same_pads(size, k, stride):
    """TensorFlow SAME padding (asymmetric for stride 2: 32->16 pads 0 before, 1 after)."""
    out = -(-size // stride)
    total = max((out - 1) * stride + k - size, 0)
    before = total // 2
    return out, before, total - before


def conv_accumulate(x, w_levels, stride):
    """Exact SAME conv: x [N,H,W,Cin] (integer valued or float), w [kh,kw,Cin,Cout] integer
    levels.  Integer inputs: float64 BLAS on exact integers (exact while |sum| < 2^53), returned
    as int64.  Float inputs: float64 accumulation, returned as float64."""
    x = np.asarray(x)
    is_int = np.issubdtype(x.dtype, np.integer)
    n, h, wd, cin = x.shape
    kh, kw, cin2, cout = w_levels.shape
    assert cin == cin2
    oh, pt, pb = same_pads(h, kh, stride)
    ow, pl, pr = same_pads(wd, kw, stride)
    xp = np.zeros((n, h + pt + pb, wd + pl + pr, cin), np.float64)
    xp[:, pt:pt + h, pl:pl + wd, :] = x
    wf = np.asarray(w_levels, np.float64)
    acc = np.zeros((n, oh, ow, cout), np.float64)
    for r in range(kh):
        for s in range(kw):
            patch = xp[:, r:r + (oh - 1) * stride + 1:stride, s:s + (ow - 1) * stride + 1:stride, :]
            acc += patch.reshape(-1, cin).dot(wf[r, s]).reshape(n, oh, ow, cout)
    if is_int:
        return np.rint(acc).astype(np.int64)
    return acc


def bn_constants(gamma, beta, mean, var, eps):
    """inv = gamma / sqrt(var + eps); shift = beta - mean*inv -- all fp32, computed once."""
    gamma, beta, mean, var = (np.asarray(a, F32) for a in (gamma, beta, mean, var))
    inv = (gamma / np.sqrt(var + F32(eps))).astype(F32)
    shift = (beta - mean * inv).astype(F32)
    return inv, shift


def acc_scale(x_scale, w_scale):
    """fp32 scale applied to an accumulator: fl32(x_scale * w_scale), product in float64."""
    return F32(np.float64(x_scale) * np.float64(w_scale))


# --------------------------------------------------------------------------- tensors
class QT:
    """A tagged tensor: kind in {'u8','i8','b1','i32','f32'}; value = data * scale for the
    integer kinds ('b1' holds +-1 levels)."""

    def __init__(self, kind, data, scale=1.0):
        self.kind, self.data, self.scale = kind, data, float(scale)

    def values(self):
        if self.kind == "f32":
            return self.data
        return (self.data.astype(F32) * F32(self.scale)).astype(F32)


def make_input(x):
    x = np.asarray(x)
    if x.dtype == np.uint8:
        return QT("u8", x.astype(np.int32), 1.0 / 255.0)     # utils/load_data.py:40
    return QT("f32", x.astype(F32))


def linear(x: QT, w_levels, w_scale, stride=1, dense=False):
    """conv (or dense) + the first epilogue step: c = float(acc) * s  (App. A.4 step 1)."""
    if dense:
        xd = x.data.reshape(x.data.shape[0], 1, 1, -1)
        wl = w_levels.reshape(1, 1, *w_levels.shape)
    else:
        xd, wl = x.data, w_levels
    if x.kind == "f32":
        acc = conv_accumulate(xd.astype(np.float64), wl, stride)
        c = acc.astype(F32) * F32(w_scale)
        iacc = None
    else:
        iacc = conv_accumulate(xd.astype(np.int64), wl, stride)
        c = iacc.astype(F32) * acc_scale(x.scale, w_scale)
    if dense:
        c = c.reshape(c.shape[0], -1)
        if iacc is not None:
            iacc = iacc.reshape(iacc.shape[0], -1)
    return c.astype(F32), iacc


def maxpool2(a):
    n, h, w, c = a.shape
    h2, w2 = h // 2, w // 2
    a = a[:, :h2 * 2, :w2 * 2, :].reshape(n, h2, 2, w2, 2, c)
    return a.max(axis=(2, 4))


def softmax64(z):
    z = np.asarray(z, np.float64)
    z = z - z.max(axis=-1, keepdims=True)
    e = np.exp(z)
    return (e / e.sum(axis=-1, keepdims=True)).astype(F32)


# --------------------------------------------------------------------------- graph evaluation
def forward(nodes, x, return_all=False, teacher=None):
    """Evaluate a netspec graph exactly.  Returns fp32 output (N, classes) (probabilities when the
    final dense has softmax -- the pre-softmax logits are in ``info['logits']``).

    ``return_all`` -> (out, list of QT per node, info)."""
    vals = [None] * len(nodes)
    info = {"acc": {}}
    for i, nd in enumerate(nodes):
        op = nd["op"]
        src = [vals[j] for j in nd["in"]]
        if op == "input":
            vals[i] = make_input(x)
        elif op == "zeropad":
            p = nd["pad"]
            t = src[0]
            vals[i] = QT(t.kind, np.pad(t.data, ((0, 0), (p, p), (p, p), (0, 0))), t.scale)
        elif op in ("conv", "dense"):
            lv, ws = weight_levels(nd["kernel"], nd["wkind"], nd["nb"], nd["H"])
            xin = src[0]
            if nd["wkind"] == "float" and xin.kind != "f32":
                # a float layer sees VALUES; pixel levels become x / 255 in fp32 as utils/load_data.py:40 computes them
                xin = QT("f32", (xin.data.astype(F32) / F32(255)).astype(F32)) if xin.kind == "u8" else QT("f32", xin.values())
            c, iacc = linear(xin, lv, ws, nd.get("stride", 1), dense=(op == "dense"))
            if iacc is not None:
                info["acc"][i] = iacc
            if nd["use_bias"]:
                c = (c + np.asarray(nd["bias"], F32)).astype(F32)
            vals[i] = QT("f32", c)
            if op == "dense" and nd.get("softmax"):
                info["logits"] = c
                vals[i] = QT("f32", softmax64(c))
        elif op == "bn":
            inv, shift = bn_constants(nd["gamma"], nd["beta"], nd["mean"], nd["var"], nd["eps"])
            p = src[0].values()
            y = (p * inv).astype(F32)          # separate RN multiply ...
            y = (y + shift).astype(F32)        # ... then RN add (no FMA)
            vals[i] = QT("f32", y)
        elif op == "add":
            a = src[0].values()
            b = src[1].values()
            z = ((a + b).astype(F32) * F32(nd["mul"])).astype(F32)
            vals[i] = QT("f32", z)
        elif op == "act":
            z = src[0].values()
            if nd["akind"] == "quant":
                vals[i] = QT("i8", act_quant_levels(z, nd["abits"]), 1.0 / float(2 ** (nd["abits"] - 1)))
            elif nd["akind"] == "binary":
                vals[i] = QT("b1", act_binary_levels(z), 1.0)
            else:
                vals[i] = QT("f32", leaky(z))
        elif op == "maxpool":
            t = src[0]
            vals[i] = QT(t.kind, maxpool2(t.data), t.scale)
        elif op == "avgpool":
            t = src[0]
            s = nd["size"]
            n, h, w, c = t.data.shape
            blk = t.data[:, :h // s * s, :w // s * s, :].reshape(n, h // s, s, w // s, s, c)
            if t.kind == "f32":
                vals[i] = QT("f32", (blk.astype(np.float64).sum(axis=(2, 4)) / float(s * s)).astype(F32))
            else:
                vals[i] = QT("i32", blk.astype(np.int64).sum(axis=(2, 4)), t.scale / float(s * s))
        elif op == "flatten":
            t = src[0]
            vals[i] = QT(t.kind, t.data.reshape(t.data.shape[0], -1), t.scale)
        else:
            raise ValueError(op)
    out = vals[-1].values()
    if "logits" not in info:
        info["logits"] = out
    if return_all:
        return out, vals, info
    return out
