"""O2 -- torch-CPU fp32 restatement of the reference forward graph (TEST INFRASTRUCTURE).

Op-for-op with what Keras/TensorFlow execute for ``model.predict`` on the reference
(SURVEY.md section 3c): kernels are re-quantised on every forward, the conv is wrapped in
the gradient-scaling identity (``trick=True`` -> O2a, exactly as written in
layers/quantized_layers.py:167-180 and layers/binary_layers.py:163-176; ``trick=False`` ->
O2b), then bias_add, BatchNormalization, activation and pooling run as separate fp32 passes.

This is also the CPU baseline timed by ``bench.py`` (it is what the reference executes).
TensorFlow/Keras semantics restated from their documentation: tf.round = half-to-even;
SAME padding (asymmetric under stride 2); tf.nn.batch_normalization =
``x*inv + (beta - mean*inv)`` with ``inv = rsqrt(var+eps)*gamma``; MaxPooling2D(2,2) valid;
AveragePooling2D(8); Flatten in H,W,C order; LeakyReLU alpha=0.3; softmax.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import netspec
from .exact import same_pads


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a, np.float32)))


def round_through(x):                       # quantized_ops.py:8-14 (forward value)
    return x + (torch.round(x) - x)         # torch.round is half-to-even like tf.round


def quantize(w, nb):                        # quantized_ops.py:49-66 / 87-100
    m = float(2 ** (nb - 1))
    return torch.clamp(round_through(w * m), -m, m - 1) / m


def binary_tanh(x):                         # binary_ops.py:16-24, 37-51
    hs = torch.clamp(0.5 * x + 0.5, 0, 1)
    return 2 * round_through(hs) - 1


def binarize(w, H=1.0):                     # binary_ops.py:54-64
    return H * binary_tanh(w / H)


def ternarize(w, H=1.0):                    # ternary_ops.py:15-41
    x = w / H
    cutoff = 0.7 * torch.mean(torch.abs(x))
    ones = torch.ones_like(x)
    wt = torch.where(x > cutoff, ones, torch.where(x <= -cutoff, -ones, torch.zeros_like(x)))
    wt = wt * H
    return w + (wt - w)


def _weights(nd):
    k = _t(nd["kernel"])
    if nd["wkind"] == "quantized":
        return quantize(k, nd["nb"])
    if nd["wkind"] == "binary":
        return binarize(k, nd["H"])
    if nd["wkind"] == "float":                     # keras Conv2D / Dense (model_factory.py:24-27): no quantiser
        return k
    return ternarize(k, nd["H"])


def _conv_same(x_nhwc, w_hwio, stride):
    n, h, w, c = x_nhwc.shape
    kh, kw = w_hwio.shape[0], w_hwio.shape[1]
    _, pt, pb = same_pads(h, kh, stride)
    _, pl, pr = same_pads(w, kw, stride)
    x = x_nhwc.permute(0, 3, 1, 2)
    x = F.pad(x, (pl, pr, pt, pb))
    wt = w_hwio.permute(3, 2, 0, 1).contiguous()
    y = F.conv2d(x, wt, stride=stride)
    return y.permute(0, 2, 3, 1).contiguous()


def _scaled_conv(x, wq, stride, klm, trick):
    """The gradient-scaling identity around K.conv2d (quantized_layers.py:167-180).  The python
    constants are float64 scalars (klm is a numpy float32; ``1./klm`` promotes to float64 under
    the numpy the reference ran on) that TF casts to fp32 when they meet the fp32 tensor."""
    if not trick:
        return _conv_same(x, wq, stride)
    klm64 = float(np.float32(klm))
    inv = 1.0 / klm64
    c_in = float(np.float32(1.0 - 1.0 / inv))
    s_in = float(np.float32(inv))
    xi = (x - c_in * x) * s_in
    o = _conv_same(xi, wq, stride)
    c_out = float(np.float32(1.0 - 1.0 / klm64))
    s_out = float(np.float32(klm64))
    return (o - c_out * o) * s_out


def prepare_input(x):
    x = np.asarray(x)
    if x.dtype == np.uint8:
        x = x.astype("float32") / 255          # utils/load_data.py:40
    return _t(x)


@torch.no_grad()
def forward(nodes, x, trick=True, return_all=False, teacher=None, dtype=torch.float32):
    """``teacher``: optional dict node-index -> fp32 numpy array that REPLACES the computed value
    of that node before its consumers read it (teacher forcing with the exact oracle's values)."""
    vals = [None] * len(nodes)
    info = {}
    for i, nd in enumerate(nodes):
        op = nd["op"]
        src = [vals[j] for j in nd["in"]]
        if op == "input":
            v = prepare_input(x).to(dtype)
        elif op == "zeropad":
            p = nd["pad"]
            v = F.pad(src[0], (0, 0, p, p, p, p))
        elif op == "conv":
            wq = _weights(nd).to(dtype)
            # TernaryConv2D.call has no scaling identity (ternary_layers.py:156-174)
            use_trick = trick and nd["wkind"] in ("quantized", "binary")
            v = _scaled_conv(src[0], wq, nd["stride"], nd["klm"], use_trick)
            if nd["use_bias"]:
                v = v + _t(nd["bias"]).to(dtype)
        elif op == "dense":
            wq = _weights(nd).to(dtype)
            v = src[0] @ wq                       # K.dot, no scaling identity (quantized_layers.py:79-88)
            if nd["use_bias"]:
                v = v + _t(nd["bias"]).to(dtype)
            if nd.get("softmax"):
                info["logits"] = v.float().numpy()
                v = torch.softmax(v, dim=-1)
        elif op == "bn":
            g, b, mu, var = (_t(nd[k]).to(dtype) for k in ("gamma", "beta", "mean", "var"))
            inv = torch.rsqrt(var + nd["eps"]) * g
            v = src[0] * inv + (b - mu * inv)
        elif op == "add":
            v = (src[0] + src[1]) * nd["mul"]
        elif op == "act":
            if nd["akind"] == "quant":
                v = quantize(src[0], nd["abits"])
            elif nd["akind"] == "binary":
                v = binary_tanh(src[0])
            else:
                v = F.leaky_relu(src[0], negative_slope=netspec.LEAKY_ALPHA)
        elif op == "maxpool":
            v = F.max_pool2d(src[0].permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1).contiguous()
        elif op == "avgpool":
            v = F.avg_pool2d(src[0].permute(0, 3, 1, 2), nd["size"]).permute(0, 2, 3, 1).contiguous()
        elif op == "flatten":
            v = src[0].reshape(src[0].shape[0], -1)
        else:
            raise ValueError(op)
        if teacher is not None and i in teacher:
            info.setdefault("own", {})[i] = v.float().numpy()
            v = _t(teacher[i]).to(dtype)
        vals[i] = v
    out = vals[-1].float().numpy()
    if "logits" not in info:
        info["logits"] = out
    if return_all:
        return out, [None if v is None else v.float().numpy() for v in vals], info
    return out
