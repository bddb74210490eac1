"""Thin host wrappers over the C ABI: torch tensors in, torch tensors out (device memory only).

``QTensor`` is the tagged activation tensor that lets consecutive layers stay in the integer
domain: kind 'u8' (pixel levels, value = level/255), 'i8' (quantized_tanh levels, value =
level*scale), 'b1' (bit-packed +-1, uint32 words on the channel axis) or 'f32'.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib as L

KIND_CODE = {"u8": L.KIND_U8, "i8": L.KIND_I8, "b1": L.KIND_B1, "f32": L.KIND_F32}
F32 = np.float32


@dataclass
class QTensor:
    kind: str                 # 'u8' | 'i8' | 'b1' | 'f32'
    data: torch.Tensor        # NHWC (or [N, F]); 'b1': int32 words, last dim = ceil(C/32)
    scale: float = 1.0        # value = level * scale for 'u8' / 'i8'
    channels: int = 0         # logical channel count (needed for 'b1')

    @property
    def shape(self):
        s = tuple(self.data.shape)
        if self.kind == "b1":
            return s[:-1] + (self.channels,)
        return s

    def to_float(self) -> torch.Tensor:
        """fp32 values (level*scale or +-1) -- what the reference layer would have output."""
        if self.kind == "f32":
            return self.data
        out = torch.empty(self.shape, dtype=torch.float32, device=self.data.device)
        count = out.numel()
        ch = self.channels if self.kind == "b1" else int(self.shape[-1])
        L.check(L.lib().qnnb_dequantize(KIND_CODE[self.kind], L.ptr(self.data), count, ch,
                                        float(F32(self.scale)), L.ptr(out), L.current_stream_ptr()))
        return out


def as_qtensor(x) -> QTensor:
    if isinstance(x, QTensor):
        return x
    if not isinstance(x, torch.Tensor):
        raise TypeError("expected a torch CUDA tensor or QTensor, got %r" % type(x))
    if x.dtype == torch.uint8:
        return QTensor("u8", x.contiguous(), 1.0 / 255.0, int(x.shape[-1]))   # utils/load_data.py:40
    if x.dtype == torch.float32:
        return QTensor("f32", x.contiguous(), 1.0, int(x.shape[-1]))
    raise TypeError("unsupported input dtype %s (uint8 pixel levels or float32 values)" % x.dtype)


# --------------------------------------------------------------------------- K0
def packed_weight_bytes(wfmt, kh, kw, cin, cout) -> int:
    return int(L.lib().qnnb_packed_weight_bytes(wfmt, kh, kw, cin, cout))


def pack_weights(kernel_hwio: torch.Tensor, mode: int, nb: int, H: float, wfmt: int) -> torch.Tensor:
    """quantize / binarize / ternarize + pack.  ``kernel_hwio``: fp32 CUDA (kh,kw,cin,cout)."""
    if kernel_hwio.dim() == 2:
        kernel_hwio = kernel_hwio.reshape(1, 1, *kernel_hwio.shape)
    kh, kw, cin, cout = (int(v) for v in kernel_hwio.shape)
    k = kernel_hwio.contiguous().float()
    nbytes = packed_weight_bytes(wfmt, kh, kw, cin, cout)
    out = torch.empty(nbytes, dtype=torch.uint8, device=k.device)
    scratch = torch.zeros(2, dtype=torch.float32, device=k.device)
    L.check(L.lib().qnnb_pack_weights(mode, int(nb), float(H), L.ptr(k), kh, kw, cin, cout, wfmt,
                                      L.ptr(out), L.ptr(scratch), L.current_stream_ptr()))
    if wfmt == L.WFMT_I8:
        return out.view(torch.int8).reshape(cout, kh, kw, (cin + 3) // 4 * 4)
    if wfmt == L.WFMT_F32:
        return out.view(torch.float32).reshape(cout, kh, kw, (cin + 3) // 4 * 4)
    return out.view(torch.int32).reshape(cout, kh, kw, (cin + 31) // 32)


# --------------------------------------------------------------------------- epilogue
def make_epilogue(acc_scale, bias=None, bn_inv=None, bn_shift=None, residual: QTensor | None = None, res_mul=1.0,
                  act=L.ACT_NONE, abits=0, leaky_alpha=0.3, pool=0) -> L.Epilogue:
    e = L.Epilogue()
    e.acc_scale = float(F32(acc_scale))
    e.bias = L.ptr(bias)
    e.bn_inv = L.ptr(bn_inv)
    e.bn_shift = L.ptr(bn_shift)
    if residual is None:
        e.res_kind = L.KIND_NONE
        e.residual = None
        e.res_scale = 1.0
    else:
        if residual.kind not in ("i8", "f32"):
            raise ValueError("residual must be an int8 or fp32 tensor, got %s" % residual.kind)
        e.res_kind = KIND_CODE[residual.kind]
        e.residual = L.ptr(residual.data)
        e.res_scale = float(F32(residual.scale))
    e.res_mul = float(res_mul)
    e.act = int(act)
    e.abits = int(abits)
    e.leaky_alpha = float(F32(leaky_alpha))
    e.pool = int(pool)
    # the struct only holds raw device pointers: keep the tensors alive as long as the struct is
    e._refs = (bias, bn_inv, bn_shift, residual)
    return e


def acc_scale(x_scale: float, w_scale: float) -> np.float32:
    """fp32 scale of an integer accumulator: fl32(x_scale * w_scale), product formed in float64."""
    return F32(np.float64(x_scale) * np.float64(w_scale))


# --------------------------------------------------------------------------- conv / dense
def _conv_desc(x: QTensor, kh, kw, cout, stride, epi: L.Epilogue, impl, max_ctas=0) -> L.ConvDesc:
    n, h, w, cin = (int(v) for v in x.shape)
    d = L.ConvDesc()
    d.n, d.h, d.w, d.cin = n, h, w, cin
    d.cout, d.kh, d.kw, d.stride = int(cout), int(kh), int(kw), int(stride)
    d.in_kind = KIND_CODE[x.kind]
    d.impl = int(impl)
    d.epi = epi
    d.w_f32 = 0
    d.max_ctas = int(max_ctas)
    return d


def conv2d_on_tensor_cores(x: QTensor, kh, kw, cout, stride, epi: L.Epilogue, impl=L.IMPL_AUTO) -> bool:
    """Would ``conv2d`` run this call on the tcgen05 kernels?  (The plan asks before choosing the int8 form of a
    +-1 activation map over the bit-packed one.)"""
    if impl == L.IMPL_GENERIC:
        return False
    d = _conv_desc(x, kh, kw, cout, stride, epi, impl)
    return bool(L.lib().qnnb_conv2d_tc_supported(C.byref(d)))


def conv2d(x: QTensor, w_packed: torch.Tensor, kh, kw, cout, stride, epi: L.Epilogue, impl=L.IMPL_AUTO,
           out: torch.Tensor | None = None, max_ctas=0) -> QTensor:
    """``w_packed`` of dtype float32 is a QNNB_WFMT_F32 kernel ('float' networks; fp32 activations only).  ``max_ctas``:
    SM share of the persistent kernels (0 = the whole device; see qnnb_conv_desc.max_ctas)."""
    n, h, w, cin = (int(v) for v in x.shape)
    d = _conv_desc(x, kh, kw, cout, stride, epi, impl, max_ctas)
    d.w_f32 = 1 if w_packed.dtype == torch.float32 else 0
    oh, ow = C.c_int32(), C.c_int32()
    L.check(L.lib().qnnb_conv2d_out_shape(C.byref(d), C.byref(oh), C.byref(ow)))
    oh, ow = oh.value, ow.value
    dev = x.data.device
    if epi.act == L.ACT_QUANT:
        shape, dtype, kind, scale = (n, oh, ow, cout), torch.int8, "i8", 1.0 / float(1 << (epi.abits - 1))
    elif epi.act == L.ACT_SIGN:
        shape, dtype, kind, scale = (n, oh, ow, (cout + 31) // 32), torch.int32, "b1", 1.0
    elif epi.act == L.ACT_SIGN_I8:
        shape, dtype, kind, scale = (n, oh, ow, cout), torch.int8, "i8", 1.0          # levels +1 / -1
    else:
        shape, dtype, kind, scale = (n, oh, ow, cout), torch.float32, "f32", 1.0
    if out is None:
        out = torch.empty(shape, dtype=dtype, device=dev)
    elif tuple(out.shape) != shape or out.dtype != dtype:
        raise ValueError("conv2d: bad output buffer %s/%s, need %s/%s" % (tuple(out.shape), out.dtype, shape, dtype))
    L.check(L.lib().qnnb_conv2d(C.byref(d), L.ptr(x.data), L.ptr(w_packed), L.ptr(out), L.current_stream_ptr()))
    return QTensor(kind, out, scale, int(cout))


def dense(x: QTensor, w_packed: torch.Tensor, units, epi: L.Epilogue, softmax=False, want_logits=False,
          out: torch.Tensor | None = None, logits: torch.Tensor | None = None, avg_positions=0, max_ctas=0):
    """``avg_positions`` = P > 1 (fp32 input): ``x`` is [n, P*fin] laid out [n][P][fin]; the layer input is the sum over
    the P positions (global average pooling folded in, the 1/P is part of ``epi.acc_scale``)."""
    n = int(x.data.shape[0])
    fin = int(np.prod(x.shape[1:]))
    d = L.DenseDesc()
    if avg_positions and avg_positions > 1:
        if fin % int(avg_positions):
            raise ValueError("dense: %d features do not split into %d positions" % (fin, avg_positions))
        fin //= int(avg_positions)
        d.avg_positions = int(avg_positions)
    d.n, d.fin, d.units = n, fin, int(units)
    d.in_kind = KIND_CODE[x.kind]
    d.softmax = 1 if softmax else 0
    d.epi = epi
    d.w_f32 = 1 if w_packed.dtype == torch.float32 else 0
    d.max_ctas = int(max_ctas)
    dev = x.data.device
    if out is None:
        out = torch.empty((n, units), dtype=torch.float32, device=dev)
    elif tuple(out.shape) != (n, int(units)) or out.dtype != torch.float32:
        raise ValueError("dense: bad output buffer %s/%s, need %s/float32" % (tuple(out.shape), out.dtype, (n, int(units))))
    if softmax and want_logits and logits is None:
        logits = torch.empty((n, units), dtype=torch.float32, device=dev)
    L.check(L.lib().qnnb_dense(C.byref(d), L.ptr(x.data), L.ptr(w_packed), L.ptr(out),
                               L.ptr(logits) if logits is not None else C.c_void_p(0), L.current_stream_ptr()))
    return out, logits


# --------------------------------------------------------------------------- whole-network launch
def vgg_desc(n, h, w, cin, convs, units, dense_w, dense_epi, max_ctas=0) -> L.VggDesc:
    """``convs``: list of (cout, pool, packed kernel, Epilogue) -- see ``qnnb_vgg_desc`` in include/qnnb200.h."""
    d = L.VggDesc()
    d.n, d.h, d.w, d.cin = int(n), int(h), int(w), int(cin)
    d.nconv = len(convs)
    if len(convs) > L.NET_MAX_CONVS:
        raise ValueError("vgg_desc: at most %d convolutions" % L.NET_MAX_CONVS)
    for i, (cout, pool, wp, epi) in enumerate(convs):
        d.conv[i].cout, d.conv[i].pool = int(cout), int(pool)
        d.conv[i].w = L.ptr(wp)
        d.conv[i].epi = epi
    d.units = int(units)
    d.dense_w = L.ptr(dense_w)
    d.dense_epi = dense_epi
    d.max_ctas = int(max_ctas)
    d._refs = (convs, dense_w, dense_epi)
    return d


def vgg_forward_supported(d: L.VggDesc) -> bool:
    return bool(L.lib().qnnb_vgg_forward_supported(C.byref(d)))


def vgg_pack(d: L.VggDesc, device) -> torch.Tensor:
    """The net's resident image (kernels in tensor-core operand order + dense kernel + epilogue constants): build once
    per set of weights, pass to every ``vgg_forward``."""
    nbytes = int(L.lib().qnnb_vgg_blob_bytes(C.byref(d)))
    if nbytes <= 0:
        raise ValueError("vgg_pack: this net is outside the whole-network kernel's scope")
    blob = torch.empty(nbytes, dtype=torch.uint8, device=device)
    L.check(L.lib().qnnb_vgg_pack(C.byref(d), L.ptr(blob), L.current_stream_ptr()))
    return blob


def vgg_forward(d: L.VggDesc, blob: torch.Tensor, x: torch.Tensor, out=None):
    """One launch: uint8 images [n, h, w, cin] -> fp32 [n, units] (models/vgg.py:15-42 end to end)."""
    if x.dtype != torch.uint8:
        raise TypeError("vgg_forward takes uint8 pixel levels")
    if out is None:
        out = torch.empty((int(d.n), int(d.units)), dtype=torch.float32, device=x.device)
    elif tuple(out.shape) != (int(d.n), int(d.units)) or out.dtype != torch.float32:
        raise ValueError("vgg_forward: bad output buffer %s/%s" % (tuple(out.shape), out.dtype))
    L.check(L.lib().qnnb_vgg_forward(C.byref(d), L.ptr(blob), L.ptr(x), L.ptr(out), L.current_stream_ptr()))
    return out


# --------------------------------------------------------------------------- stand-alone ops
def quantize_act(x: torch.Tensor, abits: int) -> QTensor:
    x = x.contiguous()
    y = torch.empty(x.shape, dtype=torch.int8, device=x.device)
    L.check(L.lib().qnnb_quantize_act(L.ACT_QUANT, int(abits), L.ptr(x), x.numel(), int(x.shape[-1]), L.ptr(y),
                                      L.current_stream_ptr()))
    return QTensor("i8", y, 1.0 / float(1 << (abits - 1)), int(x.shape[-1]))


def sign_act(x: torch.Tensor) -> QTensor:
    x = x.contiguous()
    ch = int(x.shape[-1])
    y = torch.empty(tuple(x.shape[:-1]) + ((ch + 31) // 32,), dtype=torch.int32, device=x.device)
    L.check(L.lib().qnnb_quantize_act(L.ACT_SIGN, 0, L.ptr(x), x.numel(), ch, L.ptr(y), L.current_stream_ptr()))
    return QTensor("b1", y, 1.0, ch)


def batchnorm(x: torch.Tensor, inv: torch.Tensor, shift: torch.Tensor) -> torch.Tensor:
    x = x.contiguous()
    ch = int(x.shape[-1])
    y = torch.empty_like(x)
    L.check(L.lib().qnnb_batchnorm_f32(L.ptr(x), x.numel() // ch, ch, L.ptr(inv), L.ptr(shift), L.ptr(y),
                                       L.current_stream_ptr()))
    return y


def maxpool2(x: torch.Tensor) -> torch.Tensor:
    x = x.contiguous()
    n, h, w, c = (int(v) for v in x.shape)
    y = torch.empty((n, h // 2, w // 2, c), dtype=torch.float32, device=x.device)
    L.check(L.lib().qnnb_maxpool2_f32(L.ptr(x), n, h, w, c, L.ptr(y), L.current_stream_ptr()))
    return y


def leaky(x: torch.Tensor, alpha: float) -> torch.Tensor:
    x = x.contiguous()
    y = torch.empty_like(x)
    L.check(L.lib().qnnb_leaky_f32(L.ptr(x), x.numel(), float(F32(alpha)), L.ptr(y), L.current_stream_ptr()))
    return y


def round_half_even(x: torch.Tensor) -> torch.Tensor:
    x = x.contiguous()
    y = torch.empty_like(x)
    L.check(L.lib().qnnb_round_f32(L.ptr(x), x.numel(), L.ptr(y), L.current_stream_ptr()))
    return y


def bn_constants(gamma, beta, mean, var, eps):
    """inv = gamma / sqrt(var + eps), shift = beta - mean*inv -- fp32 on the host, once
    (keras BatchNormalization inference)."""
    gamma, beta, mean, var = (np.asarray(a, F32) for a in (gamma, beta, mean, var))
    inv = (gamma / np.sqrt(var + F32(eps))).astype(F32)
    shift = (beta - mean * inv).astype(F32)
    return inv, shift
