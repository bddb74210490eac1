"""ctypes binding of ``libqnnb200.so`` (C ABI declared in ``include/qnnb200.h``).

There is NO CPU fallback: if the shared library is missing the import of any compute entry
point raises, and every non-zero status code becomes a Python exception carrying
``qnnb_last_error()``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# QNNB_LIB selects an alternative build of the same library (kernel A/B experiments); never a different backend
LIB_PATH = os.environ.get("QNNB_LIB") or os.path.join(_HERE, "libqnnb200.so")

# ---- constants mirrored from include/qnnb200.h
KIND_NONE, KIND_U8, KIND_I8, KIND_B1, KIND_F32 = -1, 0, 1, 2, 3
W_QUANT, W_BINARY, W_TERNARY, W_FLOAT = 0, 1, 2, 3
WFMT_I8, WFMT_B1, WFMT_F32 = 0, 1, 2
ACT_NONE, ACT_QUANT, ACT_SIGN, ACT_LEAKY, ACT_SIGN_I8 = 0, 1, 2, 3, 4
IMPL_AUTO, IMPL_GENERIC, IMPL_TCGEN05, IMPL_TCGEN05_V1 = 0, 1, 2, 3
EINVAL, ECUDA, EUNSUPPORTED = -1, -2, -3


class Epilogue(C.Structure):
    _fields_ = [
        ("acc_scale", C.c_float),
        ("bias", C.c_void_p),
        ("bn_inv", C.c_void_p),
        ("bn_shift", C.c_void_p),
        ("res_kind", C.c_int32),
        ("residual", C.c_void_p),
        ("res_scale", C.c_float),
        ("res_mul", C.c_float),
        ("act", C.c_int32),
        ("abits", C.c_int32),
        ("leaky_alpha", C.c_float),
        ("pool", C.c_int32),
    ]


class ConvDesc(C.Structure):
    _fields_ = [
        ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("cin", C.c_int32),
        ("cout", C.c_int32), ("kh", C.c_int32), ("kw", C.c_int32),
        ("stride", C.c_int32),
        ("in_kind", C.c_int32),
        ("impl", C.c_int32),
        ("epi", Epilogue),
        ("w_f32", C.c_int32),
        ("max_ctas", C.c_int32),
    ]


class DenseDesc(C.Structure):
    _fields_ = [
        ("n", C.c_int32), ("fin", C.c_int32), ("units", C.c_int32),
        ("in_kind", C.c_int32),
        ("softmax", C.c_int32),
        ("epi", Epilogue),
        ("avg_positions", C.c_int32),
        ("w_f32", C.c_int32),
        ("max_ctas", C.c_int32),
    ]


NET_MAX_CONVS = 6


class NetConv(C.Structure):
    _fields_ = [("cout", C.c_int32), ("pool", C.c_int32), ("w", C.c_void_p), ("epi", Epilogue)]


class VggDesc(C.Structure):
    _fields_ = [
        ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("cin", C.c_int32),
        ("nconv", C.c_int32),
        ("conv", NetConv * NET_MAX_CONVS),
        ("units", C.c_int32),
        ("dense_w", C.c_void_p),
        ("dense_epi", Epilogue),
        ("max_ctas", C.c_int32),
    ]


# name -> (restype, argtypes); the CPU test-suite checks this table against the header
PROTOTYPES = {
    "qnnb_version": (C.c_int, []),
    "qnnb_last_error": (C.c_char_p, []),
    "qnnb_device_info": (C.c_int, [C.POINTER(C.c_int32)] * 3),
    "qnnb_pack_weights": (C.c_int, [C.c_int32, C.c_int32, C.c_float, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                    C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "qnnb_packed_weight_bytes": (C.c_int64, [C.c_int32] * 5),
    "qnnb_conv2d": (C.c_int, [C.POINTER(ConvDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "qnnb_conv2d_tc_supported": (C.c_int, [C.POINTER(ConvDesc)]),
    "qnnb_conv2d_out_shape": (C.c_int, [C.POINTER(ConvDesc), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "qnnb_dense": (C.c_int, [C.POINTER(DenseDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "qnnb_vgg_forward_supported": (C.c_int, [C.POINTER(VggDesc)]),
    "qnnb_vgg_blob_bytes": (C.c_int64, [C.POINTER(VggDesc)]),
    "qnnb_vgg_pack": (C.c_int, [C.POINTER(VggDesc), C.c_void_p, C.c_void_p]),
    "qnnb_vgg_forward": (C.c_int, [C.POINTER(VggDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "qnnb_quantize_act": (C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]),
    "qnnb_batchnorm_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "qnnb_maxpool2_f32": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "qnnb_leaky_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_float, C.c_void_p, C.c_void_p]),
    "qnnb_round_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "qnnb_debug_set_trace": (C.c_int, [C.c_void_p, C.c_int64]),
    "qnnb_peer_alloc": (C.c_int, [C.c_int64, C.POINTER(C.c_void_p), C.c_void_p]),
    "qnnb_peer_open": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "qnnb_peer_close": (C.c_int, [C.c_void_p]),
    "qnnb_peer_free": (C.c_int, [C.c_void_p]),
    "qnnb_dequantize": (C.c_int, [C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_void_p, C.c_void_p]),
}


class QnnbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libqnnb200: %s (status %d)" % (msg, code))
        self.code = code


_lib = None


def lib():
    """Load the shared library (once).  Raises if it has not been built: there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "libqnnb200.so not found at %s -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C <package>/csrc`; this package has no CPU fallback" % LIB_PATH)
        h = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(h, name)          # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        if h.qnnb_version() != 100:
            raise ImportError("libqnnb200.so version mismatch: %d" % h.qnnb_version())
        _lib = h
    return _lib


def check(status):
    if status != 0:
        msg = lib().qnnb_last_error()
        raise QnnbError(status, msg.decode("utf-8", "replace") if msg else "unknown error")


def current_stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


PEER_HANDLE_BYTES = 64


class DeviceBuffer:
    """A raw device address with a shape -- what a peer-mapped (CUDA IPC) buffer looks like on the ranks that do not
    own it: kernels may write through it, torch never sees it.  Quacks like a tensor for :func:`ptr`."""
    is_cuda = True

    def __init__(self, address, shape, dtype):
        self.address, self.shape, self.dtype = int(address), tuple(int(v) for v in shape), dtype

    def data_ptr(self):
        return self.address

    def is_contiguous(self):
        return True

    def rows(self, lo, hi):
        """Row block [lo, hi) of a 2-D buffer."""
        import torch
        width = self.shape[1] * torch.empty((), dtype=self.dtype).element_size()
        return DeviceBuffer(self.address + lo * width, (hi - lo, self.shape[1]), self.dtype)


def ptr(t):
    """Device pointer of a torch tensor or DeviceBuffer (None -> NULL)."""
    if t is None:
        return C.c_void_p(0)
    if not t.is_cuda:
        raise ValueError("libqnnb200 needs CUDA tensors (got a %s tensor); there is no CPU path" % t.device)
    if not t.is_contiguous():
        raise ValueError("libqnnb200 needs contiguous tensors")
    return C.c_void_p(t.data_ptr())


def device_info():
    sm, ma, mi = C.c_int32(), C.c_int32(), C.c_int32()
    check(lib().qnnb_device_info(C.byref(sm), C.byref(ma), C.byref(mi)))
    return sm.value, ma.value, mi.value
