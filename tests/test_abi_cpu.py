"""CPU-side checks of the drop-in boundary: the shared library loads and exports every symbol
include/qnnb200.h declares; ctypes structs match the header; no compute call needs a GPU here."""
import ctypes as C
import os
import re

import pytest

import qnn_b200 as q
from qnn_b200 import _lib as L

HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "qnnb200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(qnnb_[a-z0-9_]+)\s*\(", src)))


def test_library_built_and_loads():
    assert os.path.exists(L.LIB_PATH), "libqnnb200.so must be built in-tree (python -c 'import __graft_entry__ as g; g.build()')"
    h = L.lib()
    assert h.qnnb_version() == 100


def test_every_declared_symbol_is_exported_and_bound():
    syms = declared_symbols()
    assert len(syms) >= 12
    h = C.CDLL(L.LIB_PATH)
    for s in syms:
        assert hasattr(h, s), "missing export %s" % s
        assert s in L.PROTOTYPES, "header symbol %s has no ctypes prototype" % s
    for s in L.PROTOTYPES:
        assert s in syms, "ctypes prototype %s is not declared in the header" % s


def test_struct_layout_matches_header_constants():
    src = open(HEADER).read()
    for name, val in [("QNNB_KIND_U8", L.KIND_U8), ("QNNB_KIND_I8", L.KIND_I8), ("QNNB_KIND_B1", L.KIND_B1),
                      ("QNNB_KIND_F32", L.KIND_F32), ("QNNB_ACT_QUANT", L.ACT_QUANT), ("QNNB_ACT_SIGN", L.ACT_SIGN),
                      ("QNNB_ACT_LEAKY", L.ACT_LEAKY), ("QNNB_ACT_SIGN_I8", L.ACT_SIGN_I8), ("QNNB_W_TERNARY", L.W_TERNARY), ("QNNB_WFMT_B1", L.WFMT_B1),
                      ("QNNB_IMPL_TCGEN05", L.IMPL_TCGEN05)]:
        m = re.search(r"#define\s+%s\s+(-?\d+)" % name, src)
        assert m and int(m.group(1)) == val, name
    # field order of the epilogue struct
    body = re.search(r"typedef struct qnnb_epilogue \{(.*?)\} qnnb_epilogue;", src, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"(\w+)\s*;", body)
    assert fields == [f[0] for f in L.Epilogue._fields_]
    # sizes / offsets as the C compiler sees them
    import subprocess
    import tempfile
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "qnnb200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(qnnb_epilogue), sizeof(qnnb_conv_desc), sizeof(qnnb_dense_desc),
         offsetof(qnnb_conv_desc, epi), offsetof(qnnb_dense_desc, epi), offsetof(qnnb_epilogue, residual),
         offsetof(qnnb_epilogue, pool), offsetof(qnnb_dense_desc, avg_positions));
  printf("%zu %zu %zu %zu %zu %zu %d\n", sizeof(qnnb_net_conv), sizeof(qnnb_vgg_desc), offsetof(qnnb_net_conv, epi),
         offsetof(qnnb_vgg_desc, conv), offsetof(qnnb_vgg_desc, dense_w), offsetof(qnnb_vgg_desc, dense_epi), QNNB_NET_MAX_CONVS);
  printf("%zu %zu %zu %zu %zu\n", offsetof(qnnb_conv_desc, w_f32), offsetof(qnnb_conv_desc, max_ctas), offsetof(qnnb_dense_desc, w_f32),
         offsetof(qnnb_dense_desc, max_ctas), offsetof(qnnb_vgg_desc, max_ctas));
  return 0;
}'''
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "t.c"), "w").write(prog)
        subprocess.check_call(["gcc", "-I", os.path.dirname(HEADER), "-o", os.path.join(td, "t"), os.path.join(td, "t.c")])
        got = [int(v) for v in subprocess.check_output([os.path.join(td, "t")]).split()]
    want = [C.sizeof(L.Epilogue), C.sizeof(L.ConvDesc), C.sizeof(L.DenseDesc), L.ConvDesc.epi.offset,
            L.DenseDesc.epi.offset, L.Epilogue.residual.offset, L.Epilogue.pool.offset, L.DenseDesc.avg_positions.offset,
            C.sizeof(L.NetConv), C.sizeof(L.VggDesc), L.NetConv.epi.offset, L.VggDesc.conv.offset, L.VggDesc.dense_w.offset,
            L.VggDesc.dense_epi.offset, L.NET_MAX_CONVS,
            L.ConvDesc.w_f32.offset, L.ConvDesc.max_ctas.offset, L.DenseDesc.w_f32.offset, L.DenseDesc.max_ctas.offset,
            L.VggDesc.max_ctas.offset]
    assert got == want


def test_argument_errors_surface_without_a_gpu():
    h = L.lib()
    assert h.qnnb_packed_weight_bytes(L.WFMT_I8, 3, 3, 3, 64) == 64 * 9 * 4
    assert h.qnnb_packed_weight_bytes(L.WFMT_B1, 3, 3, 64, 128) == 128 * 9 * 2 * 4
    d = L.ConvDesc()
    d.n, d.h, d.w, d.cin, d.cout, d.kh, d.kw, d.stride = 1, 8, 8, 4, 4, 5, 5, 1
    d.epi.acc_scale = 1.0
    d.epi.res_kind = L.KIND_NONE
    oh, ow = C.c_int32(), C.c_int32()
    rc = h.qnnb_conv2d_out_shape(C.byref(d), C.byref(oh), C.byref(ow))
    assert rc == L.EINVAL
    assert b"kernel 5x5" in h.qnnb_last_error()
    with pytest.raises(L.QnnbError):
        L.check(rc)
    d.kh = d.kw = 3
    d.stride = 2
    d.epi.pool = 2
    d.epi.act = L.ACT_QUANT
    d.epi.abits = 4
    assert h.qnnb_conv2d_out_shape(C.byref(d), C.byref(oh), C.byref(ow)) == 0
    assert (oh.value, ow.value) == (2, 2)


def test_no_cpu_fallback():
    import numpy as np
    import torch
    from helpers import make_cf
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = q.build_model(make_cf())
    with pytest.raises(RuntimeError):
        m.predict(np.zeros((1, 32, 32, 3), np.uint8))
    with pytest.raises(ValueError):
        L.ptr(torch.zeros(4))


def _desc(n, h, w, cin, cout, k=3, stride=1, in_kind=None, act=None, abits=4, pool=0, scale=1.0 / 64, res_kind=None):
    d = L.ConvDesc()
    d.n, d.h, d.w, d.cin, d.cout, d.kh, d.kw, d.stride = n, h, w, cin, cout, k, k, stride
    d.in_kind = L.KIND_I8 if in_kind is None else in_kind
    d.impl = L.IMPL_AUTO
    d.epi.acc_scale = scale
    d.epi.res_kind = L.KIND_NONE if res_kind is None else res_kind
    d.epi.residual = 0x1000 if res_kind is not None else None       # any non-null value: the query never dereferences it
    d.epi.act = L.ACT_QUANT if act is None else act
    d.epi.abits = abits
    d.epi.pool = pool
    d.epi.res_mul = 1.0
    return d


def test_kernel_selection_query_is_pure_host_logic():
    """qnnb_conv2d_tc_supported (which the plan uses to pick the storage form of +-1 maps) needs no GPU: check the
    shapes of BASELINE.json's configs land on the kernels DESIGN.md says they do."""
    h = L.lib()
    q = lambda d: h.qnnb_conv2d_tc_supported(C.byref(d))
    # cfg3 / cfg4 interior layers: int8 implicit GEMM on tcgen05 (K1)
    assert q(_desc(1024, 16, 16, 64, 128, pool=2)) == 1
    assert q(_desc(1024, 8, 8, 128, 256, pool=2)) == 1
    assert q(_desc(4096, 32, 32, 256, 256, abits=8, scale=1.0 / (128 * 128))) == 1
    # first layer: uint8 RGB, 32 wide (K5); and its +-1 (full-bnn) form with int8 levels out
    assert q(_desc(1024, 32, 32, 3, 64, in_kind=L.KIND_U8, pool=2, scale=1.0 / (255 * 8))) == 1
    assert q(_desc(256, 32, 32, 3, 64, in_kind=L.KIND_U8, act=L.ACT_SIGN_I8, pool=2, scale=1.0 / 255)) == 1
    assert q(_desc(256, 16, 16, 64, 128, act=L.ACT_SIGN_I8, pool=2, scale=1.0)) == 1
    # bit-packed +-1 output, MNIST maps, narrow layers, strides, residuals, non-power-of-two scales: generic kernel
    assert q(_desc(256, 16, 16, 64, 128, act=L.ACT_SIGN, pool=2, scale=1.0)) == 0
    assert q(_desc(100, 14, 14, 64, 64, abits=2, pool=2)) == 0
    assert q(_desc(100, 28, 28, 1, 64, in_kind=L.KIND_U8, abits=2, pool=2)) == 0
    assert q(_desc(1024, 32, 32, 16, 16)) == 0
    assert q(_desc(1024, 32, 32, 64, 128, stride=2)) == 0
    assert q(_desc(1024, 16, 16, 64, 128, res_kind=L.KIND_I8)) == 0
    assert q(_desc(1024, 16, 16, 64, 128, scale=0.3)) == 0
    # fp32 activations (qnn / bnn / tnn nets): bf16-split tensor-core kernel (K4) for 16..64 channels, 3x3 stride 1
    f32 = dict(in_kind=L.KIND_F32, act=L.ACT_LEAKY)
    assert q(_desc(1024, 32, 32, 16, 16, **f32)) == 1
    assert q(_desc(1024, 16, 16, 32, 32, res_kind=L.KIND_F32, **f32)) == 1
    assert q(_desc(1024, 8, 8, 64, 64, in_kind=L.KIND_F32, act=L.ACT_NONE)) == 1
    assert q(_desc(1024, 32, 32, 16, 32, stride=2, **f32)) == 0
    assert q(_desc(1024, 32, 32, 3, 16, **f32)) == 0
    assert q(_desc(1024, 16, 16, 128, 128, **f32)) == 0
    assert q(_desc(1024, 16, 16, 32, 32, pool=2, **f32)) == 0
    # malformed descriptors are "not supported", never a crash
    assert q(_desc(1024, 16, 16, 64, 128, k=5)) == 0
    assert h.qnnb_conv2d_tc_supported(None) == 0


def test_dense_rejects_pooled_integer_input_without_a_gpu():
    h = L.lib()
    d = L.DenseDesc()
    d.n, d.fin, d.units, d.in_kind, d.softmax = 4, 64, 10, L.KIND_I8, 0
    d.epi.acc_scale = 1.0
    d.epi.res_kind = L.KIND_NONE
    d.avg_positions = 64
    rc = h.qnnb_dense(C.byref(d), C.c_void_p(0x1000), C.c_void_p(0x1000), C.c_void_p(0x1000), None, None)
    assert rc == L.EINVAL and b"avg_positions" in h.qnnb_last_error()


def _vgg_desc(h, w, cin, filters, pools, abits=2, units=10, n=100):
    d = L.VggDesc()
    d.n, d.h, d.w, d.cin, d.nconv, d.units = n, h, w, cin, len(filters), units
    for i, (f, pl) in enumerate(zip(filters, pools)):
        d.conv[i].cout, d.conv[i].pool, d.conv[i].w = f, pl, 0x1000
        d.conv[i].epi.acc_scale = 1.0
        d.conv[i].epi.res_kind = L.KIND_NONE
        d.conv[i].epi.act, d.conv[i].epi.abits = L.ACT_QUANT, abits
    d.dense_w = 0x1000
    d.dense_epi.acc_scale = 1.0
    d.dense_epi.res_kind = L.KIND_NONE
    return d


def test_whole_network_scope_query_is_pure_host_logic():
    """qnnb_vgg_forward_supported: which nets run as ONE launch (csrc/net_fused.cu) -- BASELINE config 1 and the
    64/64/64 CIFAR-10 variant do, the headline 64/128/256 net (288 KB of kernels in the last conv) does not."""
    h = L.lib()
    q = lambda d: h.qnnb_vgg_forward_supported(C.byref(d))
    assert q(_vgg_desc(28, 28, 1, [64, 64, 64], [2, 2, 2])) == 1                    # cfg1
    assert q(_vgg_desc(32, 32, 3, [64, 64, 64], [2, 2, 2], abits=4)) == 1           # config/config_CIFAR-10.py
    assert q(_vgg_desc(32, 32, 3, [32, 64, 32], [2, 2, 2], abits=8)) == 1
    assert q(_vgg_desc(28, 28, 1, [32, 32, 64, 64], [0, 2, 2, 2])) == 1             # nla = 2: an un-pooled 28x28 layer
    assert q(_vgg_desc(32, 32, 3, [64, 64, 64, 64], [2, 0, 2, 2], abits=4)) == 1    # nlb = 2
    assert q(_vgg_desc(32, 32, 3, [64, 128, 256], [2, 2, 2], abits=4)) == 0         # cfg3
    assert q(_vgg_desc(32, 32, 3, [64, 64, 64, 64, 64, 64], [0, 0, 2, 0, 2, 2])) == 0   # maps + kernels > 227 KB
    assert q(_vgg_desc(64, 64, 3, [64, 64, 64], [2, 2, 2])) == 0
    assert q(_vgg_desc(28, 28, 1, [64, 64, 64], [2, 2, 2], units=40)) == 0
    d = _vgg_desc(28, 28, 1, [64, 64, 64], [2, 2, 2])
    d.conv[1].epi.act = L.ACT_LEAKY
    assert q(d) == 0
    assert h.qnnb_vgg_forward_supported(None) == 0
    # argument errors surface as status codes
    d = _vgg_desc(28, 28, 1, [64, 64, 64], [2, 2, 2])
    d.nconv = 9
    assert h.qnnb_vgg_forward(C.byref(d), C.c_void_p(0x1000), C.c_void_p(0x1000), C.c_void_p(0x1000), None) == L.EINVAL
    d = _vgg_desc(32, 32, 3, [64, 128, 256], [2, 2, 2], abits=4)
    assert h.qnnb_vgg_forward(C.byref(d), C.c_void_p(0x1000), C.c_void_p(0x1000), C.c_void_p(0x1000), None) == L.EUNSUPPORTED
    assert b"whole-network" in h.qnnb_last_error()
    assert h.qnnb_vgg_blob_bytes(C.byref(d)) == 0
    # resident image of cfg1: (1 + 18 + 18 + 1) A blocks of 2 KB, the 10 x 576 dense kernel, 3 x 64 + 10 constant rows
    d = _vgg_desc(28, 28, 1, [64, 64, 64], [2, 2, 2])
    assert h.qnnb_vgg_blob_bytes(C.byref(d)) == 38 * 2048 + 5760 + 192 * 16 + 160
