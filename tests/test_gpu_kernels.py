"""GPU parity of every kernel against the exact oracle (O1), through the C ABI.
Bit-exact for integer/bit outputs and for fp32 outputs produced from integer accumulators;
tolerance (stated per test) only where the ACCUMULATION itself is floating point."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import exact  # noqa: E402
from helpers import oracle_layer, unpack_bits  # noqa: E402

F32 = np.float32


def _seed(obj):
    import zlib
    return zlib.crc32(repr(obj).encode())


def _mods():
    import qnn_b200 as q
    from qnn_b200 import _lib as L, kernels as K
    return q, L, K


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


# --------------------------------------------------------------------------- K0 packers
@pytest.mark.parametrize("shape", [(3, 3, 3, 64), (3, 3, 64, 128), (1, 1, 16, 32), (3, 3, 5, 7), (1, 1, 4096, 10), (1, 1, 70, 3)])
@pytest.mark.parametrize("mode", ["q2", "q4", "q8", "bin", "binH", "ter"])
def test_pack_weights_matches_oracle(shape, mode):
    q, L, K = _mods()
    rng = np.random.default_rng(_seed((shape, mode)))
    w = rng.uniform(-1, 1, size=shape).astype(F32)
    # adversarial values: exact half-way points, zero, the binary threshold, saturation
    flat = w.reshape(-1)
    specials = np.array([0.0, -0.0, 2.0 ** -24, np.nextafter(F32(2.0 ** -24), F32(1)), 0.5, -0.5, 0.25, -0.25, 0.125,
                         0.0625, 0.1875, -0.1875, 1.0, -1.0, 0.99999, 0.9375, 0.96875, 3 / 256, 5 / 256, -3 / 256], F32)
    flat[: min(len(specials), flat.size)] = specials[: flat.size]
    H = 1.0
    if mode.startswith("q"):
        nb = int(mode[1])
        lv = exact.quantize_levels(w, nb)
        got = K.pack_weights(dev(w), L.W_QUANT, nb, 1.0, L.WFMT_I8)
    elif mode.startswith("bin"):
        H = 0.75 if mode == "binH" else 1.0
        lv = exact.binarize_levels(w, H)
        got = K.pack_weights(dev(w), L.W_BINARY, 1, H, L.WFMT_I8)
        gotb = K.pack_weights(dev(w), L.W_BINARY, 1, H, L.WFMT_B1).cpu().numpy()
        wantb = exact.pack_bits_lastdim(np.transpose(lv, (3, 0, 1, 2)))
        assert np.array_equal(gotb.astype(np.uint32), wantb)
    else:
        lv = exact.ternarize_levels(w, 1.0)
        got = K.pack_weights(dev(w), L.W_TERNARY, 2, 1.0, L.WFMT_I8)
    got = got.cpu().numpy()
    kh, kw, cin, cout = shape
    want = np.zeros((cout, kh, kw, (cin + 3) // 4 * 4), np.int8)
    want[..., :cin] = np.transpose(lv, (3, 0, 1, 2))
    assert np.array_equal(got, want)


def test_quantiser_ops_known_answers():
    """Hand-derivable KATs: level tables, half-to-even, binary threshold, ternary asymmetry."""
    q, L, K = _mods()
    from qnn_b200.layers import quantized_ops as qo, binary_ops as bo, ternary_ops as to
    x = dev(np.array([[-1.0, -0.75, -0.5, -0.26, -0.25, 0.0, 0.24, 0.25, 0.5, 0.75, 0.76, 1.0]], F32))
    got = qo.quantize(x, nb=2).cpu().numpy()[0]
    #  x*2: -2,-1.5,-1,-.52,-.5,0,.48,.5,1,1.5,1.52,2 -> rint half-even: -2,-2,-1,-1,-0,0,0,0,1,2,2,2 -> clip[-2,1] /2
    assert np.array_equal(got, np.array([-1, -1, -.5, -.5, 0, 0, 0, 0, .5, .5, .5, .5], F32))
    assert np.array_equal(qo.quantized_tanh(x, nb=2).cpu().numpy()[0], got)
    t = F32(2.0 ** -24)
    b = dev(np.array([[0.0, -0.0, t, np.nextafter(t, F32(1)), -1e-3, 1e-3, 1.0, -1.0]], F32))
    assert np.array_equal(bo.binary_tanh(b).cpu().numpy()[0], np.array([-1, -1, -1, 1, -1, 1, 1, -1], F32))
    assert np.array_equal(bo.binarize(b, H=1).cpu().numpy()[0], np.array([-1, -1, -1, 1, -1, 1, 1, -1], F32))
    w = np.array([[1.0, -1.0, 0.7, -0.7, 0.1, -0.1, 0.0, 0.5]], F32)        # mean|w| = 0.5125, cutoff = 0.35875
    got = to.ternarize(dev(w), H=1).cpu().numpy()[0]
    assert np.array_equal(got, np.array([1, -1, 1, -1, 0, 0, 0, 1], F32))
    w2 = np.array([[0.5, -0.5, 0.5, -0.5]], F32)                             # cutoff 0.35
    assert np.array_equal(to.ternarize(dev(w2), H=1).cpu().numpy()[0], np.array([1, -1, 1, -1], F32))
    r = dev(np.array([[0.5, 1.5, 2.5, -0.5, -1.5, 0.49999997, 1e9]], F32))
    assert np.array_equal(qo.round_through(r).cpu().numpy()[0], np.array([0, 2, 2, -0., -2, 0, 1e9], F32))


# --------------------------------------------------------------------------- fused conv
def _rand_input(rng, kind, shape, abits=4):
    if kind == "u8":
        return rng.integers(0, 256, size=shape, dtype=np.uint8), 1.0 / 255.0
    if kind == "i8":
        m = 1 << (abits - 1)
        return rng.integers(-m, m, size=shape).astype(np.int8), 1.0 / m
    if kind == "b1":
        return (rng.integers(0, 2, size=shape) * 2 - 1).astype(np.int8), 1.0
    return rng.normal(0, 1, size=shape).astype(F32), 1.0


def _to_qtensor(K, kind, x, scale, abits=4):
    if kind == "b1":
        words = exact.pack_bits_lastdim(x).astype(np.int32)
        return K.QTensor("b1", dev(words), 1.0, int(x.shape[-1]))
    return K.QTensor(kind, dev(x), scale, int(x.shape[-1]))


CONV_CASES = [
    # kind, n, h, w, cin, cout, k, stride, wkind, nb, act, pool, bn, bias, res
    ("u8", 3, 32, 32, 3, 64, 3, 1, "quantized", 4, "quant", True, True, True, None),
    ("u8", 2, 28, 28, 1, 64, 3, 1, "quantized", 2, "quant", True, True, True, None),
    ("u8", 2, 32, 32, 3, 16, 3, 1, "quantized", 4, "quant", False, True, False, None),
    ("u8", 2, 32, 32, 3, 64, 3, 1, "binary", 1, "binary", True, True, True, None),
    ("u8", 2, 32, 32, 3, 16, 3, 1, "quantized", 4, "leaky", False, True, False, None),
    ("i8", 3, 16, 16, 64, 128, 3, 1, "quantized", 4, "quant", True, True, True, None),
    ("i8", 2, 14, 14, 64, 64, 3, 1, "quantized", 2, "quant", True, True, True, None),
    ("i8", 2, 7, 7, 64, 64, 3, 1, "quantized", 2, "quant", True, True, True, None),
    ("i8", 2, 8, 8, 128, 256, 3, 1, "quantized", 8, "quant", True, True, True, None),
    ("i8", 2, 32, 32, 16, 16, 3, 1, "quantized", 4, "quant", False, True, False, "i8"),
    ("i8", 2, 32, 32, 16, 32, 3, 2, "quantized", 4, "quant", False, True, False, None),
    ("i8", 2, 32, 32, 16, 32, 1, 2, "quantized", 4, None, False, False, False, None),
    ("i8", 2, 16, 16, 32, 32, 3, 1, "quantized", 4, "quant", False, True, True, "f32"),
    ("i8", 2, 9, 11, 20, 7, 3, 1, "ternary", 2, "quant", False, True, True, None),
    ("i8", 2, 10, 10, 24, 40, 3, 2, "binary", 1, "quant", True, True, True, None),
    ("i8", 1, 8, 8, 64, 64, 3, 1, "quantized", 8, None, False, False, True, None),
    ("b1", 3, 16, 16, 64, 128, 3, 1, "binary", 1, "binary", True, True, True, None),
    ("b1", 2, 8, 8, 128, 256, 3, 1, "binary", 1, "binary", True, True, True, None),
    ("b1", 2, 12, 10, 40, 70, 3, 1, "binary", 1, "binary", False, True, False, None),
    ("b1", 2, 16, 16, 32, 32, 3, 2, "binary", 1, "quant", False, True, False, None),
    ("b1", 2, 16, 16, 32, 64, 1, 2, "binary", 1, None, False, False, True, None),
]


@pytest.mark.parametrize("case", CONV_CASES, ids=[("%s_%dx%d_%d-%d_k%ds%d_%s%d_%s%s%s" % (c[0], c[2], c[3], c[4], c[5], c[6], c[7], c[8][:3], c[9], c[10], "_pool" if c[11] else "", ("_res" + c[14]) if c[14] else "")) for c in CONV_CASES])
@pytest.mark.parametrize("impl", ["generic"])
def test_conv2d_integer_paths_bit_exact(case, impl):
    q, L, K = _mods()
    kind, n, h, w, cin, cout, k, stride, wkind, nb, act, pool, use_bn, use_bias, res = case
    rng = np.random.default_rng(_seed(case))
    abits = 8 if nb == 8 else (2 if nb == 2 else 4)
    x, xs = _rand_input(rng, kind, (n, h, w, cin), abits)
    kernel = rng.uniform(-1, 1, size=(k, k, cin, cout)).astype(F32)
    fan = k * k * cin
    bias = rng.uniform(-0.3, 0.3, size=cout).astype(F32) if use_bias else None
    bn = None
    if use_bn:
        var_scale = fan * (0.11 if kind != "b1" else 1.0)
        bn = (rng.uniform(0.3, 0.9, cout).astype(F32) * rng.choice([1, 1, -1], cout).astype(F32),
              rng.uniform(-0.2, 0.2, cout).astype(F32),
              (rng.uniform(-0.2, 0.2, cout) * np.sqrt(var_scale)).astype(F32),
              (rng.uniform(0.5, 1.5, cout) * var_scale).astype(F32))
    oh, ow = -(-h // stride), -(-w // stride)
    residual = res_q = None
    if res == "i8":
        rl = rng.integers(-8, 8, size=(n, oh, ow, cout)).astype(np.int8)
        residual = rl.astype(F32) * F32(1 / 8)
        res_q = K.QTensor("i8", dev(rl), 1 / 8, cout)
    elif res == "f32":
        residual = rng.normal(0, 0.5, size=(n, oh, ow, cout)).astype(F32)
        res_q = K.QTensor("f32", dev(residual), 1.0, cout)
    want, _ = oracle_layer(x, kind, xs, kernel, wkind, nb, 1.0, stride, bias=bias, bn=bn, eps=1e-4, residual=residual,
                           res_mul=0.5, act=act, abits=abits, pool=pool)
    mode = {"quantized": L.W_QUANT, "binary": L.W_BINARY, "ternary": L.W_TERNARY}[wkind]
    wfmt = L.WFMT_B1 if kind == "b1" else L.WFMT_I8
    wp = K.pack_weights(dev(kernel), mode, nb, 1.0, wfmt)
    wscale = 1.0 / (1 << (nb - 1)) if wkind == "quantized" else 1.0
    inv = shift = None
    if bn is not None:
        i_, s_ = K.bn_constants(*bn, 1e-4)
        inv, shift = dev(i_), dev(s_)
    actc = {"quant": L.ACT_QUANT, "binary": L.ACT_SIGN, "leaky": L.ACT_LEAKY, None: L.ACT_NONE}[act]
    epi = K.make_epilogue(K.acc_scale(xs, wscale), bias=dev(bias) if bias is not None else None, bn_inv=inv, bn_shift=shift,
                          residual=res_q, res_mul=0.5, act=actc, abits=abits, leaky_alpha=0.3, pool=2 if pool else 0)
    implc = {"generic": L.IMPL_GENERIC, "auto": L.IMPL_AUTO}[impl]
    y = K.conv2d(_to_qtensor(K, kind, x, xs, abits), wp, k, k, cout, stride, epi, impl=implc)
    torch.cuda.synchronize()
    got = y.data.cpu().numpy()
    if act == "binary":
        got = unpack_bits(got, cout)
    assert got.shape == want.shape
    if got.dtype == np.float32:
        assert np.array_equal(got, want.astype(F32)), "max abs diff %g" % np.abs(got - want).max()
    else:
        assert np.array_equal(got.astype(np.int32), want.astype(np.int32)), "mismatching levels: %d" % (got != want).sum()


@pytest.mark.parametrize("case", [
    (2, 32, 32, 3, 16, 3, 1, "quantized", 4, "leaky", False),
    (2, 16, 16, 16, 32, 3, 2, "quantized", 4, "leaky", False),
    (2, 16, 16, 16, 32, 1, 2, "ternary", 2, None, False),
    (2, 16, 16, 32, 32, 3, 1, "binary", 1, "leaky", True),
    (2, 8, 8, 64, 64, 3, 1, "quantized", 4, "quant", False),
])
def test_conv2d_float_activations_tolerance(case):
    """fp32 activations x exact integer weights: accumulation order differs from the float64 oracle, so the
    bar is the north_star tolerance: max abs error <= 1e-4 relative to the largest |output|."""
    q, L, K = _mods()
    n, h, w, cin, cout, k, stride, wkind, nb, act, pool = case
    rng = np.random.default_rng(_seed(case))
    x = rng.normal(0, 1, size=(n, h, w, cin)).astype(F32)
    kernel = rng.uniform(-1, 1, size=(k, k, cin, cout)).astype(F32)
    fan = k * k * cin
    bn = (rng.uniform(0.3, 0.9, cout).astype(F32), rng.uniform(-0.2, 0.2, cout).astype(F32),
          rng.uniform(-0.2, 0.2, cout).astype(F32), (rng.uniform(0.5, 1.5, cout) * fan * 0.3).astype(F32))
    want, _ = oracle_layer(x, "f32", 1.0, kernel, wkind, nb, 1.0, stride, bn=bn, eps=1e-3, act=act, abits=4, pool=pool)
    mode = {"quantized": L.W_QUANT, "binary": L.W_BINARY, "ternary": L.W_TERNARY}[wkind]
    wp = K.pack_weights(dev(kernel), mode, nb, 1.0, L.WFMT_I8)
    wscale = 1.0 / (1 << (nb - 1)) if wkind == "quantized" else 1.0
    i_, s_ = K.bn_constants(*bn, 1e-3)
    actc = {"quant": L.ACT_QUANT, "leaky": L.ACT_LEAKY, None: L.ACT_NONE}[act]
    epi = K.make_epilogue(F32(wscale), bn_inv=dev(i_), bn_shift=dev(s_), act=actc, abits=4, pool=2 if pool else 0)
    y = K.conv2d(K.QTensor("f32", dev(x), 1.0, cin), wp, k, k, cout, stride, epi, impl=L.IMPL_GENERIC)
    got = y.data.cpu().numpy()
    if act == "quant":
        # a re-quantised output may flip by one level where the fp32 sum lands on a rounding boundary
        diff = np.abs(got.astype(np.int32) - want.astype(np.int32))
        assert diff.max() <= 1 and (diff > 0).mean() <= 1e-3
    else:
        tol = 1e-4 * np.abs(want).max()
        assert np.abs(got - want).max() <= tol


# --------------------------------------------------------------------------- dense head
@pytest.mark.parametrize("kind,fin,units,wkind,nb,bn,softmax", [
    ("i8", 4096, 10, "quantized", 4, True, False),
    ("i8", 576, 10, "quantized", 2, True, False),
    ("i8", 4096, 10, "quantized", 8, True, False),
    ("i8", 64, 10, "quantized", 4, False, True),
    ("i8", 1024, 17, "ternary", 2, True, False),
    ("b1", 4096, 10, "binary", 1, True, False),
    ("b1", 256, 32, "binary", 1, False, False),
    ("f32", 64, 10, "quantized", 4, False, True),
    ("f32", 300, 10, "binary", 1, True, False),
])
def test_dense_head(kind, fin, units, wkind, nb, bn, softmax):
    q, L, K = _mods()
    rng = np.random.default_rng(fin * 31 + units)
    n = 37
    abits = nb if wkind == "quantized" else 4
    x, xs = _rand_input(rng, kind, (n, fin), abits)
    kernel = rng.uniform(-1, 1, size=(fin, units)).astype(F32)
    bias = rng.uniform(-0.3, 0.3, size=units).astype(F32)
    bnw = None
    if bn:
        bnw = (rng.uniform(0.5, 1.5, units).astype(F32), rng.uniform(-0.2, 0.2, units).astype(F32),
               rng.uniform(-1, 1, units).astype(F32), (rng.uniform(0.5, 1.5, units) * fan_var(fin, kind)).astype(F32))
    want, _ = oracle_layer(x, kind, xs, kernel, wkind, nb, 1.0, 1, bias=bias, bn=bnw, eps=1e-4, dense=True)
    mode = {"quantized": L.W_QUANT, "binary": L.W_BINARY, "ternary": L.W_TERNARY}[wkind]
    wfmt = L.WFMT_B1 if kind == "b1" else L.WFMT_I8
    wp = K.pack_weights(dev(kernel), mode, nb, 1.0, wfmt)
    wscale = 1.0 / (1 << (nb - 1)) if wkind == "quantized" else 1.0
    inv = shift = None
    if bnw is not None:
        i_, s_ = K.bn_constants(*bnw, 1e-4)
        inv, shift = dev(i_), dev(s_)
    epi = K.make_epilogue(K.acc_scale(xs if kind != "f32" else 1.0, wscale), bias=dev(bias), bn_inv=inv, bn_shift=shift)
    out, logits = K.dense(_to_qtensor(K, kind, x, xs, abits), wp, units, epi, softmax=softmax, want_logits=softmax)
    got = out.cpu().numpy()
    if softmax:
        lg = logits.cpu().numpy()
        if kind == "f32":
            assert np.abs(lg - want).max() <= 1e-4 * np.abs(want).max()
        else:
            assert np.array_equal(lg, want)
        assert np.abs(got - exact.softmax64(lg)).max() <= 2e-6
        assert np.array_equal(got.argmax(1), lg.argmax(1))
    elif kind == "f32":
        assert np.abs(got - want).max() <= 1e-4 * np.abs(want).max()
    else:
        assert np.array_equal(got, want)


def fan_var(fin, kind):
    return fin * (1.0 if kind == "b1" else 0.11)


# --------------------------------------------------------------------------- stand-alone fp32 ops
def test_standalone_ops():
    q, L, K = _mods()
    rng = np.random.default_rng(3)
    x = rng.normal(0, 0.6, size=(5, 6, 8, 70)).astype(F32)
    x.reshape(-1)[:6] = [0.0, 2.0 ** -24, 0.0625, -0.0625, 0.1875, 5.0]
    for ab in (2, 4, 8):
        got = K.quantize_act(dev(x), ab)
        assert np.array_equal(got.data.cpu().numpy().astype(np.int32), exact.act_quant_levels(x, ab))
        assert np.array_equal(got.to_float().cpu().numpy(), exact.act_quant_levels(x, ab).astype(F32) * F32(1 / (1 << (ab - 1))))
    sb = K.sign_act(dev(x))
    assert np.array_equal(sb.data.cpu().numpy().astype(np.uint32), exact.pack_bits_lastdim(exact.act_binary_levels(x)))
    assert np.array_equal(sb.to_float().cpu().numpy(), exact.act_binary_levels(x).astype(F32))
    inv = rng.uniform(0.5, 1.5, 70).astype(F32)
    shift = rng.uniform(-1, 1, 70).astype(F32)
    want = ((x * inv).astype(F32) + shift).astype(F32)
    assert np.array_equal(K.batchnorm(dev(x), dev(inv), dev(shift)).cpu().numpy(), want)
    assert np.array_equal(K.maxpool2(dev(x)).cpu().numpy(), exact.maxpool2(x))
    assert np.array_equal(K.leaky(dev(x), 0.3).cpu().numpy(), exact.leaky(x))


def test_empty_batch_and_errors():
    q, L, K = _mods()
    wp = K.pack_weights(dev(np.zeros((3, 3, 4, 8), F32)), L.W_QUANT, 4, 1.0, L.WFMT_I8)
    epi = K.make_epilogue(1.0, act=L.ACT_QUANT, abits=4)
    y = K.conv2d(K.QTensor("i8", torch.zeros((0, 8, 8, 4), dtype=torch.int8, device="cuda"), 0.125, 4), wp, 3, 3, 8, 1, epi)
    assert tuple(y.data.shape) == (0, 8, 8, 8)
    with pytest.raises(L.QnnbError):
        K.conv2d(K.QTensor("i8", torch.zeros((1, 8, 8, 4), dtype=torch.int8, device="cuda"), 0.125, 4), wp, 3, 3, 8, 3, epi)
    with pytest.raises(L.QnnbError):
        K.pack_weights(dev(np.zeros((3, 3, 4, 8), F32)), L.W_QUANT, 16, 1.0, L.WFMT_I8)
    with pytest.raises(L.QnnbError) as ei:
        K.conv2d(K.QTensor("i8", torch.zeros((1, 9, 9, 4), dtype=torch.int8, device="cuda"), 0.125, 4), wp, 3, 3, 8, 1, epi,
                 impl=L.IMPL_TCGEN05)
    assert ei.value.code == L.EUNSUPPORTED


# --------------------------------------------------------------------------- tcgen05 implicit-GEMM conv (K1)
TC_CASES = [
    # n, h, w, cin, cout, nb, abits, pool, f32_out
    (1, 16, 16, 64, 128, 4, 4, False, True),      # raw accumulators, one tile
    (3, 16, 16, 64, 128, 4, 4, True, False),      # cfg3 layer 2
    (5, 8, 8, 128, 256, 4, 4, True, False),       # cfg3 layer 3, ragged image group (5 = 4 + 1)
    (2, 32, 32, 256, 256, 8, 8, False, False),    # cfg4 block A interior layer
    (2, 32, 32, 256, 256, 8, 8, True, False),     # cfg4 block A last layer
    (3, 16, 16, 256, 256, 8, 8, False, False),
    (9, 8, 8, 256, 256, 8, 8, True, False),
    (2, 32, 32, 64, 128, 2, 2, True, False),
    (40, 16, 16, 128, 128, 4, 4, False, False),   # more tiles than one wave at grid = #SM? (40 tiles) exercises the ring
    (700, 8, 8, 64, 128, 4, 4, True, False),      # 175 tiles > 148 SMs: persistent loop, both TMEM accumulators
]


TC_CASES_V2_ONLY = [
    (3, 24, 40, 64, 128, 4, 4, True, False),      # any multiple of 8: 24 rows -> 8-row blocks of 4 images, 5 column tiles
    (2, 64, 16, 128, 128, 4, 4, False, False),    # tall map: two 32-row tiles per image
    (5, 16, 16, 192, 256, 8, 8, True, False),     # Cin = 192 (three 64-byte chunks), ragged image pair
    (1, 8, 8, 64, 128, 4, 4, False, True),        # raw accumulators, 8x8 (image group of 4 with 3 missing)
]


import functools  # noqa: E402


@functools.lru_cache(maxsize=None)
def _has_v1():
    """Does the loaded library contain the v1 kernel?  (A production build answers QNNB_EUNSUPPORTED for any v1 request.)"""
    q, L, K = _mods()
    x = K.QTensor("i8", torch.zeros((1, 8, 32, 64), dtype=torch.int8, device="cuda"), 0.125, 64)
    wp = torch.zeros((128, 3, 3, 64), dtype=torch.int8, device="cuda")
    try:
        K.conv2d(x, wp, 3, 3, 128, 1, K.make_epilogue(1.0, act=L.ACT_NONE), impl=L.IMPL_TCGEN05_V1)
        return True
    except L.QnnbError as exc:
        if exc.code == L.EUNSUPPORTED:
            return False
        raise


@pytest.mark.parametrize("impl", ["v2", "v1"])
@pytest.mark.parametrize("case", TC_CASES + TC_CASES_V2_ONLY, ids=["n%d_%dx%d_%d-%d_w%da%d%s%s" % (c[0], c[1], c[2], c[3], c[4], c[5], c[6], "_pool" if c[7] else "", "_f32" if c[8] else "") for c in TC_CASES + TC_CASES_V2_ONLY])
def test_conv2d_tcgen05_bit_exact(case, impl):
    q, L, K = _mods()
    if impl == "v1" and case in TC_CASES_V2_ONLY:
        pytest.skip("shape only covered by the halo-resident kernel")
    IMPL = L.IMPL_TCGEN05 if impl == "v2" else L.IMPL_TCGEN05_V1
    if impl == "v1" and not _has_v1():
        pytest.skip("production build: the first-generation kernel is only compiled with make EXTRA=-DQNNB_WITH_V1")
    n, h, w, cin, cout, nb, abits, pool, f32_out = case
    rng = np.random.default_rng(_seed(case))
    x, xs = _rand_input(rng, "i8", (n, h, w, cin), abits)
    kernel = rng.uniform(-1, 1, size=(3, 3, cin, cout)).astype(F32)
    fan = 9 * cin
    wp = K.pack_weights(dev(kernel), L.W_QUANT, nb, 1.0, L.WFMT_I8)
    if f32_out:
        # scale 1, no bias / BN: the fp32 output IS the int32 accumulator (|acc| < 2^24 here)
        epi = K.make_epilogue(1.0, act=L.ACT_NONE)
        y = K.conv2d(K.QTensor("i8", dev(x), xs, cin), wp, 3, 3, cout, 1, epi, impl=IMPL)
        torch.cuda.synchronize()
        got = y.data.cpu().numpy()
        lv = exact.quantize_levels(kernel, nb)
        want = exact.conv_accumulate(x.astype(np.int64), lv, 1).astype(F32)
        bad = np.argwhere(got != want)
        assert bad.shape[0] == 0, "accumulator mismatches: %d of %d, first %s got %s want %s" % (
            bad.shape[0], want.size, bad[:5].tolist(), [got[tuple(b)] for b in bad[:5]], [want[tuple(b)] for b in bad[:5]])
        return
    bias = rng.uniform(-0.3, 0.3, size=cout).astype(F32)
    bn = (rng.uniform(0.3, 0.9, cout).astype(F32) * rng.choice([1, 1, -1], cout).astype(F32),
          rng.uniform(-0.2, 0.2, cout).astype(F32),
          (rng.uniform(-0.2, 0.2, cout) * np.sqrt(fan * 0.11)).astype(F32),
          (rng.uniform(0.5, 1.5, cout) * fan * 0.11).astype(F32))
    want, _ = oracle_layer(x, "i8", xs, kernel, "quantized", nb, 1.0, 1, bias=bias, bn=bn, eps=1e-4, act="quant", abits=abits, pool=pool)
    i_, s_ = K.bn_constants(*bn, 1e-4)
    epi = K.make_epilogue(K.acc_scale(xs, 1.0 / (1 << (nb - 1))), bias=dev(bias), bn_inv=dev(i_), bn_shift=dev(s_),
                          act=L.ACT_QUANT, abits=abits, pool=2 if pool else 0)
    y = K.conv2d(K.QTensor("i8", dev(x), xs, cin), wp, 3, 3, cout, 1, epi, impl=IMPL)
    torch.cuda.synchronize()
    got = y.data.cpu().numpy().astype(np.int32)
    assert got.shape == want.shape
    bad = np.argwhere(got != want)
    assert bad.shape[0] == 0, "level mismatches: %d of %d, first %s" % (bad.shape[0], want.size, bad[:8].tolist())


# --------------------------------------------------------------------------- tcgen05 first layer (K5: uint8 RGB, im2col producers)
K5_CASES = [
    # n, h, cout, nb, abits, pool, f32_out
    (1, 32, 64, 4, 4, False, True),
    (3, 32, 64, 4, 4, True, False),       # cfg3 layer 1 (two pixel groups per tile)
    (2, 32, 64, 8, 8, False, False),      # two pixel groups, un-pooled
    (2, 16, 64, 4, 4, True, False),       # 16 rows = exactly one two-group tile per image
    (2, 32, 256, 8, 8, False, False),     # cfg4 layer 1 (two channel tiles)
    (2, 32, 32, 4, 4, False, False),      # narrow layer: one active lane quarter, runtime row pitch
    (3, 16, 128, 2, 2, True, False),      # non-square map, 32 wide
    (150, 32, 64, 4, 4, True, False),     # 600 tiles: persistent loop
]


@pytest.mark.parametrize("case", K5_CASES, ids=["n%d_h%d_c%d_w%da%d%s%s" % (c[0], c[1], c[2], c[3], c[4], "_pool" if c[5] else "", "_f32" if c[6] else "") for c in K5_CASES])
def test_first_layer_tcgen05_bit_exact(case):
    q, L, K = _mods()
    n, h, cout, nb, abits, pool, f32_out = case
    rng = np.random.default_rng(_seed(("k5",) + case))
    x = rng.integers(0, 256, size=(n, h, 32, 3), dtype=np.uint8)
    x[0, 0, :4] = 255
    x[0, -1, -4:] = 255
    kernel = rng.uniform(-1, 1, size=(3, 3, 3, cout)).astype(F32)
    wp = K.pack_weights(dev(kernel), L.W_QUANT, nb, 1.0, L.WFMT_I8)
    xq = K.QTensor("u8", dev(x), 1.0 / 255.0, 3)
    if f32_out:
        epi = K.make_epilogue(1.0, act=L.ACT_NONE)
        got = K.conv2d(xq, wp, 3, 3, cout, 1, epi, impl=L.IMPL_TCGEN05).data.cpu().numpy()
        want = exact.conv_accumulate(x.astype(np.int64), exact.quantize_levels(kernel, nb), 1).astype(F32)
        bad = np.argwhere(got != want)
        assert bad.shape[0] == 0, "accumulator mismatches: %d of %d, first %s got %s want %s" % (
            bad.shape[0], want.size, bad[:6].tolist(), [got[tuple(b)] for b in bad[:6]], [want[tuple(b)] for b in bad[:6]])
        return
    bias = rng.uniform(-0.3, 0.3, size=cout).astype(F32)
    bn = (rng.uniform(0.3, 0.9, cout).astype(F32) * rng.choice([1, 1, -1], cout).astype(F32),
          rng.uniform(-0.2, 0.2, cout).astype(F32), (rng.uniform(-0.2, 0.2, cout) * np.sqrt(3.0)).astype(F32),
          (rng.uniform(0.5, 1.5, cout) * 3.0).astype(F32))
    want, _ = oracle_layer(x, "u8", 1.0 / 255.0, kernel, "quantized", nb, 1.0, 1, bias=bias, bn=bn, eps=1e-4, act="quant", abits=abits, pool=pool)
    i_, s_ = K.bn_constants(*bn, 1e-4)
    epi = K.make_epilogue(K.acc_scale(1.0 / 255.0, 1.0 / (1 << (nb - 1))), bias=dev(bias), bn_inv=dev(i_), bn_shift=dev(s_),
                          act=L.ACT_QUANT, abits=abits, pool=2 if pool else 0)
    got = K.conv2d(xq, wp, 3, 3, cout, 1, epi, impl=L.IMPL_TCGEN05).data.cpu().numpy().astype(np.int32)
    assert got.shape == want.shape
    bad = np.argwhere(got != want)
    assert bad.shape[0] == 0, "level mismatches: %d of %d, first %s" % (bad.shape[0], want.size, bad[:8].tolist())


# --------------------------------------------------------------------------- +-1 maps as int8 levels (QNNB_ACT_SIGN_I8)
SIGN_I8_CASES = [
    # in_kind, n, h, w, cin, cout, pool, impl
    ("u8", 3, 32, 32, 3, 64, True, "tc"),         # cfg2 layer 1 on the first-layer tcgen05 kernel (two pixel groups)
    ("u8", 2, 32, 32, 3, 128, False, "tc"),
    ("u8", 2, 32, 32, 3, 32, True, "tc"),         # runtime row pitch
    ("i8", 3, 16, 16, 64, 128, True, "tc"),       # cfg2 layer 2: +-1 x +-1 on tcgen05 kind::i8
    ("i8", 5, 8, 8, 128, 256, True, "tc"),        # cfg2 layer 3
    ("i8", 2, 32, 32, 256, 128, False, "tc"),
    ("i8", 2, 12, 10, 40, 70, False, "generic"),  # any shape on the CUDA-core kernel
    ("i8", 2, 16, 16, 64, 128, True, "generic"),
]


@pytest.mark.parametrize("case", SIGN_I8_CASES, ids=["%s_n%d_%dx%d_%d-%d%s_%s" % (c[0], c[1], c[2], c[3], c[4], c[5], "_pool" if c[6] else "", c[7]) for c in SIGN_I8_CASES])
def test_conv2d_sign_to_int8_levels_bit_exact(case):
    """BinaryConv2D + BN + binary_tanh (+ pool) with the +-1 result stored as int8 levels: same decisions as the
    bit-packed form (oracle), on the tcgen05 kernels and on the generic kernel."""
    q, L, K = _mods()
    kind, n, h, w, cin, cout, pool, impl = case
    rng = np.random.default_rng(_seed(("sign8",) + case))
    if kind == "u8":
        x, xs = _rand_input(rng, "u8", (n, h, w, cin))
    else:
        x, xs = _rand_input(rng, "b1", (n, h, w, cin))          # int8 +-1 levels, scale 1
    kernel = rng.uniform(-1, 1, size=(3, 3, cin, cout)).astype(F32)
    fan = 9 * cin
    bias = rng.uniform(-0.3, 0.3, size=cout).astype(F32)
    var_scale = fan * (0.11 if kind == "u8" else 1.0)
    bn = (rng.uniform(0.3, 0.9, cout).astype(F32) * rng.choice([1, 1, -1], cout).astype(F32),
          rng.uniform(-0.2, 0.2, cout).astype(F32),
          (rng.uniform(-0.2, 0.2, cout) * np.sqrt(var_scale)).astype(F32),
          (rng.uniform(0.5, 1.5, cout) * var_scale).astype(F32))
    want, _ = oracle_layer(x, kind if kind == "u8" else "b1", xs, kernel, "binary", 1, 1.0, 1, bias=bias, bn=bn, eps=1e-4,
                           act="binary", abits=4, pool=pool)
    wp = K.pack_weights(dev(kernel), L.W_BINARY, 1, 1.0, L.WFMT_I8)
    i_, s_ = K.bn_constants(*bn, 1e-4)
    epi = K.make_epilogue(K.acc_scale(xs, 1.0), bias=dev(bias), bn_inv=dev(i_), bn_shift=dev(s_), act=L.ACT_SIGN_I8,
                          pool=2 if pool else 0)
    xq = K.QTensor(kind, dev(x), xs, cin)
    IMPL = L.IMPL_TCGEN05 if impl == "tc" else L.IMPL_GENERIC
    assert K.conv2d_on_tensor_cores(xq, 3, 3, cout, 1, epi, IMPL) == (impl == "tc")
    y = K.conv2d(xq, wp, 3, 3, cout, 1, epi, impl=IMPL)
    torch.cuda.synchronize()
    assert y.kind == "i8" and y.scale == 1.0
    got = y.data.cpu().numpy().astype(np.int32)
    assert got.shape == want.shape
    assert set(np.unique(got)) <= {-1, 1}
    bad = np.argwhere(got != want.astype(np.int32))
    assert bad.shape[0] == 0, "sign mismatches: %d of %d, first %s" % (bad.shape[0], want.size, bad[:8].tolist())


# --------------------------------------------------------------------------- K4: fp32 activations on the tensor cores (bf16 x 3 split)
F32_TC_CASES = [
    # n, h, w, cin, cout, wkind, nb, act, residual, bias
    (2, 32, 32, 16, 16, "quantized", 4, "leaky", True, False),     # ResNet stack 0 block end
    (3, 16, 16, 32, 32, "quantized", 4, "leaky", False, True),     # ResNet stack 1 first conv of a block
    (5, 8, 8, 64, 64, "ternary", 2, "leaky", True, False),         # stack 2, two images per tile, ragged pair
    (2, 16, 24, 16, 48, "binary", 1, None, False, True),           # Cout != Cin, no activation
    (1, 32, 8, 48, 16, "quantized", 8, "leaky", False, False),     # three channel chunks
    (40, 32, 32, 16, 16, "quantized", 4, "leaky", True, True),     # 320 tiles: persistent loop, ring wrap-around
]


@pytest.mark.parametrize("case", F32_TC_CASES, ids=["n%d_%dx%d_%d-%d_%s%d_%s%s" % (c[0], c[1], c[2], c[3], c[4], c[5][:3], c[6], c[7], "_res" if c[8] else "") for c in F32_TC_CASES])
def test_conv2d_f32_tcgen05_tolerance(case):
    """fp32 activations x exact integer kernels on tcgen05 kind::f16 with the activations split into three bf16
    terms: products are exact, accumulation is fp32 in TMEM -> same tolerance class as the FFMA kernel."""
    q, L, K = _mods()
    n, h, w, cin, cout, wkind, nb, act, use_res, use_bias = case
    rng = np.random.default_rng(_seed(("f32tc",) + case))
    x = rng.normal(0, 1, size=(n, h, w, cin)).astype(F32)
    x[0, 0, 0, :] = 0.0
    x[0, -1, -1, :] = F32(1e-6)           # tiny values: the low split terms matter
    x[0, 1, 1, :] = F32(123.456)
    kernel = rng.uniform(-1, 1, size=(3, 3, cin, cout)).astype(F32)
    fan = 9 * cin
    bias = rng.uniform(-0.3, 0.3, size=cout).astype(F32) if use_bias else None
    bn = (rng.uniform(0.3, 0.9, cout).astype(F32) * rng.choice([1, 1, -1], cout).astype(F32), rng.uniform(-0.2, 0.2, cout).astype(F32),
          rng.uniform(-0.2, 0.2, cout).astype(F32), (rng.uniform(0.5, 1.5, cout) * fan * 0.3).astype(F32))
    residual = rng.normal(0, 0.5, size=(n, h, w, cout)).astype(F32) if use_res else None
    want, _ = oracle_layer(x, "f32", 1.0, kernel, wkind, nb, 1.0, 1, bias=bias, bn=bn, eps=1e-3, residual=residual, res_mul=0.5,
                           act=act, abits=4)
    mode = {"quantized": L.W_QUANT, "binary": L.W_BINARY, "ternary": L.W_TERNARY}[wkind]
    wp = K.pack_weights(dev(kernel), mode, nb, 1.0, L.WFMT_I8)
    wscale = 1.0 / (1 << (nb - 1)) if wkind == "quantized" else 1.0
    i_, s_ = K.bn_constants(*bn, 1e-3)
    res_q = K.QTensor("f32", dev(residual), 1.0, cout) if use_res else None
    actc = {"leaky": L.ACT_LEAKY, None: L.ACT_NONE}[act]
    epi = K.make_epilogue(F32(wscale), bias=dev(bias) if use_bias else None, bn_inv=dev(i_), bn_shift=dev(s_), residual=res_q,
                          res_mul=0.5, act=actc, leaky_alpha=0.3)
    xq = K.QTensor("f32", dev(x), 1.0, cin)
    assert K.conv2d_on_tensor_cores(xq, 3, 3, cout, 1, epi, L.IMPL_TCGEN05)
    y = K.conv2d(xq, wp, 3, 3, cout, 1, epi, impl=L.IMPL_TCGEN05)
    torch.cuda.synchronize()
    got = y.data.cpu().numpy()
    assert got.shape == want.shape
    err = np.abs(got - want)
    assert err.max() <= 1e-5 * np.abs(want).max(), "max abs err %g (rel %g) at %s" % (err.max(), err.max() / np.abs(want).max(), np.unravel_index(err.argmax(), err.shape))
    # and it agrees with the CUDA-core kernel to the same tolerance
    y2 = K.conv2d(xq, wp, 3, 3, cout, 1, epi, impl=L.IMPL_GENERIC).data.cpu().numpy()
    assert np.abs(got - y2).max() <= 1e-5 * np.abs(want).max()


@pytest.mark.parametrize("n,pos,ch,units,softmax", [(5, 64, 64, 10, True), (3, 16, 96, 7, False), (40, 64, 64, 10, True)])
def test_dense_with_global_average_pool_fp32(n, pos, ch, units, softmax):
    """AveragePooling2D + Flatten + Dense on an fp32 map as ONE launch (qnnb_dense_desc.avg_positions): tolerance class
    of the fp32 path, against float64."""
    q, L, K = _mods()
    rng = np.random.default_rng(_seed(("avgdense", n, pos, ch, units)))
    x = rng.normal(0, 1, size=(n, pos, ch)).astype(F32)
    kernel = rng.uniform(-1, 1, size=(ch, units)).astype(F32)
    lv = exact.quantize_levels(kernel.reshape(1, 1, ch, units), 4).reshape(ch, units).astype(np.float64) / 8.0
    z = (x.astype(np.float64).sum(axis=1) / pos) @ lv
    want = z
    if softmax:
        e = np.exp(z - z.max(axis=1, keepdims=True))
        want = e / e.sum(axis=1, keepdims=True)
    wp = K.pack_weights(dev(kernel), L.W_QUANT, 4, 1.0, L.WFMT_I8)
    epi = K.make_epilogue(K.acc_scale(1.0 / pos, 1.0 / 8))
    out, logits = K.dense(K.QTensor("f32", dev(x.reshape(n, -1)), 1.0, pos * ch), wp, units, epi, softmax=softmax,
                          want_logits=softmax, avg_positions=pos)
    torch.cuda.synchronize()
    assert np.abs(out.cpu().numpy() - want).max() <= 1e-5 * max(np.abs(want).max(), 1e-30)
    if softmax:
        assert np.abs(logits.cpu().numpy() - z).max() <= 1e-5 * np.abs(z).max()
