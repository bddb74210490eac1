"""CPU checks of the bit-exact REWRITES the CUDA epilogues rely on, restated in NumPy fp32:
each kernel-side shortcut must give exactly the value of the oracle's fixed op order.  No GPU, no library calls."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import exact  # noqa: E402

F32 = np.float32


def _consts(rng, n):
    bias = rng.uniform(-0.3, 0.3, n).astype(F32)
    inv = (rng.uniform(0.01, 2.0, n) * rng.choice([1.0, 1.0, -1.0], n)).astype(F32)
    shift = rng.uniform(-0.5, 0.5, n).astype(F32)
    return bias, inv, shift


def _fixed_order(acc, s, bias, inv, shift, qm):
    """oracle order: c = float(acc)*s; p = c + bias; y = p*inv + shift (separate RN ops); zq = y*qm"""
    c = acc.astype(F32) * F32(s)
    p = (c + bias).astype(F32)
    y = ((p * inv).astype(F32) + shift).astype(F32)
    return (y * F32(qm)).astype(F32)


@pytest.mark.parametrize("abits,wbits", [(4, 4), (8, 8), (2, 2), (4, 8)])
def test_power_of_two_folding_is_exact(abits, wbits):
    """FOLD: ((f + bias/s) * (inv*s*qm)) + shift*qm == (((f*s) + bias) * inv + shift) * qm for power-of-two s, qm."""
    rng = np.random.default_rng(abits * 16 + wbits)
    n = 200000
    s = 1.0 / (2 ** (abits - 1) * 2 ** (wbits - 1))
    qm = float(2 ** (abits - 1))
    acc = rng.integers(-300000, 300000, n).astype(np.int32)
    bias, inv, shift = _consts(rng, n)
    want = _fixed_order(acc, s, bias, inv, shift, qm)
    a = (bias * F32(1.0 / s)).astype(F32)
    b = ((inv * F32(s)).astype(F32) * F32(qm)).astype(F32)
    c = (shift * F32(qm)).astype(F32)
    got = (((acc.astype(F32) + a).astype(F32) * b).astype(F32) + c).astype(F32)
    assert np.array_equal(got, want)


def test_pooling_raw_accumulators_equals_pooling_activations():
    """2x2 max-pool of quantised activations == quantise(max or min of the raw accumulators, by the sign of the BN
    slope): every step of the pipeline is monotone in acc for fixed channel constants."""
    rng = np.random.default_rng(7)
    n = 100000
    s, qm = 1.0 / 64, 8.0
    acc = rng.integers(-4000, 4000, (n, 4)).astype(np.int32)
    bias, inv, shift = _consts(rng, n)
    z = _fixed_order(acc, s, bias[:, None], inv[:, None], shift[:, None], 1.0)
    want = exact.act_quant_levels(z, 4).max(axis=1)
    pick = np.where(inv < 0, acc.min(axis=1), acc.max(axis=1))
    got = exact.act_quant_levels(_fixed_order(pick, s, bias, inv, shift, 1.0), 4)
    assert np.array_equal(got, want)
    # and for binary_tanh
    want_b = exact.act_binary_levels(z).max(axis=1)
    got_b = exact.act_binary_levels(_fixed_order(pick, s, bias, inv, shift, 1.0))
    assert np.array_equal(got_b, want_b)


def test_magic_number_rounding_equals_rint_with_clamp():
    """quant_scaled: low byte of float_as_int(clamp(zq) + 1.5*2^23) == int8(clamp(rint(zq)))."""
    rng = np.random.default_rng(3)
    for qm in (2.0, 8.0, 128.0):
        zq = np.concatenate([rng.uniform(-2 * qm, 2 * qm, 300000), np.arange(-qm - 2, qm + 2, 0.5), [np.nan]]).astype(F32)
        clamped = np.minimum(np.maximum(zq, F32(-qm)), F32(qm - 1)).astype(F32)      # fminf(fmaxf(NaN, lo), hi) = lo
        clamped = np.where(np.isnan(zq), F32(-qm), clamped).astype(F32)
        magic = (clamped + F32(12582912.0)).astype(F32)
        low = (magic.view(np.uint32) & 0xFF).astype(np.uint8).view(np.int8).astype(np.int32)
        want = np.clip(np.rint(np.where(np.isnan(zq), F32(-qm), zq)), -qm, qm - 1).astype(np.int32)
        assert np.array_equal(low, want)


def test_three_way_bf16_truncation_split_is_exact():
    """K4: hi = top16(x), mid = top16(x - hi), lo = top16(x - hi - mid); hi + mid + lo == x for every fp32 with
    |x| >= 2^-100 (each bf16 term keeps 8 significant bits, every subtraction is exact); smaller magnitudes: |error| < 2^-126."""
    rng = np.random.default_rng(11)
    x = np.concatenate([rng.normal(0, 1, 500000), rng.normal(0, 1, 100000) * 1e-20, rng.normal(0, 1, 100000) * 1e20,
                        [0.0, -0.0, 1.0, -1.0, 123.456, 1e-38, 3.0e38, np.float32(2 ** -126), np.float32(2 ** -140)]]).astype(F32)
    def top(v):
        return (v.view(np.uint32) & np.uint32(0xFFFF0000)).view(F32)
    hi = top(x)
    r = (x - hi).astype(F32)
    mid = top(r)
    q = (r - mid).astype(F32)
    lo = top(q)
    total = hi.astype(np.float64) + mid.astype(np.float64) + lo.astype(np.float64)
    normal = (np.abs(x) >= F32(2.0 ** -100)) | (x == 0)          # residuals stay normal numbers
    assert np.array_equal((q - lo).astype(F32)[normal], np.zeros_like(x)[normal]), "a fourth term would be needed"
    assert np.array_equal(total[normal], x.astype(np.float64)[normal])
    # below 2^-100 the residuals run into the subnormal range: the split is then exact to an ABSOLUTE 2^-126
    assert np.abs(total - x.astype(np.float64)).max() <= 2.0 ** -126
    # the kernel levels (|k| <= 128) are exact in bf16
    k = np.arange(-128, 129).astype(F32)
    assert np.array_equal(top(k), k)


def test_multiply_shift_tile_decode_is_exact():
    """FastDiv: (x * ceil(2^40 / d)) >> 40 == x // d for x < 2^24, d < 2^15."""
    rng = np.random.default_rng(5)
    for d in list(range(1, 70)) + [127, 128, 255, 1000, 4096, 32767]:
        m = ((1 << 40) + d - 1) // d
        xs = np.concatenate([rng.integers(0, 1 << 24, 20000), [0, 1, d - 1, d, d + 1, (1 << 24) - 1]]).astype(np.uint64)
        q = (xs * np.uint64(m)) >> np.uint64(40)
        assert np.array_equal(q, xs // np.uint64(d)), d


def test_int8_conv_is_linear_and_sign_levels_reproduce_xnor_popcount():
    """Size-independent properties of the exact accumulator: linearity in the input, and the +-1 x +-1 dot product
    equals valid*cin - 2*popc(x xor w) on the bit-packed form (the two storage forms of a binary map agree)."""
    rng = np.random.default_rng(9)
    x1 = rng.integers(-8, 8, (2, 6, 6, 8)).astype(np.int64)
    x2 = rng.integers(-8, 8, (2, 6, 6, 8)).astype(np.int64)
    w = rng.integers(-8, 8, (3, 3, 8, 5)).astype(np.int64)
    assert np.array_equal(exact.conv_accumulate(x1 + x2, w, 1), exact.conv_accumulate(x1, w, 1) + exact.conv_accumulate(x2, w, 1))
    xb = (rng.integers(0, 2, (1, 1, 1, 64)) * 2 - 1).astype(np.int64)
    wb = (rng.integers(0, 2, (1, 1, 64, 3)) * 2 - 1).astype(np.int64)
    acc = exact.conv_accumulate(xb, wb, 1)[0, 0, 0]
    xw = exact.pack_bits_lastdim(xb[0, 0, 0][None])[0]
    for u in range(3):
        ww = exact.pack_bits_lastdim(wb[0, 0, :, u][None])[0]
        popc = sum(bin(int(a) ^ int(b)).count("1") for a, b in zip(xw, ww))
        assert acc[u] == 64 - 2 * popc
