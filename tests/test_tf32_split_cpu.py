"""The 3xTF32 operand split of the tensor-core GEMMs (omni-pq_b200/csrc/pn2_sm100.cuh, round_tf32 / split_tf32) restated
with numpy: rounding to tf32 on the bit pattern -- add half an ulp of the 10-bit mantissa, clear the 13 dropped bits --
must equal round-to-nearest, ties away from zero (what cvt.rna.tf32.f32 computes) for every finite value, and
hi + lo must reproduce the operand to 2^-21 relative."""
import numpy as np


def round_tf32_bits(v):
    b = v.astype(np.float32).view(np.uint32)
    return ((b + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def round_tf32_reference(v):
    """nearest multiple of the tf32 ulp of v's binade, ties away from zero, in exact float64 arithmetic"""
    v64 = v.astype(np.float64)
    mant, expo = np.frexp(np.abs(v64))                    # |v| = mant * 2^expo, mant in [0.5, 1)
    expo = np.maximum(expo, -125)                         # denormals share the smallest normal binade's ulp
    ulp = np.ldexp(1.0, expo - 11)                        # 10 explicit mantissa bits + the implicit one
    q = np.floor(np.abs(v64) / ulp + 0.5)                 # ties away from zero on the magnitude
    return (np.sign(v64) * q * ulp).astype(np.float32)


def test_bit_pattern_rounding_is_round_to_nearest_ties_away():
    rng = np.random.default_rng(0)
    bits = rng.integers(0, 1 << 32, 400000, dtype=np.uint64).astype(np.uint32)
    v = bits.view(np.float32)
    v = v[np.isfinite(v) & (np.abs(v) < 3.0e38)]          # (the last binade rounds up to inf in both forms)
    ties = (rng.integers(0, 1 << 19, 4096, dtype=np.uint64).astype(np.uint32) << np.uint32(13) | np.uint32(0x1000)).view(np.float32)
    small = np.array([0.0, -0.0, 1e-45, -1e-45, 1.17549435e-38, 5.9e-39, 1.0, -1.0, 1.0 + 2.0 ** -11, 1.0 + 2.0 ** -12], np.float32)
    v = np.concatenate([v, ties[np.isfinite(ties)], small])
    got, want = round_tf32_bits(v), round_tf32_reference(v)
    assert np.array_equal(got.view(np.uint32) & np.uint32(0x7FFFFFFF), want.view(np.uint32) & np.uint32(0x7FFFFFFF))
    assert np.array_equal(np.signbit(got), np.signbit(v))
    assert not np.any(got.view(np.uint32) & np.uint32(0x1FFF))      # representable in tf32


def test_hi_plus_lo_reproduces_the_operand():
    rng = np.random.default_rng(1)
    v = (rng.standard_normal(200000) * np.exp(rng.uniform(-20, 20, 200000))).astype(np.float32)
    hi = round_tf32_bits(v)
    lo = round_tf32_bits((v - hi).astype(np.float32))               # v - hi is exact in fp32
    err = np.abs(v.astype(np.float64) - hi.astype(np.float64) - lo.astype(np.float64))
    assert np.all(err <= np.abs(v.astype(np.float64)) * 2.0 ** -21)
    assert np.all(np.abs(lo.astype(np.float64)) <= np.abs(v.astype(np.float64)) * 2.0 ** -10)
