"""The arithmetic contract of the tensor-core interaction forward (csrc/interact_warp.cu), restated in
numpy so it is checked without a GPU: every fp32 operand v is split into a TF32 head (mantissa
truncated to 10 bits) and a TF32 tail of the exact remainder, and a product a*b is replaced by
a_lo*b_hi + a_hi*b_lo + a_hi*b_hi."""
import numpy as np

MASK = np.uint32(0xFFFFE000)


def split(v: np.ndarray):
    v = v.astype(np.float32)
    hi = (v.view(np.uint32) & MASK).view(np.float32)
    rem = (v - hi).astype(np.float32)              # exact: at most 13 significant bits
    lo = (rem.view(np.uint32) & MASK).view(np.float32)
    return hi, lo, rem


def test_remainder_is_exact_and_tail_error_is_below_2_pow_minus_20():
    rng = np.random.default_rng(0)
    v = np.concatenate([rng.standard_normal(100000), np.exp(10 * rng.standard_normal(100000)) * rng.choice([-1, 1], 100000),
                        rng.uniform(-1e-3, 1e-3, 100000)]).astype(np.float32)
    hi, lo, rem = split(v)
    assert np.array_equal(hi.astype(np.float64) + rem.astype(np.float64), v.astype(np.float64))   # v - hi is exact in fp32
    err = np.abs(v.astype(np.float64) - hi.astype(np.float64) - lo.astype(np.float64))
    assert np.all(err <= 2.0 ** -20 * np.abs(v.astype(np.float64)))
    # heads and tails are valid TF32 values: their low 13 mantissa bits are zero
    assert not np.any(hi.view(np.uint32) & ~MASK) and not np.any(lo.view(np.uint32) & ~MASK)


def test_three_term_product_error_and_integer_exactness():
    rng = np.random.default_rng(1)
    a = rng.standard_normal((2000, 128)).astype(np.float32)
    b = rng.standard_normal((2000, 128)).astype(np.float32)
    ah, al, _ = split(a)
    bh, bl, _ = split(b)
    f = np.float64
    approx = (al.astype(f) * bh.astype(f) + ah.astype(f) * bl.astype(f) + ah.astype(f) * bh.astype(f)).sum(1)
    exact = (a.astype(f) * b.astype(f)).sum(1)
    scale = (np.abs(a).astype(f) * np.abs(b).astype(f)).sum(1)
    assert np.max(np.abs(approx - exact) / scale) < 3 * 2.0 ** -20          # dropped a_lo*b_lo + two tail truncations
    # small integers have no tail at all: the Gram matrix of integer features is exact
    ints = rng.integers(-1024, 1025, size=100000).astype(np.float32)
    hi, lo, _ = split(ints)
    assert np.array_equal(hi, ints) and not np.any(lo)
