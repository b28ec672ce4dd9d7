"""The arithmetic model behind the FP32-accurate tensor-core product (csrc/gemm_tcgen05.cu, SPLIT3; cg_set_tf32(2)), on
the CPU: what the split into hi / lo operands leaves out, and why the contraction is accumulated in chunks.

Two facts about tcgen05.mma kind::tf32 are taken from measurements on the B200: every f32 operand is
TRUNCATED to 10 mantissa bits, and every instruction adds its 8-deep partial sum to the f32 accumulator rounding TOWARDS
ZERO. The model below restates both in numpy (products and the 8-deep partial sums exact, in float64) and checks
  * a single accumulator for the whole contraction ends hundreds of ulps short at k = 4096 — the first version of the
    kernel, which failed its own GPU test by 2.4x for exactly this reason;
  * the shipped scheme — small terms in an accumulator of their own, 256-deep chunks summed round-to-nearest — stays
    within the element-wise bound the GPU test uses (1.5e-6 * sum |a||b|) and within a few tens of ulps;
  * the operand split itself is exact (x == hi + lo) and integers / operands that need the dropped bits come out exact.
No GPU, no product code: this pins the numbers quoted in DESIGN.md and in the GPU test's docstring."""
import numpy as np

KC = 8 * 32  # SPLIT_KC k-blocks of BK = 32


def trunc_tf32(x):
    return (np.asarray(x, dtype=np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def rz_f32(v):
    """float64 -> float32 rounding towards zero."""
    t = v.astype(np.float32)
    over = np.abs(t.astype(np.float64)) > np.abs(v)
    t[over] = np.nextafter(t[over], np.float32(0))
    return t


def mma_accumulate(acc, a, b):
    """One kind::tf32 MMA: operands truncated, the 8-deep partial sum exact, the accumulate rounded towards zero."""
    part = trunc_tf32(a).astype(np.float64) @ trunc_tf32(b).astype(np.float64)
    return rz_f32(acc.astype(np.float64) + part)


def product_single_accumulator(a, b):
    acc = np.zeros((a.shape[0], b.shape[1]), np.float32)
    a_lo, b_lo = a - trunc_tf32(a), b - trunc_tf32(b)
    for k0 in range(0, a.shape[1], 8):
        s = slice(k0, k0 + 8)
        acc = mma_accumulate(acc, a_lo[:, s], b[s])
        acc = mma_accumulate(acc, a[:, s], b_lo[s])
        acc = mma_accumulate(acc, a[:, s], b[s])
    return acc


def product_chunked(a, b):
    a_lo, b_lo = a - trunc_tf32(a), b - trunc_tf32(b)
    total = None
    for c0 in range(0, a.shape[1], KC):
        hi = np.zeros((a.shape[0], b.shape[1]), np.float32)
        lo = np.zeros_like(hi)
        for k0 in range(c0, min(c0 + KC, a.shape[1]), 8):
            s = slice(k0, k0 + 8)
            lo = mma_accumulate(lo, a_lo[:, s], b[s])
            lo = mma_accumulate(lo, a[:, s], b_lo[s])
            hi = mma_accumulate(hi, a[:, s], b[s])
        part = hi + lo  # f32, round to nearest: the epilogue's __fadd_rn
        total = part if total is None else total + part
    return total


def ulps(got, want):
    return np.abs(got.astype(np.float64) - want) / np.spacing(np.abs(want).astype(np.float32)).astype(np.float64)


def test_split_is_exact_and_lo_fits_the_second_truncation():
    rng = np.random.default_rng(0)
    x = rng.uniform(-4, 4, 100000).astype(np.float32)
    hi, lo = trunc_tf32(x), x - trunc_tf32(x)
    assert np.array_equal(hi.astype(np.float64) + lo.astype(np.float64), x.astype(np.float64))
    assert np.all(np.abs(lo) < np.abs(x) * 2.0 ** -10 + 1e-45)
    # what the tensor core drops of lo when it truncates it in turn: < 2^-21 of the operand
    assert np.all(np.abs(lo - trunc_tf32(lo)) <= np.abs(x) * 2.0 ** -21)


def test_single_accumulator_is_hundreds_of_ulps_short_and_chunks_are_not():
    rng = np.random.default_rng(1)
    m, n, k = 24, 24, 4096
    a = rng.uniform(-1, 1, (m, k)).astype(np.float32)
    b = rng.uniform(-1, 1, (k, n)).astype(np.float32)
    want = a.astype(np.float64) @ b.astype(np.float64)
    bound = 1.5e-6 * (np.abs(a).astype(np.float64) @ np.abs(b).astype(np.float64))
    one = product_single_accumulator(a, b)
    big = np.abs(want) > 8  # ulps of a cancelled result say nothing
    assert np.median(ulps(one, want)[big]) > 150
    assert np.max(np.abs(one - want) / bound) > 1.0  # fails the GPU test's bound, as the first kernel did
    assert np.mean(np.abs(one[big]) < np.abs(want[big])) > 0.8  # a bias towards zero, not noise
    ch = product_chunked(a, b)
    assert np.max(np.abs(ch - want) / bound) < 0.5
    assert np.max(ulps(ch, want)[big]) < 64
    # against the plain f32 k-ascending FMA-free loop the reference runs (cpu/ops.cpp:103-114): same order of magnitude
    ref = np.zeros((m, n), np.float32)
    for kk in range(k):
        ref = ref + np.outer(a[:, kk], b[kk])
    assert np.max(ulps(ch, want)[big]) < 4 * max(8.0, float(np.max(ulps(ref, want)[big])))


def test_integers_and_dropped_bits_are_exact():
    rng = np.random.default_rng(2)
    a = rng.integers(-4, 5, (16, 512)).astype(np.float32)
    b = rng.integers(-4, 5, (512, 16)).astype(np.float32)
    assert np.array_equal(product_chunked(a, b).astype(np.float64), a.astype(np.float64) @ b.astype(np.float64))
    a = (1.0 + rng.integers(0, 2, (16, 512)) * 2.0 ** -12).astype(np.float32)  # plain TF32 would return the sum of b
    b = rng.integers(-1, 2, (512, 16)).astype(np.float32)
    want = a.astype(np.float64) @ b.astype(np.float64)
    assert np.array_equal(product_chunked(a, b).astype(np.float64), want)
    plain = np.zeros((16, 16), np.float32)
    for k0 in range(0, 512, 8):
        plain = mma_accumulate(plain, a[:, k0:k0 + 8], b[k0:k0 + 8])
    assert not np.array_equal(plain.astype(np.float64), want)
