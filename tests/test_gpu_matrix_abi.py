"""The reference's 26 matrix symbols served by libcg_b200.so, diffed symbol by symbol against the oracle
(reference-compiled CPU backend when oracle/_ref is present). Bit-exact except sigmoid/tanh (<= 4 ulp:
CUDA expf/tanhf vs glibc) and the TF32 / FMA products (tolerances stated per test)."""
import ctypes as C

import numpy as np
import pytest

from helpers import rel_err, ulp_diff
from matrix_abi import BINS, MAPS, OracleMatrixAbi, ProductMatrixAbi

pytestmark = pytest.mark.gpu

SHAPES = [(1, 1), (2, 2), (3, 5), (17, 33), (128, 96), (257, 1031)]


@pytest.fixture()
def abis(cg, oracle):
    return ProductMatrixAbi(cg.lib), OracleMatrixAbi(oracle)


def special_values(rng, h, w):
    a = rng.uniform(-3, 3, (h, w)).astype(np.float32)
    flat = a.reshape(-1)
    specials = np.array([0.0, -0.0, 1.0, -1.0, 0.5, np.inf, -np.inf, 1e-30, -1e-30, 88.0, -88.0, 100.0, -100.0],
                        dtype=np.float32)
    n = min(len(flat), len(specials))
    flat[:n] = specials[:n]
    return a


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("op", MAPS)
def test_map_symbols(abis, shape, op):
    p, o = abis
    rng = np.random.default_rng(hash((shape, op)) % 2**31)
    a = special_values(rng, *shape)
    mp, rp = p.new(*shape, a), p.new(*shape)
    mo, ro = o.new(*shape, a), o.new(*shape)
    getattr(p.lib, "map_" + op)(C.byref(mp), C.byref(rp))
    o.lib.orc_sym_map(MAPS.index(op), C.byref(mo), C.byref(ro))
    got, want = p.get(rp), o.get(ro)
    if op in ("sigmoid", "tanh"):
        assert np.all(ulp_diff(got, want) <= 4), (op, np.max(ulp_diff(got, want)))
    else:
        assert np.array_equal(got.view(np.int32), want.view(np.int32)), op  # bit-exact, incl. -0.0 and inf
    # in place (m aliases result), as the reference allows
    getattr(p.lib, "map_" + op)(C.byref(mp), C.byref(mp))
    assert np.array_equal(p.get(mp).view(np.int32), got.view(np.int32))
    p.free(mp), p.free(rp)


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("op", BINS)
def test_binary_symbols(abis, shape, op):
    p, o = abis
    rng = np.random.default_rng(hash((shape, op, 1)) % 2**31)
    a, b = special_values(rng, *shape), rng.uniform(-3, 3, shape).astype(np.float32)
    v = np.float32(rng.uniform(-2, 2))
    ap, bp, rp = p.new(*shape, a), p.new(*shape, b), p.new(*shape)
    ao, bo, ro = o.new(*shape, a), o.new(*shape, b), o.new(*shape)
    k = BINS.index(op)
    with np.errstate(invalid="ignore"):
        getattr(p.lib, "elemwise_" + op)(C.byref(ap), C.byref(bp), C.byref(rp))
        o.lib.orc_sym_elemwise(k, C.byref(ao), C.byref(bo), C.byref(ro))
        assert np.array_equal(p.get(rp).view(np.int32), o.get(ro).view(np.int32))
        getattr(p.lib, "broadcast_" + op)(C.byref(ap), v, C.byref(rp))
        o.lib.orc_sym_broadcast(k, C.byref(ao), v, C.byref(ro))
        assert np.array_equal(p.get(rp).view(np.int32), o.get(ro).view(np.int32))
        getattr(p.lib, "broadcast_" + op + "_rev")(v, C.byref(ap), C.byref(rp))
        o.lib.orc_sym_broadcast_rev(k, v, C.byref(ao), C.byref(ro))
        assert np.array_equal(p.get(rp).view(np.int32), o.get(ro).view(np.int32))
    for m in (ap, bp, rp):
        p.free(m)


@pytest.mark.parametrize("shape", SHAPES)
def test_memory_and_reduction_symbols(abis, shape):
    p, o = abis
    rng = np.random.default_rng(5)
    a = rng.uniform(-1, 1, shape).astype(np.float32)
    mp, mo = p.new(*shape, a), o.new(*shape, a)
    assert np.array_equal(p.get(mp), a)  # set_array(transpose=true) + get_array round trip
    raw = np.zeros(a.size, np.float32)
    p.lib.get_array(C.byref(mp), raw.ctypes.data_as(C.POINTER(C.c_float)))
    assert np.array_equal(raw, a.T.reshape(-1))  # storage is column-major (cpu/matrix.cpp:26-36)
    p.lib.set_array(C.byref(mp), raw.ctypes.data_as(C.POINTER(C.c_float)), False)  # raw storage order
    assert np.array_equal(p.get(mp), a)
    cp = p.new(*shape)
    p.lib.copy_matrix(C.byref(mp), C.byref(cp))
    assert np.array_equal(p.get(cp), a)
    # reduce_sum: sequential f32 (bit-exact) up to 4096 elements, pairwise beyond (1e-6 relative)
    got, want = p.lib.reduce_sum(C.byref(mp)), o.lib.orc_sym_reduce_sum(C.byref(mo))
    if a.size <= 4096:
        assert np.float32(got) == np.float32(want)
    else:
        assert abs(got - want) <= 1e-6 * max(1.0, float(np.sum(np.abs(a))))
    assert p.lib.all_less_than(C.byref(mp), 1.0) and not p.lib.all_less_than(C.byref(mp), float(a.max()))
    assert not p.lib.all_equal(C.byref(mp), 0.25) or a.size == 0
    p.lib.fill_matrix(C.byref(cp), 0.25)
    assert p.lib.all_equal(C.byref(cp), 0.25) and np.all(p.get(cp) == np.float32(0.25))
    p.free(mp), p.free(cp)


@pytest.mark.parametrize("t1,t2", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("dims", [(2, 2, 2), (5, 7, 3), (30, 30, 30), (64, 17, 40)])
def test_gemm_strict_bit_exact(abis, dims, t1, t2):
    """Products up to 64 in every dimension reproduce the reference's naive k-ascending f32 loop bit for bit
    (cpu/ops.cpp:103-114); square ones are checked against the reference-compiled gemm itself."""
    p, o = abis
    m, n, k = dims
    rng = np.random.default_rng(m * 1000 + n * 10 + k)
    a = rng.uniform(-1, 1, (k, m) if t1 else (m, k)).astype(np.float32)
    b = rng.uniform(-1, 1, (n, k) if t2 else (k, n)).astype(np.float32)
    ap, bp, rp = p.new(*a.shape, a), p.new(*b.shape, b), p.new(m, n)
    ao, bo, ro = o.new(*a.shape, a), o.new(*b.shape, b), o.new(m, n)
    p.lib.gemm(C.byref(ap), t1, C.byref(bp), t2, C.byref(rp))
    o.lib.orc_sym_gemm(C.byref(ao), t1, C.byref(bo), t2, C.byref(ro))
    assert np.array_equal(p.get(rp).view(np.int32), o.get(ro).view(np.int32))
    for x in (ap, bp, rp):
        p.free(x)


@pytest.mark.parametrize("t1,t2", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("dims", [(100, 130, 70), (256, 384, 160), (129, 65, 257)])
def test_gemm_tiled_fp32(abis, dims, t1, t2):
    """Tiled FP32 SIMT kernel (FMA accumulation, k ascending): <= 1e-5 relative to the oracle product."""
    p, o = abis
    m, n, k = dims
    rng = np.random.default_rng(m + n + k)
    a = rng.uniform(-1, 1, (k, m) if t1 else (m, k)).astype(np.float32)
    b = rng.uniform(-1, 1, (n, k) if t2 else (k, n)).astype(np.float32)
    ap, bp, rp = p.new(*a.shape, a), p.new(*b.shape, b), p.new(m, n)
    p.lib.gemm(C.byref(ap), t1, C.byref(bp), t2, C.byref(rp))
    want = (a.T if t1 else a).astype(np.float64) @ (b.T if t2 else b).astype(np.float64)
    assert rel_err(p.get(rp), want) < 1e-5
    for x in (ap, bp, rp):
        p.free(x)


@pytest.mark.parametrize("t1,t2", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("dims", [(256, 256, 128), (384, 512, 320), (1024, 768, 512), (256, 64, 32), (512, 832, 96),
                                  (1024, 1408, 160), (2048, 4096, 64)])
@pytest.mark.parametrize("pair", [False, True])
def test_gemm_tcgen05_tf32(cg, abis, dims, t1, t2, pair):
    """tcgen05 kind::tf32 product, all four operand-major combinations, on the single-CTA kernel and (pair = True,
    shapes with m % 256 == 0 and n % 64 == 0) the 2-SM cta_group::2 kernel. Small-integer inputs are exact in TF32
    and their dot products exact in f32 => the result must be EXACT (catches any layout/descriptor error);
    uniform inputs are held to 1e-3 relative (TF32 keeps 10 mantissa bits; north_star tolerance)."""
    p, _ = abis
    cg.set_tf32(True)
    cg.set_tf32_pair(pair)
    try:
        m, n, k = dims
        rng = np.random.default_rng(m + 3 * n + 7 * k)
        for exact in (True, False):
            if exact:
                a = rng.integers(-4, 5, (k, m) if t1 else (m, k)).astype(np.float32)
                b = rng.integers(-4, 5, (n, k) if t2 else (k, n)).astype(np.float32)
            else:
                a = rng.uniform(-1, 1, (k, m) if t1 else (m, k)).astype(np.float32)
                b = rng.uniform(-1, 1, (n, k) if t2 else (k, n)).astype(np.float32)
            ap, bp, rp = p.new(*a.shape, a), p.new(*b.shape, b), p.new(m, n)
            p.lib.fill_matrix(C.byref(rp), np.float32(np.nan))
            p.lib.gemm(C.byref(ap), t1, C.byref(bp), t2, C.byref(rp))
            want = (a.T if t1 else a).astype(np.float64) @ (b.T if t2 else b).astype(np.float64)
            got = p.get(rp)
            if exact:
                assert np.array_equal(got.astype(np.float64), want), (dims, t1, t2, float(np.max(np.abs(got - want))))
            else:
                assert rel_err(got, want) < 1e-3
            for x in (ap, bp, rp):
                p.free(x)
    finally:
        cg.set_tf32(False)
        cg.set_tf32_pair(True)


@pytest.mark.parametrize("t1,t2", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("dims", [(256, 256, 128), (1024, 768, 512), (512, 832, 96 + 32), (1024, 1408, 160), (2048, 4096, 128),
                                  (256, 4096, 4096)])
def test_gemm_tcgen05_tf32x3(cg, abis, dims, t1, t2):
    """FP32-accurate product on the tensor core (cg_set_tf32(2): every operand split in shared memory into the part TF32
    keeps and the part it drops, three tcgen05.mma per k-step, the contraction accumulated in 256-deep chunks that are
    summed round-to-nearest in registers; CTA-pair kernel). Small-integer inputs must be EXACT (any layout / barrier
    error of the split tiles shows); operands that NEED the dropped bits (1 + 2^-12 against 1: plain TF32 returns k)
    must be exact too; uniform inputs are held, element by element, to 1.5e-6 of sum |a||b| (the term left out,
    lo_a lo_b, is < 2^-20 |a||b| and 2^-22 on average; the first version, one accumulator for the whole contraction,
    failed this at k = 4096 by 2.4x through the tensor core's truncating accumulation) and to 1e-5 in the max norm."""
    p, _ = abis
    cg.set_tf32(2)
    try:
        m, n, k = dims
        rng = np.random.default_rng(m + 3 * n + 7 * k)
        for case in ("int", "lowbits", "uniform"):
            if case == "int":
                a = rng.integers(-4, 5, (k, m) if t1 else (m, k)).astype(np.float32)
                b = rng.integers(-4, 5, (n, k) if t2 else (k, n)).astype(np.float32)
            elif case == "lowbits":
                a = (1.0 + rng.integers(0, 2, (k, m) if t1 else (m, k)) * 2.0 ** -12).astype(np.float32)
                b = rng.integers(-1, 2, (n, k) if t2 else (k, n)).astype(np.float32)
            else:
                a = rng.uniform(-1, 1, (k, m) if t1 else (m, k)).astype(np.float32)
                b = rng.uniform(-1, 1, (n, k) if t2 else (k, n)).astype(np.float32)
            ap, bp, rp = p.new(*a.shape, a), p.new(*b.shape, b), p.new(m, n)
            p.lib.fill_matrix(C.byref(rp), np.float32(np.nan))
            p.lib.gemm(C.byref(ap), t1, C.byref(bp), t2, C.byref(rp))
            a64, b64 = (a.T if t1 else a).astype(np.float64), (b.T if t2 else b).astype(np.float64)
            want = a64 @ b64
            got = p.get(rp).astype(np.float64)
            if case != "uniform":
                assert np.array_equal(got, want), (case, dims, t1, t2, float(np.max(np.abs(got - want))))
            else:
                bound = 1.5e-6 * (np.abs(a64) @ np.abs(b64))
                assert np.all(np.abs(got - want) <= bound), (dims, t1, t2, float(np.max(np.abs(got - want) / bound)))
                assert rel_err(got, want) < 1e-5
            for x in (ap, bp, rp):
                p.free(x)
    finally:
        cg.set_tf32(False)
