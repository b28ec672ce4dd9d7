"""The CPU oracle pinned against everything the reference offers for this path (no GPU needed):
  * the reference's own 15 threshold tests (src/main.rs:85-166),
  * the C1 golden trajectory derivable with pure IEEE arithmetic,
  * the restated arithmetic ("port") against the REFERENCE-COMPILED CPU backend (oracle/_ref), symbol by symbol,
  * the committed golden vectors (tests/golden, generated through the reference-compiled backend),
  * the structural facts of the reference's tree walk (gemm counts per iteration, Q1/Q2/Q3/Q4 quirks).
"""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import c2_graph, leaves_dict, readme_lstm, rel_err
from matrix_abi import BINS, MAPS, Matrix, OracleMatrixAbi
from reference_cases import CASES, run_case

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.int32)


@pytest.mark.parametrize("name", list(CASES))
def test_reference_thresholds(oracle, name):
    res, goal = run_case(oracle, name, trace=True)
    for f, _ in res:
        assert f.all_less_than(goal), name


@pytest.mark.parametrize("name", list(CASES))
def test_reference_suite_golden(oracle, name):
    gold = np.load(os.path.join(GOLD, "reference_suite.npz"))
    res, _ = run_case(oracle, name, trace=True)
    for i, (f, tr) in enumerate(res):
        assert np.array_equal(bits(tr), bits(gold[f"{name}/{i}/trace"]))
        assert np.array_equal(bits(f.value()), bits(gold[f"{name}/{i}/final"]))


@pytest.mark.parametrize("backend", ["ref", "port"])
def test_port_equals_reference_compiled_backend_on_suite(oracle, backend):
    """The restated arithmetic gives the same bits as the reference's own compiled CPU backend."""
    gold = np.load(os.path.join(GOLD, "reference_suite.npz"))
    if backend == "port":
        oracle.use_port()
    try:
        for name in CASES:
            res, _ = run_case(oracle, name, trace=True)
            for i, (f, tr) in enumerate(res):
                assert np.array_equal(bits(tr), bits(gold[f"{name}/{i}/trace"])), (backend, name)
    finally:
        if not oracle.use_ref("O0"):
            oracle.use_port()


def test_c1_golden(oracle):
    F = oracle.Function
    x = F.scalar_param("x", 1.0)
    f = (x + F.scalar(3.0)).sq()
    tr = f.minimize({}, 0.01, 1000, trace=True)
    golden = {0: 16.0, 1: 15.366400718688965, 2: 14.757889747619629, 9: 11.122164726257324,
              99: 0.2930106520652771, 999: 3.2741809263825417e-11}
    for i, v in golden.items():
        assert np.float32(tr[i]) == np.float32(v)
    assert f.leaves()[0][1] == np.float32(-2.9999942779541016)
    g = np.load(os.path.join(GOLD, "c1_scalar.npz"))
    assert np.array_equal(bits(tr), bits(g["trace"]))


def test_tanh_trajectory_of_survey(oracle):
    """tanh_test's 10-step forward trajectory through the reference-compiled map_tanh etc."""
    want = [0.4621172, -0.278862, -0.8362842, -0.906817, -0.9337634, -0.9483457, -0.9575633, -0.9639423, -0.9686293,
            -0.9722237]
    res, _ = run_case(oracle, "tanh_test", trace=True)
    for f, tr in res:
        assert np.allclose(np.asarray(tr).reshape(10, -1)[:, 0], want, rtol=0, atol=5e-7)


@pytest.mark.parametrize("name,builder", [("c2_elementwise_32", lambda o: (c2_graph(o, 32), 20, "exact")),
                                          ("c3_lstm_d5", lambda o: (readme_lstm(o, 5), 10, "exact")),
                                          ("c3_lstm_d10", lambda o: (readme_lstm(o, 10), 10, "exact")),
                                          ("lstm_dag_d6_T3", lambda o: (readme_lstm(o, 6, T=3, seed=77), 10, "dag"))])
def test_graph_goldens(oracle, name, builder):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    mode = builder(oracle)[2]
    oracle.set_mode(mode)
    try:
        f, iters, _ = builder(oracle)
        tr = f.minimize({}, 0.01, iters, trace=True)
        assert np.array_equal(bits(tr), bits(g["trace"]))
        for k, v in leaves_dict(f).items():
            assert np.array_equal(bits(v), bits(g[k.replace(":", "_")])), k
    finally:
        oracle.set_mode("exact")


def _have_ref(oracle):
    ok = oracle.use_ref("O0")
    return ok


@pytest.mark.parametrize("shape", [(1, 1), (2, 2), (5, 3), (17, 33), (64, 64)])
def test_port_symbols_equal_reference_symbols(oracle, shape):
    """Every one of the matrix symbols: restated port == reference-compiled .so, bit for bit (incl. libm calls,
    which resolve to the same glibc). The square gemm is the reference's own (valid only for square, Q7)."""
    if not _have_ref(oracle):
        pytest.skip("oracle/_ref not built (no /root/reference and no prebuilt copy)")
    rng = np.random.default_rng(shape[0] * 131 + shape[1])
    a = rng.uniform(-3, 3, shape).astype(np.float32)
    b = rng.uniform(-3, 3, shape).astype(np.float32)
    a.reshape(-1)[0] = -0.0
    results = {}
    for backend in ("ref", "port"):
        if backend == "port":
            oracle.use_port()
        else:
            oracle.use_ref("O0")
        o = OracleMatrixAbi(oracle)
        ma, mb, r = o.new(*shape, a), o.new(*shape, b), o.new(*shape)
        out = []
        for k in range(len(MAPS)):
            o.lib.orc_sym_map(k, C.byref(ma), C.byref(r))
            out.append(o.get(r))
        for k in range(len(BINS)):
            o.lib.orc_sym_elemwise(k, C.byref(ma), C.byref(mb), C.byref(r))
            out.append(o.get(r))
            o.lib.orc_sym_broadcast(k, C.byref(ma), 1.25, C.byref(r))
            out.append(o.get(r))
            o.lib.orc_sym_broadcast_rev(k, 1.25, C.byref(ma), C.byref(r))
            out.append(o.get(r))
        out.append(np.float32(o.lib.orc_sym_reduce_sum(C.byref(ma))))
        out.append(np.float32(o.lib.orc_sym_all_less_than(C.byref(ma), 3.5)))
        out.append(np.float32(o.lib.orc_sym_all_equal(C.byref(ma), float(a[0, 0]))))
        if shape[0] == shape[1]:
            for t1 in (False, True):
                for t2 in (False, True):
                    o.lib.orc_sym_gemm(C.byref(ma), t1, C.byref(mb), t2, C.byref(r))
                    out.append(o.get(r))
        results[backend] = out
    oracle.use_ref("O0")
    assert len(results["ref"]) == len(results["port"])
    for x, y in zip(results["ref"], results["port"]):
        assert np.array_equal(bits(x), bits(y))


def test_o0_and_o3_reference_builds_agree(oracle):
    """the -O0 build (the reference's build.rs flags) and -O3 -ffp-contract=off give the same bits."""
    if not oracle.use_ref("O3"):
        pytest.skip("oracle/_ref not built")
    try:
        f = readme_lstm(oracle, 5)
        t3 = f.minimize({}, 0.01, 5, trace=True)
    finally:
        oracle.use_ref("O0")
    f = readme_lstm(oracle, 5)
    t0 = f.minimize({}, 0.01, 5, trace=True)
    assert np.array_equal(bits(t0), bits(t3))


def test_tree_walk_gemm_counts(oracle):
    """Q2: the un-memoised walk does 39 forward + 78 backward gemms per iteration at T=2
    (15 unique products), 216 + 432 at T=3; growth ~5.4x per step."""
    for T, fwd in ((1, 6), (2, 39), (3, 216)):
        f = readme_lstm(oracle, 5, T=T)
        size, dots = oracle.tree_size(f)
        assert dots == fwd
        oracle.reset_counts()
        f.minimize({}, 0.01, 1)
        assert oracle.gemm_count() == 3 * fwd
    f = readme_lstm(oracle, 5, T=2)
    assert len(f.leaves()) == 28


def test_quirks(oracle):
    F = oracle.Function
    # Q4: sigmoid(x) with x == 0.5 is eliminated at construction; 0 - x is x
    x = F.scalar_param("x", 0.5)
    assert oracle.lib.orc_kind(x.sigmoid()._h) == oracle.lib.orc_kind(x._h)
    z = F.scalar(0.0)
    assert oracle.lib.orc_kind((z - x)._h) == oracle.lib.orc_kind(x._h)
    # Q1: the two uses of x in (x - 2) * (x + b) diverge
    f = ((x - F.scalar(2.)) * (x + F.matrix(2, 2, [1.2, -1., 0., 1.5]))).sq()
    f.minimize({}, 0.04, 3)
    (n0, v0), (n1, v1) = f.leaves()
    assert n0 == n1 == "x" and v0 != v1
    # Q3: sigmoid hands its child the local derivative: d/dw of 10 * sigmoid(w) moves w by lr * s'(w), not 10x that
    # (w starts at 2: a node's stale construction-time value equals its child's, and 1 is Mul's identity — Q4)
    w = F.scalar_param("w", 2.0)
    g = F.scalar(10.0) * w.sigmoid()
    g.minimize({}, 1.0, 1)
    s = 1.0 / (1.0 + np.exp(-2.0))
    assert abs(float(g.leaves()[0][1]) - (2.0 - s * (1 - s))) < 1e-6
    # ... and the stale-value elimination itself: sigmoid(w) with w == 1 carries the value 1 => 10 * it is just 10
    w1 = F.scalar_param("w", 1.0)
    assert (F.scalar(10.0) * w1.sigmoid()).leaves() == []
    # Q5: scalar <- matrix gradient is averaged: x + ones(2,2)*c with root error ones => x -= lr * avg(1) = lr
    xs = F.scalar_param("x", 0.5)
    h = xs + F.matrix(2, 2, [1., 2., 3., 4.])
    h.minimize({}, 0.1, 1)
    assert np.float32(h.leaves()[0][1]) == np.float32(0.5) - np.float32(0.1)


def test_dag_equals_exact_without_shared_nodes(oracle):
    out = []
    for mode in ("exact", "dag"):
        oracle.set_mode(mode)
        f = readme_lstm(oracle, 7, T=1, seed=3)
        tr = f.minimize({}, 0.01, 5, trace=True)
        out.append((tr, leaves_dict(f)))
    oracle.set_mode("exact")
    assert np.array_equal(bits(out[0][0]), bits(out[1][0]))
    for k in out[0][1]:
        assert np.array_equal(bits(out[0][1][k]), bits(out[1][1][k]))


def _activation_chain(api, xv, wv, bv, tv):
    F = api.Function
    x, t = F.constant(api.Constant(xv)), F.constant(api.Constant(tv))
    w, b = F.param("W", api.Constant(wv)), F.param("b", api.Constant(bv))
    return ((api.dot(x, w) + b).sigmoid().tanh() - t).sq()


def test_corrected_activation_gradients_match_finite_differences(oracle):
    """SURVEY §8 row f2. The reference's Sigmoid / Tanh backward rules hand the child the LOCAL derivative and drop
    the upstream error (Q3, src/optimize.rs:168-179) — reproduced by default because it is the reference's semantics.
    With the correction switched on (fix_act_grad) the update of every parameter must be -lr * dL/dtheta for
    L = sum of the root's elements (root error == 1, optimize.rs:69), checked here against central finite differences
    of the oracle's own forward pass; with the reference rule it must NOT be (that is what Q3 means)."""
    rng = np.random.default_rng(8)
    xv = rng.uniform(-1, 1, (4, 3)).astype(np.float32)
    wv = rng.uniform(-1, 1, (3, 5)).astype(np.float32)
    bv = rng.uniform(-1, 1, (4, 5)).astype(np.float32)
    tv = rng.uniform(-1, 1, (4, 5)).astype(np.float32)
    lr, eps = 1e-2, 1e-2

    def loss(w, b):
        f = _activation_chain(oracle, xv, w, b, tv)
        return float(np.asarray(f.minimize({}, 0.0, 1, trace=True), dtype=np.float64).sum())

    def numeric(which):
        base = {"W": wv, "b": bv}[which]
        g = np.zeros(base.shape)
        for idx in np.ndindex(base.shape):
            hi, lo = base.copy(), base.copy()
            hi[idx] += eps
            lo[idx] -= eps
            g[idx] = ((loss(hi, bv) if which == "W" else loss(wv, hi)) - (loss(lo, bv) if which == "W" else loss(wv, lo))) / (2 * eps)
        return g

    want = {"W": numeric("W"), "b": numeric("b")}
    got = {}
    for fix in (1, 0):
        oracle.lib.orc_set_fix_act_grad(fix)
        try:
            f = _activation_chain(oracle, xv, wv, bv, tv)
            f.minimize({}, lr, 1)
            leaves = dict(f.leaves())
            got[fix] = {"W": (wv.astype(np.float64) - leaves["W"]) / lr, "b": (bv.astype(np.float64) - leaves["b"]) / lr}
        finally:
            oracle.lib.orc_set_fix_act_grad(0)
    for k in ("W", "b"):
        scale = np.max(np.abs(want[k]))
        assert np.max(np.abs(got[1][k] - want[k])) < 2e-2 * scale, k      # corrected rule: the true gradient
        assert np.max(np.abs(got[0][k] - want[k])) > 0.3 * scale, k       # reference rule (Q3): not a gradient


def test_random_graphs_port_vs_reference_backend_and_exact_vs_dag(oracle):
    """Random expression trees (tests/plan_fuzz.py's generator: every operator, scalar / matrix mixes, broadcasts,
    scalar <- matrix means, square products) through the oracle: the port's arithmetic must equal the reference-compiled
    backend's bit for bit, and — a tree has no shared non-leaf — dag mode must equal exact mode bit for bit."""
    from plan_fuzz import random_graph
    if not oracle.use_ref("O0"):
        pytest.skip("oracle/_ref is not built")
    try:
        for g in range(25):
            runs = []
            for backend, mode in (("ref", "exact"), ("port", "exact"), ("port", "dag")):
                oracle.use_ref("O0") if backend == "ref" else oracle.use_port()
                oracle.set_mode(mode)
                f = random_graph(oracle, np.random.default_rng(1000 + g), 1 + g % 5)
                tr = np.asarray(f.minimize({}, 0.01, 3, trace=True))
                runs.append((tr, [np.asarray(v) for _, v in f.leaves()]))
            for other in runs[1:]:
                assert np.array_equal(runs[0][0].view(np.int32), other[0].view(np.int32)), g
                assert len(runs[0][1]) == len(other[1])
                for a, b in zip(runs[0][1], other[1]):
                    assert np.array_equal(a.view(np.int32), b.view(np.int32)), g
    finally:
        oracle.set_mode("exact")
        if not oracle.use_ref("O0"):
            oracle.use_port()


@pytest.mark.parametrize("T,batch,d_in,d", [(1, 6, 5, 7), (2, 10, 10, 10), (3, 12, 9, 8), (4, 33, 17, 20)])
def test_float64_dag_restatement_is_pinned_to_the_oracle(oracle, T, batch, d_in, d):
    """tests/f64_graph.py (the multithreaded float64 restatement used for full-size parity on the GPU box) against
    cg_oracle.c's dag mode: same leaf list (names, order, count — so no identity elimination fired in either), same loss
    trajectory and same leaves after 5 iterations, within f32 rounding of the oracle (1e-5, per element)."""
    from f64_graph import get_f64
    from helpers import close_err
    f64 = get_f64()
    oracle.set_mode("dag")
    try:
        fo = readme_lstm(oracle, d, T=T, seed=70 + T, batch=batch, d_in=d_in)
        fd = readme_lstm(f64, d, T=T, seed=70 + T, batch=batch, d_in=d_in)
        to = np.asarray(fo.minimize({}, 0.01, 5, trace=True))
        td = np.asarray(fd.minimize({}, 0.01, 5, trace=True))
        assert close_err(to, td, 1e-5) <= 1.0
        lo, ld = fo.leaves(), fd.leaves()
        assert [n for n, _ in lo] == [n for n, _ in ld]
        for (n, vo), (_, vd) in zip(lo, ld):
            assert close_err(vo, vd, 1e-5) <= 1.0, n
    finally:
        oracle.set_mode("exact")


def test_close_err_catches_a_wrong_small_entry():
    """The per-element criterion (helpers.close_err) rejects what the max-norm `rel_err` lets through: a wrong small
    entry next to one large entry."""
    from helpers import close_err
    b = np.full((8, 8), 1e-3, dtype=np.float32)
    b[0, 0] = 100.0
    a = b.copy()
    a[3, 3] = 2e-3                      # 100 % wrong, but 1e-5 of the largest entry
    assert rel_err(a, b) < 2e-5
    assert close_err(a, b, 1e-5) > 1.0
    assert close_err(b * np.float32(1 + 5e-6), b, 1e-5) <= 1.0


def test_corrected_broadcast_sum_and_sub_sign_match_finite_differences(oracle):
    """SURVEY §8f-2, the two other corrected-semantics switches, each against central finite differences of the loss the
    loop actually descends: L = sum of the root's elements (the root error is ones, src/optimize.rs:69).
      * broadcast_sum: a scalar under a matrix-shaped error receives the SUM of it (d L / d x), not the reference's mean;
      * sub_sign: `0 - x` is `-x` (the reference's shared identity rule turns it into `x`, Q4)."""
    F, Cn = oracle.Function, oracle.Constant
    B = np.array([[0.5, -1.0, 2.0], [0.25, 1.5, -0.75]], dtype=np.float32)
    lr = 1e-3

    def build(x0):
        x = F.scalar_param("x", x0)
        return ((x * F.scalar(3.0) + F.constant(Cn(B))).sq() + F.scalar(0.5)).sq()  # x is used ONCE (each use is its own leaf, Q1)

    def loss(x0):
        f = build(x0)
        return float(np.sum(np.asarray(f.minimize({}, lr, 1, trace=True), dtype=np.float64)))

    x0, h = 0.3, 1e-3
    fd = (loss(x0 + h) - loss(x0 - h)) / (2 * h)
    got = {}
    for fix in (0, 1):
        oracle.lib.orc_set_fix_broadcast_sum(fix)
        try:
            f = build(x0)
            f.minimize({}, lr, 1)
            got[fix] = (x0 - float(f.leaves()[0][1])) / lr
        finally:
            oracle.lib.orc_set_fix_broadcast_sum(0)
    # the leaf sits under nested matrix errors, so the reference's means compound; the corrected sum is the derivative
    assert abs(got[1] - fd) <= 2e-3 * abs(fd)
    assert abs(got[0] - fd) > 0.5 * abs(fd)

    def step(fix):
        oracle.lib.orc_set_fix_sub_sign(fix)
        try:
            m = F.param("m", Cn(np.full((2, 2), 0.5, dtype=np.float32)))
            f = F.constant(Cn(np.zeros((2, 2), dtype=np.float32))) - m
            f.minimize({}, 0.1, 1)
            return np.asarray(f.leaves()[0][1])
        finally:
            oracle.lib.orc_set_fix_sub_sign(0)
    assert np.allclose(step(0), 0.4)   # reference: 0 - m IS m, so m descends
    assert np.allclose(step(1), 0.6)   # corrected: d(sum(-m))/dm = -1, so m ascends


def test_tied_parameters_match_finite_differences(oracle):
    """SURVEY §8f-2, tied parameters: with fix_tied_params every use of a parameter NAME is one parameter — the backward
    walk accumulates one gradient per name from un-updated values and every copy takes the same update. On the
    reference's own `complex_test` graph ((x-2)*(x+B)-10)^2 (src/main.rs:144-154), where x is used twice, the update then
    equals -lr * dL/dx of L = sum of the root's elements (with the broadcast sum fixed too), in exact AND dag mode, and
    both copies of x stay identical; without the switch the two copies drift apart (Q1)."""
    F, Cn = oracle.Function, oracle.Constant
    B = np.array([[1.0, 2.0], [3.0, 4.0]], dtype=np.float32)
    lr = 1e-4

    def build(x0):
        x = F.scalar_param("x", x0)
        return ((x - F.scalar(2.0)) * (x + F.constant(Cn(B))) - F.scalar(10.0)).sq()

    def loss(x0):
        return float(np.sum(np.asarray(build(x0).minimize({}, lr, 1, trace=True), dtype=np.float64)))

    x0, h = 0.5, 1e-3
    fd = (loss(x0 + h) - loss(x0 - h)) / (2 * h)
    for mode in ("exact", "dag"):
        oracle.set_mode(mode)
        oracle.lib.orc_set_fix_tied_params(1)
        oracle.lib.orc_set_fix_broadcast_sum(1)
        try:
            f = build(x0)
            f.minimize({}, lr, 1)
            leaves = [float(v) for _, v in f.leaves()]
            assert len(leaves) == 2 and leaves[0] == leaves[1]
            assert abs((x0 - leaves[0]) / lr - fd) <= 2e-3 * abs(fd), (mode, (x0 - leaves[0]) / lr, fd)
            f.minimize({}, lr, 3)
            leaves = [float(v) for _, v in f.leaves()]
            assert leaves[0] == leaves[1]
        finally:
            oracle.lib.orc_set_fix_tied_params(0)
            oracle.lib.orc_set_fix_broadcast_sum(0)
            oracle.set_mode("exact")
    f = build(x0)
    f.minimize({}, lr, 3)
    leaves = [float(v) for _, v in f.leaves()]
    assert leaves[0] != leaves[1]  # the reference: two private copies (Q1)
