"""Parity of the CUDA graph path (libcg_b200.so through its C ABI) against the CPU oracle.

Tolerances: everything without a transcendental is BIT-EXACT (integer compare of the f32 bits);
graphs with sigmoid/tanh are held to 1e-5 relative (north_star) because CUDA expf/tanhf differ from
glibc's by <= 2 ulp. TF32 runs use 1e-3 and say so.
"""
import numpy as np
import pytest

from helpers import c2_graph, close_err, leaves_dict, readme_lstm, rel_err
from reference_cases import CASES, run_case

pytestmark = pytest.mark.gpu

TRANSCENDENTAL = {"tanh_test"}  # sigmoid_test is identity-eliminated (Q4): no sigmoid is ever evaluated


@pytest.mark.parametrize("name", list(CASES))
def test_reference_suite_matches_oracle(cg, oracle, name):
    """src/main.rs:85-166: the threshold holds AND the 10-iteration trajectory equals the oracle's."""
    got, goal = run_case(cg, name, trace=True)
    want, _ = run_case(oracle, name, trace=True)
    assert len(got) == len(want)
    for (fg, tg), (fo, to) in zip(got, want):
        assert fg.all_less_than(goal), f"{name}: threshold {goal} not met"
        if name in TRANSCENDENTAL:
            assert rel_err(tg, to) < 1e-5
            assert rel_err(fg.value(), fo.value()) < 1e-5
        else:
            assert np.array_equal(np.asarray(tg).view(np.int32), np.asarray(to).view(np.int32)), name
            assert np.array_equal(np.asarray(fg.value()).view(np.int32), np.asarray(fo.value()).view(np.int32))
            for (ng, vg), (no, vo) in zip(fg.leaves(), fo.leaves()):
                assert ng == no
                assert np.array_equal(np.asarray(vg).view(np.int32), np.asarray(vo).view(np.int32))


def test_c1_readme_scalar_golden(cg):
    """BASELINE config 1: minimize (x+3)^2, x=1, lr=.01, 1000 iters — golden values of SURVEY §8c."""
    F = cg.Function
    x = F.scalar_param("x", 1.0)
    f = (x + F.scalar(3.0)).sq()
    tr = f.minimize({}, 0.01, 1000, trace=True)
    golden = {0: 16.0, 1: 15.366400718688965, 2: 14.757889747619629, 9: 11.122164726257324,
              99: 0.2930106520652771, 999: 3.2741809263825417e-11}
    for i, v in golden.items():
        assert np.float32(tr[i]) == np.float32(v), (i, tr[i], v)
    assert f.leaves()[0][1] == np.float32(-2.9999942779541016)


@pytest.mark.parametrize("fuse,graph", [(True, True), (False, True), (True, False)])
def test_c2_elementwise_small(cg, oracle, fuse, graph):
    """BASELINE config 2 at 64x64 / 100 steps: fused, unfused and eager launches agree with the oracle."""
    cg.set_fuse(fuse)
    cg.set_cuda_graph(graph)
    fg, fo = c2_graph(cg, 64), c2_graph(oracle, 64)
    tg = fg.minimize({}, 0.01, 100, trace=True)
    to = fo.minimize({}, 0.01, 100, trace=True)
    assert rel_err(tg, to) < 1e-5
    lg, lo = leaves_dict(fg), leaves_dict(fo)
    assert list(lg) == list(lo) and len(lg) == 3  # W (sigmoid use), V, W (abs use): Q1
    for k in lg:
        assert rel_err(lg[k], lo[k]) < 1e-5, k


def test_c2_fused_is_one_kernel(cg):
    f = c2_graph(cg, 256)
    st = cg.plan_stats(f, 0.01)
    assert st["kernels_per_iter"] == 1 and st["gemms"] == 0
    # 3 leaves read + 3 written + root value written: 7 * 4 bytes per element
    assert st["ew_bytes_per_iter"] == 7 * 4 * 256 * 256


@pytest.mark.parametrize("d", [5, 10, 30])
def test_c3_readme_lstm_exact_mode(cg, oracle, d):
    """BASELINE config 3: README LSTM, T=2, exact (tree-walk) mode, loss trajectory + all 28 leaves."""
    iters = 20
    fg, fo = readme_lstm(cg, d), readme_lstm(oracle, d)
    tg = fg.minimize({}, 0.01, iters, trace=True)
    to = fo.minimize({}, 0.01, iters, trace=True)
    assert np.all(np.isfinite(to))
    assert rel_err(tg, to) < 1e-5
    lg, lo = leaves_dict(fg), leaves_dict(fo)
    assert list(lg) == list(lo) and len(lg) == 28
    for k in lg:
        assert rel_err(lg[k], lo[k]) < 1e-5, k


def test_c3_gradients_first_step(cg, oracle):
    """Per-parameter gradients (p_before - p_after)/lr of one iteration, exact mode, d=10."""
    lr = 0.01
    fg, fo = readme_lstm(cg, 10), readme_lstm(oracle, 10)
    bg, bo = leaves_dict(fg), leaves_dict(fo)
    fg.minimize({}, lr, 1)
    fo.minimize({}, lr, 1)
    ag, ao = leaves_dict(fg), leaves_dict(fo)
    for k in bg:
        gg, go = (bg[k] - ag[k]) / lr, (bo[k] - ao[k]) / lr
        # a gradient recovered from two f32 parameter values carries +-ulp(p)/lr of rounding noise
        noise = 4 * np.finfo(np.float32).eps * float(np.max(np.abs(bo[k]))) / lr
        assert np.max(np.abs(gg - go)) <= 1e-5 * float(np.max(np.abs(go))) + noise, k


@pytest.mark.parametrize("T,batch,d_in,d", [(1, 6, 5, 7), (3, 12, 9, 8), (4, 33, 17, 20)])
def test_lstm_dag_mode_nonsquare(cg, oracle, T, batch, d_in, d):
    """dag mode on shapes the reference cannot run (batch != input != hidden, Q7) against the oracle's dag
    restatement."""
    cg.set_mode("dag")
    oracle.set_mode("dag")
    try:
        fg = readme_lstm(cg, d, T=T, seed=40 + T, batch=batch, d_in=d_in)
        fo = readme_lstm(oracle, d, T=T, seed=40 + T, batch=batch, d_in=d_in)
        tg = fg.minimize({}, 0.01, 10, trace=True)
        to = fo.minimize({}, 0.01, 10, trace=True)
        assert rel_err(tg, to) < 1e-5
        lg, lo = leaves_dict(fg), leaves_dict(fo)
        assert list(lg) == list(lo)
        for k in lg:
            assert rel_err(lg[k], lo[k]) < 1e-5, k
    finally:
        oracle.set_mode("exact")


def test_dag_equals_exact_without_shared_nodes(cg):
    """SURVEY §8 mode table: the two modes are bit-identical on graphs without a shared non-leaf (LSTM T=1)."""
    res = []
    for mode in ("exact", "dag"):
        cg.set_mode(mode)
        f = readme_lstm(cg, 12, T=1, seed=7)
        tr = f.minimize({}, 0.01, 5, trace=True)
        res.append((tr, leaves_dict(f)))
    assert np.array_equal(res[0][0].view(np.int32), res[1][0].view(np.int32))
    for k in res[0][1]:
        assert np.array_equal(res[0][1][k].view(np.int32), res[1][1][k].view(np.int32)), k


@pytest.mark.parametrize("tf32,shape", [(False, (3, 12, 9, 8)), (False, (2, 80, 72, 96)), (True, (2, 256, 256, 256))])
def test_hoisting_is_bit_identical(cg, tf32, shape):
    """cg_set_hoist only reorders the tape along its data dependencies (plan.cpp hoist_backward): the root values of
    every iteration and every leaf must come out bit for bit the same with and without it."""
    T, batch, d_in, d = shape
    cg.set_mode("dag")
    cg.set_tf32(tf32)
    res = []
    for hoist in (False, True):
        cg.set_hoist(hoist)
        f = readme_lstm(cg, d, T=T, seed=77, batch=batch, d_in=d_in, scale=1.0 / np.sqrt(d))
        tr = f.minimize({}, 0.01, 4, trace=True)
        res.append((np.asarray(tr), leaves_dict(f), cg.plan_describe(f, 0.01)))
    assert res[0][2] != res[1][2]  # the kernel order did change
    assert np.array_equal(res[0][0].view(np.int32), res[1][0].view(np.int32))
    for k in res[0][1]:
        assert np.array_equal(res[0][1][k].view(np.int32), res[1][1][k].view(np.int32)), k


@pytest.mark.parametrize("case", ["c3_exact_d5", "c3_exact_d30", "c3_dag_d30", "broadcast_avg", "inputs"])
def test_resident_executor_is_bit_identical_to_the_graph(cg, case):
    """The resident executors — one CTA running the whole tape with the iteration loop inside the kernel, either
    interpreting it out of L1/L2 (vm_kernel.cu) or compiled with every value in shared memory (ew_jit.cpp tape_jit) —
    must give the same bits as the per-step kernels replayed as a CUDA graph: root value of every iteration, every
    leaf."""
    res = []
    for executor in ("graph", "interpreted", "compiled"):
        dim = 0 if executor == "graph" else 48
        cg.set_vm_max_dim(dim)
        cg.set_tape_jit(executor == "compiled")
        args = {}
        if case.startswith("c3"):
            _, mode, d = case.split("_")
            cg.set_mode(mode)
            f = readme_lstm(cg, int(d[1:]))
            lr, iters = 0.01, 12
        elif case == "broadcast_avg":
            # scalar <- matrix gradients (mean, Q5) and scalar-matrix broadcasts: main.rs:133-154
            F = cg.Function
            x = F.scalar_param("x", 0.5)
            m = F.param("m", cg.Constant(np.arange(12, dtype=np.float32).reshape(3, 4) / 7 - 0.6))
            b = F.matrix(3, 4, list(np.linspace(-1, 1.5, 12)))
            f = (((x - F.scalar(2.)) * (m + b) - F.scalar(10.)).sq() + (x * m).tanh()).abs()
            lr, iters = 0.004, 15
        else:
            F = cg.Function
            w = F.param("w", cg.Constant(np.linspace(-1, 1, 30, dtype=np.float32).reshape(6, 5)))
            f = ((cg.dot(F.input("x", [8, 6]), w)).sigmoid() * F.input("s", []) - F.input("t", [8, 5])).sq()
            rng = np.random.default_rng(3)
            args = {"x": cg.Constant(rng.uniform(-1, 1, (8, 6)).astype(np.float32)), "s": cg.Constant.Scalar(1.5),
                    "t": cg.Constant(rng.uniform(-1, 1, (8, 5)).astype(np.float32))}
            lr, iters = 0.05, 9
        st = cg.plan_stats(f, lr)
        small = case != "c3_exact_d30" and case != "c3_dag_d30"  # the interpreter only takes plans with d <= 20
        want = {"graph": 0, "interpreted": 1 if small else 0, "compiled": 2}[executor]
        assert st["vm_resident"] == want, st
        assert (st["tape_smem_bytes"] > 0) == (executor == "compiled")
        l0 = cg.launch_count()
        tr = np.array(f.minimize(args, lr, iters, trace=True))
        launched = cg.launch_count() - l0
        tr2 = np.array(f.minimize(args, lr, 3, 2, trace=True))  # print_freq = 2: the loop is cut at every print
        res.append((tr, tr2, leaves_dict(f), np.asarray(f.value())))
        if want:
            assert launched <= 1 + len(args)  # ONE kernel for all iterations (+ the layout kernel of a matrix input)
    cg.set_mode("exact")
    cg.set_tape_jit(True)
    a = res[0]
    for b in res[1:]:
        assert a[0].shape == b[0].shape and np.array_equal(a[0].view(np.int32), b[0].view(np.int32))
        assert np.array_equal(a[1].view(np.int32), b[1].view(np.int32))
        assert np.array_equal(a[3].view(np.int32), b[3].view(np.int32))
        for k in a[2]:
            assert np.array_equal(np.asarray(a[2][k]).view(np.int32), np.asarray(b[2][k]).view(np.int32)), k


def test_root_store_elision_is_unobservable(cg, oracle):
    """cg_set_elide_root: iterations whose root value nobody can observe run without storing it. Leaves, the cached
    root value after the call (Q12: the forward value of the LAST iteration), traces and prints must be what they are
    with the store in every iteration, and equal to the oracle."""
    n = 96  # 9216 elements: above the resident executor's limit, so this is the CUDA-graph path
    res = []
    for elide in (False, True):
        cg.set_elide_root(elide)
        f = c2_graph(cg, n)
        st = cg.plan_stats(f, 0.01)
        assert st["vm_resident"] == 0 and st["quiet_variant"] == (1 if elide else 0)
        if elide:
            assert st["ew_bytes_per_iter"] == 28 * n * n and st["quiet_ew_bytes_per_iter"] == 24 * n * n
        f.minimize({}, 0.01, 7)
        v1 = np.array(f.value())
        tr = np.array(f.minimize({}, 0.01, 4, trace=True))
        f.minimize({}, 0.01, 1)
        v2 = np.array(f.value())
        res.append((v1, tr, v2, leaves_dict(f)))
    cg.set_elide_root(True)
    a, b = res
    for x, y in zip(a[:3], b[:3]):
        assert np.array_equal(x.view(np.int32), y.view(np.int32))
    for k in a[3]:
        assert np.array_equal(a[3][k].view(np.int32), b[3][k].view(np.int32)), k
    fo = c2_graph(oracle, n)
    fo.minimize({}, 0.01, 7)
    assert rel_err(b[0], np.asarray(fo.value())) < 1e-5


@pytest.mark.parametrize("mode", ["exact", "dag"])
def test_corrected_activation_gradients_match_the_oracle(cg, oracle, mode):
    """with cg_set_fix_act_grad(1) Sigmoid / Tanh pass error * derivative to their child (the true
    chain rule instead of Q3). Same switch on the oracle (whose corrected rule is checked against finite differences in
    tests/test_oracle.py); the LSTM then trains through its gates, and both sides must still agree to 1e-5."""
    cg.set_mode(mode)
    oracle.set_mode(mode)
    cg.set_fix_act_grad(True)
    oracle.lib.orc_set_fix_act_grad(1)
    try:
        fg, fo = readme_lstm(cg, 8, T=2, seed=5), readme_lstm(oracle, 8, T=2, seed=5)
        tg = fg.minimize({}, 0.01, 8, trace=True)
        to = fo.minimize({}, 0.01, 8, trace=True)
        assert rel_err(tg, to) < 1e-5
        lg, lo = leaves_dict(fg), leaves_dict(fo)
        for k in lg:
            assert rel_err(lg[k], lo[k]) < 1e-5, k
        # and the corrected rule is a different computation from the reference's (Q3)
        cg.set_fix_act_grad(False)
        fq = readme_lstm(cg, 8, T=2, seed=5)
        fq.minimize({}, 0.01, 8)
        lq = leaves_dict(fq)
        assert any(rel_err(lq[k], lg[k]) > 1e-3 for k in lg)
    finally:
        cg.set_fix_act_grad(False)
        oracle.lib.orc_set_fix_act_grad(0)
        oracle.set_mode("exact")
        cg.set_mode("exact")


def test_eval_and_symbolic_grad(cg):
    """SURVEY §8 row f4 — `eval` and the symbolic `grad` of src/optimize.rs:8-56 on the same kernels (the README's
    "backprop vs naive" comparison). eval against closed forms; grad against the analytic derivative, element by
    element, and against finite differences of eval; and the README's claim itself: one backprop step (with the chain
    rule completed, cg_set_fix_act_grad) moves the parameter by -lr * mean(grad.eval()) in total — "in total" because
    every use site of x is its own copy (Q1) and receives its own share, and "mean" because that is the reference's
    scalar <- matrix rule (Q5). Neither call may change a parameter."""
    F = cg.Function
    rng = np.random.default_rng(12)
    mv = rng.uniform(-1, 1, (5, 7)).astype(np.float32)
    x0 = np.float32(0.37)

    def build():
        x = F.scalar_param("x", float(x0))
        m = F.constant(cg.Constant(mv))
        return ((x * m + x.sq()).tanh() - (x * m).sigmoid()) * x.abs(), x

    def closed(xv):
        xv = np.float64(xv)
        u = xv * mv.astype(np.float64) + xv * xv
        s = 1.0 / (1.0 + np.exp(-xv * mv.astype(np.float64)))
        val = (np.tanh(u) - s) * abs(xv)
        du = (1 - np.tanh(u) ** 2) * (mv + 2 * xv) - s * (1 - s) * mv
        return val, du * abs(xv) + (np.tanh(u) - s) * np.sign(xv)

    f, x = build()
    val, dval = closed(x0)
    got = np.asarray(f.eval({}))
    assert got.shape == (5, 7) and rel_err(got, val) < 1e-5
    g = f.grad("x")
    gv = np.asarray(g.eval({}))
    assert rel_err(gv, dval) < 1e-4
    # finite differences of eval (float32 forward: a loose bound)
    eps = 1e-2
    fp, _ = build()
    hi = np.asarray((F.scalar(float(x0 + eps)) * F.constant(cg.Constant(mv))).eval({}))  # eval of a constant-only graph
    assert rel_err(hi, (x0 + eps) * mv) < 1e-6
    assert rel_err(gv, (closed(x0 + eps)[0] - closed(x0 - eps)[0]) / (2 * eps)) < 1e-2
    # nothing moved
    assert all(float(v) == float(x0) for n, v in f.leaves() if n == "x")
    # backprop vs naive: the same update
    cg.set_fix_act_grad(True)
    try:
        f2, _ = build()
        f2.minimize({}, 0.01, 1)
        moved = sum((float(x0) - float(v)) / 0.01 for n, v in f2.leaves() if n == "x")
    finally:
        cg.set_fix_act_grad(False)
    assert abs(moved - float(gv.astype(np.float64).mean())) < 2e-3 * max(1.0, abs(float(gv.mean())))
    # an input named like the parameter, products, signum: the reference's panics
    w = F.param("w", cg.Constant(mv))
    for bad, msg in ((cg.dot(F.constant(cg.Constant(mv.T.copy())), w), "not implemented"), (w.signum(), "nondifferentiable")):
        with pytest.raises(Exception, match=msg):
            bad.grad("w")
    assert float(np.asarray(w.sq().grad("nope").eval({}))) == 0.0  # a parameter the function does not mention


def test_medium_lstm_tiled_gemm(cg, oracle):
    """Products above the strict-kernel limit take the tiled FP32 SIMT kernel (FMA accumulation): 1e-5."""
    cg.set_mode("dag")
    oracle.set_mode("dag")
    try:
        fg = readme_lstm(cg, 96, T=2, seed=11, batch=80, d_in=72, scale=0.1)
        fo = readme_lstm(oracle, 96, T=2, seed=11, batch=80, d_in=72, scale=0.1)
        tg = fg.minimize({}, 0.01, 3, trace=True)
        to = fo.minimize({}, 0.01, 3, trace=True)
        assert rel_err(tg, to) < 1e-5
        lg, lo = leaves_dict(fg), leaves_dict(fo)
        for k in lg:
            assert rel_err(lg[k], lo[k]) < 1e-5, k
    finally:
        oracle.set_mode("exact")


def test_inputs_and_maximize(cg, oracle):
    """Function::input fed through args (optimize.rs:87-90) and maximize = minimize(-f) (optimize.rs:74-81)."""
    rng = np.random.default_rng(3)
    xv = rng.uniform(-1, 1, (4, 3)).astype(np.float32)
    wv = rng.uniform(-1, 1, (3, 5)).astype(np.float32)
    out = []
    for api in (cg, oracle):
        x = api.Function.input("x", [4, 3])
        w = api.Function.param("w", api.Constant(wv))
        f = api.dot(x, w).sq()
        tr = f.maximize({"x": api.Constant(xv)}, 0.01, 5, trace=True)
        out.append((tr, leaves_dict(f)))
    assert np.array_equal(np.asarray(out[0][0]).view(np.int32), np.asarray(out[1][0]).view(np.int32))
    for k in out[0][1]:
        assert np.array_equal(out[0][1][k].view(np.int32), out[1][1][k].view(np.int32))


def test_errors_surface_as_reference_messages(cg):
    from computational_graph_b200 import GraphError
    F = cg.Function
    a = F.param("a", cg.Constant(np.ones((2, 3), np.float32) * 2))
    b = F.param("b", cg.Constant(np.ones((4, 3), np.float32) * 2))
    with pytest.raises(GraphError, match="dims"):
        a + b
    with pytest.raises(GraphError, match="inner_dim"):
        cg.dot(a, b)
    with pytest.raises(GraphError, match="scalars"):
        cg.dot(a, F.scalar(2.0))
    with pytest.raises(GraphError, match="not differentiable"):
        a.signum().minimize({}, 0.1, 1)
    x = F.input("x", [2, 2])
    with pytest.raises(GraphError, match="missing arg"):
        (x + F.param("p", cg.Constant(np.ones((2, 2), np.float32)))).minimize({}, 0.1, 1)


@pytest.mark.parametrize("pair", [False, True])
def test_lstm_tf32_tensor_core_path(cg, oracle, pair):
    """LSTM in dag mode with every product on the tcgen05 TF32 kernel (pair = True: the 2-SM cta_group::2 kernel) (forward NN, dX NT, dW TN with the fused
    SGD epilogue, dX accumulation epilogue). TF32 keeps 10 mantissa bits: tolerance 1e-3 (north_star)."""
    cg.set_mode("dag")
    cg.set_tf32(True)
    cg.set_tf32_pair(pair)
    oracle.set_mode("dag")
    try:
        fg = readme_lstm(cg, 256, T=2, seed=21, batch=256, d_in=256, scale=1.0 / 16)
        fo = readme_lstm(oracle, 256, T=2, seed=21, batch=256, d_in=256, scale=1.0 / 16)
        st = cg.plan_stats(fg, 0.01)
        assert st["gemms"] > 0 and st["gemms_fused_sgd"] > 0
        tg = fg.minimize({}, 0.01, 2, trace=True)
        to = fo.minimize({}, 0.01, 2, trace=True)
        assert rel_err(tg, to) < 1e-3
        lg, lo = leaves_dict(fg), leaves_dict(fo)
        for k in lg:
            assert rel_err(lg[k], lo[k]) < 1e-3, k
    finally:
        oracle.set_mode("exact")
        cg.set_tf32_pair(True)


def test_host_transfers_inside_the_iteration_graph(cg, oracle):
    """`minimize(&args, lr, 1)` with page-locked host buffers runs H2D + layout + iteration + D2H as ONE graph
    (exec.cpp launch_io_graph). Results must be bit-identical to the pageable path (upload, then iterate), across
    calls that change the host pointers and the input values, and equal to the oracle."""
    rng = np.random.default_rng(11)
    B, I, H = 8, 6, 5
    wv = rng.uniform(-1, 1, (I, H)).astype(np.float32)
    bv = rng.uniform(-1, 1, (B, H)).astype(np.float32)
    xs = [rng.uniform(-1, 1, (B, I)).astype(np.float32) for _ in range(4)]
    ts = [rng.uniform(-1, 1, (B, H)).astype(np.float32) for _ in range(4)]

    def build(api):
        x = api.Function.input("x", [B, I])
        t = api.Function.input("t", [B, H])
        s = api.Function.input("s", [])  # a scalar input rides the same path (no layout kernel)
        w = api.Function.param("w", api.Constant(wv))
        b = api.Function.param("b", api.Constant(bv))
        return ((api.dot(x, w) + b).tanh() * s - t).sq()

    def run(api, pinned, io):
        if api is cg:
            cg.set_io_graph(io)
        f = build(api)
        traces = []
        for k in range(4):
            if pinned:
                # fresh page-locked buffers on every call for k < 2 (pointer update), the same ones afterwards
                xa, ta = cg.pinned_empty((B, I)), cg.pinned_empty((B, H))
                xa[...] = xs[k]
                ta[...] = ts[k]
                keep.append((xa, ta))
            else:
                xa, ta = xs[k], ts[k]
            args = {"x": api.Constant(xa), "t": api.Constant(ta), "s": api.Constant.Scalar(0.5 + k)}
            traces.append(np.array(f.minimize(args, 0.05, 1, trace=True)))
        more = np.array(f.minimize(args, 0.05, 3, trace=True))  # iters > 1: upload once, then the plain graph
        return traces, more, leaves_dict(f)

    keep = []
    cg.set_vm_max_dim(0)  # these shapes would otherwise run on the resident executor, which has no transfer graph
    launches0 = cg.launch_count()
    try:
        got = run(cg, True, True)
        io_launches = cg.launch_count() - launches0
        ref = run(cg, False, True)       # pageable buffers: never eligible
        off = run(cg, True, False)       # page-locked but switched off
    finally:
        cg.set_io_graph(True)
    want = run(oracle, False, False)
    assert io_launches > 0
    for other in (ref, off):
        for a, b in zip(got[0], other[0]):
            assert np.array_equal(a.view(np.int32), b.view(np.int32))
        assert np.array_equal(got[1].view(np.int32), other[1].view(np.int32))
        for k in got[2]:
            assert np.array_equal(got[2][k].view(np.int32), other[2][k].view(np.int32)), k
    for a, b in zip(got[0], want[0]):
        assert rel_err(a, b) < 1e-5
    for k in got[2]:
        assert rel_err(got[2][k], want[2][k]) < 1e-5, k


@pytest.mark.parametrize("shape", [(301, 303), (256, 512)])
def test_specialised_elementwise_kernels_equal_the_interpreter(cg, shape):
    """ew_jit.cpp: the straight-line kernel NVRTC compiles from a fused program is bit-identical to the bytecode
    interpreter running the same program — C2 graph (sigmoid, tanh, abs, SGD) on a ragged and an aligned size."""
    rng = np.random.default_rng(5)
    wv = rng.uniform(-1, 1, shape).astype(np.float32)
    vv = rng.uniform(-1, 1, shape).astype(np.float32)

    def run(jit_min):
        cg.set_ew_jit_min(jit_min)
        W = cg.Function.param("W", cg.Constant(wv))
        V = cg.Function.param("V", cg.Constant(vv))
        f = (W.sigmoid() - V.sq()).tanh() + W.abs()
        st = cg.plan_stats(f, 0.01)
        tr = np.array(f.minimize({}, 0.01, 4, trace=True))
        return st, tr, leaves_dict(f)

    try:
        st_j, tr_j, lv_j = run(0)
        st_i, tr_i, lv_i = run(-1)
    finally:
        cg.set_ew_jit_min(32768)
    assert st_j["ew_specialised"] == st_j["ew_kernels"] > 0 and st_i["ew_specialised"] == 0
    assert np.array_equal(tr_j.view(np.int32), tr_i.view(np.int32))
    for k in lv_j:
        assert np.array_equal(lv_j[k].view(np.int32), lv_i[k].view(np.int32)), k


def test_specialised_elementwise_kernels_lstm_dag(cg, oracle):
    """Every elementwise group of a dag-mode LSTM on the specialised path: bit-identical to the interpreter and
    within 1e-5 of the oracle (transcendentals)."""
    cg.set_mode("dag")
    oracle.set_mode("dag")
    try:
        out = []
        for jit_min in (0, -1):
            cg.set_ew_jit_min(jit_min)
            f = readme_lstm(cg, 24, T=3, seed=77, batch=40, d_in=16)
            out.append((np.array(f.minimize({}, 0.01, 3, trace=True)), leaves_dict(f), cg.plan_stats(f, 0.01)))
        fo = readme_lstm(oracle, 24, T=3, seed=77, batch=40, d_in=16)
        to = fo.minimize({}, 0.01, 3, trace=True)
        lo = leaves_dict(fo)
    finally:
        cg.set_ew_jit_min(32768)
        oracle.set_mode("exact")
    assert out[0][2]["ew_specialised"] > 0
    assert np.array_equal(out[0][0].view(np.int32), out[1][0].view(np.int32))
    for k in out[0][1]:
        assert np.array_equal(out[0][1][k].view(np.int32), out[1][1][k].view(np.int32)), k
        assert rel_err(out[0][1][k], lo[k]) < 1e-5, k
    assert rel_err(out[0][0], to) < 1e-5


@pytest.mark.parametrize("executor", ["resident", "graph"])
def test_leaf_checkpoint_restore_replays_bit_identically(cg, executor):
    """save every leaf (`leaves()`), keep training, perturb, restore (`load_leaves` -> cg_leaf_set), and the
    trajectory from the restored state is bit-identical to the one first taken from it — on both executors."""
    cg.set_mode("exact")
    if executor == "graph":
        cg.set_vm_max_dim(0)
    f = readme_lstm(cg, 10)
    f.minimize({}, 0.01, 5)
    saved = f.leaves()
    first = np.asarray(f.minimize({}, 0.01, 5, trace=True))
    for i, (_, v) in enumerate(saved):
        f.set_leaf(i, np.full_like(v, 0.123))
    assert all(np.all(v == np.float32(0.123)) for _, v in f.leaves())
    f.load_leaves(saved)
    for (n0, v0), (n1, v1) in zip(saved, f.leaves()):
        assert n0 == n1 and np.array_equal(np.asarray(v0).view(np.int32), np.asarray(v1).view(np.int32))
    again = np.asarray(f.minimize({}, 0.01, 5, trace=True))
    assert np.array_equal(first.view(np.int32), again.view(np.int32))
    x = cg.Function.scalar_param("x", 1.0)
    g = (x + cg.Function.scalar(3.0)).sq()
    g.minimize({}, 0.01, 3)
    g.set_leaf(0, 1.0)
    assert g.leaves()[0][1] == np.float32(1.0)
    with pytest.raises(Exception):
        g.set_leaf(5, 1.0)


@pytest.mark.parametrize("executor", ["resident", "graph"])
def test_print_freq_loss_log(cg, capfd, executor):
    """`minimize(&args, lr, iters, print_freq)` prints the root value every print_freq iterations (src/optimize.rs:66-68):
    the lines are the traced values of those iterations, in order — read back from the device-side ring, not with a
    blocking copy per print."""
    if executor == "graph":
        cg.set_vm_max_dim(0)

    def build():
        F = cg.Function
        m = F.param("m", cg.Constant(np.array([[0.5, -1.0], [2.0, 0.25]], dtype=np.float32)))
        return (m * m + F.scalar(1.0)).sq()

    tr = np.asarray(build().minimize({}, 0.01, 7, trace=True))
    capfd.readouterr()
    build().minimize({}, 0.01, 7, print_freq=2)
    out = capfd.readouterr().out.strip().splitlines()
    assert len(out) == 3 * 2  # three prints of a 2 x 2 matrix, one row per line
    for k, it in enumerate((1, 3, 5)):
        rows = np.array([[float(x) for x in out[2 * k + r].split()] for r in range(2)], dtype=np.float32)
        assert np.allclose(rows, tr[it], rtol=1e-5), (it, rows, tr[it])


def test_corrected_broadcast_sum_and_sub_sign_match_the_oracle(cg, oracle):
    """The same two switches on the CUDA path against the oracle (which the CPU tier checks against finite differences):
    bit-exact, no transcendental is involved. The sum takes the CUDA-graph executor (the resident one keeps the mean)."""
    B = np.array([[0.5, -1.0, 2.0], [0.25, 1.5, -0.75]], dtype=np.float32)

    def build(api):
        x = api.Function.scalar_param("x", 0.3)
        return ((x * api.Function.scalar(3.0) + api.Function.constant(api.Constant(B))).sq() + api.Function.scalar(0.5)).sq()

    cg.set_fix_broadcast_sum(True)
    oracle.lib.orc_set_fix_broadcast_sum(1)
    try:
        fg, fo = build(cg), build(oracle)
        tg = np.asarray(fg.minimize({}, 1e-3, 5, trace=True))
        to = np.asarray(fo.minimize({}, 1e-3, 5, trace=True))
        assert np.array_equal(tg.view(np.int32), to.view(np.int32))
        assert np.float32(fg.leaves()[0][1]) == np.float32(fo.leaves()[0][1])
        assert cg.plan_stats(fg, 1e-3)["vm_resident"] == 0
    finally:
        cg.set_fix_broadcast_sum(False)
        oracle.lib.orc_set_fix_broadcast_sum(0)
    for fix, want in ((False, 0.4), (True, 0.6)):
        cg.set_fix_sub_sign(fix)
        try:
            m = cg.Function.param("m", cg.Constant(np.full((2, 2), 0.5, dtype=np.float32)))
            f = cg.Function.constant(cg.Constant(np.zeros((2, 2), dtype=np.float32))) - m
            f.minimize({}, 0.1, 1)
            assert np.allclose(np.asarray(f.leaves()[0][1]), want)
        finally:
            cg.set_fix_sub_sign(False)


@pytest.mark.parametrize("mode", ["exact", "dag"])
def test_tied_parameters_match_the_oracle(cg, oracle, mode):
    """cg_set_fix_tied_params on the CUDA path against the oracle (checked against finite differences in the CPU tier):
    the reference's complex_test graph with both uses of x tied (bit-exact: no transcendental), and an LSTM whose time
    steps share their weights — trajectory and leaves within 1e-5, copies of one name bit-identical."""
    B = np.array([[1.0, 2.0], [3.0, 4.0]], dtype=np.float32)

    def build(api):
        x = api.Function.scalar_param("x", 0.5)
        return ((x - api.Function.scalar(2.0)) * (x + api.Function.constant(api.Constant(B))) - api.Function.scalar(10.0)).sq()

    cg.set_mode(mode)
    oracle.set_mode(mode)
    cg.set_fix_tied_params(True)
    oracle.lib.orc_set_fix_tied_params(1)
    try:
        fg, fo = build(cg), build(oracle)
        tg = np.asarray(fg.minimize({}, 1e-4, 6, trace=True))
        to = np.asarray(fo.minimize({}, 1e-4, 6, trace=True))
        assert np.array_equal(tg.view(np.int32), to.view(np.int32))
        lg, lo = fg.leaves(), fo.leaves()
        assert [np.float32(v) for _, v in lg] == [np.float32(v) for _, v in lo] and lg[0][1] == lg[1][1]
        fg, fo = readme_lstm(cg, 12, T=3, seed=5), readme_lstm(oracle, 12, T=3, seed=5)
        tg = fg.minimize({}, 1e-3, 5, trace=True)
        to = fo.minimize({}, 1e-3, 5, trace=True)
        assert close_err(tg, to, 1e-5) <= 1.0
        by_name = {}
        for (n, vg), (no, vo) in zip(fg.leaves(), fo.leaves()):
            assert n == no and close_err(vg, vo, 1e-5) <= 1.0, n
            by_name.setdefault(n, []).append(np.asarray(vg))
        assert len(by_name["Wi"]) == 3
        for n, copies in by_name.items():
            for c in copies[1:]:
                assert np.array_equal(c.view(np.int32), copies[0].view(np.int32)), n
    finally:
        cg.set_fix_tied_params(False)
        oracle.lib.orc_set_fix_tied_params(0)
        oracle.set_mode("exact")
