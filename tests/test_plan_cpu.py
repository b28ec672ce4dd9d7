"""Host logic of the tape compiler (csrc/plan.cpp), checked WITHOUT a GPU through plan-only mode
(cg_set_plan_only: placeholder addresses, nothing computed): kernel counts per BASELINE config, the fusions, the
hoisting pass (its own dependency verifier raises on a bad order), arena sizing — and that the mode can never be used
to compute. Runs tests/plan_probe.py in a process of its own (the mode is chosen once per process)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def plans():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "plan_probe.py"), "--c4"], capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def test_plan_only_mode_never_computes(plans):
    assert plans["minimize_raises"] is True


def test_c1_and_c2_are_one_fused_kernel(plans):
    """C1 (x+3)^2: forward, derivative and update in one program. C2 tanh(sigmoid(W)-sq(V))+abs(W): one pass over the
    three live leaves, 24 bytes per element (+4 for the root value) — SURVEY §8d."""
    assert plans["c1"]["kernels_per_iter"] == 1 and plans["c1"]["arena_bytes"] == 0
    c2 = plans["c2"]
    assert c2["kernels_per_iter"] == 1 and c2["arena_bytes"] == 0
    assert c2["kernels"][0].split()[1] == "EW"
    assert c2["ew_bytes_per_iter"] == 28 * 64 * 64
    # iterations whose root value is unobservable (Q12) skip its store: exactly SURVEY §8d's 24 bytes per element
    assert c2["quiet_variant"] == 1 and c2["quiet_ew_bytes_per_iter"] == 24 * 64 * 64
    assert plans["c4_hoist1"]["quiet_variant"] == 0


def test_c3_kernel_counts(plans):
    """README LSTM T=2: exact mode replays the tree walk (39 forward dots memoised to 15... + 57 backward products:
    72 launched), dag mode visits each node once (35). Every dW product carries its SGD update."""
    ex, dag = plans["c3_d30_exact"], plans["c3_d30_dag"]
    assert (ex["gemms"], ex["gemms_fused_sgd"], ex["kernels_per_iter"]) == (72, 39, 103)
    assert (dag["gemms"], dag["gemms_fused_sgd"], dag["kernels_per_iter"]) == (35, 15, 52)
    assert dag["gemm_flop_per_iter"] == 35 * 2 * 30 ** 3
    # latency-bound plans go to the resident tape executor (one CTA, one launch for all iterations); C4 never does
    # and are compiled into one shared-memory-resident kernel (NVRTC runs without a GPU): d = 30 exact needs 176 KB
    assert ex["vm_resident"] == 2 and dag["vm_resident"] == 2 and plans["c1"]["vm_resident"] == 2
    assert 150_000 < ex["tape_smem_bytes"] <= 224 * 1024 and dag["tape_smem_bytes"] < ex["tape_smem_bytes"]
    assert plans["c4_hoist1"]["vm_resident"] == 0


def test_c4_plan(plans):
    """C4: 9 dots x 8 steps minus the dead last-step output gate = 69 forward products, 69 weight gradients (each with
    its update in the epilogue), 11 live dX products (7 through Vo into the cell chain, 4 into h0)."""
    for key in ("c4_hoist0", "c4_hoist1"):
        p = plans[key]
        assert (p["gemms"], p["gemms_fused_sgd"]) == (149, 69), key
        kinds = [l.split()[1] for l in p["kernels"]]
        assert kinds.count("GEMM") == 149 and len(kinds) == p["kernels_per_iter"]
        dw = [l for l in p["kernels"] if "epi=1" in l]
        assert len(dw) == 69 and all("m=4096 n=4096 k=1024 ta=1 tb=0" in l for l in dw)
        assert p["gemm_flop_per_iter"] == 149 * 2 * 1024 * 4096 * 4096


def test_hoisting_moves_loss_independent_backward_work_forward(plans):
    """Q3 makes every gate derivative and weight gradient independent of the loss: with hoisting the first weight
    update follows the first gate of step 0 instead of the whole forward pass, the same kernels are launched, and
    activations die earlier (smaller arena, fewer elementwise bytes)."""
    a, b = plans["c4_hoist0"], plans["c4_hoist1"]
    first = lambda p: next(i for i, l in enumerate(p["kernels"]) if "epi=1" in l)
    assert first(a) > 69 and first(b) < 20
    assert sorted(l.split(" ", 1)[1] for l in a["kernels"] if "GEMM" in l) == \
        sorted(l.split(" ", 1)[1] for l in b["kernels"] if "GEMM" in l)
    assert b["arena_bytes"] < a["arena_bytes"] and b["ew_bytes_per_iter"] <= a["ew_bytes_per_iter"]
    # the loss-carrying chain stays at the end: the last kernels are the dX-accumulate products through Vo
    assert any("tb=1 epi=2" in l for l in b["kernels"][-8:])
    s0, s1 = plans["lstm_T3_hoist0"], plans["lstm_T3_hoist1"]
    assert s0["gemms"] == s1["gemms"] and s0["live_instrs"] == s1["live_instrs"]


def test_data_parallel_tape(plans):
    """SURVEY §8e as a tape: every weight gradient X^T.E is followed by one exchange — ncclAllReduce + an elementwise
    update, or (peer memory available) ONE fused all-reduce + SGD kernel; row-local parameters (biases, h0, C0) need
    none. With hoisting the exchanges are spread over the whole iteration instead of queueing behind the forward."""
    for hoist in (0, 1):
        nccl, fused = plans[f"c4_hoist{hoist}_dp2_nccl"], plans[f"c4_hoist{hoist}_dp2_fused"]
        assert (nccl["allreduces"], nccl["dp_fused_updates"], nccl["gemms_fused_sgd"]) == (69, 0, 0)
        assert (fused["allreduces"], fused["dp_fused_updates"], fused["gemms_fused_sgd"]) == (0, 69, 0)
        assert fused["kernels_per_iter"] == plans[f"c4_hoist{hoist}"]["kernels_per_iter"] + 69
        assert fused["ew_kernels"] == plans[f"c4_hoist{hoist}"]["ew_kernels"]  # no separate update pass
        assert nccl["ew_kernels"] > fused["ew_kernels"]
        ex = [i for i, l in enumerate(fused["kernels"]) if "DPUPDATE" in l]
        assert len(ex) == 69 and all("n=16777216" in fused["kernels"][i] for i in ex)
        # every exchange comes after the product that made its gradient (which no longer carries the update)
        dw = [i for i, l in enumerate(fused["kernels"]) if "ta=1 tb=0 epi=0" in l]
        assert len(dw) == 69 and all(d < e for d, e in zip(dw, ex))
        if hoist:
            assert ex[0] < 20 and ex[34] < 160  # half of the exchanges are issued in the first half of the iteration
        else:
            assert ex[0] > 69


def test_data_parallel_in_switch_tape_and_exact_mode_replicas(plans):
    """From three ranks up the exchange is planned through the multicast heap (NVLS): the same 69 fused exchanges, every
    one marked `multimem`, with the gradient slots in an arena of their own. Exact mode under data parallelism plans NO
    exchange at all — C1-C3 and the tree walk run as replicas, the sum would multiply every update."""
    mc, ipc = plans["c4_dp8_multimem"], plans["c4_hoist1_dp2_fused"]
    assert (mc["allreduces"], mc["dp_fused_updates"], mc["dp_multimem_updates"]) == (0, 69, 69)
    assert ipc["dp_multimem_updates"] == 0
    ex = [l for l in mc["kernels"] if "DPUPDATE" in l]
    assert len(ex) == 69 and all(l.endswith("multimem") for l in ex)
    assert mc["arena_bytes"] == ipc["arena_bytes"]  # same slots, split over two arenas
    rep = plans["c3_exact_dp2"]
    assert (rep["allreduces"], rep["dp_fused_updates"]) == (0, 0)
    assert rep["kernels_per_iter"] == plans["c3_d30_exact"]["kernels_per_iter"]


def test_random_graphs_plan_and_compile():
    """tests/plan_fuzz.py: random expression trees (every operator, scalar / matrix mixes, broadcasts, scalar <- matrix
    means, square products) plan in both modes with and without hoisting — the hoisting pass re-verifies every
    dependency itself — and the small ones come out as compiled resident kernels that NVRTC accepts for sm_100a."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "plan_fuzz.py"), "7", "30"], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["errors"] == [] and out["graphs"] == 30 and out["plans"] == 120
    assert out["compiled"] >= 90 and out["with_means"] > 0
