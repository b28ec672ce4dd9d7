"""Random expression graphs through the tape compiler in PLAN-ONLY mode (no GPU, nothing computed): every graph must
plan in both execution modes, pass the hoisting pass's own dependency verifier, and — being small — come out as a
compiled resident kernel whose generated CUDA source NVRTC accepts for sm_100a. Prints one JSON summary line.
Run in a process of its own (tests/test_plan_cpu.py drives it):  python tests/plan_fuzz.py <seed> <count>"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def random_graph(api, rng, depth):
    """A random tree over scalar / matrix params and constants with every operator of src/ops.rs."""
    F = api.Function
    h, w = int(rng.integers(1, 7)), int(rng.integers(1, 7))
    counter = [0]

    def leaf(want_matrix):
        counter[0] += 1
        kind = rng.integers(0, 3)
        if want_matrix:
            v = rng.uniform(-1, 1, (h, w)).astype(np.float32)
            return F.param(f"m{counter[0]}", api.Constant(v)) if kind else F.constant(api.Constant(v))
        x = float(rng.uniform(-1, 1)) or 0.25
        return F.scalar_param(f"s{counter[0]}", x) if kind else F.scalar(x)

    def node(d, want_matrix):
        if d == 0 or rng.random() < 0.2:
            return leaf(want_matrix)
        op = rng.integers(0, 10)
        if op < 5:
            f = node(d - 1, want_matrix)
            return [lambda: -f, f.sq, f.abs, f.sigmoid, f.tanh][op]()
        if op < 9:
            a = node(d - 1, want_matrix)
            b = node(d - 1, want_matrix and rng.random() < 0.7)  # matrix (+) scalar broadcasts; scalar <- matrix means
            if rng.random() < 0.5:
                a, b = b, a
            return [lambda: a + b, lambda: a - b, lambda: a * b, lambda: a * b][op - 5]()
        if want_matrix and h == w:
            return api.dot(node(d - 1, True), node(d - 1, True))
        return node(d - 1, want_matrix)

    return node(depth, bool(rng.random() < 0.8))


def main():
    seed, count = int(sys.argv[1]), int(sys.argv[2])
    import computational_graph_b200 as pkg
    lib = pkg.load_library()
    assert lib.cg_set_plan_only(1) == 0
    api = pkg.get_api()
    rng = np.random.default_rng(seed)
    out = {"graphs": 0, "plans": 0, "compiled": 0, "kernels": 0, "with_products": 0, "with_means": 0, "errors": []}
    g = 0
    while g < count:
        try:
            f = random_graph(api, rng, int(rng.integers(1, 6)))
        except Exception as e:  # noqa: BLE001
            # constant (+) constant folds by COMPUTING at construction (src/ops.rs:209-213): not in plan-only mode
            if "plan-only" in str(e):
                continue
            raise
        g += 1
        out["graphs"] += 1
        for mode in ("exact", "dag"):
            api.set_mode(mode)
            for hoist in (False, True):
                api.set_hoist(hoist)
                try:
                    st = api.plan_stats(f, 0.01)
                except Exception as e:  # noqa: BLE001
                    out["errors"].append(f"graph {g} mode {mode} hoist {hoist}: {e}")
                    continue
                out["plans"] += 1
                out["kernels"] += st["kernels_per_iter"]
                out["compiled"] += 1 if st["vm_resident"] == 2 else 0
                out["with_products"] += 1 if st["gemms"] else 0
                out["with_means"] += 1 if st["reduces"] else 0
    print(json.dumps(out))


if __name__ == "__main__":
    main()
