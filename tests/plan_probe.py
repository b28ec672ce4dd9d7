"""Builds graphs in PLAN-ONLY mode (no GPU: placeholder addresses, nothing computed — cg_set_plan_only) and prints
what the tape compiler made of them as one JSON object. Run in a process of its own (the mode must be chosen before
any other call into the library); tests/test_plan_cpu.py drives it."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


class CheapRng:
    """Stands in for numpy's Generator when only shapes matter (a 4096 x 4096 draw per weight would dominate)."""

    def uniform(self, lo, hi, size=None):
        a = np.empty(size, dtype=np.float32)
        a.fill(0.25)
        a.reshape(-1)[::7] = -0.125
        return a


def main():
    import computational_graph_b200 as pkg
    from computational_graph_b200.lstm import lstm_loss, lstm_params
    from helpers import c2_graph, readme_lstm
    lib = pkg.load_library()
    assert lib.cg_set_plan_only(1) == 0
    api = pkg.get_api()
    out = {}

    def record(name, f, lr, describe=False):
        st = api.plan_stats(f, lr)
        if describe:
            st["kernels"] = api.plan_describe(f, lr)
        out[name] = st

    x = api.Function.scalar_param("x", 1.0)
    record("c1", (x + api.Function.scalar(3.0)).sq(), 0.01, describe=True)
    record("c2", c2_graph(api, 64), 0.01, describe=True)
    for mode in ("exact", "dag"):
        api.set_mode(mode)
        record(f"c3_d30_{mode}", readme_lstm(api, 30), 0.01)
    api.set_mode("dag")
    for hoist in (0, 1):
        api.set_hoist(bool(hoist))
        record(f"lstm_T3_hoist{hoist}", readme_lstm(api, 24, T=3, seed=1, batch=40, d_in=16), 0.01, describe=True)
    if "--c4" in sys.argv:
        H = I = 4096
        B, T = 1024, 8
        rng = CheapRng()
        for hoist in (0, 1):
            api.set_hoist(bool(hoist))
            api.set_tf32(True)
            params = lstm_params(I, H, B, rng, scale=1.0)
            xs = [api.Constant(rng.uniform(-.1, .1, (B, I))) for _ in range(T)]
            tgt = api.Constant(rng.uniform(-.1, .1, (B, H)))
            f = lstm_loss(api, xs, tgt, params)
            record(f"c4_hoist{hoist}", f, 0.01, describe=True)
            # the same graph as one of 2 data-parallel ranks: NCCL exchange, then the fused peer-memory exchange
            for fused in (0, 1):
                assert lib.cg_dp_plan_only(2, fused) == 0
                record(f"c4_hoist{hoist}_dp2_{'fused' if fused else 'nccl'}", f, 0.01, describe=True)
            if hoist:
                # 8 ranks with the in-switch exchange (fused = 2): same tape, every exchange through the multicast heap
                assert lib.cg_dp_plan_only(8, 2) == 0
                record("c4_dp8_multimem", f, 0.01, describe=True)
                # exact mode runs as replicas: no exchange is ever planned (its in-line SGD is interleaved with the walk)
                api.set_mode("exact")
                assert lib.cg_dp_plan_only(2, 1) == 0
                record("c3_exact_dp2", readme_lstm(api, 30), 0.01)
                api.set_mode("dag")
            assert lib.cg_dp_plan_only(1, 0) == 0
    try:
        x.minimize({}, 0.01, 1)
        out["minimize_raises"] = False
    except Exception as e:  # noqa: BLE001
        out["minimize_raises"] = "plan-only" in str(e)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
