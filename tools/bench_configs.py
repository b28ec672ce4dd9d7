"""Device-resident timings of the BASELINE configs that are not the bench.py line (C1, C2, C3): parity-test cases,
reported for the roofline table in DESIGN.md. One JSON object per line.

    python tools/bench_configs.py [--c2-n 8192] [--reps 20]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import computational_graph_b200 as pkg  # noqa: E402
from helpers import c2_graph, readme_lstm  # noqa: E402


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path))["hbm_gbs"], "measured"
    return 6650.0, "fallback"


def c2(api, n, reps, peak, src):
    f = c2_graph(api, n)
    f.minimize({}, 0.01, 2)
    ms = api.time_iterations(f, 0.01, 3, reps)
    st = api.plan_stats(f, 0.01)
    alg = 24.0 * n * n
    planned = st["ew_bytes_per_iter"]
    if st.get("quiet_variant"):
        planned = (st["quiet_ew_bytes_per_iter"] * (reps - 1) + st["ew_bytes_per_iter"]) / reps
    print(json.dumps({"config": f"C2 tanh(sigmoid(W)-sq(V))+abs(W), {n}x{n}", "ms_per_step": ms, "steps_per_s": 1e3 / ms,
                      "algorithmic_bytes": alg, "achieved_gbs": alg / (ms * 1e-3) / 1e9,
                      "frac_of_hbm": alg / (ms * 1e-3) / 1e9 / peak, "peak_gbs": peak, "peak_source": src,
                      # bytes the kernels of a timed step actually move: the last of `reps` steps also stores the root
                      # value (28 n), the others do not when the quiet variant is in use (24 n)
                      "planned_bytes": planned,
                      "planned_gbs": planned / (ms * 1e-3) / 1e9, "plan": st}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--c2-n", type=int, default=8192)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--only", default="", help="c1 | c2 | c3 (default: all)")
    args = ap.parse_args()
    want = lambda c: not args.only or args.only == c
    api = pkg.get_api()
    peak, src = hbm_peak()

    # C1: README scalar example
    api.set_mode("exact")
    F = api.Function
    if want("c1"):
        x = F.scalar_param("x", 1.0)
        f = (x + F.scalar(3.0)).sq()
        f.minimize({}, 0.01, 10)
        ms = api.time_iterations(f, 0.01, 20, 1000)
        print(json.dumps({"config": "C1 (x+3)^2 scalar", "us_per_iter": ms * 1e3, "iters_per_s": 1e3 / ms,
                          "plan": api.plan_stats(f, 0.01)}))

    # C2: elementwise-only matrix graph, HBM-bound: 24 n algorithmic bytes per step
    n = args.c2_n
    if want("c2"):
        c2(api, n, args.reps, peak, src)
    if not want("c3"):
        return

    # C3: README LSTM, T = 2, d sweep, exact mode (the reference's tree walk) and dag mode
    for mode in ("exact", "dag"):
        api.set_mode(mode)
        for d in (5, 10, 15, 20, 25, 30):
            f = readme_lstm(api, d)
            f.minimize({}, 0.01, 5)
            ms = api.time_iterations(f, 0.01, 20, 500)
            st = api.plan_stats(f, 0.01)
            print(json.dumps({"config": f"C3 README LSTM T=2 d={d} ({13 * d * d} params), {mode} mode",
                              "us_per_iter": ms * 1e3, "iters_per_s": 1e3 / ms, "kernels_per_iter": st["kernels_per_iter"],
                              "gemms": st["gemms"], "ew_kernels": st["ew_kernels"]}))
    api.set_mode("exact")


if __name__ == "__main__":
    main()
