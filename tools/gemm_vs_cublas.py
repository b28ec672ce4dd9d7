"""The three products of the hot path (forward X.W, dX = E.W^T, dW = X^T.E; src/optimize.rs:135,197-198) through the
26-symbol ABI's `gemm` (tcgen05 TF32 kernel) against cuBLAS TF32 (torch.matmul, allow_tf32) on the same shapes, BOTH
sustained: each case runs back to back for `--seconds` and the second half is timed with CUDA events while nvidia-smi
samples the SM clock — a dense GEMM on a 1 kW part is power-capped, so a best-of-N burst and a sustained rate differ.

    python tools/gemm_vs_cublas.py [--rows 1024 8192] [--seconds 1.5] [--pair 0|1]
One JSON object per line.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import computational_graph_b200 as pkg  # noqa: E402
from matrix_abi import ProductMatrixAbi  # noqa: E402


class Clock:
    def __enter__(self):
        self.v = []
        self.p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "50", "-i", "0"],
                                  stdout=subprocess.PIPE, text=True)
        threading.Thread(target=lambda: [self.v.append(l) for l in self.p.stdout], daemon=True).start()
        return self

    def __exit__(self, *a):
        self.p.terminate()

    def median(self):
        c = [float(l.split(",")[0]) for l in self.v if "," in l]
        w = [float(l.split(",")[1]) for l in self.v if "," in l]
        return (float(np.median(c[len(c) // 2:])) if c else None, float(np.median(w[len(w) // 2:])) if w else None)


def sustained(fn, sync, seconds):
    fn(); sync()
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < seconds / 2:   # heat up
        for _ in range(20):
            fn()
        sync()
        n += 20
    reps = max(20, n)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with Clock() as clk:
        e0.record(torch.cuda.current_stream())
        sync_needed = False
        t1 = time.perf_counter()
        for _ in range(reps):
            fn()
        sync()
        dt = (time.perf_counter() - t1) / reps
        mhz, watts = clk.median()
    return dt, mhz, watts


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, nargs="+", default=[1024, 8192])
    ap.add_argument("--h", type=int, default=4096)
    ap.add_argument("--seconds", type=float, default=1.5)
    ap.add_argument("--pair", type=int, default=0)
    args = ap.parse_args()
    api = pkg.get_api()
    api.set_tf32(True)
    api.set_tf32_pair(bool(args.pair))
    p = ProductMatrixAbi(api.lib)
    torch.backends.cuda.matmul.allow_tf32 = True
    H = args.h
    rng = np.random.default_rng(0)
    for B in args.rows:
        X = p.new(B, H, rng.uniform(-1, 1, (B, H)).astype(np.float32))
        W = p.new(H, H, rng.uniform(-1, 1, (H, H)).astype(np.float32))
        E = p.new(B, H, rng.uniform(-1, 1, (B, H)).astype(np.float32))
        Y, dX, dW = p.new(B, H), p.new(B, H), p.new(H, H)
        tx, tw, te = (torch.randn(B, H, device="cuda"), torch.randn(H, H, device="cuda"), torch.randn(B, H, device="cuda"))
        flop = 2.0 * B * H * H
        cases = [("NN forward X.W", (X, False, W, False, Y), lambda: torch.matmul(tx, tw)),
                 ("NT dX = E.W^T", (E, False, W, True, dX), lambda: torch.matmul(te, tw.t())),
                 ("TN dW = X^T.E", (X, True, E, False, dW), lambda: torch.matmul(tx.t(), te))]
        for name, (a, ta, b, tb, r), cub in cases:
            ours = lambda: p.lib.gemm(C.byref(a), ta, C.byref(b), tb, C.byref(r))
            dt_o, mhz_o, w_o = sustained(ours, api.synchronize, args.seconds)
            dt_c, mhz_c, w_c = sustained(cub, torch.cuda.synchronize, args.seconds)
            print(json.dumps({"rows": B, "case": name, "pair_kernel": bool(args.pair),
                              "ours_us": dt_o * 1e6, "ours_tflops": flop / dt_o / 1e12, "ours_sm_mhz": mhz_o, "ours_watts": w_o,
                              "cublas_us": dt_c * 1e6, "cublas_tflops": flop / dt_c / 1e12, "cublas_sm_mhz": mhz_c, "cublas_watts": w_c,
                              "ours_over_cublas": dt_c / dt_o}), flush=True)
        for m in (X, W, E, Y, dX, dW):
            p.free(m)
        del tx, tw, te
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
