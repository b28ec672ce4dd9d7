"""Times the three products of the hot path (forward X.W, dX = E.W^T, dW = X^T.E; optimize.rs:135,197-198) through the
26-symbol ABI's `gemm` at the BASELINE config-4 shapes. Also the small command `ncu --set full` is pointed at.

    python tools/gemm_bench.py [--reps 50] [--product tf32|tf32x3|f32] [--b 1024] [--h 4096]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import computational_graph_b200 as pkg  # noqa: E402
from matrix_abi import ProductMatrixAbi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=50)
    ap.add_argument("--product", default="tf32")
    ap.add_argument("--b", type=int, default=1024)
    ap.add_argument("--h", type=int, default=4096)
    args = ap.parse_args()
    api = pkg.get_api()
    api.set_tf32({"f32": 0, "tf32": 1, "tf32x3": 2}[args.product])
    p = ProductMatrixAbi(api.lib)
    B, H = args.b, args.h
    rng = np.random.default_rng(0)
    X = p.new(B, H, rng.uniform(-1, 1, (B, H)).astype(np.float32))
    W = p.new(H, H, rng.uniform(-1, 1, (H, H)).astype(np.float32))
    E = p.new(B, H, rng.uniform(-1, 1, (B, H)).astype(np.float32))
    Y, dX, dW = p.new(B, H), p.new(B, H), p.new(H, H)
    cases = [("forward X.W", X, False, W, False, Y), ("dX = E.W^T", E, False, W, True, dX), ("dW = X^T.E", X, True, E, False, dW)]
    flop = 2.0 * B * H * H
    out = {}
    for name, a, ta, b, tb, r in cases:
        for _ in range(3):
            p.lib.gemm(C.byref(a), ta, C.byref(b), tb, C.byref(r))
        api.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.reps):
            p.lib.gemm(C.byref(a), ta, C.byref(b), tb, C.byref(r))
        api.synchronize()
        dt = (time.perf_counter() - t0) / args.reps
        out[name] = {"us": dt * 1e6, "tflops": flop / dt / 1e12}
    print(json.dumps({"product": args.product, "B": B, "H": H, "gflop_each": flop / 1e9, "cases": out}))


if __name__ == "__main__":
    main()
