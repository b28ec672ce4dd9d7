"""time(K) = a + b K/32 for the forward product (m x K) . (K x n): separates launch/prologue/epilogue from the
per-k-block cost of the main loop.   python tools/gemm_sweep.py [--m 1024] [--n 4096]"""
import argparse, ctypes as C, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import computational_graph_b200 as pkg
from matrix_abi import ProductMatrixAbi

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=1024)
ap.add_argument("--n", type=int, default=4096)
ap.add_argument("--reps", type=int, default=40)
args = ap.parse_args()
api = pkg.get_api(); api.set_tf32(True)
p = ProductMatrixAbi(api.lib)
rows = []
for k in (256, 512, 1024, 2048, 4096, 8192):
    X, W, Y = p.new(args.m, k), p.new(k, args.n), p.new(args.m, args.n)
    p.lib.fill_matrix(C.byref(X), C.c_float(0.5)); p.lib.fill_matrix(C.byref(W), C.c_float(0.25))
    for _ in range(3): p.lib.gemm(C.byref(X), False, C.byref(W), False, C.byref(Y))
    api.synchronize(); t0 = time.perf_counter()
    for _ in range(args.reps): p.lib.gemm(C.byref(X), False, C.byref(W), False, C.byref(Y))
    api.synchronize(); us = (time.perf_counter() - t0) / args.reps * 1e6
    rows.append((k, us))
    for x in (X, W, Y): p.free(x)
ks = np.array([r[0] / 32 for r in rows]); ts = np.array([r[1] for r in rows])
b, a = np.polyfit(ks[2:], ts[2:], 1)
print(json.dumps({"m": args.m, "n": args.n, "us_by_k": rows, "fixed_us": a, "us_per_kblock": b,
                  "ideal_us_per_kblock_at_1.9GHz": 512 / 1900.0 * (1 if args.n >= 256 else 1)}))
