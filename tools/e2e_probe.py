"""Breaks the end-to-end `minimize(&args, lr, 1)` call of bench.py into its parts (wall clock, per call)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import computational_graph_b200 as pkg  # noqa: E402

api = pkg.get_api()
api.set_mode("dag")
api.set_tf32(True)
f, host_args = bench.build_lstm(api, bench.B_PER_GPU, bench.I, bench.H, bench.T, 4, with_inputs=True)
pinned = {}
for k, v in host_args.items():
    a = api.pinned_empty(v.shape)
    a[...] = v
    pinned[k] = api.Constant(a)
LR = bench.LR
f.minimize(pinned, LR, 1)


def timeit(name, fn, n=6):
    fn()
    fn()
    api.synchronize()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        api.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    print(f"{name:55s} min {min(ts):7.2f} ms  median {sorted(ts)[len(ts) // 2]:7.2f} ms  all {[round(t, 2) for t in ts]}", flush=True)


print("device-resident, back to back:", api.time_iterations(f, LR, 3, 10), "ms/iter", flush=True)
timeit("minimize_async(1) + synchronize", lambda: api.minimize_async(f, LR, 1))
timeit("minimize(pinned args, 1)            [io graph]", lambda: f.minimize(pinned, LR, 1))
timeit("minimize(pinned args, 1, trace)     [io graph]", lambda: f.minimize(pinned, LR, 1, trace=True))
api.set_io_graph(False)
timeit("minimize(pinned args, 1)            [upload first]", lambda: f.minimize(pinned, LR, 1))
timeit("minimize(pinned args, 1, trace)     [upload first]", lambda: f.minimize(pinned, LR, 1, trace=True))
api.set_io_graph(True)
pageable = {k: api.Constant(np.array(v)) for k, v in host_args.items()}
timeit("minimize(pageable args, 1, trace)", lambda: f.minimize(pageable, LR, 1, trace=True))
