import os, sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo/tools")
os.environ["CG_DEBUG_TIMING"] = "1"
import numpy as np, bench
import computational_graph_b200 as pkg
api = pkg.get_api(); api.set_mode("dag"); api.set_tf32(True)
f, host_args = bench.build_lstm(api, bench.B_PER_GPU, bench.I, bench.H, bench.T, 4, with_inputs=True)
pinned = {}
for k, v in host_args.items():
    a = api.pinned_empty(v.shape); a[...] = v; pinned[k] = api.Constant(a)
ios = [int(x) for x in os.environ.get("IOS", "1,0").split(",")]
for io in ios:
    api.set_io_graph(bool(io))
    for _ in range(3):
        t0 = time.perf_counter(); f.minimize(pinned, 0.01, 1, trace=True); print("py wall", (time.perf_counter() - t0) * 1e3, flush=True)
