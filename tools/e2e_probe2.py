"""Where does the end-to-end call lose time against the device-resident iteration? Config-5-like LSTM (B = 8192 rows, T given):
device-resident ms, then `minimize(&args, lr, 1)` with page-locked host buffers in four variants. One JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import computational_graph_b200 as pkg  # noqa: E402
import bench  # noqa: E402


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    rows = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    api = pkg.get_api()
    api.set_mode("dag")
    api.set_tf32(True)
    out = {"T": T, "rows": rows}
    f, host_args = bench.build_lstm(api, rows, 4096, 4096, T, 5, with_inputs=True)
    pinned, keep = {}, []
    for k, v in host_args.items():
        t = torch.empty(v.shape, dtype=torch.float32, pin_memory=True)
        t.numpy()[...] = v
        keep.append(t)
        pinned[k] = api.Constant(t.numpy())
    out["h2d_bytes"] = sum(int(np.asarray(v.value).nbytes) for v in pinned.values())
    f.minimize(pinned, 0.01, 1)
    api.minimize_async(f, 0.01, 3)
    api.synchronize()
    out["resident_ms"] = api.time_iterations(f, 0.01, 0, 5)
    # the same iteration launched one at a time, waiting for each (no transfers): what a launch costs when nothing hides it
    t0 = time.perf_counter()
    for _ in range(5):
        api.minimize_async(f, 0.01, 1)
        api.synchronize()
    out["resident_one_at_a_time_ms"] = (time.perf_counter() - t0) * 1e3 / 5

    def e2e(trace, reps=5):
        for _ in range(2):
            f.minimize(pinned, 0.01, 1, trace=trace)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            f.minimize(pinned, 0.01, 1, trace=trace)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) * 1e3 / reps
    out["e2e_trace_ms"] = e2e(True)
    out["e2e_notrace_ms"] = e2e(False)
    # the same call with ONE input fed per call and the others kept on the device would need API changes; instead: how
    # long do the transfers alone take when nothing overlaps them (io graph off => upload, then iterate)?
    api.set_io_graph(False)
    out["e2e_serial_upload_trace_ms"] = e2e(True)
    api.set_io_graph(True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
