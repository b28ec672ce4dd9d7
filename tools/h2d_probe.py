"""How fast do the inputs of one `minimize(&args, lr, 1)` call reach the GPU? Times the H2D copies of config 5's 33 inputs
(134 MB each at N = 1) from page-locked host memory, alone and while a product runs, with torch's pinned allocator and with
cg_host_alloc. One JSON line."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import computational_graph_b200 as pkg  # noqa: E402


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    n_in = int(sys.argv[2]) if len(sys.argv) > 2 else 33
    api = pkg.get_api()
    shape = (rows, 4096)
    nbytes = rows * 4096 * 4
    dev = [torch.empty(shape, device="cuda") for _ in range(n_in)]
    out = {"rows": rows, "inputs": n_in, "bytes_each": nbytes}
    for kind in ("torch_pinned", "cg_host_alloc"):
        if kind == "torch_pinned":
            host = [torch.empty(shape, dtype=torch.float32, pin_memory=True) for _ in range(n_in)]
            for h in host:
                h.numpy()[...] = 1.0
        else:
            arrs = [api.pinned_empty(shape) for _ in range(n_in)]
            for a in arrs:
                a[...] = 1.0
            host = [torch.from_numpy(a) for a in arrs]
        s = torch.cuda.Stream()
        for busy in (False, True):
            a = torch.randn(8192, 4096, device="cuda")
            w = torch.randn(4096, 4096, device="cuda")
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if busy:
                for _ in range(200):
                    torch.matmul(a, w)
            with torch.cuda.stream(s):
                e0.record()
                for d, h in zip(dev, host):
                    d.copy_(h, non_blocking=True)
                e1.record()
            e1.synchronize()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            out[f"{kind}_{'beside_products' if busy else 'alone'}_gbs"] = n_in * nbytes / (ms * 1e-3) / 1e9
        del host
    print(json.dumps(out))


if __name__ == "__main__":
    main()
