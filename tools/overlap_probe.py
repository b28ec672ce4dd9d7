"""Does an elementwise chain overlap a tcgen05 product when the iteration graph has them on parallel branches?"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import computational_graph_b200 as pkg
api = pkg.get_api()
api.set_mode("dag"); api.set_tf32(True)
F, C = api.Function, api.Constant
rng = np.random.default_rng(0)
B, H = 1024, 4096
def mk(kind):
    x = F.constant(C(rng.uniform(-.1, .1, (B, H)).astype(np.float32)))
    terms = []
    gem = []
    if kind in ("gemm", "both"):
        for i in range(4):
            w = F.param(f"w{i}", C((rng.uniform(-1, 1, (H, H)) / 64).astype(np.float32)))
            gem.append(api.dot(x, w))
    if kind in ("ew", "both"):
        for i in range(4):
            p = F.param(f"p{i}", C(rng.uniform(-1, 1, (B, H)).astype(np.float32)))
            q = F.param(f"q{i}", C(rng.uniform(-1, 1, (B, H)).astype(np.float32)))
            terms.append((p.sigmoid() * q.tanh()).tanh())
    terms = terms + gem   # elementwise chains first in tape order: their fused kernel does not depend on any product
    f = terms[0]
    for t in terms[1:]:
        f = f + t
    return f
for streams in (0, 4):
    api.set_graph_streams(streams)
    for kind in ("gemm", "ew", "both"):
        f = mk(kind)
        st = api.plan_stats(f, 0.01)
        ms = api.time_iterations(f, 0.01, 3, 20)
        print(f"streams={streams} {kind:5s}: {ms*1e3:8.1f} us/iter  kernels={st['kernels_per_iter']} gemms={st['gemms']} ew={st['ew_kernels']}")
