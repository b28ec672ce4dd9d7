"""Bring-up probe for the tcgen05 TF32 product: one-hot operands reveal which (row, k) the tensor core reads
from each shared-memory position. Usage: python tools/tf32_probe.py  (env CG_TF32_* override descriptor fields)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import computational_graph_b200 as pkg
from matrix_abi import ProductMatrixAbi

api = pkg.get_api()
api.set_tf32(True)
p = ProductMatrixAbi(api.lib)
m, n, k = 256, 256, 128


def run(a, t1, b, t2):
    ap, bp, rp = p.new(*a.shape, a), p.new(*b.shape, b), p.new(m, n)
    p.lib.fill_matrix(C.byref(rp), np.float32(-7777.0))
    p.lib.gemm(C.byref(ap), t1, C.byref(bp), t2, C.byref(rp))
    out = p.get(rp)
    for x in (ap, bp, rp):
        p.free(x)
    return out


print("env:", {k_: v for k_, v in os.environ.items() if k_.startswith("CG_TF32")})
# A MN-major (t1 = False), B K-major (t2 = False, known good): B[kk, j] = kk + 1
b = np.tile(np.arange(1, k + 1, dtype=np.float32)[:, None], (1, n))
for (i0, k0) in [(0, 0), (1, 0), (4, 0), (31, 0), (32, 0), (64, 0), (127, 0), (128, 0), (0, 1), (0, 7), (0, 8), (0, 31), (0, 32), (5, 9), (37, 45)]:
    a = np.zeros((m, k), np.float32)
    a[i0, k0] = 1.0
    out = run(a, False, b, False)
    rows = np.nonzero(np.any(out != 0, axis=1))[0]
    desc = [(int(r), sorted(set(np.round(out[r]).astype(int).tolist()))[:4]) for r in rows[:6]]
    print(f"A onehot (i={i0},k={k0}) expect row {i0} value {k0 + 1}: nonzero rows {len(rows)} -> {desc}")
# B MN-major (t2 = True), A K-major (t1 = True): A^T stored (k, m): op(A)[i, kk] = kk + 1
a = np.tile(np.arange(1, k + 1, dtype=np.float32)[:, None], (1, m))  # shape (k, m), transposed use
for (k0, j0) in [(0, 0), (0, 1), (0, 31), (0, 32), (0, 255), (1, 0), (8, 0), (31, 0), (32, 0), (9, 5), (45, 37)]:
    bt = np.zeros((n, k), np.float32)  # stored (n, k); op(B) = B^T
    bt[j0, k0] = 1.0
    out = run(a, True, bt, True)
    cols = np.nonzero(np.any(out != 0, axis=0))[0]
    desc = [(int(c), sorted(set(np.round(out[:, c]).astype(int).tolist()))[:4]) for c in cols[:6]]
    print(f"B onehot (k={k0},j={j0}) expect col {j0} value {k0 + 1}: nonzero cols {len(cols)} -> {desc}")
