import sys, ctypes as C
sys.path.insert(0,'/root/repo')
import computational_graph_b200 as pkg
api = pkg.get_api(); api.set_tf32(True)
out = (C.c_double*3)()
for i in range(2):
    rc = api.lib.cg_debug_overlap(out)
    print(rc, "gemm x8 alone %.3f ms, ew x2 alone %.3f ms, together %.3f ms" % tuple(out))
