"""ctypes view of the reference's 26-symbol matrix ABI (include/cg_matrix.h == src/matrix.rs:4-18, src/ops.rs:8-33),
bound either to the product (libcg_b200.so, device pointers) or to the oracle's `orc_sym_*` hooks (host pointers)."""
import ctypes as C

import numpy as np


class Matrix(C.Structure):
    _fields_ = [("height", C.c_uint), ("width", C.c_uint), ("array", C.POINTER(C.c_float))]


_FP = C.POINTER(C.c_float)
_MP = C.POINTER(Matrix)

MAPS = ["neg", "sq", "abs", "signum", "sigmoid", "tanh", "one_minus"]
BINS = ["add", "sub", "mul"]


class ProductMatrixAbi:
    """The product library's legacy symbols. Every matrix lives in device memory."""

    def __init__(self, lib):
        self.lib = lib
        for n in ("alloc_matrix", "free_matrix"):
            getattr(lib, n).argtypes, getattr(lib, n).restype = [_MP], None
        lib.fill_matrix.argtypes, lib.fill_matrix.restype = [_MP, C.c_float], None
        lib.copy_matrix.argtypes, lib.copy_matrix.restype = [_MP, _MP], None
        lib.get_array.argtypes, lib.get_array.restype = [_MP, _FP], None
        lib.set_array.argtypes, lib.set_array.restype = [_MP, _FP, C.c_bool], None
        for n in MAPS:
            f = getattr(lib, "map_" + n)
            f.argtypes, f.restype = [_MP, _MP], None
        for n in BINS:
            f = getattr(lib, "elemwise_" + n)
            f.argtypes, f.restype = [_MP, _MP, _MP], None
            f = getattr(lib, "broadcast_" + n)
            f.argtypes, f.restype = [_MP, C.c_float, _MP], None
            f = getattr(lib, "broadcast_" + n + "_rev")
            f.argtypes, f.restype = [C.c_float, _MP, _MP], None
        lib.gemm.argtypes, lib.gemm.restype = [_MP, C.c_bool, _MP, C.c_bool, _MP], None
        lib.all_equal.argtypes, lib.all_equal.restype = [_MP, C.c_float], C.c_bool
        lib.all_less_than.argtypes, lib.all_less_than.restype = [_MP, C.c_float], C.c_bool
        lib.reduce_sum.argtypes, lib.reduce_sum.restype = [_MP], C.c_float

    def new(self, h, w, rowmajor=None):
        m = Matrix(h, w, None)
        self.lib.alloc_matrix(C.byref(m))
        if rowmajor is not None:
            a = np.ascontiguousarray(rowmajor, dtype=np.float32)
            self.lib.set_array(C.byref(m), a.ctypes.data_as(_FP), True)
        return m

    def free(self, m):
        self.lib.free_matrix(C.byref(m))

    def get(self, m):
        """[row, col] view of the column-major storage."""
        buf = np.zeros(m.height * m.width, dtype=np.float32)
        self.lib.get_array(C.byref(m), buf.ctypes.data_as(_FP))
        return buf.reshape(m.width, m.height).T.copy()


class OracleMatrixAbi:
    """The same operations on the CPU oracle's active backend (reference-compiled or port); host memory."""

    def __init__(self, orc):
        lib = self.lib = orc.lib
        lib.orc_sym_map.argtypes = [C.c_int, _MP, _MP]
        lib.orc_sym_elemwise.argtypes = [C.c_int, _MP, _MP, _MP]
        lib.orc_sym_broadcast.argtypes = [C.c_int, _MP, C.c_float, _MP]
        lib.orc_sym_broadcast_rev.argtypes = [C.c_int, C.c_float, _MP, _MP]
        lib.orc_sym_gemm.argtypes = [_MP, C.c_bool, _MP, C.c_bool, _MP]
        lib.orc_sym_set_array.argtypes = [_MP, _FP, C.c_bool]
        lib.orc_sym_get_array.argtypes = [_MP, _FP]
        lib.orc_sym_fill.argtypes = [_MP, C.c_float]
        lib.orc_sym_copy.argtypes = [_MP, _MP]
        lib.orc_sym_reduce_sum.argtypes, lib.orc_sym_reduce_sum.restype = [_MP], C.c_float
        lib.orc_sym_all_equal.argtypes, lib.orc_sym_all_equal.restype = [_MP, C.c_float], C.c_bool
        lib.orc_sym_all_less_than.argtypes, lib.orc_sym_all_less_than.restype = [_MP, C.c_float], C.c_bool
        self._keep = []

    def new(self, h, w, rowmajor=None):
        store = np.zeros(max(h * w, 1), dtype=np.float32)
        self._keep.append(store)
        m = Matrix(h, w, store.ctypes.data_as(_FP))
        if rowmajor is not None:
            a = np.ascontiguousarray(rowmajor, dtype=np.float32)
            self.lib.orc_sym_set_array(C.byref(m), a.ctypes.data_as(_FP), True)
        return m

    def get(self, m):
        buf = np.zeros(m.height * m.width, dtype=np.float32)
        self.lib.orc_sym_get_array(C.byref(m), buf.ctypes.data_as(_FP))
        return buf.reshape(m.width, m.height).T.copy()
