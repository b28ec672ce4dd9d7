"""oracle/oracle.py — TEST INFRASTRUCTURE. ctypes binding of oracle/liboracle.so (the CPU restatement
of the reference's hot path, oracle/cg_oracle.c) through the same `Api` wrapper the product uses, with
the `orc_` symbol prefix. Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import
this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))

from computational_graph_b200.api import Api  # noqa: E402

LIB_PATH = os.path.join(_HERE, "liboracle.so")
REF_O0 = os.path.join(_HERE, "_ref", "libmatrix_ref_O0.so")
REF_O3 = os.path.join(_HERE, "_ref", "libmatrix_ref_O3.so")


def build() -> None:
    """Compile the restatement and, when /root/reference is present, the reference's CPU backend."""
    subprocess.run(["make", "-C", _HERE, "--no-print-directory"], check=True, stdout=subprocess.DEVNULL)


class OracleApi(Api):
    def __init__(self):
        if not os.path.exists(LIB_PATH):
            build()
        lib = ctypes.CDLL(LIB_PATH)
        super().__init__(lib, "orc_")
        lib.orc_use_ref.restype, lib.orc_use_ref.argtypes = ctypes.c_int, [ctypes.c_char_p]
        lib.orc_use_port.restype = None
        lib.orc_using_ref.restype = ctypes.c_int
        lib.orc_count_gemm.restype = ctypes.c_long
        lib.orc_tree_size.restype = ctypes.c_long
        lib.orc_tree_size.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_long)]
        lib.orc_set_fix_act_grad.argtypes = [ctypes.c_int]
        lib.orc_set_fix_broadcast_sum.argtypes = [ctypes.c_int]
        lib.orc_set_fix_sub_sign.argtypes = [ctypes.c_int]
        lib.orc_set_fix_tied_params.argtypes = [ctypes.c_int]

    def use_ref(self, which: str = "O0") -> bool:
        """Route every matrix op through the reference-compiled CPU backend. Returns False when
        oracle/_ref has not been built (no /root/reference and no prebuilt copy)."""
        path = {"O0": REF_O0, "O3": REF_O3}[which]
        if not os.path.exists(path):
            return False
        return self.lib.orc_use_ref(path.encode()) == 0

    def use_port(self) -> None:
        self.lib.orc_use_port()

    def gemm_count(self) -> int:
        return int(self.lib.orc_count_gemm())

    def reset_counts(self) -> None:
        self.lib.orc_reset_counts()

    def tree_size(self, f):
        dots = ctypes.c_long()
        n = self.lib.orc_tree_size(f._h, ctypes.byref(dots))
        return int(n), int(dots.value)


_API = None


def get_oracle() -> OracleApi:
    global _API
    if _API is None:
        _API = OracleApi()
    return _API
