"""Loader for the product's C-ABI library. There is no CPU fallback: a missing library is an error."""
from __future__ import annotations

import ctypes
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB: Optional[ctypes.CDLL] = None
_API = None


def library_path() -> str:
    return os.path.join(_HERE, "csrc", "libcg_b200.so")


def load_library() -> ctypes.CDLL:
    """dlopen csrc/libcg_b200.so (built by `make -C computational-graph_b200/csrc` or
    `__graft_entry__.build()`). Raises — never falls back — when it is absent."""
    global _LIB
    if _LIB is None:
        path = library_path()
        if not os.path.exists(path):
            raise RuntimeError(
                f"{path} is missing: build the sm_100a CUDA library first "
                "(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback.")
        _LIB = ctypes.CDLL(path, mode=ctypes.RTLD_GLOBAL)
    return _LIB


def get_api():
    """The `Api` bound to the product library's `cg_*` symbols."""
    global _API
    if _API is None:
        from .api import Api
        _API = Api(load_library(), "cg_")
    return _API
