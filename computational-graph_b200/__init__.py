"""computational-graph_b200 — B200-native (sm_100a) execution backend for the hot path of
ethanabrooks/computational-graph: `Function` tree forward -> generalized backprop -> in-place SGD.

Layout:
  csrc/        hand-written CUDA kernels + the C-ABI shared library (libcg_b200.so)
  api.py       host-side mirror of the reference's Function/Constant surface (ctypes over `cg_*`)
  lstm.py      src/lstm.rs transliterated over that surface
  _lib.py      loader for libcg_b200.so — raises if the CUDA library has not been built

The directory name carries a hyphen (it is the reference's name); import it as
`computational_graph_b200` (the sibling shim package re-exports this one).
"""
from .api import Api, Constant, Function, GraphError  # noqa: F401
from ._lib import load_library, library_path, get_api  # noqa: F401
from . import lstm  # noqa: F401
from . import dp  # noqa: F401
