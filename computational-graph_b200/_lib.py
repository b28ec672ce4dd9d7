"""Loader for the product's C-ABI library. There is no CPU fallback: a missing library is an error."""
from __future__ import annotations

import ctypes as C
import os as _os

# more hardware work queues, so that the H2D copies of an end-to-end call do not queue behind kernels of the iteration graph
# (csrc/graph.cpp says why); must be set before the CUDA context exists
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import os
import weakref
from typing import Dict, List, Optional

import numpy as np

from .api import Api

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB: Optional[C.CDLL] = None
_API = None


def library_path() -> str:
    return os.path.join(_HERE, "csrc", "libcg_b200.so")


def load_library() -> C.CDLL:
    """dlopen csrc/libcg_b200.so (built by `make -C computational-graph_b200/csrc` or
    `__graft_entry__.build()`). Raises — never falls back — when it is absent."""
    global _LIB
    if _LIB is None:
        path = library_path()
        if not os.path.exists(path):
            raise RuntimeError(
                f"{path} is missing: build the sm_100a CUDA library first "
                "(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback.")
        _LIB = C.CDLL(path)
    return _LIB


class PlanStats(C.Structure):
    """`cg_plan_stats` (include/cg_graph.h)."""
    _fields_ = [(n, C.c_ulonglong) for n in ("tape_instrs", "live_instrs", "kernels_per_iter", "ew_kernels", "gemms",
                                             "gemms_fused_sgd", "reduces", "arena_bytes")] + \
               [("gemm_flop_per_iter", C.c_double), ("ew_bytes_per_iter", C.c_double), ("ew_specialised", C.c_ulonglong),
                ("allreduces", C.c_ulonglong), ("dp_fused_updates", C.c_ulonglong), ("vm_resident", C.c_ulonglong), ("tape_smem_bytes", C.c_ulonglong),
                ("quiet_variant", C.c_ulonglong), ("quiet_ew_bytes_per_iter", C.c_double),
                ("dp_multimem_updates", C.c_ulonglong)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class ProductApi(Api):
    """`Api` over the `cg_*` symbols plus the product-only switches and introspection."""

    def __init__(self, lib: C.CDLL):
        super().__init__(lib, "cg_")
        for name in ("set_tf32", "set_fix_act_grad", "set_fix_broadcast_sum", "set_fix_sub_sign", "set_fix_tied_params", "set_fuse", "set_cuda_graph", "set_strict_gemm_max",
                     "set_graph_streams", "set_io_graph", "set_tf32_pair", "set_hoist", "set_vm_max_dim", "set_tape_jit", "set_elide_root"):
            fn = self._sym(name)
            fn.restype, fn.argtypes = None, [C.c_int]
        lib.cg_launch_count.restype = C.c_ulonglong
        lib.cg_plan_stats_get.restype, lib.cg_plan_stats_get.argtypes = C.c_int, [C.c_void_p, C.c_float, C.POINTER(PlanStats)]
        lib.cg_time_iterations.restype = C.c_double
        lib.cg_time_iterations.argtypes = [C.c_void_p, C.c_float, C.c_int, C.c_int]
        lib.cg_minimize_async.restype, lib.cg_minimize_async.argtypes = C.c_int, [C.c_void_p, C.c_float, C.c_int]
        lib.cg_synchronize.restype = C.c_int
        lib.cg_leaf_set.restype, lib.cg_leaf_set.argtypes = C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_float)]
        lib.cg_set_device.restype, lib.cg_set_device.argtypes = C.c_int, [C.c_int]
        lib.cg_dp_unique_id.restype, lib.cg_dp_unique_id.argtypes = C.c_int, [C.c_char_p]
        lib.cg_dp_init.restype, lib.cg_dp_init.argtypes = C.c_int, [C.c_int, C.c_int, C.c_char_p]
        lib.cg_dp_shutdown.restype = C.c_int
        lib.cg_dp_fused.restype = C.c_int
        lib.cg_dp_multimem.restype = C.c_int
        lib.cg_dp_set_fused_min.restype, lib.cg_dp_set_fused_min.argtypes = C.c_int, [C.c_longlong]
        lib.cg_dp_multimem_supported.restype = C.c_int
        lib.cg_dp_multimem_reason.restype = C.c_char_p
        lib.cg_dp_set_multimem.restype, lib.cg_dp_set_multimem.argtypes = C.c_int, [C.c_int]
        lib.cg_grad.restype, lib.cg_grad.argtypes = C.c_void_p, [C.c_void_p, C.c_char_p]
        lib.cg_eval.restype = C.c_int
        lib.cg_eval.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.POINTER(C.c_float)),
                                C.POINTER(C.c_uint), C.POINTER(C.c_uint), C.POINTER(C.c_float)]
        lib.cg_set_ew_jit_min.restype, lib.cg_set_ew_jit_min.argtypes = None, [C.c_longlong]
        lib.cg_ew_jit_selftest.restype = C.c_longlong
        lib.cg_host_alloc.restype, lib.cg_host_alloc.argtypes = C.c_void_p, [C.c_size_t]
        lib.cg_host_free.restype, lib.cg_host_free.argtypes = None, [C.c_void_p]
        self._pinned_free: Dict[int, List[int]] = {}  # nbytes -> page-locked blocks ready for reuse

    # page-locked host buffers ---------------------------------------------------------------------
    _PINNED_KEEP = 4  # blocks kept per size (page-locking is slow: ~ms per 16 MB)

    def pinned_empty(self, shape) -> np.ndarray:
        """float32 array in page-locked host memory (cg_host_alloc). Its block returns to a small reuse pool
        when the array (and every view of it) has been garbage-collected."""
        n = int(np.prod(shape))
        nbytes = max(n, 1) * 4
        free = self._pinned_free.setdefault(nbytes, [])
        ptr = free.pop() if free else self.lib.cg_host_alloc(nbytes)
        if not ptr:
            self._check(-1)
        arr = np.ctypeslib.as_array((C.c_float * max(n, 1)).from_address(ptr))[:n].reshape(shape)
        base = arr
        while isinstance(base.base, np.ndarray):
            base = base.base
        weakref.finalize(base, self._pinned_release, ptr, nbytes)
        return arr

    def _pinned_release(self, ptr: int, nbytes: int):
        free = self._pinned_free.setdefault(nbytes, [])
        if len(free) < self._PINNED_KEEP:
            free.append(ptr)
        else:
            self.lib.cg_host_free(ptr)

    def _trace_buffer(self, iters: int, size: int) -> np.ndarray:
        # page-locked, so a 1-iteration call can read the root value back from inside the iteration graph
        return self.pinned_empty((iters, size))

    def _check(self, rc: int):
        if rc != 0:
            from .api import GraphError
            raise GraphError(self._last_error().decode())

    def set_tf32(self, on): self.lib.cg_set_tf32(int(on))  # 0 FP32 SIMT, 1 TF32, 2 FP32-accurate 3xTF32
    def set_fix_act_grad(self, on: bool): self.lib.cg_set_fix_act_grad(int(on))
    def set_fix_broadcast_sum(self, on: bool): self.lib.cg_set_fix_broadcast_sum(int(on))
    def set_fix_sub_sign(self, on: bool): self.lib.cg_set_fix_sub_sign(int(on))
    def set_fix_tied_params(self, on: bool): self.lib.cg_set_fix_tied_params(int(on))
    def set_fuse(self, on: bool): self.lib.cg_set_fuse(int(on))
    def set_cuda_graph(self, on: bool): self.lib.cg_set_cuda_graph(int(on))
    def set_strict_gemm_max(self, dim: int): self.lib.cg_set_strict_gemm_max(int(dim))
    def set_graph_streams(self, n: int): self.lib.cg_set_graph_streams(int(n))
    def set_ew_jit_min(self, min_elems: int): self.lib.cg_set_ew_jit_min(int(min_elems))
    def set_tf32_pair(self, on: bool): self.lib.cg_set_tf32_pair(int(on))
    def set_io_graph(self, on: bool): self.lib.cg_set_io_graph(int(on))
    def set_hoist(self, on: bool): self.lib.cg_set_hoist(int(on))
    def set_vm_max_dim(self, dim: int): self.lib.cg_set_vm_max_dim(int(dim))
    def set_tape_jit(self, on: bool): self.lib.cg_set_tape_jit(int(on))
    def set_elide_root(self, on: bool): self.lib.cg_set_elide_root(int(on))
    def set_device(self, dev: int): self._check(self.lib.cg_set_device(int(dev)))
    def launch_count(self) -> int: return int(self.lib.cg_launch_count())
    def synchronize(self): self._check(self.lib.cg_synchronize())

    def plan_stats(self, f, learn_rate: float) -> dict:
        st = PlanStats()
        self._check(self.lib.cg_plan_stats_get(f._h, float(learn_rate), C.byref(st)))
        return st.as_dict()

    def plan_describe(self, f, learn_rate: float) -> List[str]:
        """One line per kernel of the compiled iteration, in launch order (cg_plan_describe)."""
        fn = self.lib.cg_plan_describe
        fn.restype, fn.argtypes = C.c_longlong, [C.c_void_p, C.c_float, C.c_char_p, C.c_ulonglong]
        need = fn(f._h, float(learn_rate), None, 0)
        if need < 0:
            self._check(-1)
        buf = C.create_string_buffer(int(need))
        fn(f._h, float(learn_rate), buf, need)
        return buf.value.decode().splitlines()

    def time_iterations(self, f, learn_rate: float, warmup: int, reps: int) -> float:
        """ms per iteration, CUDA events on the library stream (inputs already resident in HBM)."""
        ms = self.lib.cg_time_iterations(f._h, float(learn_rate), int(warmup), int(reps))
        if ms < 0:
            self._check(-1)
        return ms

    def minimize_async(self, f, learn_rate: float, iters: int):
        self._check(self.lib.cg_minimize_async(f._h, float(learn_rate), int(iters)))

    # optimize.rs:8-56 ---------------------------------------------------------------------------
    def grad(self, f, param: str):
        """`f.grad(param)`: a new Function, d f / d param, built symbolically (the README's "naive" gradient)."""
        return self._wrap(self.lib.cg_grad(f._h, param.encode()))

    def eval(self, f, args=None):
        """`f.eval(&args)`: the forward value at the current parameters; nothing cached changes."""
        args = args or {}
        n = len(args)
        names = (C.c_char_p * max(n, 1))()
        ptrs = (C.POINTER(C.c_float) * max(n, 1))()
        hs, ws = (C.c_uint * max(n, 1))(), (C.c_uint * max(n, 1))()
        keep = []
        for i, (k, v) in enumerate(args.items()):
            names[i] = k.encode()
            arr = np.ascontiguousarray(np.atleast_1d(np.asarray(v.value, dtype=np.float32)))
            keep.append(arr)
            ptrs[i] = arr.ctypes.data_as(C.POINTER(C.c_float))
            hs[i], ws[i] = (v.value.shape if v.is_matrix else (0, 0))
        d = f.dims()
        out = np.zeros(d[0] * d[1] if d else 1, dtype=np.float32)
        self._check(self.lib.cg_eval(f._h, n, names, ptrs, hs, ws, out.ctypes.data_as(C.POINTER(C.c_float))))
        return out.reshape(d[1], d[0]).T.copy() if d else np.float32(out[0])

    # data parallel ------------------------------------------------------------------------------
    def dp_unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        self._check(self.lib.cg_dp_unique_id(buf))
        return buf.raw

    def dp_init(self, rank: int, world: int, uid: bytes):
        self._check(self.lib.cg_dp_init(int(rank), int(world), uid))

    def dp_shutdown(self):
        self._check(self.lib.cg_dp_shutdown())

    def dp_fused(self) -> bool:
        """True when weight gradients are exchanged by the fused peer-memory kernel rather than ncclAllReduce."""
        return bool(self.lib.cg_dp_fused())

    def dp_set_fused_min(self, min_elems: int):
        """Weights with fewer elements stay on ncclAllReduce + update (default 2^18; < 0 restores it)."""
        self._check(self.lib.cg_dp_set_fused_min(int(min_elems)))

    def dp_multimem(self) -> bool:
        """True when plans built now exchange weight gradients inside the NVSwitch (multimem.ld_reduce / multimem.st)."""
        return bool(self.lib.cg_dp_multimem())

    def dp_multimem_supported(self) -> bool:
        return bool(self.lib.cg_dp_multimem_supported())

    def dp_multimem_reason(self) -> str:
        """Why the multicast heap could not be created (empty when it was)."""
        return (self.lib.cg_dp_multimem_reason() or b"").decode()

    def dp_set_multimem(self, mode: int):
        """0 never, 1 (default) from three ranks up, 2 always."""
        self._check(self.lib.cg_dp_set_multimem(int(mode)))

    def dp_exchange_name(self) -> str:
        if self.dp_multimem():
            return "fused in-switch all-reduce + SGD update + broadcast kernel (NVLS multimem.ld_reduce / multimem.st over a symmetric multicast heap)"
        if self.dp_fused():
            return "fused all-reduce + SGD update kernel over NVLink peer memory (CUDA IPC)"
        return "ncclAllReduce + elementwise update"


def get_api() -> ProductApi:
    """The `Api` bound to the product library's `cg_*` symbols."""
    global _API
    if _API is None:
        _API = ProductApi(load_library())
    return _API
