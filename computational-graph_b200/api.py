"""Host-side mirror of the reference's `Function` / `Constant` surface over a graph-level C ABI.

Mirrors (reference file:line):
  * `Function::{param,constant,input,scalar_param,matrix_param,single_val_param,random_param,
    random_matrix,scalar,matrix,single_val_matrix}`           src/function.rs:100-155
  * `.sq() .abs() .signum() .sigmoid() .tanh()`, `-f`, `+ - *`, `dot`   src/ops.rs:331-343,363-407
  * `minimize / maximize(&args, lr, iters, print_freq)`        src/optimize.rs:59-81
  * `all_less_than / all_equal`                                 src/ops.rs:309-329
  * `Constant::{Scalar, matrix, single_val, random}`            src/constant.rs:30-76

The binding is generic over a symbol prefix: the product library exports `cg_*` (include/cg_graph.h);
the test-only CPU oracle exports the same signatures as `orc_*` and binds through `oracle/oracle.py`.
This module never imports the oracle.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, List, Optional, Sequence, Tuple, Union

import numpy as np

_F = C.c_void_p
_FP = C.POINTER(C.c_float)
_UP = C.POINTER(C.c_uint)


class Constant:
    """`enum Constant {Scalar(f32), Matrix(Matrix)}` (src/constant.rs:5-8). Matrix values are host
    numpy arrays in ROW-major order, exactly what `Constant::matrix(h, w, vals)` takes."""

    __slots__ = ("value",)

    def __init__(self, value):
        if np.isscalar(value):
            self.value = np.float32(value)
        else:
            v = np.ascontiguousarray(np.asarray(value, dtype=np.float32))
            if v.ndim != 2:
                raise ValueError("not supported")  # constant.rs:40,58
            self.value = v

    # constructors -------------------------------------------------------------------------------
    @staticmethod
    def Scalar(x: float) -> "Constant":
        return Constant(np.float32(x))

    @staticmethod
    def matrix(height: int, width: int, vals: Sequence[float]) -> "Constant":
        v = np.asarray(vals, dtype=np.float32)
        assert v.size == height * width, "wrong number of values"  # matrix.rs:60
        return Constant(v.reshape(height, width))

    @staticmethod
    def single_val(dims: Sequence[int], val: float) -> "Constant":
        if len(dims) == 0:
            return Constant.Scalar(val)
        if len(dims) == 2:
            return Constant(np.full(tuple(dims), val, dtype=np.float32))
        raise ValueError("not supported")

    @staticmethod
    def random(dims: Sequence[int], lo: float, hi: float, rng: Optional[np.random.Generator] = None) -> "Constant":
        """constant.rs:45-60 draws from an unseeded thread_rng; here the caller injects a seeded
        numpy Generator so CPU oracle and GPU runs see identical values."""
        rng = rng if rng is not None else np.random.default_rng()
        if len(dims) == 0:
            return Constant.Scalar(rng.uniform(lo, hi))
        if len(dims) == 2:
            return Constant(rng.uniform(lo, hi, size=tuple(dims)).astype(np.float32))
        raise ValueError("not supported")

    # accessors ----------------------------------------------------------------------------------
    @property
    def is_matrix(self) -> bool:
        return isinstance(self.value, np.ndarray)

    def dims(self) -> List[int]:
        return list(self.value.shape) if self.is_matrix else []

    def width(self) -> int:
        return self.value.shape[1] if self.is_matrix else 1

    def height(self) -> int:
        return self.value.shape[0] if self.is_matrix else 1


class GraphError(RuntimeError):
    """A reference `panic!` / `check()` surfaced through the C ABI."""


class Api:
    """One bound graph library (product `cg_*` or the oracle's `orc_*`)."""

    def _trace_buffer(self, iters: int, size: int) -> np.ndarray:
        """Host array receiving the per-iteration root values (the product overrides it with page-locked memory)."""
        return np.zeros((iters, size), dtype=np.float32)

    def __init__(self, lib: C.CDLL, prefix: str):
        self.lib = lib
        self.prefix = prefix
        g = self._sym
        for name in ("scalar_param",):
            g(name).restype, g(name).argtypes = _F, [C.c_char_p, C.c_float]
        g("matrix_param").restype, g("matrix_param").argtypes = _F, [C.c_char_p, C.c_uint, C.c_uint, _FP]
        g("single_val_param").restype, g("single_val_param").argtypes = _F, [C.c_char_p, C.c_uint, C.c_uint, C.c_float]
        g("scalar").restype, g("scalar").argtypes = _F, [C.c_float]
        g("matrix").restype, g("matrix").argtypes = _F, [C.c_uint, C.c_uint, _FP]
        g("single_val_matrix").restype, g("single_val_matrix").argtypes = _F, [C.c_uint, C.c_uint, C.c_float]
        g("input").restype, g("input").argtypes = _F, [C.c_char_p, C.c_int, C.c_uint, C.c_uint]
        for name in ("neg", "sq", "abs", "signum", "sigmoid", "tanh", "clone"):
            g(name).restype, g(name).argtypes = _F, [_F]
        for name in ("add", "sub", "mul"):
            g(name).restype, g(name).argtypes = _F, [_F, _F]
        g("dot").restype, g("dot").argtypes = _F, [_F, C.c_int, _F, C.c_int]
        g("release").restype, g("release").argtypes = None, [_F]
        for name in ("minimize", "maximize"):
            g(name).restype = C.c_int
            g(name).argtypes = [_F, C.c_int, C.POINTER(C.c_char_p), C.POINTER(_FP), _UP, _UP, C.c_float, C.c_int,
                                C.c_int, _FP]
        g("value_dims").restype, g("value_dims").argtypes = C.c_int, [_F, _UP, _UP]
        g("value_get").restype, g("value_get").argtypes = None, [_F, _FP]
        g("all_less_than").restype, g("all_less_than").argtypes = C.c_bool, [_F, C.c_float]
        g("all_equal").restype, g("all_equal").argtypes = C.c_bool, [_F, C.c_float]
        g("num_leaves").restype, g("num_leaves").argtypes = C.c_int, [_F]
        g("leaf_info").restype, g("leaf_info").argtypes = C.c_int, [_F, C.c_int, C.c_char_p, C.c_int, _UP, _UP]
        g("leaf_get").restype, g("leaf_get").argtypes = C.c_int, [_F, C.c_int, _FP]
        g("set_mode").restype, g("set_mode").argtypes = None, [C.c_int]
        self._last_error = getattr(lib, prefix + "last_error", None)
        if self._last_error is not None:
            self._last_error.restype = C.c_char_p
        api = self

        class BoundFunction(Function):
            _api = api

        BoundFunction.__name__ = "Function"
        self.Function = BoundFunction
        self.Constant = Constant

    def _sym(self, name: str):
        return getattr(self.lib, self.prefix + name)

    def _wrap(self, handle) -> "Function":
        if not handle:
            msg = self._last_error().decode() if self._last_error is not None else "graph construction failed"
            raise GraphError(msg)
        return self.Function(handle)

    # execution mode: "exact" = the reference's tree walk, "dag" = memoised ---------
    def set_mode(self, mode: str) -> None:
        self._sym("set_mode")({"exact": 0, "dag": 1}[mode])

    # free functions of the reference surface ----------------------------------------------------
    def dot(self, f1: "Function", f2: "Function", t1: bool = False, t2: bool = False) -> "Function":
        """`dot!` (src/macros.rs:51-64) / `ops::dot` (src/ops.rs:395-407)."""
        return self._wrap(self._sym("dot")(f1._h, int(t1), f2._h, int(t2)))

    def sq(self, f: "Function") -> "Function":  # README.md:60 spelling
        return f.sq()


def _fptr(a: np.ndarray):
    return a.ctypes.data_as(_FP)


class Function:
    """`struct Function` (src/function.rs:10-16). Instances are created through `Api.Function`."""

    _api: Api = None  # bound by Api

    def __init__(self, handle):
        self._h = handle

    def __del__(self):
        try:
            if self._h:
                self._api._sym("release")(self._h)
                self._h = None
        except Exception:
            pass

    # constructors (src/function.rs:100-155) ------------------------------------------------------
    @classmethod
    def param(cls, name: str, value: Constant) -> "Function":
        a = cls._api
        if value.is_matrix:
            h, w = value.value.shape
            return a._wrap(a._sym("matrix_param")(name.encode(), h, w, _fptr(value.value)))
        return a._wrap(a._sym("scalar_param")(name.encode(), float(value.value)))

    @classmethod
    def constant(cls, value: Constant) -> "Function":
        a = cls._api
        if value.is_matrix:
            h, w = value.value.shape
            return a._wrap(a._sym("matrix")(h, w, _fptr(value.value)))
        return a._wrap(a._sym("scalar")(float(value.value)))

    @classmethod
    def input(cls, name: str, dims: Sequence[int]) -> "Function":
        a = cls._api
        if len(dims) == 0:
            return a._wrap(a._sym("input")(name.encode(), 0, 0, 0))
        return a._wrap(a._sym("input")(name.encode(), 1, dims[0], dims[1]))

    @classmethod
    def scalar_param(cls, name: str, value: float) -> "Function":
        return cls.param(name, Constant.Scalar(value))

    @classmethod
    def matrix_param(cls, name: str, height: int, width: int, vals) -> "Function":
        return cls.param(name, Constant.matrix(height, width, vals))

    @classmethod
    def single_val_param(cls, name: str, dims: Sequence[int], val: float) -> "Function":
        a = cls._api
        if len(dims) == 0:
            return cls.scalar_param(name, val)
        return a._wrap(a._sym("single_val_param")(name.encode(), dims[0], dims[1], float(val)))

    @classmethod
    def random_param(cls, name: str, dims, lo: float, hi: float, rng=None) -> "Function":
        return cls.param(name, Constant.random(dims, lo, hi, rng))

    @classmethod
    def random_matrix(cls, dims, lo: float, hi: float, rng=None) -> "Function":
        return cls.constant(Constant.random(dims, lo, hi, rng))

    @classmethod
    def scalar(cls, x: float) -> "Function":
        return cls.constant(Constant.Scalar(x))

    @classmethod
    def matrix(cls, height: int, width: int, values) -> "Function":
        return cls.constant(Constant.matrix(height, width, values))

    @classmethod
    def single_val_matrix(cls, height: int, width: int, value: float) -> "Function":
        a = cls._api
        return a._wrap(a._sym("single_val_matrix")(height, width, float(value)))

    # ops (src/ops.rs) ------------------------------------------------------------------------------
    def _un(self, name: str) -> "Function":
        return self._api._wrap(self._api._sym(name)(self._h))

    def _bin(self, name: str, other: "Function") -> "Function":
        return self._api._wrap(self._api._sym(name)(self._h, other._h))

    # optimize.rs:8-56 (bound by the product only; the CPU oracle restates the hot loop, not these)
    def grad(self, param: str) -> "Function":
        """`f.grad(param)`: a new Function, d f / d param, built symbolically."""
        return self._api.grad(self, param)

    def eval(self, args=None):
        """`f.eval(&args)`: the forward value at the current parameters."""
        return self._api.eval(self, args)

    def sq(self): return self._un("sq")
    def abs(self): return self._un("abs")
    def signum(self): return self._un("signum")
    def sigmoid(self): return self._un("sigmoid")
    def tanh(self): return self._un("tanh")
    def __neg__(self): return self._un("neg")
    def __add__(self, o): return self._bin("add", o)
    def __sub__(self, o): return self._bin("sub", o)
    def __mul__(self, o): return self._bin("mul", o)
    def clone(self): return self._un("clone")

    # optimisation (src/optimize.rs:59-81) ---------------------------------------------------------
    def _run(self, which: str, args: Optional[Dict[str, Constant]], learn_rate: float, iters: int,
             print_freq: Optional[int], trace: bool):
        args = args or {}
        n = len(args)
        names = (C.c_char_p * max(n, 1))()
        ptrs = (_FP * max(n, 1))()
        hs = (C.c_uint * max(n, 1))()
        ws = (C.c_uint * max(n, 1))()
        keep = []
        for i, (k, v) in enumerate(args.items()):
            names[i] = k.encode()
            arr = np.ascontiguousarray(np.atleast_1d(np.asarray(v.value, dtype=np.float32)))
            keep.append(arr)
            ptrs[i] = _fptr(arr)
            hs[i], ws[i] = (v.value.shape if v.is_matrix else (0, 0))
        out = None
        tptr = None
        if trace:
            h, w = C.c_uint(), C.c_uint()
            ism = self._api._sym("value_dims")(self._h, C.byref(h), C.byref(w))
            size = h.value * w.value if ism else 1
            out = self._api._trace_buffer(iters, size)
            tptr = _fptr(out)
        pf = print_freq if print_freq is not None else iters + 1
        rc = self._api._sym(which)(self._h, n, names, ptrs, hs, ws, float(learn_rate), int(iters), int(pf), tptr)
        if rc != 0:
            msg = self._api._last_error().decode() if self._api._last_error is not None else f"rc={rc}"
            raise GraphError(msg)
        if trace:
            h, w = C.c_uint(), C.c_uint()
            if self._api._sym("value_dims")(self._h, C.byref(h), C.byref(w)):
                # storage order is column-major: expose as (iters, h, w) row-indexable arrays
                out = out.reshape(iters, w.value, h.value).transpose(0, 2, 1)
            else:
                out = out.reshape(iters)
        return out

    def minimize(self, args=None, learn_rate: float = 0.01, iters: int = 1, print_freq: Optional[int] = None,
                 trace: bool = False):
        """`f.minimize(&args, lr, iters, print_freq)`; the README's 3-argument spelling is accepted.
        With `trace=True` returns the root value of every iteration (the value the reference would
        print: the forward value *before* that iteration's update)."""
        return self._run("minimize", args, learn_rate, iters, print_freq, trace)

    def maximize(self, args=None, learn_rate: float = 0.01, iters: int = 1, print_freq: Optional[int] = None,
                 trace: bool = False):
        return self._run("maximize", args, learn_rate, iters, print_freq, trace)

    # read-back -----------------------------------------------------------------------------------------
    def dims(self) -> List[int]:
        h, w = C.c_uint(), C.c_uint()
        ism = self._api._sym("value_dims")(self._h, C.byref(h), C.byref(w))
        return [h.value, w.value] if ism else []

    def value(self) -> Union[float, np.ndarray]:
        d = self.dims()
        if not d:
            buf = np.zeros(1, dtype=np.float32)
            self._api._sym("value_get")(self._h, _fptr(buf))
            return np.float32(buf[0])
        buf = np.zeros(d[0] * d[1], dtype=np.float32)
        self._api._sym("value_get")(self._h, _fptr(buf))
        return buf.reshape(d[1], d[0]).T.copy()  # column-major storage -> [row, col]

    def all_less_than(self, x: float) -> bool:
        return bool(self._api._sym("all_less_than")(self._h, float(x)))

    def all_equal(self, x: float) -> bool:
        return bool(self._api._sym("all_equal")(self._h, float(x)))

    def set_leaf(self, i: int, value) -> None:
        """Checkpoint restore (the reference can only `Display` its parameters, src/print.rs:71-80):
        overwrite leaf `i` of `leaves()` with `value` — a [rows, cols] array (or a scalar) as `leaves()` returns it."""
        arr = np.ascontiguousarray(np.atleast_1d(np.asarray(value, dtype=np.float32)))
        fn = self._api._sym("leaf_set")
        fn.restype, fn.argtypes = C.c_int, [_F, C.c_int, _FP]
        if fn(self._h, int(i), _fptr(arr)) != 0:
            msg = self._api._last_error().decode() if self._api._last_error is not None else "leaf_set failed"
            raise GraphError(msg or f"no leaf {i}")

    def load_leaves(self, saved) -> None:
        """Restore every leaf from a list as returned by `leaves()` (names must match, in order)."""
        cur = self.leaves()
        if [n for n, _ in cur] != [n for n, _ in saved]:
            raise GraphError("checkpoint does not match this function's parameter leaves")
        for i, (_, v) in enumerate(saved):
            self.set_leaf(i, v)

    def leaves(self) -> List[Tuple[str, Union[np.float32, np.ndarray]]]:
        """Every parameter leaf of the tree in memoised post-order. Each *use site* of a param is its
        own leaf (the reference clones the value into every use, Q1 / src/function.rs:10-16)."""
        a = self._api
        out = []
        for i in range(a._sym("num_leaves")(self._h)):
            name = C.create_string_buffer(128)
            h, w = C.c_uint(), C.c_uint()
            ism = a._sym("leaf_info")(self._h, i, name, 128, C.byref(h), C.byref(w))
            if ism:
                buf = np.zeros(h.value * w.value, dtype=np.float32)
                a._sym("leaf_get")(self._h, i, _fptr(buf))
                out.append((name.value.decode(), buf.reshape(w.value, h.value).T.copy()))
            else:
                buf = np.zeros(1, dtype=np.float32)
                a._sym("leaf_get")(self._h, i, _fptr(buf))
                out.append((name.value.decode(), np.float32(buf[0])))
        return out
