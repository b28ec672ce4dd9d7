"""TEST INFRASTRUCTURE — a float64, multithreaded (numpy / BLAS) restatement of the path in DAG mode, for parity at sizes
the C oracle cannot finish (one iteration of BASELINE config 4 is 6 TFLOP: minutes on one core, seconds here).

It is a second, independent restatement of the same rules `oracle/cg_oracle.c` follows (SURVEY §8 "Execution-mode
definitions", dag row), written over a duck-typed mirror of the `Api` surface so that `computational_graph_b200.lstm`
builds the reference's graph on it unchanged:

  * graph construction: src/function.rs:10-16,100-155 and src/ops.rs:199-264,395-407 — every use of a param leaf is its
    own copy (Q1: `clone` deep-copies the value, shares the body); NO value-based identity elimination (Q4) — valid for
    graphs where none fires, which the pinning test checks by comparing leaf lists with the oracle;
  * forward: src/optimize.rs:84-138, memoised per shared body;
  * backward: every unique node once, in reverse first-evaluation order; error into a node = sum of its consumers'
    contributions; the per-node rules of src/optimize.rs:153-200 including Q3 (Sigmoid / Tanh hand the child the LOCAL
    derivative only); leaf update `p -= lr * e` (src/optimize.rs:149-152);
  * root error = ones (src/optimize.rs:69); the traced value is the forward value before the update (Q12).

Matrices only (the LSTM has no scalars). tests/test_oracle.py pins it to cg_oracle.c's dag mode at small sizes.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np

LEAF_KINDS = ("param", "constant")


class Constant:
    def __init__(self, value):
        self.value = np.asarray(value)
        self.is_matrix = self.value.ndim == 2


class Node:
    """`Expr` (src/function.rs:18-33): shared between clones of the Function that owns it."""

    def __init__(self, kind: str, f1: Optional["Function"] = None, f2: Optional["Function"] = None, name: str = ""):
        self.kind, self.f1, self.f2, self.name = kind, f1, f2, name


class Function:
    api: "F64Api" = None

    def __init__(self, body: Node, value: Optional[np.ndarray], has_params: bool):
        self.body, self.value, self.has_params = body, value, has_params

    # src/function.rs:100-155
    @classmethod
    def param(cls, name: str, value: Constant) -> "Function":
        return cls(Node("param", name=name), np.array(value.value, dtype=np.float64), True)

    @classmethod
    def constant(cls, value: Constant) -> "Function":
        return cls(Node("constant"), np.array(value.value, dtype=np.float64), False)

    def clone(self) -> "Function":
        # #[derive(Clone)]: the value is copied, the body shared; a leaf's value IS its state, so each use site owns one
        v = self.value.copy() if self.body.kind in LEAF_KINDS else None
        return Function(self.body, v, self.has_params)

    def _unary(self, kind):
        return Function(Node(kind, self.clone()), None, self.has_params)

    def _binary(self, kind, other):
        return Function(Node(kind, self.clone(), other.clone()), None, self.has_params or other.has_params)

    def sigmoid(self): return self._unary("sigmoid")
    def tanh(self): return self._unary("tanh")
    def sq(self): return self._unary("sq")
    def __neg__(self): return self._unary("neg")
    def __add__(self, o): return self._binary("add", o)
    def __sub__(self, o): return self._binary("sub", o)
    def __mul__(self, o): return self._binary("mul", o)

    # ---- src/optimize.rs:59-72, dag mode ----------------------------------------------------------------------------
    def minimize(self, args=None, learn_rate: float = 0.01, iters: int = 1, trace: bool = False):
        out = []
        for _ in range(iters):
            val: Dict[int, np.ndarray] = {}
            order: List[Function] = []

            def value_of(f: Function) -> np.ndarray:
                return f.value if f.body.kind in LEAF_KINDS else val[id(f.body)]

            def forward(f: Function):
                e = f.body
                if e.kind in LEAF_KINDS or id(e) in val:
                    return
                if e.f1 is not None:
                    forward(e.f1)
                if e.f2 is not None:
                    forward(e.f2)
                a = value_of(e.f1)
                b = value_of(e.f2) if e.f2 is not None else None
                if e.kind == "sigmoid": v = 1.0 / (1.0 + np.exp(-a))
                elif e.kind == "tanh": v = np.tanh(a)
                elif e.kind == "sq": v = a * a
                elif e.kind == "neg": v = -a
                elif e.kind == "add": v = a + b
                elif e.kind == "sub": v = a - b
                elif e.kind == "mul": v = a * b
                elif e.kind == "dot": v = a @ b
                else: raise ValueError(e.kind)
                val[id(e)] = v
                order.append(f)

            forward(self)
            root = value_of(self)
            if trace:
                out.append(np.array(root, dtype=np.float64))
            err: Dict[int, np.ndarray] = {id(self.body): np.ones_like(root)}

            def deliver(child: Function, g: np.ndarray):
                if not child.has_params:
                    return
                k = child.body.kind
                if k == "param":
                    child.value -= learn_rate * g          # optimize.rs:149-152
                elif k not in LEAF_KINDS:
                    key = id(child.body)
                    err[key] = g if key not in err else err[key] + g

            for f in reversed(order):
                e = f.body
                if id(e) not in err or not f.has_params:
                    continue
                g = err.pop(id(e))
                v = val[id(e)]
                if e.kind == "neg": deliver(e.f1, -g)
                elif e.kind == "sq": deliver(e.f1, g * 2.0 * value_of(e.f1))
                elif e.kind == "sigmoid": deliver(e.f1, (1.0 - v) * v)          # Q3: the local derivative only
                elif e.kind == "tanh": deliver(e.f1, 1.0 - v * v)               # Q3
                elif e.kind == "add":
                    deliver(e.f1, g)
                    deliver(e.f2, g)
                elif e.kind == "sub":
                    deliver(e.f1, g)
                    deliver(e.f2, -g)
                elif e.kind == "mul":
                    a, b = value_of(e.f1), value_of(e.f2)
                    deliver(e.f1, g * b)
                    deliver(e.f2, a * g)
                elif e.kind == "dot":
                    a, b = value_of(e.f1), value_of(e.f2)
                    # both are formed before either child is touched (optimize.rs:196-200)
                    ga = g @ b.T if e.f1.has_params else None
                    gb = a.T @ g if e.f2.has_params else None
                    if ga is not None: deliver(e.f1, ga)
                    if gb is not None: deliver(e.f2, gb)
            self.value = root
        return out if trace else None

    def leaves(self):
        """Param leaves in memoised post-order, f1 before f2 (graph.cpp collect_param_leaves)."""
        seen, out = set(), []

        def collect(f: Function):
            e = f.body
            if e.kind == "param":
                out.append((e.name, f.value))
                return
            if e.kind in LEAF_KINDS or id(e) in seen:
                return
            seen.add(id(e))
            if e.f1 is not None:
                collect(e.f1)
            if e.f2 is not None:
                collect(e.f2)
        collect(self)
        return out


class F64Api:
    """The slice of `Api` that lstm.py uses."""
    Function = Function
    Constant = Constant

    @staticmethod
    def dot(f1: Function, f2: Function) -> Function:
        return Function(Node("dot", f1.clone(), f2.clone()), None, f1.has_params or f2.has_params)


def get_f64() -> F64Api:
    return F64Api()
