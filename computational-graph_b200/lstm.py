"""`lstm_custom_params` / `lstm` of the reference (src/lstm.rs:7-77), transliterated line for line over
the `Api` surface, so the graph — including every clone the Rust operators make (Q1), the `h' = o *
tanh(C_old)` quirk and the "return final C" quirk (Q6) — is the reference's graph.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np

from .api import Api, Constant, Function

PARAM_NAMES = ["Wi", "Ui", "bi", "Wc", "Uc", "bc", "Wf", "Uf", "bf", "Wo", "Uo", "Vo", "bo", "C", "h"]


def lstm_custom_params(api: Api, inputs: Sequence, Wi, Ui, bi, Wc, Uc, bc, Wf, Uf, bf, Wo, Uo, Vo, bo, C, h):
    """src/lstm.rs:7-43. `inputs` holds `Constant`s (wrapped with `Function::constant`, as the
    reference does) or ready-made `Function`s (e.g. `Function::input` nodes fed per call)."""
    dot = api.dot
    for x_in in inputs:  # the tail recursion of lstm.rs:26-42, unrolled
        x = x_in if isinstance(x_in, Function) else api.Function.constant(x_in)
        i = (dot(x, Wi) + (dot(h, Ui) + bi)).sigmoid()          # lstm.rs:32
        c = (dot(x, Wc) + (dot(h, Uc) + bc)).tanh()             # lstm.rs:33
        f = (dot(x, Wf) + (dot(h, Uf) + bf)).sigmoid()          # lstm.rs:34
        C_new = i * c + (f * C)                                 # lstm.rs:35
        o = (dot(x, Wo) + dot(h, Uo) + (dot(C_new, Vo) + bo)).sigmoid()  # lstm.rs:36
        h_new = o * C.tanh()                                    # lstm.rs:37 (old C — Q6)
        C, h = C_new, h_new
    return C                                                    # lstm.rs:27


def lstm_params(d_in: int, d_hidden: int, batch: int, rng: np.random.Generator, lo: float = -1.0, hi: float = 1.0,
                scale: Optional[float] = None) -> Dict[str, np.ndarray]:
    """Seeded stand-in for the 15 `Function::random_param` draws of src/lstm.rs:60-76. The reference
    uses d x d for everything (batch == d_in == d_hidden == d); the general shapes are W*: [d_in, H],
    U*, Vo: [H, H], b*, C, h: [batch, H] (biases are full matrices because M+M needs equal dims,
    src/ops.rs:225). `scale` multiplies the weight draws (the 1/sqrt(H) variant of SURVEY §8d)."""
    shapes = {"W": (d_in, d_hidden), "U": (d_hidden, d_hidden), "V": (d_hidden, d_hidden)}
    out = {}
    for name in PARAM_NAMES:
        shape = shapes.get(name[0], (batch, d_hidden))
        v = rng.uniform(lo, hi, size=shape).astype(np.float32)
        if scale is not None and name[0] in "WUV":
            v = (v * np.float32(scale)).astype(np.float32)
        out[name] = v
    return out


def lstm(api: Api, inputs: Sequence, params: Dict[str, np.ndarray]):
    """src/lstm.rs:57-77 with injected (seeded) initial values instead of `random_param`."""
    fs = [api.Function.param(n, Constant(params[n])) for n in PARAM_NAMES]
    return lstm_custom_params(api, inputs, *fs)


def lstm_loss(api: Api, inputs: Sequence, target, params: Dict[str, np.ndarray]):
    """`(lstm(inputs) - target).sq()` — the README benchmark program (src/main.rs:177-191)."""
    t = target if isinstance(target, Function) else api.Function.constant(target)
    return (lstm(api, inputs, params) - t).sq()
