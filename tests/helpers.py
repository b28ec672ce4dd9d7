"""Shared builders for the parity tests: the same seeded graphs on the oracle and on the product."""
import numpy as np

from computational_graph_b200.lstm import lstm_loss, lstm_params


def ulp_diff(a, b):
    """Distance in units in the last place between two float32 arrays (same sign assumed near zero)."""
    a = np.ascontiguousarray(a, dtype=np.float32).view(np.int32).astype(np.int64)
    b = np.ascontiguousarray(b, dtype=np.float32).view(np.int32).astype(np.int64)
    a = np.where(a < 0, -(a & 0x7FFFFFFF), a)
    b = np.where(b < 0, -(b & 0x7FFFFFFF), b)
    return np.abs(a - b)


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    denom = max(float(np.max(np.abs(b))), 1e-30)
    return float(np.max(np.abs(a - b))) / denom


def close_err(a, b, tol):
    """Per-element closeness: max over elements of |a - b| / (tol * |b| + tol * rms(b)); <= 1 passes.
    Unlike `rel_err` (max-norm relative), a wrong SMALL entry in an array with one large entry does not pass: every
    element is held to `tol` relative to its own magnitude plus `tol` of the array's typical magnitude."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.shape != b.shape:
        raise AssertionError(f"shape mismatch {a.shape} vs {b.shape}")
    if a.size == 0:
        return 0.0
    rms = float(np.sqrt(np.mean(b * b)))
    bound = tol * np.abs(b) + tol * max(rms, 1e-30)
    return float(np.max(np.abs(a - b) / bound))


def readme_lstm(api, d, T=2, seed=None, batch=None, d_in=None, scale=None):
    """src/main.rs:177-191 (`run_lstm`): T random d x d inputs and a target ~U(-.1,.1), 15 params ~U(-1,1),
    loss (lstm(inputs) - target).sq(). Seeded (seed 300+d)."""
    rng = np.random.default_rng(300 + d if seed is None else seed)
    batch = d if batch is None else batch
    d_in = d if d_in is None else d_in
    params = lstm_params(d_in, d, batch, rng, scale=scale)
    xs = [api.Constant(rng.uniform(-.1, .1, (batch, d_in)).astype(np.float32)) for _ in range(T)]
    tgt = api.Constant(rng.uniform(-.1, .1, (batch, d)).astype(np.float32))
    return lstm_loss(api, xs, tgt, params)


def c2_graph(api, n, seed=2):
    """BASELINE config 2: tanh(sigmoid(W) - sq(V)) + abs(W), W,V ~ U(-1,1) n x n params."""
    rng = np.random.default_rng(seed)
    W = api.Function.param("W", api.Constant(rng.uniform(-1, 1, (n, n)).astype(np.float32)))
    V = api.Function.param("V", api.Constant(rng.uniform(-1, 1, (n, n)).astype(np.float32)))
    return (W.sigmoid() - V.sq()).tanh() + W.abs()


def leaves_dict(f):
    out = {}
    for i, (name, val) in enumerate(f.leaves()):
        out[f"{i}:{name}"] = np.asarray(val)
    return out
