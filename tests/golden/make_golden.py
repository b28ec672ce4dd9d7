"""Generates tests/golden/*.npz with the CPU oracle routed through the REFERENCE-COMPILED backend
(oracle/_ref/libmatrix_ref_O0.so = /root/reference/src/cpu/{matrix,ops,util}.cpp built with the reference's own
flags). Run in the build container (needs /root/reference to build oracle/_ref):  python tests/golden/make_golden.py
The vectors travel to the GPU box, where /root/reference does not exist."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle.oracle import get_oracle  # noqa: E402
from helpers import c2_graph, leaves_dict, readme_lstm  # noqa: E402
from reference_cases import CASES, run_case  # noqa: E402


def main():
    orc = get_oracle()
    assert orc.use_ref("O0"), "oracle/_ref is missing: run `make -C oracle` where /root/reference exists"
    orc.set_mode("exact")
    out = {}
    # the reference's own 15 tests (src/main.rs:85-166): every function's 10-iteration trajectory + final value
    for name in CASES:
        res, _ = run_case(orc, name, trace=True)
        for i, (f, tr) in enumerate(res):
            out[f"{name}/{i}/trace"] = np.asarray(tr, dtype=np.float32)
            out[f"{name}/{i}/final"] = np.asarray(f.value(), dtype=np.float32)
    np.savez_compressed(os.path.join(HERE, "reference_suite.npz"), **out)

    # C1: README scalar example
    x = orc.Function.scalar_param("x", 1.0)
    f = (x + orc.Function.scalar(3.0)).sq()
    tr = f.minimize({}, 0.01, 1000, trace=True)
    np.savez_compressed(os.path.join(HERE, "c1_scalar.npz"), trace=np.asarray(tr, np.float32),
                        x_final=np.float32(f.leaves()[0][1]))

    # C2 at 32x32, 20 steps
    f = c2_graph(orc, 32)
    tr = f.minimize({}, 0.01, 20, trace=True)
    np.savez_compressed(os.path.join(HERE, "c2_elementwise_32.npz"), trace=np.asarray(tr, np.float32),
                        **{k.replace(":", "_"): v for k, v in leaves_dict(f).items()})

    # C3: README LSTM d=5 and d=10, T=2, exact mode, 10 iterations: trajectory + all 28 leaves
    for d in (5, 10):
        f = readme_lstm(orc, d)
        tr = f.minimize({}, 0.01, 10, trace=True)
        np.savez_compressed(os.path.join(HERE, f"c3_lstm_d{d}.npz"), trace=np.asarray(tr, np.float32),
                            **{k.replace(":", "_"): v for k, v in leaves_dict(f).items()})
    # dag mode, T=3, square d=6 (square => the reference-compiled gemm itself is used)
    orc.set_mode("dag")
    f = readme_lstm(orc, 6, T=3, seed=77)
    tr = f.minimize({}, 0.01, 10, trace=True)
    np.savez_compressed(os.path.join(HERE, "lstm_dag_d6_T3.npz"), trace=np.asarray(tr, np.float32),
                        **{k.replace(":", "_"): v for k, v in leaves_dict(f).items()})
    orc.set_mode("exact")
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
