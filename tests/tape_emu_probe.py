"""Executes the CUDA source the product GENERATES for the compiled resident executor (csrc/ew_jit.cpp tape_jit) on the
CPU and compares it with the oracle — without a GPU. Plan-only mode dumps, next to the source, the kernel's pointer
table with the initial contents of every persistent buffer (CG_TAPE_JIT_DUMP); tests/tape_emu compiles the source with
g++ behind a shim (512 OS threads for the CTA, std::barrier for __syncthreads(), an array for the shared memory) and
runs it. Same libm on both sides, same operation order: the results must be BIT-IDENTICAL to the oracle's.
Prints one JSON line; run in a process of its own (tests/test_tape_emulator.py drives it)."""
import ctypes as C
import json
import os
import struct
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
EMU = os.path.join(ROOT, "tests", "tape_emu")


def read_table(path):
    raw = open(path, "rb").read()
    count, root_off, root_n, root_is_leaf, smem = struct.unpack_from("<5Q", raw, 0)
    pos, entries = 40, []
    for _ in range(count):
        addr, n, off, written = struct.unpack_from("<4Q", raw, pos)
        pos += 32 + 4 * n
        entries.append((addr, n))
    return root_n, entries


def emulate(src, iters, workdir, sanitize=False):
    """Compiles the generated source behind the shim and runs it. sanitize=True builds with ThreadSanitizer and returns
    the number of data-race reports instead: the CTA's threads are real threads here, so a barrier the generator failed
    to place between two dependent steps is a reported race (checked: removing one __syncthreads() is caught)."""
    exe = os.path.join(workdir, "emu_tsan" if sanitize else "emu")
    subprocess.run(["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-pthread", "-Wno-unknown-pragmas", "-I", EMU,
                    "-include", os.path.join(EMU, "shim.h"), f'-DTAPE_SOURCE="{src}"', os.path.join(EMU, "glue.cpp"),
                    os.path.join(EMU, "driver.cpp"), "-o", exe] + (["-g", "-fsanitize=thread"] if sanitize else []), check=True)
    out = os.path.join(workdir, "out.bin")
    if sanitize:
        r = subprocess.run([exe, src + ".table", str(iters), out], capture_output=True, text=True, timeout=900)
        return r.stderr.count("WARNING: ThreadSanitizer")
    subprocess.run([exe, src + ".table", str(iters), out], check=True, timeout=600)
    root_n, entries = read_table(src + ".table")
    data = np.fromfile(out, dtype=np.float32)
    trace = data[:iters * root_n].reshape(iters, root_n)
    pos, finals = iters * root_n, {}
    for addr, n in entries:
        finals[addr] = data[pos:pos + n]
        pos += n
    return trace, finals


def cases(api):
    """name -> (builder(api) -> Function, mode, lr, iters)"""
    from helpers import readme_lstm

    def broadcast_avg(a):
        F = a.Function
        x = F.scalar_param("x", 0.5)
        m = F.param("m", a.Constant(np.arange(12, dtype=np.float32).reshape(3, 4) / 7 - 0.6))
        b = F.matrix(3, 4, list(np.linspace(-1, 1.5, 12)))
        return (((x - F.scalar(2.)) * (m + b) - F.scalar(10.)).sq() + (x * m).tanh()).abs()

    from plan_fuzz import random_graph
    fuzz = {}
    for g in (8, 29, 49, 50, 53, 59):  # seeds whose graphs fold no constants at construction and take >= 3 kernels
        fuzz[f"random_{g}"] = (lambda a, g=g: random_graph(a, np.random.default_rng(5000 + g), 1 + g % 5),
                               "exact" if g % 2 else "dag", 0.01, 3)
    return {
        **fuzz,
        "c3_exact_d5": (lambda a: readme_lstm(a, 5), "exact", 0.01, 6),
        "c3_dag_d6": (lambda a: readme_lstm(a, 6), "dag", 0.01, 6),
        "c3_exact_d17": (lambda a: readme_lstm(a, 17), "exact", 0.01, 3),   # odd size: blocked product edges, vec-4 tails
        "c3_dag_d30": (lambda a: readme_lstm(a, 30), "dag", 0.01, 2),
        "lstm_T3_nonsquare": (lambda a: readme_lstm(a, 8, T=3, seed=41, batch=12, d_in=9), "dag", 0.01, 4),
        "broadcast_avg": (broadcast_avg, "exact", 0.004, 8),
        "c1": (lambda a: (a.Function.scalar_param("x", 1.0) + a.Function.scalar(3.0)).sq(), "exact", 0.01, 50),
    }


def main():
    import computational_graph_b200 as pkg
    from oracle.oracle import get_oracle
    lib = pkg.load_library()
    assert lib.cg_set_plan_only(1) == 0
    api = pkg.get_api()
    lib.cg_leaf_address.restype, lib.cg_leaf_address.argtypes = C.c_ulonglong, [C.c_void_p, C.c_int]
    orc = get_oracle()
    orc.use_port()
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, (build, mode, lr, iters) in cases(api).items():
            api.set_mode(mode)
            orc.set_mode(mode)
            f = build(api)
            src = os.path.join(tmp, name + ".cu")
            os.environ["CG_TAPE_JIT_DUMP"] = src
            st = api.plan_stats(f, lr)
            del os.environ["CG_TAPE_JIT_DUMP"]
            assert st["vm_resident"] == 2, (name, st)
            n_leaves = lib.cg_num_leaves(f._h)
            addrs = [lib.cg_leaf_address(f._h, i) for i in range(n_leaves)]
            trace, finals = emulate(src, iters, tmp)
            fo = build(orc)
            to = np.asarray(fo.minimize({}, lr, iters, trace=True), dtype=np.float32)
            d = fo.dims()
            to = to.transpose(0, 2, 1).reshape(iters, -1) if d else to.reshape(iters, 1)  # storage (column-major) order
            leaves_o = fo.leaves()
            assert len(leaves_o) == n_leaves
            worst, exact = 0.0, bool(np.array_equal(trace.view(np.int32), to.view(np.int32)))
            worst = max(worst, float(np.max(np.abs(trace.astype(np.float64) - to)) / max(np.max(np.abs(to)), 1e-30)))
            for (lname, lv), addr in zip(leaves_o, addrs):
                want = np.asarray(lv, dtype=np.float32)
                want = want.T.reshape(-1) if want.ndim == 2 else want.reshape(1)
                got = finals[addr]
                exact = exact and bool(np.array_equal(got.view(np.int32), want.view(np.int32)))
                worst = max(worst, float(np.max(np.abs(got.astype(np.float64) - want)) / max(np.max(np.abs(want)), 1e-30)))
            out[name] = {"bit_identical": exact, "worst_rel_err": worst, "kernels": st["kernels_per_iter"],
                         "smem_bytes": st["tape_smem_bytes"], "leaves": n_leaves, "iters": iters}
            if name in ("c3_exact_d5", "c3_dag_d6", "lstm_T3_nonsquare", "broadcast_avg", "random_59"):
                out[name]["races"] = emulate(src, 2, tmp, sanitize=True)
    orc.set_mode("exact")
    print(json.dumps(out))


if __name__ == "__main__":
    main()
