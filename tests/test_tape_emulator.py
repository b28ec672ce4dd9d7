"""CPU-tier functional check of the code GENERATOR behind the compiled resident executor (csrc/ew_jit.cpp tape_jit): the
CUDA source it emits for a plan is compiled with g++ behind tests/tape_emu/shim.h and executed on the CPU (512 threads,
std::barrier, an array for the shared memory), then compared with the oracle — trace of root values and every leaf,
bit for bit (same libm, same operation order). tests/tape_emu_probe.py does the work in a process of its own."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emulated():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tape_emu_probe.py")], capture_output=True, text=True,
                       timeout=900)
    assert r.returncode == 0, (r.stdout[-1000:], r.stderr[-3000:])
    return json.loads(r.stdout.strip().splitlines()[-1])


@pytest.mark.parametrize("case", ["c1", "broadcast_avg", "c3_exact_d5", "c3_dag_d6", "c3_exact_d17", "c3_dag_d30",
                                  "lstm_T3_nonsquare", "random_8", "random_29", "random_49", "random_50", "random_53",
                                  "random_59"])
def test_generated_resident_kernel_equals_the_oracle_bit_for_bit(emulated, case):
    r = emulated[case]
    assert r["bit_identical"], r
    assert r["worst_rel_err"] == 0.0 and r["smem_bytes"] > 0 and r["leaves"] >= 1


@pytest.mark.parametrize("case", ["c3_exact_d5", "c3_dag_d6", "lstm_T3_nonsquare", "broadcast_avg", "random_59"])
def test_generated_resident_kernel_has_no_data_race(emulated, case):
    """The generator schedules steps by dependency level and puts a __syncthreads() only between levels; under
    ThreadSanitizer (the CTA's 512 threads are OS threads in the emulator) a missing barrier is a reported data race."""
    assert emulated[case]["races"] == 0, emulated[case]
