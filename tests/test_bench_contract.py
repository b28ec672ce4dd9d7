"""The driver-facing contract of bench.py that can be checked without a GPU: the reference arm (`--impl reference`)
times the reference's own CPU path (oracle/_ref when built, else the port) on a bounded sample and prints ONE JSON
line with the keys the driver reads; the product arm refuses to run without a CUDA device."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--no-same-config"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "LSTM fwd+backprop+SGD iterations/sec"
    assert d["unit"] == "iterations/s" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["steps"] == 1 and d["warmup"] == 0  # the arm honours the driver's K and W
    assert d["config"]["workload"].startswith("BASELINE config 5") and d["scaling"] == "strong"
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] == 1 and cb["value"] == d["value"] and cb["sample"]
    # ms_per_step is what one step of THIS arm took on the clock (a pass over the bounded sample); value scales it to the
    # whole workload by the sample's flop fraction, and the line says so
    assert abs(d["ms_per_step"] - 1e3 * cb["sample_seconds_per_pass"]) < 1e-6
    assert abs(d["value"] - cb["sample_flop_fraction"] / cb["sample_seconds_per_pass"]) <= 1e-9 * d["value"]
    assert "EXTRAPOLATED" in cb["sample"] and d["step_definition"]
    assert d["e2e"] == {"value": d["value"], "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], env=env,
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("this is the CPU tier's check")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "CUDA" in (r.stderr + r.stdout)
