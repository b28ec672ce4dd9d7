"""Multi-GPU parity of the data-parallel path (needs >= 2 GPUs; skipped otherwise): tests/dp_check.py under
torchrun, NCCL over NVLink, against the full-batch CPU oracle."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("exchange", ["fused", "nccl"])
@pytest.mark.parametrize("tf32,shape", [(0, (64, 48, 40)), (1, (512, 256, 256))])
def test_dp_two_ranks_match_full_batch_oracle(tf32, shape, exchange):
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    b, h, i = shape
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(ROOT, "tests", "dp_check.py"), "--tf32", str(tf32), "--batch", str(b),
           "--hidden", str(h), "--inp", str(i), "--steps", "2", "--iters", "3", "--exchange", exchange]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "OK" in r.stdout
