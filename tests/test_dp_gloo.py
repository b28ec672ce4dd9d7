"""World-size-2 check (gloo, CPU) of the data-parallel contract of SURVEY §8e, with the oracle as the arithmetic:
rows of inputs / target / biases / h0 / C0 are sharded, weights replicated; after one iteration the SUM over ranks
of the per-rank weight updates equals the full-batch update, and every row-local parameter equals the matching rows
of the full-batch run. This is exactly what the CUDA path's NCCL sum-allreduce implements."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from computational_graph_b200.dp import ROW_LOCAL, row_block, shard_lstm_problem
    from computational_graph_b200.lstm import PARAM_NAMES, lstm_loss, lstm_params
    from oracle.oracle import get_oracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = get_oracle()
    orc.use_port()
    orc.set_mode("dag")
    B, I, H, T, lr = 8, 5, 6, 3, 0.01
    rng = np.random.default_rng(123)
    params = lstm_params(I, H, B, rng, scale=0.5)
    xs = [rng.uniform(-.1, .1, (B, I)).astype(np.float32) for _ in range(T)]
    tgt = rng.uniform(-.1, .1, (B, H)).astype(np.float32)

    def run(p, x, t):
        f = lstm_loss(orc, [orc.Constant(a) for a in x], orc.Constant(t), p)
        before = [np.asarray(v, dtype=np.float64) for _, v in f.leaves()]
        names = [n for n, _ in f.leaves()]
        f.minimize({}, lr, 1)
        after = [np.asarray(v, dtype=np.float64) for _, v in f.leaves()]
        return names, before, after

    names, fb, fa = run(params, xs, tgt)                       # full batch (every rank computes it redundantly)
    p_r, x_r, t_r = shard_lstm_problem(params, xs, tgt, rank, world)
    names_r, sb, sa = run(p_r, x_r, t_r)                       # this rank's shard
    assert names == names_r
    lo, hi = row_block(B, rank, world)
    worst_w, worst_r = 0.0, 0.0
    for n, b_full, a_full, b_s, a_s in zip(names, fb, fa, sb, sa):
        if n in ROW_LOCAL:
            worst_r = max(worst_r, float(np.max(np.abs(a_s - a_full[lo:hi]))))
        else:
            delta = torch.from_numpy(b_s - a_s)                # lr * local gradient
            dist.all_reduce(delta, op=dist.ReduceOp.SUM)       # the exchange step
            full = b_full - a_full
            scale = max(float(np.max(np.abs(full))), 1e-12)
            # an update recovered from two f32 parameter values carries +-ulp(p) of rounding per rank
            noise = (world + 1) * np.finfo(np.float32).eps * float(np.max(np.abs(b_full)))
            err = max(float(np.max(np.abs(delta.numpy() - full))) - noise, 0.0)
            worst_w = max(worst_w, err / scale)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), np.array([worst_w, worst_r]))
    dist.destroy_process_group()


def test_sum_allreduce_of_shard_gradients_equals_full_batch(tmp_path):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        worst_w, worst_r = np.load(tmp_path / f"r{r}.npy")
        assert worst_w < 1e-5, worst_w     # weight updates: sum over ranks == full batch (f32 summation order only)
        assert worst_r < 1e-6, worst_r     # row-local parameters: identical rows, no communication


def test_row_block_partition():
    from computational_graph_b200.dp import row_block, shard_rows
    blocks = [row_block(8192, r, 8) for r in range(8)]
    assert blocks[0] == (0, 1024) and blocks[-1] == (7168, 8192)
    assert all(blocks[i][1] == blocks[i + 1][0] for i in range(7))
    with pytest.raises(ValueError):
        row_block(10, 0, 4)
    a = np.arange(12, dtype=np.float32).reshape(6, 2)
    assert np.array_equal(np.concatenate([shard_rows(a, r, 3) for r in range(3)]), a)
