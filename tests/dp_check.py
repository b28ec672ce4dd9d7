"""Run under torchrun with N ranks (one per GPU): every rank trains its batch-row shard of one global LSTM problem
through the CUDA path with the NCCL sum-allreduce, and the result is compared with the full-batch run of the CPU
oracle (dag mode): replicated weights on every rank, row-local parameters against their rows. Tolerance 1e-5
(f32 products, summation order differs) or 1e-3 with --tf32."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tf32", type=int, default=0)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--hidden", type=int, default=48)
    ap.add_argument("--inp", type=int, default=40)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--iters", type=int, default=5)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    import computational_graph_b200 as pkg
    from computational_graph_b200.dp import ROW_LOCAL, row_block, shard_lstm_problem
    from computational_graph_b200.lstm import lstm_loss, lstm_params
    from helpers import rel_err
    from oracle.oracle import get_oracle

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    api = pkg.get_api()
    api.set_device(local)
    pkg.dp.init_from_torch_distributed(api, local)
    api.set_mode("dag")
    api.set_tf32(bool(a.tf32))

    rng = np.random.default_rng(99)
    params = lstm_params(a.inp, a.hidden, a.batch, rng, scale=1.0 / np.sqrt(a.hidden))
    xs = [rng.uniform(-.1, .1, (a.batch, a.inp)).astype(np.float32) for _ in range(a.steps)]
    tgt = rng.uniform(-.1, .1, (a.batch, a.hidden)).astype(np.float32)
    p_r, x_r, t_r = shard_lstm_problem(params, xs, tgt, rank, world)
    f = lstm_loss(api, [api.Constant(x) for x in x_r], api.Constant(t_r), p_r)
    st = api.plan_stats(f, 0.01)
    tr = f.minimize({}, 0.01, a.iters, trace=True)

    orc = get_oracle()
    orc.use_port()
    orc.set_mode("dag")
    fo = lstm_loss(orc, [orc.Constant(x) for x in xs], orc.Constant(tgt), params)
    to = fo.minimize({}, 0.01, a.iters, trace=True)
    lo, hi = row_block(a.batch, rank, world)
    tol = 1e-3 if a.tf32 else 1e-5
    worst = rel_err(tr, np.asarray(to)[:, lo:hi, :])
    for (n, vg), (no, vo) in zip(f.leaves(), fo.leaves()):
        assert n == no
        want = vo[lo:hi] if n in ROW_LOCAL else vo
        worst = max(worst, rel_err(vg, want))
    ok = torch.tensor([1.0 if worst < tol else 0.0, worst], device="cuda")
    dist.all_reduce(ok[:1], op=dist.ReduceOp.MIN)
    dist.all_reduce(ok[1:], op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"dp_check world={world} tf32={a.tf32} worst_rel_err={float(ok[1]):.3e} tol={tol} "
              f"kernels/iter={st['kernels_per_iter']} -> {'OK' if float(ok[0]) == 1.0 else 'FAIL'}")
    api.dp_shutdown()
    dist.destroy_process_group()
    sys.exit(0 if float(ok[0]) == 1.0 else 1)


if __name__ == "__main__":
    main()
