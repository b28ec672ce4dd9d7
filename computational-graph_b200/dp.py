"""Data-parallel host logic for the batch-sharded LSTM. The reference has no distributed code; this
is the one new exchange step the path gets: every rank owns a contiguous block of batch rows of the inputs, the
target and the row-shaped parameters (biases, h0, C0 — full [batch, H] matrices in the reference, src/lstm.rs:63-75),
weights are replicated, and each weight gradient X^T.E is SUMMED over the ranks before its update — by the library's
fused all-reduce + update kernel over NVLink peer memory, or by ncclAllReduce (csrc/dp.cpp) — (root error == 1,
src/optimize.rs:69: gradients add, they are not averaged)."""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np

# parameters of src/lstm.rs:60-76 whose rows are batch rows (everything that is not a weight matrix)
ROW_LOCAL = ("bi", "bc", "bf", "bo", "C", "h")


def row_block(n_rows: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous row block [lo, hi) of rank `rank`; rows must divide evenly (the reference has no ragged batch)."""
    if n_rows % world:
        raise ValueError(f"batch of {n_rows} rows does not divide over {world} ranks")
    per = n_rows // world
    return rank * per, (rank + 1) * per


def shard_rows(a: np.ndarray, rank: int, world: int) -> np.ndarray:
    lo, hi = row_block(a.shape[0], rank, world)
    return np.ascontiguousarray(a[lo:hi])


def shard_lstm_problem(params: Dict[str, np.ndarray], inputs, target, rank: int, world: int):
    """Per-rank view of a global LSTM problem: row-local tensors are sliced, weights are shared."""
    p = {k: (shard_rows(v, rank, world) if k in ROW_LOCAL else v) for k, v in params.items()}
    xs = [shard_rows(x, rank, world) for x in inputs]
    return p, xs, shard_rows(target, rank, world)


def init_from_torch_distributed(api, device_index: int) -> None:
    """Create the library's NCCL communicator from an initialised torch.distributed process group: rank 0 makes the
    NCCL unique id, the group broadcasts its 128 bytes, every rank joins (include/cg_graph.h cg_dp_*)."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        uid = torch.frombuffer(bytearray(api.dp_unique_id()), dtype=torch.uint8).clone()
    if dist.get_backend() == "nccl":
        uid = uid.cuda(device_index)
    dist.broadcast(uid, 0)
    api.dp_init(rank, world, uid.cpu().numpy().tobytes())
