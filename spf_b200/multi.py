"""Multi-GPU plumbing for the bootstrapping path.

The path shards by independent units: every circuit bootstrap / keyswitch of a dependency level
shares only the read-only compute key.  So: one process per GPU, the 149 MB compute key is
replicated once (one broadcast from the rank that holds it -- NCCL over NVLink on GPUs, gloo in
the CPU tests), each level's ready batch is split into contiguous index ranges, and there is no
data-path collective.  torch.distributed is plumbing only.
"""
from __future__ import annotations

from typing import Sequence


def shard(batch: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous [start, start+count) range of a batch owned by `rank`: the first batch % world
    ranks get one extra item, so counts differ by at most one and the ranges tile the batch."""
    if world <= 0 or not (0 <= rank < world) or batch < 0:
        raise ValueError("bad shard arguments")
    base, extra = divmod(batch, world)
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def broadcast_compute_key(tensors: Sequence, src: int = 0) -> None:
    """Replicate ComputeKey {bs_key, ks_key, ss_key, auto_key} (crypto/keys.rs:306-318), already
    allocated with identical shapes on every rank, from `src` to all ranks.  In place."""
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    for t in tensors:
        dist.broadcast(t, src=src)


def max_over_ranks(value: float, device=None) -> float:
    """Timing reduction: multi-GPU numbers are the max over ranks of device-measured time."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
