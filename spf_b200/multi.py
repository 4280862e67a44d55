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


def cbs_chunk(n: int, world: int) -> int:
    """Items per rank of a sharded CircuitBootstrap level: `world` equal chunks, the last ones
    possibly short or empty (graph.cuh::cbs_chunk_items)."""
    if n < 0 or world <= 0:
        raise ValueError("bad chunk arguments")
    return (n + world - 1) // world


class _DevMem:
    """Device memory [ptr, ptr+nbytes) exposed through __cuda_array_interface__ so that torch can
    wrap it without copying."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3,
                                         "strides": None}


def all_gather_chunks(buf, chunk_bytes: int, world: int, rank: int, group=None) -> None:
    """The level exchange on a torch uint8 tensor `buf` of world * chunk_bytes: on entry chunk
    `rank` is valid, on return every chunk is valid on every rank.  Works on CUDA tensors (NCCL over
    NVLink / NVSwitch) and on CPU tensors (gloo; used by the CPU tests)."""
    import torch.distributed as dist

    if world == 1:
        return
    assert buf.numel() == world * chunk_bytes
    mine = buf[rank * chunk_bytes:(rank + 1) * chunk_bytes].clone()  # NCCL forbids aliasing input and output
    if buf.is_cuda:
        dist.all_gather_into_tensor(buf, mine, group=group)
    else:
        parts = [buf[r * chunk_bytes:(r + 1) * chunk_bytes] for r in range(world)]
        tmp = [p.clone() for p in parts]
        dist.all_gather(tmp, mine, group=group)
        for p, t in zip(parts, tmp):
            p.copy_(t)


class NcclExchange:
    """exchange callable for spf_b200.CompiledGraph(world > 1): all-gathers one CircuitBootstrap
    level's GGSWs (256 KiB each) in place in the graph's device arena, stream-ordered on the
    executor's stream (no host synchronisation)."""

    def __init__(self, rank: int, group=None):
        self.rank, self.group = rank, group
        self.calls, self.bytes = 0, 0

    def __call__(self, d_buf: int, chunk_bytes: int, world: int, stream: int) -> None:
        import torch

        buf = torch.as_tensor(_DevMem(d_buf, world * chunk_bytes), device=torch.device("cuda", torch.cuda.current_device()))
        with torch.cuda.stream(torch.cuda.ExternalStream(stream)):
            all_gather_chunks(buf, chunk_bytes, world, self.rank, self.group)
        self.calls += 1
        self.bytes += world * chunk_bytes


def open_peer_arenas(graph, group=None) -> None:
    """Peer-memory exchange for a sharded CompiledGraph (one process per GPU): all-gathers the CUDA IPC handles
    of the ranks' arenas over torch.distributed (host side, once per graph) and maps them; afterwards graph.run()
    needs no exchange callable -- GGSWs are stored into every rank's arena by the scheme-switch kernel itself
    (NVLink P2P) and levels are separated by flag barriers through peer memory."""
    import torch.distributed as dist

    handles = [None] * graph.world
    dist.all_gather_object(handles, graph.ipc_handle(), group=group)
    graph.open_peers(handles)
    dist.barrier(group=group)  # nobody runs before every rank has mapped every arena
