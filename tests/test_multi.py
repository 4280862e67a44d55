"""The N > 1 host path (key replication + batch sharding) on CPU with the gloo backend,
world_size 2 (the GPU path uses the same code with NCCL; bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest


def test_shard_tiles_the_batch():
    from spf_b200.multi import shard

    for batch in (0, 1, 7, 8, 4096, 65536, 4097):
        for world in (1, 2, 4, 8):
            ranges = [shard(batch, world, r) for r in range(world)]
            assert ranges[0][0] == 0
            for (s0, c0), (s1, _) in zip(ranges, ranges[1:]):
                assert s0 + c0 == s1
            assert ranges[-1][0] + ranges[-1][1] == batch
            counts = [c for _, c in ranges]
            assert max(counts) - min(counts) <= 1
    with pytest.raises(ValueError):
        shard(8, 2, 2)


def _worker(rank, world, port, out_q):
    import torch
    import torch.distributed as dist

    from spf_b200.multi import broadcast_compute_key, max_over_ranks, shard

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)
        ref = [rng.standard_normal(4096), rng.integers(0, 1 << 62, 1000), rng.standard_normal(64), rng.standard_normal(128)]
        if rank == 0:
            key = [torch.from_numpy(np.array(a)) for a in ref]
        else:
            key = [torch.zeros(len(a), dtype=torch.float64 if a.dtype == np.float64 else torch.int64) for a in ref]
        broadcast_compute_key(key, src=0)
        same = all(np.array_equal(k.numpy(), a) for k, a in zip(key, ref))
        start, count = shard(11, world, rank)
        slow = max_over_ranks(1.0 + rank)
        out_q.put((rank, same, start, count, slow))
    finally:
        dist.destroy_process_group()


def test_key_broadcast_and_sharding_world2():
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0] == (0, True, 0, 6, 2.0)
    assert res[1] == (1, True, 6, 5, 2.0)


def test_cbs_chunks_cover_the_level():
    from spf_b200.multi import cbs_chunk

    for n in (0, 1, 7, 64, 96, 148, 149, 4096):
        for world in (1, 2, 3, 4, 8):
            ch = cbs_chunk(n, world)
            assert ch * world >= n and (ch - 1) * world < n or n == 0
            covered = sum(min(max(n - r * ch, 0), ch) for r in range(world))
            assert covered == n


def _exchange_worker(rank, world, port, out_q):
    import torch
    import torch.distributed as dist

    from spf_b200.multi import all_gather_chunks

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        chunk = 4096 + 64
        full = (torch.arange(world * chunk, dtype=torch.int64) * 2654435761 % 251).to(torch.uint8)
        buf = torch.full((world * chunk,), 0xEE, dtype=torch.uint8)   # garbage everywhere ...
        buf[rank * chunk:(rank + 1) * chunk] = full[rank * chunk:(rank + 1) * chunk]  # ... but my own chunk
        all_gather_chunks(buf, chunk, world, rank)
        out_q.put((rank, bool(torch.equal(buf, full))))
    finally:
        dist.destroy_process_group()


def test_level_exchange_world2_gloo():
    """The per-level GGSW all-gather of a sharded graph run on CPU tensors."""
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_exchange_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True)]


class _FakeGraph:
    """Stands in for a sharded CompiledGraph: records what open_peer_arenas hands to open_peers."""

    def __init__(self, rank, world):
        self.rank, self.world, self.opened = rank, world, None

    def ipc_handle(self):
        return bytes([self.rank]) * 64

    def open_peers(self, handles):
        self.opened = list(handles)


def _peer_worker(rank, world, port, out_q):
    import torch.distributed as dist

    from spf_b200.multi import open_peer_arenas

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = _FakeGraph(rank, world)
        open_peer_arenas(g)
        out_q.put((rank, g.opened == [bytes([r]) * 64 for r in range(world)]))
    finally:
        dist.destroy_process_group()


def test_peer_arena_handles_are_gathered_world2_gloo():
    """Host plumbing of the peer-memory exchange: every rank ends up with every rank's 64-byte IPC handle, in
    rank order (the CUDA side -- cudaIpcOpenMemHandle, P2P stores, flag barriers -- is covered by the GPU tests)."""
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_peer_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True)]
