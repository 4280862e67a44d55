"""Run `programs` multi-level encrypted programs -- add then greater-than, or (kind = mul) BASELINE
config 4's multiply then greater-than -- as ONE graph sharded over the GPUs of a box
and report latency.
launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P tools/sharded_graph_run.py [width] [programs] [runs] [add|mul] [nccl|peer]
exchange: nccl = all-gathers between levels (NcclExchange); peer = P2P stores from the producing kernels into every
rank's arena + flag barriers (spf_b200_graph_open_peers)."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O  # key / input generation and decryption only
import spf_b200
from spf_b200.circuits import add_then_greater_than, multiply_then_greater_than
from spf_b200.multi import NcclExchange, broadcast_compute_key, open_peer_arenas

w = int(sys.argv[1]) if len(sys.argv) > 1 else 32
programs = int(sys.argv[2]) if len(sys.argv) > 2 else 8
runs = int(sys.argv[3]) if len(sys.argv) > 3 else 3
kind = sys.argv[4] if len(sys.argv) > 4 else "add"
xmode = sys.argv[5] if len(sys.argv) > 5 else "nccl"
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

# every rank derives the same keys and inputs from the harness seeds (in production: one broadcast)
keys = O.Keys()
client = O.Client(keys)
kt = [torch.from_numpy(np.ascontiguousarray(a).view(np.float64 if a.dtype == np.complex128 else np.int64)).cuda()
      for a in (keys.bsk_fft, keys.ksk, keys.ssk_fft, keys.ak_fft)]
t0 = time.perf_counter()
broadcast_compute_key(kt, src=0)
torch.cuda.synchronize()
bcast_ms = 1e3 * (time.perf_counter() - t0)
ev = spf_b200.Evaluation(*[t.data_ptr() for t in kt], device=local, on_device=True)

rng = np.random.default_rng(2024)
vals = [(int(rng.integers(0, 1 << w)), int(rng.integers(0, 1 << w)), int(rng.integers(0, 1 << w))) for _ in range(programs)]
# every ciphertext buffer is a row of ONE page-locked slab (spf_b200_host_alloc): graph IO is plain DMA
slab = spf_b200.pinned_zeros((programs * (4 * w + 1), keys.glwe_len))
rows = iter(slab)


def enc(v):
    out = []
    for i in range(w):
        r = next(rows)
        r[:] = client.encrypt_glwe_l1([(v >> i) & 1])
        out.append(r)
    return out


a, b, c = ([enc(v[k]) for v in vals] for k in range(3))
out_sum = [[next(rows) for _ in range(w)] for _ in range(programs)]
out_gt = [next(rows) for _ in range(programs)]
t0 = time.perf_counter()
circ = (multiply_then_greater_than if kind == "mul" else add_then_greater_than)(a, b, c, out_sum, out_gt, programs)
build_ms = 1e3 * (time.perf_counter() - t0)
ex = NcclExchange(rank) if world > 1 and xmode == "nccl" else None
t0 = time.perf_counter()
g = spf_b200.CompiledGraph(ev, circ, world=world, rank=rank, exchange=ex)
compile_ms = 1e3 * (time.perf_counter() - t0)
if world > 1 and xmode == "peer":
    open_peer_arenas(g)
times = []
for _ in range(runs + 1):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    g.run()
    times.append(1e3 * (time.perf_counter() - t0))
# every output is checked on the rank that holds it (a MUX tree's outputs live on the tree's owner)
out_nodes = [i for i, nd in enumerate(circ.nodes) if nd[0] == spf_b200.OP["OutputGlwe1"]]
assert len(out_nodes) == programs * (w + 1)
ok, checked = True, 0
for p_, (x, y, z) in enumerate(vals):
    want = ((x * y) if kind == "mul" else (x + y)) % (1 << w)
    bufs = out_sum[p_] + [out_gt[p_]]
    wants = [(want >> i) & 1 for i in range(w)] + [int(want > z)]
    for node, buf, bit in zip(out_nodes[p_ * (w + 1):(p_ + 1) * (w + 1)], bufs, wants):
        if g.output_rank(node) in (-1, rank):
            ok &= int(client.decrypt_glwe_l1(buf)[0]) == bit
            checked += 1
t = torch.tensor([min(times[1:])], dtype=torch.float64, device="cuda")
oks = torch.tensor([int(ok)], device="cuda")
n_checked = torch.tensor([checked], device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(oks, op=dist.ReduceOp.MIN)
    dist.all_reduce(n_checked, op=dist.ReduceOp.SUM)
if rank == 0:
    n_op = lambda name: sum(1 for n in circ.nodes if n[0] == spf_b200.OP[name])
    print(json.dumps({"workload": f"{programs} x ({kind}{w} then greater-than) in one graph", "n_gpus": world, "exchange": xmode if world > 1 else None,
                      "nodes": len(circ.nodes), "cmux": n_op("CMux"), "circuit_bootstraps": n_op("CircuitBootstrap"),
                      "host_graph_build_ms": build_ms, "compile_ms": compile_ms, "levels": g.levels, "launches": g.launches,
                      "graph_ms_max_over_ranks": float(t.item()), "correct_on_all_ranks": bool(oks.item()), "outputs_checked_over_ranks": int(n_checked.item()), "outputs": len(out_nodes),
                      "key_broadcast_ms": bcast_ms, "exchanges_per_run": (ex.calls // (runs + 1)) if ex else 0,
                      "exchange_bytes_per_run": (ex.bytes // (runs + 1)) if ex else 0}))
if world > 1:
    dist.destroy_process_group()
