"""BASELINE config 5: bootstrap throughput sweep over batch sizes on this rank's GPU (run under torchrun
for N GPUs: every rank runs the same batch, throughput is the whole-job sum / max-over-ranks time).
usage: python tools/throughput_sweep.py [--batches 1,4,16,...] [--pbs-only] [--cpu]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O  # keygen / checker / CPU baseline only
import spf_b200
from bench import cbs_lut, cpu_cbs_rate, encrypt_lwe0_numpy

ap = argparse.ArgumentParser()
ap.add_argument("--batches", default="1,4,16,64,256,1024,4096,16384,65536")
ap.add_argument("--cpu", action="store_true", help="also time the CPU port on rank 0 (bounded sample)")
args = ap.parse_args()
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
keys = O.Keys()
ev = spf_b200.Evaluation(keys.bsk_fft, keys.ksk, keys.ssk_fft, keys.ak_fft, device=local)
p = keys.params
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
lut = torch.from_numpy(cbs_lut().view(np.int64)).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
rows = []
for B in [int(x) for x in args.batches.split(",")]:
    bits = np.random.default_rng(7 + rank).integers(0, 2, B)
    cts = encrypt_lwe0_numpy(keys.lwe0_sk, bits, p.lwe_std, 7 + rank)
    d_in = torch.from_numpy(cts.view(np.int64)).to(dev)
    d_out = torch.empty(B * ev.len_ggsw * 2, dtype=torch.float64, device=dev)
    d_glwe = torch.empty(B * ev.len_glwe, dtype=torch.int64, device=dev)
    res = {}
    for name, fn in (("cbs", lambda: ev.dev_circuit_bootstrap(d_out.data_ptr(), d_in.data_ptr(), B, reference_scale=False, stream=stream.cuda_stream)),
                     ("pbs", lambda: ev.dev_programmable_bootstrap(d_glwe.data_ptr(), d_in.data_ptr(), lut.data_ptr(), 0, 2, B, stream=stream.cuda_stream))):
        fn()
        reps = 3 if B <= 4096 else 2
        ms = []
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        t = torch.tensor([min(ms)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[name] = float(t.item())
    row = {"batch_per_gpu": B, "n_gpus": world, "cbs_ms": res["cbs"], "pbs_ms": res["pbs"],
           "cbs_per_s": world * B / res["cbs"] * 1e3, "pbs_per_s": world * B / res["pbs"] * 1e3}
    if rank == 0:
        # checker: the first and last output of the batch decrypt to the input bits
        tmp = torch.empty(ev.len_ggsw * 2, dtype=torch.float64, device=dev)
        client = O.Client(keys)
        ok = True
        for i in {0, B - 1}:
            ev.dev_fft_rescale(tmp.data_ptr(), d_out.data_ptr() + i * ev.len_ggsw * 16, ev.len_ggsw, to_device=False, stream=stream.cuda_stream)
            torch.cuda.synchronize()
            ok &= client.decrypt_ggsw_l1(tmp.cpu().numpy().view(np.complex128)) == int(bits[i])
        row["decrypt_ok"] = bool(ok)
        rows.append(row)
        print(json.dumps(row), flush=True)
    del d_in, d_out, d_glwe
if rank == 0 and args.cpu:
    nt = O.hw_threads()
    bits = np.random.default_rng(7).integers(0, 2, 8 * nt)
    cts = encrypt_lwe0_numpy(keys.lwe0_sk, bits, p.lwe_std, 7)
    cpu_cbs_rate(keys, cts[:nt], nt)
    rate, dt = cpu_cbs_rate(keys, cts, nt)
    t0 = time.perf_counter(); O.circuit_bootstrap(keys, cts[0]); one = time.perf_counter() - t0
    print(json.dumps({"cpu_port": {"cbs_per_s_all_threads": rate, "threads": nt, "sample": len(cts),
                                   "single_thread_ms_per_cbs": 1e3 * one}}), flush=True)
if world > 1:
    dist.destroy_process_group()
