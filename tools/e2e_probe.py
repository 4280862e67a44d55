"""Where the end-to-end step (LWE in -> CBS -> CMUX -> GLWE out through the graph API) spends its time next to the
device-resident step: single blocking runs (with SPF_B200_GRAPH_TIMING=1 the executor prints per-op device times),
then streamed runs of n steps to separate the per-step cost from the exposed first H2D / last D2H.

  python tools/e2e_probe.py [--batch 4096]
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import spf_b200  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--mode", default="double")
    ap.add_argument("--depth", type=int, default=3)
    args = ap.parse_args()
    import oracle as O

    rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo")
    tag = "[rank %d] " % rank
    keys = O.Keys()
    p = spf_b200.default_128()
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    t = [torch.from_numpy(a.view(np.float64 if a.dtype.kind == "c" else np.int64)).to(dev)
         for a in (keys.bsk_fft, keys.ksk, keys.ssk_fft, keys.ak_fft)]
    ev = spf_b200.Evaluation(*[x.data_ptr() for x in t], params=p, device=rank, on_device=True)
    B = args.batch
    bits = np.random.default_rng(1).integers(0, 2, B)
    cts = bench.encrypt_lwe0_numpy(keys.lwe0_sk.view(np.uint64), bits, p.lwe_std, 7)
    h_in = torch.from_numpy(cts.view(np.int64)).pin_memory()
    h_lwe = h_in.numpy().view(np.uint64).reshape(B, p.lwe_n + 1)
    d_in = h_in.to(dev)
    d_out = torch.empty(B * ev.len_ggsw * 2, dtype=torch.float64, device=dev)
    s = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(s)
    for _ in range(3):
        ev.dev_circuit_bootstrap(d_out.data_ptr(), d_in.data_ptr(), B, reference_scale=False, stream=s.cuda_stream)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        ev.dev_circuit_bootstrap(d_out.data_ptr(), d_in.data_ptr(), B, reference_scale=False, stream=s.cuda_stream)
    torch.cuda.synchronize()
    print(tag + "device-resident step (no L2 flush): %.2f ms" % ((time.perf_counter() - t0) / 5 * 1e3))

    # raw copy rates of this rank while all ranks copy at the same time
    h_big = spf_b200.pinned_zeros((B, ev.len_glwe))
    d_big = torch.empty(B * ev.len_glwe, dtype=torch.int64, device=dev)
    h_t = torch.from_numpy(h_big.view(np.int64).reshape(-1))
    if world > 1:
        dist.barrier()
    for name, fn in (("D2H", lambda: h_t.copy_(d_big, non_blocking=True)), ("H2D", lambda: d_big.copy_(h_t, non_blocking=True))):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 10
        print(tag + "%s of %d MB: %.2f ms (%.1f GB/s)" % (name, h_t.numel() * 8 >> 20, dt * 1e3, h_t.numel() * 8 / dt / 1e9))
    del d_big
    if world > 1:
        dist.barrier()
    pipe = bench.E2EGraphPipeline(ev, h_lwe, ev.len_glwe, 148 * 3, args.mode, args.depth)
    pipe.step()
    g = pipe.graphs[0]
    for _ in range(1):
        t0 = time.perf_counter()
        g.run()
        print(tag + "blocking run of one graph (%d inputs): %.2f ms" % (B if args.mode == "double" else -1, (time.perf_counter() - t0) * 1e3))
    for _ in range(3):
        t0 = time.perf_counter()
        g.spawn()
        t1 = time.perf_counter()
        g.wait()
        print(tag + "spawn %.2f ms, spawn + wait %.2f ms" % ((t1 - t0) * 1e3, (time.perf_counter() - t0) * 1e3))
    for n in (8, 16, 16):
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        pipe.run(n)
        dt = time.perf_counter() - t0
        print(tag + "streamed %2d steps: %.2f ms per step (%.0f CBS/s)" % (n, dt / n * 1e3, B * n / dt))
    pipe.close()


if __name__ == "__main__":
    main()
