"""CMUX kernel choice: device time of one batched CMUX launch for several batch sizes (one selector per op).
Run once per setting of SPF_B200_CMUX_WIDE_MAX (the largest batch served by the 8-team wide kernel).
usage: python tools/cmux_sweep.py 150,200,296,444,592,740,888,1184,2000"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O  # key / input generation only
import spf_b200

sizes = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "150,296,444,592,740,1184,2000").split(",")]
keys = O.Keys()
client = O.Client(keys)
ev = spf_b200.Evaluation(keys.bsk_fft, keys.ksk, keys.ssk_fft, keys.ak_fft)
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)  # torch events see only torch's current stream: launch on it
s = stream.cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
rng = np.random.default_rng(3)
B = max(sizes)
one = client.encrypt_ggsw_l1(1)
sel = torch.from_numpy(np.ascontiguousarray(one).view(np.float64)).to(dev)
d_sel = torch.empty(B * ev.len_ggsw * 2, dtype=torch.float64, device=dev)
ev.dev_fft_rescale(d_sel.data_ptr(), sel.data_ptr(), ev.len_ggsw, to_device=True, stream=s)
d_sel.view(B, -1)[1:] = d_sel.view(B, -1)[0]
rand = lambda: torch.from_numpy(rng.integers(0, 1 << 63, (B, ev.len_glwe), dtype=np.int64)).to(dev)
a, b, out = rand(), rand(), rand()
for n in sizes:
    ms = []
    for rep in range(4):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ev.dev_cmux(out.data_ptr(), d_sel.data_ptr(), ev.len_ggsw, a.data_ptr(), b.data_ptr(), n, stream=s)
        e1.record()
        torch.cuda.synchronize()
        if rep:
            ms.append(e0.elapsed_time(e1))
    print(json.dumps({"wide_max": os.environ.get("SPF_B200_CMUX_WIDE_MAX", "sm_count"), "batch": n, "us": 1e3 * min(ms)}))
