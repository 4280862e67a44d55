"""Time pbs_kernel alone on random key material (no oracle, no keygen: for A/B and ablation builds whose results are not
checked).  usage: SPF_B200_LIB=... python tools/pbs_time.py [batches=444,4096] [reps=5]  -> one line 'name b:ms ...'"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C

import spf_b200

batches = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "444,4096").split(",")]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
p = spf_b200.default_128()
l = spf_b200.lib()
rng = np.random.default_rng(1)
mk = lambda n: (rng.standard_normal(2 * n) * 2.0 ** 40).view(np.complex128)
bsk = mk(l.spf_b200_len_bsk(C.byref(p)))
ssk = mk(l.spf_b200_len_ssk(C.byref(p)))
ak = mk(l.spf_b200_len_ak(C.byref(p)))
ksk = np.zeros(l.spf_b200_len_ksk(C.byref(p)), dtype=np.uint64)
ev = spf_b200.Evaluation(bsk, ksk, ssk, ak)
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
lut = torch.from_numpy(rng.integers(0, 1 << 63, 4096, dtype=np.int64)).to(dev)
out = []
for B in batches:
    d_in = torch.from_numpy(rng.integers(-(1 << 63), 1 << 63, (B, 638), dtype=np.int64)).to(dev)
    d_glwe = torch.empty(B * 4096, dtype=torch.int64, device=dev)
    fn = lambda: ev.dev_programmable_bootstrap(d_glwe.data_ptr(), d_in.data_ptr(), lut.data_ptr(), 0, 2, B, stream=stream.cuda_stream)
    fn(); fn()
    ms = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    out.append(f"{B}:{min(ms):.3f}")
print(os.path.basename(os.environ.get("SPF_B200_LIB", "default")), " ".join(out), flush=True)
ev.close()
