"""Time the two kernels of a circuit bootstrap separately on random key material (A/B builds).
usage: SPF_B200_LIB=... python tools/cbs_time.py [batch=4096] [reps=4]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spf_b200

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
p = spf_b200.default_128()
l = spf_b200.lib()
rng = np.random.default_rng(1)
mk = lambda n: (rng.standard_normal(2 * n) * 2.0 ** 40).view(np.complex128)
ev = spf_b200.Evaluation(mk(l.spf_b200_len_bsk(C.byref(p))), np.zeros(l.spf_b200_len_ksk(C.byref(p)), dtype=np.uint64),
                         mk(l.spf_b200_len_ssk(C.byref(p))), mk(l.spf_b200_len_ak(C.byref(p))))
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
d_in = torch.from_numpy(rng.integers(-(1 << 63), 1 << 63, (B, 638), dtype=np.int64)).to(dev)
d_out = torch.empty(B * ev.len_ggsw * 2, dtype=torch.float64, device=dev)
d_glwe = torch.empty(B * 4096, dtype=torch.int64, device=dev)
lut = torch.from_numpy(rng.integers(0, 1 << 63, 4096, dtype=np.int64)).to(dev)
res = {}
for name, fn in (("cbs", lambda: ev.dev_circuit_bootstrap(d_out.data_ptr(), d_in.data_ptr(), B, reference_scale=False, stream=stream.cuda_stream)),
                 ("pbs", lambda: ev.dev_programmable_bootstrap(d_glwe.data_ptr(), d_in.data_ptr(), lut.data_ptr(), 0, 2, B, stream=stream.cuda_stream))):
    fn(); fn()
    ms = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    res[name] = min(ms)
print(os.path.basename(os.environ.get("SPF_B200_LIB", "default")), f"batch {B}: cbs {res['cbs']:.3f} ms  pbs {res['pbs']:.3f} ms  trace+ss {res['cbs'] - res['pbs']:.3f} ms", flush=True)
ev.close()
