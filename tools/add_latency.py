"""Run the encrypted w-bit add graph a few times on cuda:0 (for ncu launch lists / latency work).
usage: python tools/add_latency.py [width] [runs]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O  # input generation / decryption only
import spf_b200
from spf_b200.circuits import ripple_carry_adder

w = int(sys.argv[1]) if len(sys.argv) > 1 else 8
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 3
keys = O.Keys()
client = O.Client(keys)
ev = spf_b200.Evaluation(keys.bsk_fft, keys.ksk, keys.ssk_fft, keys.ak_fft)
a, b = 0xDEADBEEF & ((1 << w) - 1), 0x12345679 & ((1 << w) - 1)
ab = [client.encrypt_glwe_l1([(a >> i) & 1]) for i in range(w)]
bb = [client.encrypt_glwe_l1([(b >> i) & 1]) for i in range(w)]
outs = [np.zeros(keys.glwe_len, dtype=np.uint64) for _ in range(w + 1)]
g = spf_b200.CircuitProcessor(ev).compile(ripple_carry_adder(ab, bb, outs))
for _ in range(runs):
    t0 = time.perf_counter()
    g.run()
    print(f"add{w}: {1e3 * (time.perf_counter() - t0):.2f} ms, levels {g.levels}, launches {g.launches}")
got = sum(int(client.decrypt_glwe_l1(o)[0]) << i for i, o in enumerate(outs))
print("correct:", got == a + b)
