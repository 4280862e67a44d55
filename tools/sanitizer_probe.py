"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): DEFAULT_128 rings with a short LWE
dimension (16 blind-rotation steps instead of 637) through every kernel: PBS (quad + pair teams), trace /
scheme switch, CMUX (wide + bulk), keyswitch (tensor-core + IMAD), sample extract, graph executor.
usage: compute-sanitizer --tool racecheck python tools/sanitizer_probe.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O
import spf_b200

op = O.default_128()
op.lwe_n = 16
keys = O.Keys(op, seed=7)
client = O.Client(keys)
p = spf_b200.default_128()
p.lwe_n = 16
ev = spf_b200.Evaluation(keys.bsk_fft, keys.ksk, keys.ssk_fft, keys.ak_fft, params=p)
nbig = int(os.environ.get("PROBE_BIG", "160"))
bits = np.random.default_rng(1).integers(0, 2, nbig).tolist()
cts = client.encrypt_lwe_l0_batch(bits)
small = ev.circuit_bootstrap(cts[:3])            # quad PBS + trace_ss
big = ev.circuit_bootstrap(cts)                  # pair-team PBS (2 per SM) + trace_ss
ok = [client.decrypt_ggsw_l1(g) for g in small] == bits[:3] and all(
    client.decrypt_ggsw_l1(big[i]) == bits[i] for i in (0, 1, nbig // 2, nbig - 1))
a, b = client.encrypt_glwe_l1([0, 1]), client.encrypt_glwe_l1([1, 1])
w = ev.cmux(small, np.stack([a] * 3), np.stack([b] * 3))                       # wide CMUX
bulk = ev.cmux(big, np.stack([a] * nbig), np.stack([b] * nbig))                # bulk CMUX
ok &= all(int(client.decrypt_glwe_l1(w[i])[0]) == bits[i] for i in range(3))
ok &= all(int(client.decrypt_glwe_l1(bulk[i])[0]) == bits[i] for i in (0, nbig - 1))
l1 = ev.sample_extract_l1(bulk, 0)
l0 = ev.keyswitch_lwe_l1_lwe_l0(l1)                                            # tensor-core keyswitch
ok &= np.array_equal(l0[5], O.keyswitch_lwe(keys, l1[5]))
ok &= [client.decrypt_lwe_l0(x) for x in l0[:8]] == bits[:8]
os.environ["SPF_B200_KS_NO_TC"] = "1"
l0b = ev.keyswitch_lwe_l1_lwe_l0(l1[:20])                                      # IMAD keyswitch
ok &= np.array_equal(l0b, l0[:20])
print("sanitizer probe:", "ok" if ok else "MISMATCH")
ev.close()
sys.exit(0 if ok else 1)
