"""BASELINE config 1, the reference's examples/basic_add: an encrypted 8-bit add (2 + 7) through the graph
executor on cuda:0.  Client side (key generation, encryption, decryption) is the CPU oracle -- the reference does
that part on the CPU too; everything between the ciphertexts is spf_b200.

    python examples/basic_add.py [a] [b] [width]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O        # client side only
import spf_b200
from spf_b200.circuits import InstructionCache

a = int(sys.argv[1]) if len(sys.argv) > 1 else 2
b = int(sys.argv[2]) if len(sys.argv) > 2 else 7
w = int(sys.argv[3]) if len(sys.argv) > 3 else 8

keys = O.Keys()                                   # SecretKey + ComputeKey at DEFAULT_128 (parasol_runtime/src/params.rs:107-134)
client = O.Client(keys)
ev = spf_b200.Evaluation(keys.bsk_fft, keys.ksk, keys.ssk_fft, keys.ak_fft, device=0)   # Evaluation::new

# ciphertext buffers: rows of one page-locked slab (2w inputs, w + 1 outputs), one L1 GLWE per bit
slab = spf_b200.pinned_zeros((3 * w + 1, keys.glwe_len))
a_bits, b_bits, out_bits = list(slab[:w]), list(slab[w:2 * w]), list(slab[2 * w:])
for i in range(w):
    a_bits[i][:] = client.encrypt_glwe_l1([(a >> i) & 1])
    b_bits[i][:] = client.encrypt_glwe_l1([(b >> i) & 1])

isa = InstructionCache(ev)                        # one compiled graph per (instruction, width)
isa.add(a_bits, b_bits, out_bits)                 # first call builds and compiles the graph
t0 = time.perf_counter()
g = isa.add(a_bits, b_bits, out_bits)             # cached: re-binds the buffers and runs
ms = 1e3 * (time.perf_counter() - t0)
result = sum(int(client.decrypt_glwe_l1(o)[0]) << i for i, o in enumerate(out_bits))
print(f"{a} + {b} = {result} ({w}-bit encrypted add, {g.levels} dependency levels, {g.launches} kernel launches, {ms:.2f} ms)")
assert result == a + b
ev.close()
