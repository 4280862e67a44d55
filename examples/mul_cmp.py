"""BASELINE config 4 on one GPU: an encrypted 32-bit multiply (low word) followed by greater-than, the ISA's `Mul` then
`CmpGt`, expanded from BDD-derived MUX circuits (16x16 multiplier blocks + grade-school reduction + comparison) into one
graph of ~46 k nodes with three circuit-bootstrap levels.  Client side (keys, encryption, decryption) is the CPU oracle.

    python examples/mul_cmp.py [a] [b] [c]        # checks (a * b mod 2^32, a * b mod 2^32 > c)
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O        # client side only
import spf_b200
from spf_b200.circuits import multiply_then_greater_than

w = 32
a = int(sys.argv[1]) if len(sys.argv) > 1 else 0xDEADBEEF
b = int(sys.argv[2]) if len(sys.argv) > 2 else 0x12345679
c = int(sys.argv[3]) if len(sys.argv) > 3 else 0x40000000

keys = O.Keys()
client = O.Client(keys)
ev = spf_b200.Evaluation(keys.bsk_fft, keys.ksk, keys.ssk_fft, keys.ak_fft, device=0)

slab = spf_b200.pinned_zeros((4 * w + 1, keys.glwe_len))       # 3w input bits, w product bits, 1 comparison bit
rows = iter(slab)


def encrypt(v):
    out = [next(rows) for _ in range(w)]
    for i, r in enumerate(out):
        r[:] = client.encrypt_glwe_l1([(v >> i) & 1])
    return out


ea, eb, ec = encrypt(a), encrypt(b), encrypt(c)
prod_bits, gt_bit = [next(rows) for _ in range(w)], next(rows)
t0 = time.perf_counter()
circuit = multiply_then_greater_than([ea], [eb], [ec], [prod_bits], [gt_bit], 1)
t1 = time.perf_counter()
graph = spf_b200.CircuitProcessor(ev).compile(circuit)
t2 = time.perf_counter()
graph.run()                                                     # warm-up
t3 = time.perf_counter()
graph.run()
t4 = time.perf_counter()
prod = sum(int(client.decrypt_glwe_l1(o)[0]) << i for i, o in enumerate(prod_bits))
gt = int(client.decrypt_glwe_l1(gt_bit)[0])
print(f"{a:#x} * {b:#x} mod 2^32 = {prod:#x}; > {c:#x}: {bool(gt)}")
print(f"{len(circuit)} nodes, {graph.levels} levels, {graph.launches} launches; MUX generation + expansion {1e3 * (t1 - t0):.0f} ms, "
      f"compile {1e3 * (t2 - t1):.0f} ms, run {1e3 * (t4 - t3):.2f} ms")
assert prod == (a * b) % (1 << w) and gt == int(prod > c)
ev.close()
