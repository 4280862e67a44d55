"""Debug: GLEV-mode MUX circuit pieces on the GPU with intermediate outputs."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O
import spf_b200
from spf_b200 import mux_circuits as M

keys = O.Keys(); client = O.Client(keys)
ev = spf_b200.Evaluation(keys.bsk_fft, keys.ksk, keys.ssk_fft, keys.ak_fft)
proc = spf_b200.CircuitProcessor(ev)
p = keys.params
def glev_bits(glev):
    g = glev.reshape(p.cbs.count, keys.glwe_len)
    return [int(O.decode(client.decrypt_glwe_l1_raw(g[j]), (j + 1) * p.cbs.radix_log)[0]) for j in range(p.cbs.count)]
for bit_a, bit_b in ((1, 1), (1, 0), (0, 1)):
    c = spf_b200.FheCircuit()
    front = lambda v: M.insert_ciphertext_conversion(c, c.add("InputGlwe1", io=client.encrypt_glwe_l1([v])), "glwe", "ggsw")
    sa, sb = front(bit_a), front(bit_b)
    inner = c.add("GlevCMux", sb, c.add("ZeroGlev1"), c.add("OneGlev1"))
    outer = c.add("GlevCMux", sa, c.add("ZeroGlev1"), inner)
    o_inner, o_outer = np.zeros(keys.glev_len, np.uint64), np.zeros(keys.glev_len, np.uint64)
    o_one = np.zeros(keys.glev_len, np.uint64)
    c.add("OutputGlev1", inner, io=o_inner); c.add("OutputGlev1", outer, io=o_outer); c.add("OutputGlev1", c.add("OneGlev1"), io=o_one)
    ss = c.add("SchemeSwitch", outer)
    o_ggsw = np.zeros(keys.ggsw_fft_len, np.complex128)
    c.add("OutputGgsw1", ss, io=o_ggsw)
    fin = c.add("CMux", ss, c.add("ZeroGlwe1"), c.add("OneGlwe1"))
    o_fin = np.zeros(keys.glwe_len, np.uint64)
    c.add("OutputGlwe1", fin, io=o_fin)
    proc.run_graph_blocking(c)
    print(bit_a, bit_b, "one", glev_bits(o_one), "inner", glev_bits(o_inner), "outer", glev_bits(o_outer), "ggsw", client.decrypt_ggsw_l1(o_ggsw),
          "final", int(client.decrypt_glwe_l1(o_fin)[0]))

print("---- the failing shape: AND (glev mode) -> SchemeSwitch -> OR (glwe mode)")
for a, b, d in ((0b11, 0b01, 0b10), (0b10, 0b01, 0b00)):
    c = spf_b200.FheCircuit()
    front = lambda v: [M.insert_ciphertext_conversion(c, c.add("InputGlwe1", io=client.encrypt_glwe_l1([(v >> i) & 1])), "glwe", "ggsw") for i in range(2)]
    sa, sb, sd = front(a), front(b), front(d)
    and_glev = M.insert_mux_circuit(c, M.make_and_circuit(2), [sa[0], sb[0], sa[1], sb[1]], mux_mode="glev")
    o_and = [np.zeros(keys.glev_len, np.uint64) for _ in range(2)]
    for n, buf in zip(and_glev, o_and):
        c.add("OutputGlev1", n, io=buf)
    and_sel = [M.insert_ciphertext_conversion(c, n, "glev", "ggsw") for n in and_glev]
    o_sel = [np.zeros(keys.ggsw_fft_len, np.complex128) for _ in range(2)]
    for n, buf in zip(and_sel, o_sel):
        c.add("OutputGgsw1", n, io=buf)
    or_out = M.insert_mux_circuit(c, M.make_or_circuit(2), [and_sel[0], sd[0], and_sel[1], sd[1]])
    outs = [np.zeros(keys.glwe_len, np.uint64) for _ in range(2)]
    for n, buf in zip(or_out, outs):
        c.add("OutputGlwe1", n, io=buf)
    print([spf_b200.OPS[n[0]] for n in c.nodes][-14:])
    proc.run_graph_blocking(c)
    print(a, b, d, "and glev", [glev_bits(x) for x in o_and], "sel", [client.decrypt_ggsw_l1(x) for x in o_sel],
          "or", [int(client.decrypt_glwe_l1(o)[0]) for o in outs], "want", [((a & b) | d) >> i & 1 for i in range(2)])
