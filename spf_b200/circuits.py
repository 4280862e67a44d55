"""Test/bench helper circuits.  The reference derives its MUX trees from BDDs
(mux_circuits/src/add.rs:13-56, biodivine-lib-bdd) above the drop-in boundary; that generator is out
of scope here, so this module hand-builds a functionally equivalent
ripple-carry adder MUX tree with the same front end as FheCircuit::insert_mux_circuit_and_connect_
inputs (fhe_circuit.rs:473-494): per input bit InputGlwe1 -> SampleExtract(0) -> KeyswitchL1toL0 ->
CircuitBootstrap, then CMux / Not / ZeroGlwe1 / OneGlwe1 nodes and one OutputGlwe1 per result bit.
It is NOT node-for-node the reference's graph (the BDD may share sub-terms differently); its
decryptions are identical."""
from __future__ import annotations

import numpy as np

from . import FheCircuit


def ripple_carry_adder(a_bits: list[np.ndarray], b_bits: list[np.ndarray], out_bits: list[np.ndarray]) -> FheCircuit:
    """a_bits/b_bits: width L1 GLWE ciphertexts (bit in coefficient 0); out_bits: width+1 output
    buffers (sum bits then carry).  Depth: 2 CMUX levels per bit (notes/leveled_computation.md:34)."""
    w = len(a_bits)
    assert len(b_bits) == w and len(out_bits) == w + 1
    c = FheCircuit()

    def front(ct):
        x = c.add("InputGlwe1", io=ct)
        x = c.add("SampleExtract", x, arg=0)
        x = c.add("KeyswitchL1toL0", x)
        return c.add("CircuitBootstrap", x)

    sa = [front(x) for x in a_bits]
    sb = [front(x) for x in b_bits]
    zero, one = c.add("ZeroGlwe1"), c.add("OneGlwe1")
    carry, ncarry = zero, one
    for i in range(w):
        # sum = a ? (b ? c : !c) : (b ? !c : c)
        s0 = c.add("CMux", sb[i], carry, ncarry)   # a = 0: b ? !c : c
        s1 = c.add("CMux", sb[i], ncarry, carry)   # a = 1: b ? c : !c
        s = c.add("CMux", sa[i], s0, s1)
        c.add("OutputGlwe1", s, io=out_bits[i])
        # carry' = a ? (b ? 1 : c) : (b ? c : 0)
        c0 = c.add("CMux", sb[i], zero, carry)
        c1 = c.add("CMux", sb[i], carry, one)
        carry = c.add("CMux", sa[i], c0, c1)
        if i + 1 < w:
            ncarry = c.add("Not", carry)
    c.add("OutputGlwe1", carry, io=out_bits[w])
    return c
