"""The asynchronous executor and ciphertext handles: CircuitProcessor::spawn_graph + CompletionHandler
+ flow control (parasol_runtime/src/circuit_processor/mod.rs:130-253,573-623, completion_handler.rs:14-56) and task outputs
that stay on the device between graphs (circuit_processor/task.rs:10-16; parasol_cpu issues one graph per instruction,
parasol_cpu/src/proc/ops/add.rs:13-80)."""
import threading
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture()
def proc(evaluation):
    import spf_b200

    return spf_b200.CircuitProcessor(evaluation)


def _enc_bits(client, v, w):
    return [client.encrypt_glwe_l1([(v >> i) & 1]) for i in range(w)]


def test_two_dependent_graphs_without_host_copies(keys, client, evaluation, proc):
    """Instruction 1 (add, w = 6) writes its sum bits into DEVICE ciphertext handles; instruction 2 (greater-than) reads
    them from there: between the two graphs nothing crosses PCIe (checked on the library's own copy counters: the
    handles are device memory, so a host copy of them is impossible by construction) and the second graph is ordered
    behind the first on the device (`after`), not by the host."""
    import spf_b200
    from spf_b200 import DeviceCiphertext, FheCircuit
    from spf_b200.circuits import _adder_nodes, _front, _greater_than_node

    w, a, b, cmp = 6, 37, 21, 40
    glwe_bytes = keys.glwe_len * 8
    s_handles = [DeviceCiphertext.alloc(evaluation, glwe_bytes) for _ in range(w)]
    c1 = FheCircuit()
    sa = [_front(c1, x) for x in _enc_bits(client, a, w)]
    sb = [_front(c1, x) for x in _enc_bits(client, b, w)]
    for node, h in zip(_adder_nodes(c1, sa, sb)[:w], s_handles):
        c1.add("OutputGlwe1", node, io=h)
    out = np.zeros(keys.glwe_len, dtype=np.uint64)
    c2 = FheCircuit()
    ss = [_front(c2, h) for h in s_handles]            # InputGlwe1 on a device handle
    sc = [_front(c2, x) for x in _enc_bits(client, cmp, w)]
    c2.add("OutputGlwe1", _greater_than_node(c2, ss, sc), io=out)
    events = []
    g1 = proc.spawn_graph(c1, on_completion=lambda e: events.append(("add", e, time.perf_counter())))
    g2 = proc.spawn_graph(c2, on_completion=lambda e: events.append(("cmp", e, time.perf_counter())), after=[g1])
    g2.wait()
    g1.wait()
    assert [n for n, _, _ in events] == ["add", "cmp"] and all(e is None for _, e, _ in events)
    assert int(client.decrypt_glwe_l1(out)[0]) == int(((a + b) % (1 << w)) > cmp)
    # the intermediate really is the sum: read the handles back (test only) and decrypt
    import torch

    from spf_b200.multi import _DevMem

    got = 0
    for i, h in enumerate(s_handles):
        host = torch.as_tensor(_DevMem(h.ptr, glwe_bytes), device="cuda").cpu().numpy().view(np.uint64)
        got |= int(client.decrypt_glwe_l1(host)[0]) << i
    assert got == (a + b) % (1 << w)
    g1.close(); g2.close()


def test_completion_handler_reports_first_error_and_dependents_become_noops(keys, client, evaluation, proc):
    """faults.rs through the asynchronous path: a malformed graph delivers its error through the completion handler,
    not by raising; a run whose dependency failed retires as a no-op with an error."""
    import spf_b200
    from spf_b200 import FheCircuit

    errs = []
    bad = FheCircuit()
    bad.add("SampleExtract", bad.add("ZeroLwe0"), arg=1)     # wrong ciphertext kind
    assert proc.spawn_graph(bad, on_completion=errs.append) is None
    assert len(errs) == 1 and isinstance(errs[0], spf_b200.SpfError) and errs[0].code == -4
    bad = FheCircuit()
    bad.add("Retire")
    proc.spawn_graph(bad, on_completion=errs.append)
    assert len(errs) == 2 and "Retire" in str(errs[1])
    with pytest.raises(spf_b200.SpfError):
        proc.spawn_graph(bad)                                # no handler: the error is raised, as run_graph_blocking does


def test_flow_control_bounds_in_flight_graphs(keys, client, evaluation, proc):
    """At most max_in_flight spawned graphs are between dispatch and completion: with a limit of 2, the third spawn
    returns only after the first completion has fired (mod.rs:146 `flow_control.recv()`)."""
    from spf_b200 import FheCircuit

    evaluation.set_max_in_flight(2)
    try:
        bits = [1, 0, 1, 1, 0, 1, 0, 0]
        graphs, outs, done_at, spawned_at = [], [], {}, {}
        cts = client.encrypt_lwe_l0_batch(bits * 8)
        a, b = client.encrypt_glwe_l1([0]), client.encrypt_glwe_l1([1])
        for k in range(4):
            c = FheCircuit()
            na, nb = c.add("InputGlwe1", io=a), c.add("InputGlwe1", io=b)
            o = [np.zeros(keys.glwe_len, dtype=np.uint64) for _ in range(16)]
            for i in range(16):
                sel = c.add("CircuitBootstrap", c.add("InputLwe0", io=cts[16 * k + i]))
                c.add("OutputGlwe1", c.add("CMux", sel, na, nb), io=o[i])
            graphs.append(proc.compile(c))
            outs.append(o)
        for k, g in enumerate(graphs):
            g.spawn(on_complete=lambda e, k=k: done_at.__setitem__(k, time.perf_counter()))
            spawned_at[k] = time.perf_counter()
        for g in graphs:
            g.wait()
        assert len(done_at) == 4
        assert spawned_at[2] >= done_at[0] - 1e-4, "the third spawn returned before the first graph completed"
        assert spawned_at[1] < done_at[0], "the second spawn should not have waited"
        for k in range(4):
            for i in range(16):
                assert int(client.decrypt_glwe_l1(outs[k][i])[0]) == (bits * 8)[16 * k + i]
        for g in graphs:
            g.close()
    finally:
        evaluation.set_max_in_flight(4)


def test_concurrent_graphs_overlap_and_pageable_set_io(keys, client, evaluation, proc):
    """Independent graphs spawned back to back run on their own streams; set_io with pageable buffers goes through the
    graph's page-locked staging slab (no cudaHostRegister per call) and still delivers the right bytes."""
    from spf_b200 import FheCircuit

    src = client.encrypt_glwe_l1([1, 0, 1])
    c = FheCircuit()
    out0 = np.zeros(keys.glwe_len, dtype=np.uint64)
    x = c.add("InputGlwe1", io=src)
    c.add("OutputGlwe1", c.add("Not", x), io=out0)
    g = proc.compile(c)
    g.run()
    assert client.decrypt_glwe_l1(out0)[:3].tolist() == [0, 0, 1]
    for trial in range(3):
        src2 = client.encrypt_glwe_l1([trial & 1, 1, 0])    # fresh pageable buffers every call
        out2 = np.zeros(keys.glwe_len, dtype=np.uint64)
        g.set_io(0, src2)
        g.set_io(2, out2)
        fired = threading.Event()
        g.spawn(on_complete=lambda e: fired.set())
        g.wait()
        assert fired.is_set()
        assert client.decrypt_glwe_l1(out2)[:3].tolist() == [1 - (trial & 1), 1, 0]   # Not flips coefficient 0 only
    with pytest.raises(Exception):
        g.set_io(0, np.zeros(keys.glwe_len - 1, dtype=np.uint64))   # wrong size: refused by the host mirror
    g.close()


def test_graph_outlives_its_evaluation(keys, client):
    """Evaluation.close() while a CompiledGraph is still referenced (InstructionCache, fixtures): the context is
    reference-counted by its graphs, so closing the graph later is not a use-after-free."""
    import spf_b200
    from spf_b200 import FheCircuit

    ev = spf_b200.Evaluation(keys.bsk_fft, keys.ksk, keys.ssk_fft, keys.ak_fft)
    src = client.encrypt_glwe_l1([1])
    out = np.zeros(keys.glwe_len, dtype=np.uint64)
    c = FheCircuit()
    c.add("OutputGlwe1", c.add("Not", c.add("InputGlwe1", io=src)), io=out)
    g = spf_b200.CircuitProcessor(ev).compile(c)
    ev.close()
    g.run()                      # the graph keeps the context (keys, streams) alive
    assert int(client.decrypt_glwe_l1(out)[0]) == 0
    g.close()                    # last reference: the context is freed here
