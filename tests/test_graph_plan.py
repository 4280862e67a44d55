"""The host-only half of the graph executor (spf_b200_graph_plan): validation as Task::validate
(parasol_runtime/src/circuit_processor/task.rs:24-179, tests/faults.rs), levelisation with bootstrap-stage
alignment, and the ownership partition of a run sharded over several GPUs.  No GPU needed."""
import collections

import numpy as np
import pytest

import spf_b200
from spf_b200 import OP, FheCircuit, SpfError, plan_graph
from spf_b200.circuits import add_then_greater_than, multiply_then_greater_than

LOCAL = {OP[k] for k in ("CMux", "GlevCMux", "MultiplyGgswGlwe", "Not", "GlweAdd", "MulXN", "SampleExtract", "KeyswitchL1toL0")}


def _bufs(programs, w):
    buf = lambda: np.zeros(4096, np.uint64)
    mk = lambda: [[buf() for _ in range(w)] for _ in range(programs)]
    return mk(), mk(), mk(), mk(), [buf() for _ in range(programs)]


def _check_partition(c, level, owner, world):
    for v, (op, arg, ins, io) in enumerate(c.nodes):
        if op not in LOCAL:
            assert owner[v] == -1, (v, op)
            continue
        assert -1 <= owner[v] < world
        for e, src in enumerate(ins):
            if src < 0:
                continue
            assert level[src] < level[v]
            if c.nodes[src][0] in LOCAL:   # data edge inside a tree: same rank
                assert owner[src] == owner[v], (v, src)


def test_levels_and_stage_alignment():
    a, b, c_, out_sum, out_gt = _bufs(1, 8)
    c = add_then_greater_than(a, b, c_, out_sum, out_gt, 1)
    level, owner = plan_graph(c, 1)
    assert (owner == -1).all()
    cbs = [v for v, nd in enumerate(c.nodes) if nd[0] == OP["CircuitBootstrap"]]
    by_level = collections.Counter(int(level[v]) for v in cbs)
    assert sorted(by_level.values()) == [8, 24]   # the refresh of the 8 sum bits bootstraps as ONE batch
    for v, (op, arg, ins, io) in enumerate(c.nodes):
        assert all(level[s] < level[v] for s in ins if s >= 0)


@pytest.mark.parametrize("world", [2, 8])
def test_ownership_of_adder_program(world):
    programs, w = 8, 8
    a, b, c_, out_sum, out_gt = _bufs(programs, w)
    c = add_then_greater_than(a, b, c_, out_sum, out_gt, programs)
    level, owner = plan_graph(c, world)
    _check_partition(c, level, owner, world)
    cmux_per_rank = collections.Counter(int(owner[v]) for v, nd in enumerate(c.nodes) if nd[0] == OP["CMux"])
    assert set(cmux_per_rank) == set(range(world))               # every rank owns trees
    assert max(cmux_per_rank.values()) <= 2 * min(cmux_per_rank.values())


def test_ownership_of_multiply_program_is_balanced():
    programs, w = 2, 32
    a, b, c_, out_p, out_gt = _bufs(programs, w)
    c = multiply_then_greater_than(a, b, c_, out_p, out_gt, programs)
    level, owner = plan_graph(c, 2)
    _check_partition(c, level, owner, 2)
    per_rank = collections.Counter(int(o) for o in owner if o >= 0)
    assert abs(per_rank[0] - per_rank[1]) <= 0.1 * per_rank[0]
    assert level.max() + 1 > 600   # 510-deep 16x16 blocks + reduction + compare


def test_single_tree_has_one_owner():
    a, b, c_, out_sum, out_gt = _bufs(1, 8)
    c = add_then_greater_than(a, b, c_, out_sum, out_gt, 1)
    level, owner = plan_graph(c, 4)
    adder = {int(owner[v]) for v, nd in enumerate(c.nodes) if nd[0] in (OP["CMux"], OP["Not"])}
    assert len(adder) <= 2 and -1 not in adder   # the adder tree and the comparison tree: one rank each


def test_tree_feeding_a_scheme_switch_stays_replicated():
    glev = np.zeros(16384, np.uint64)
    ggsw = np.zeros(16384, np.complex128)
    c = FheCircuit()
    sel = c.add("InputGgsw1", io=ggsw)
    x = c.add("GlevCMux", sel, c.add("InputGlev1", io=glev), c.add("OneGlev1"))
    ss = c.add("SchemeSwitch", x)
    c.add("OutputGgsw1", ss, io=np.zeros(16384, np.complex128))
    level, owner = plan_graph(c, 2)
    assert owner[x] == -1 and owner[ss] == -1


def test_malformed_graphs_are_rejected_without_a_gpu():
    """circuit_processor/tests/faults.rs: wrong ciphertext kind, missing input, illegal sample-extract index; plus a cycle."""
    glwe, lwe0 = np.zeros(4096, np.uint64), np.zeros(638, np.uint64)

    def expect(c, text):
        with pytest.raises(SpfError) as e:
            plan_graph(c)
        assert e.value.code == -4 and text in str(e.value)

    c = FheCircuit()
    c.add("KeyswitchL1toL0", c.add("InputLwe0", io=lwe0))
    expect(c, "wrong ciphertext kind")
    c = FheCircuit()
    c.add("CMux", c.add("ZeroGgsw1"), c.add("ZeroGlwe1"))
    expect(c, "missing ciphertext input")
    c = FheCircuit()
    c.add("SampleExtract", c.add("InputGlwe1", io=glwe), arg=2048)
    expect(c, "illegal sample extract index")
    c = FheCircuit()
    c.add("Not", 1)
    c.add("Not", 0)
    expect(c, "cycle")
    c = FheCircuit()
    c.nodes.append((99, 0, (-1, -1, -1), None))
    expect(c, "unknown op")
    # faults.rs wrong_inputs_none_expected: an edge into ANY zero-input op is an error, Nop included (task.rs:100-118)
    for op in ("Nop", "ZeroGlwe1", "OneGgsw1", "InputGlwe1"):
        c = FheCircuit()
        c.nodes.append((OP[op], 0, (c.add("ZeroLwe0"), -1, -1), glwe if op == "InputGlwe1" else None))
        expect(c, "unexpected extra input edge")
    # faults.rs illegal_retire_op: user graphs never contain Retire (mod.rs:606-611)
    c = FheCircuit()
    c.add("Retire")
    expect(c, "illegal Retire")
    # the host mirror refuses io buffers of the wrong size or element type before they reach the executor
    c = FheCircuit()
    with pytest.raises(SpfError):
        c.add("InputGlwe1", io=np.zeros(4095, np.uint64))
    with pytest.raises(SpfError):
        c.add("OutputLwe0", c.add("ZeroLwe0"), io=np.zeros(638, np.float64))
    with pytest.raises(SpfError):
        c.add("Not", c.add("ZeroGlwe1"), io=glwe)
    with pytest.raises(SpfError):
        plan_graph(FheCircuit(), world=0)


def test_fhe_circuit_node_table_matches_the_c_struct():
    """FheCircuit stores nodes in spf_node's layout; the list-like view, bulk append and the hand-over agree."""
    import ctypes as C

    from spf_b200 import _Node

    assert FheCircuit._DT.itemsize == C.sizeof(_Node) == 32
    assert FheCircuit._DT.fields["io"][1] == _Node.io.offset and FheCircuit._DT.fields["inp"][1] == _Node.inp.offset
    buf = np.zeros(4096, np.uint64)
    c = FheCircuit()
    x = c.add("InputGlwe1", io=buf)
    s = c.add("SampleExtract", x, arg=7)
    blk = c.add_block(np.full(100, OP["Not"], np.uint32), np.stack([np.full(100, x), np.full(100, -1), np.full(100, -1)], axis=1))
    assert blk.tolist() == list(range(2, 102)) and len(c) == len(c.nodes) == 102
    assert c.nodes[1] == (OP["SampleExtract"], 7, (x, -1, -1), None) and c.nodes[-1][0] == OP["Not"]
    assert c.nodes[0][3] is buf
    assert [n[0] for n in c.nodes][:3] == [OP["InputGlwe1"], OP["SampleExtract"], OP["Not"]]
    c.nodes.append((OP["OutputGlwe1"], 0, (blk[-1], -1, -1), buf))
    arr = c._pack()
    assert (arr[0].op, arr[1].arg, arr[1].inp[0], arr[102].inp[0]) == (OP["InputGlwe1"], 7, x, 101)
    assert arr[0].io == arr[102].io == buf.ctypes.data and not arr[1].io
    with pytest.raises(IndexError):
        c.nodes[103]
    with pytest.raises(SpfError):
        c.add("InputGlwe1", io=np.zeros((4, 4), np.uint64)[:, 0])   # not C-contiguous
    level, owner = plan_graph(c)
    assert level[x] == 0 and level[102] == 2
