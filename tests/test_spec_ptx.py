"""CPU-only checks of the specialised mode's generated PTX (complogic_b200/csrc/spec.cpp), evaluated by the numpy
PTX interpreter in tests/ptx_interp.py against the oracle VM: the straight-line kernels at every K, and the chunked
chain long programs are cut into (live registers handed over through the scratch buffer)."""
import os

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

import complogic_b200 as host
from complogic_b200 import circuits
from oracle import oracle
from tests import ptx_interp
from tests.helpers import mask_tail, random_sliced


def feedback_dag(levels=1500, width=48, n_imm=64, seed=33):
    """A long thin DAG whose first level also reads the LAST level's registers: in program order those are read
    before they are written, so they carry the previous run's values (state rows, rewritten in place)."""
    dag = circuits.random_dag(levels=levels, width=width, n_immediates=n_imm, window=2, seed=seed)
    ops = dag.sim.ops_array.copy()
    last = n_imm + (levels - 1) * width
    ops["b"][n_imm:n_imm + width] = np.arange(last, last + width)
    return host.Simulation(registers=[False] * len(dag.sim.registers), _ops_array=ops), list(range(last, last + width))


@pytest.mark.parametrize("k", [1, 2, 4])
def test_straight_line_kernel_matches_oracle(k):
    for circ in (circuits.ripple_adder(32), circuits.array_multiplier(16)):
        fp = host.FlatProgram.levelize(circ.sim, circ.n_immediates, circ.out_regs)
        nv, stride = 32 * 8, 8
        imm = random_sliced(circ.n_immediates, stride, seed=7 + k)
        want = oracle.run_batch_sliced(circ.sim, len(circ.sim.registers), imm, nv, circ.out_regs)
        mem = {"imm": imm, "out": np.zeros((len(circ.out_regs), stride), np.uint32), "state": np.zeros((1, stride), np.uint32)}
        ptx_interp.run_kernel(fp.spec_ptx(k), k, stride // k, mem)
        assert np.array_equal(mem["out"], want)
        got = ptx_interp.run_program(fp, k, imm, None, len(circ.out_regs), fp.info["lane_slots"])
        assert np.array_equal(got, want)


@pytest.mark.parametrize("k", [1, 4])
def test_chunked_chain_hands_registers_over(k):
    """Cuts in the middle of a DAG: each kernel's prologue must reload exactly what later code reads."""
    circ = circuits.random_dag(levels=1500, width=48, n_immediates=64, window=2, seed=31)
    fp = host.FlatProgram.levelize(circ.sim, 64, circ.out_regs)
    _, n = fp.spec_chunk_ptx(k, 0)
    assert fp.info["lane_instrs"] > 16384 and n >= -(-fp.info["lane_instrs"] // 16384)  # cuts move to where few registers are live
    nv, stride = 32 * 4, 4
    imm = random_sliced(64, stride, seed=99)
    want = oracle.run_batch_sliced(circ.sim, len(circ.sim.registers), imm, nv, circ.out_regs)
    got = ptx_interp.run_program(fp, k, imm, None, len(circ.out_regs), fp.info["lane_slots"])
    assert np.array_equal(got, want)
    with pytest.raises(host.ClgError):
        fp.spec_chunk_ptx(k, n)


def test_batch_chain_cuts_between_instances_and_in_equal_parts():
    """The second chain (big batches): one instance of a flattened program per kernel where an instance fits the limit
    -- all instances then share ONE kernel and ONE launch and nothing is handed over; an instance longer than the limit
    is cut into equal parts, so the parts of every instance share their kernels as well and run side by side, part by part."""
    for n, m, first, second in (
        (32, 6, {"chunks": 2, "kernels": 2, "launches": 2}, {"chunks": 6, "kernels": 1, "launches": 1, "handed_over": 0}),
        (40, 4, {"chunks": 2, "kernels": 2, "launches": 2}, {"chunks": 4, "kernels": 1, "launches": 1, "handed_over": 0}),
        (16, 16, {"chunks": 1, "kernels": 1, "launches": 1}, {"chunks": 8, "kernels": 1, "launches": 1, "handed_over": 0}),  # (two 784-instruction instances per kernel)
        # (the halves of all instances side by side, position by position: first halves in one launch, second halves in the next)
        (48, 3, {"chunks": 2, "kernels": 2, "launches": 2}, {"chunks": 6, "kernels": 2, "launches": 2}),
        (64, 2, {"chunks": 2, "kernels": 1, "launches": 1}, {"chunks": 4, "kernels": 2, "launches": 2}),
    ):
        circ = circuits.flatten_instances(circuits.array_multiplier(n), m)
        fp = host.FlatProgram.levelize(circ.sim, circ.n_immediates, circ.out_regs)
        assert fp.spec_chain_info(1) == first, (n, m)
        got = fp.spec_batch_chain_info(1)
        if "handed_over" not in second:
            assert got.pop("handed_over") > 0
        assert got == second, (n, m)
    # the generated kernels of such a chain, run by the PTX interpreter: an instance per kernel (mul16, two per kernel),
    # and instances cut in two equal parts that hand ~100 registers over (a 1 000-instruction limit)
    for circ, k, limit in ((circuits.flatten_instances(circuits.array_multiplier(16), 10), 1, None),
                           (circuits.flatten_instances(circuits.random_dag(40, 64, 32, 2, 11), 3), 2, "1000")):
        if limit:
            os.environ["CLG_SPEC_TP_CHUNK"] = limit
        try:
            fp = host.FlatProgram.levelize(circ.sim, circ.n_immediates, circ.out_regs)
            assert fp.spec_chunk_ptx(k, 0, second=True)[1] == fp.spec_batch_chain_info(k)["chunks"] >= 5
            nv, stride = 32 * 4, 4
            imm = random_sliced(circ.n_immediates, stride, seed=5)
            want = oracle.run_batch_sliced(circ.sim, len(circ.sim.registers), imm, nv, circ.out_regs)
            got = ptx_interp.run_program(fp, k, imm, None, len(circ.out_regs), fp.info["lane_slots"], second=True)
            assert np.array_equal(got, want)
        finally:
            os.environ.pop("CLG_SPEC_TP_CHUNK", None)
    # a DAG is cut where few registers are live, as in the first chain, only shorter
    dag = circuits.random_dag(210, 160, 64, 2, 7)
    fp = host.FlatProgram.levelize(dag.sim, 64, dag.out_regs)
    info = fp.spec_batch_chain_info(1)
    assert info["chunks"] == info["kernels"] == info["launches"] == 4 and info["handed_over"] > 300


def test_chunked_chain_stateful_steps():
    sim, outs = feedback_dag()
    fp = host.FlatProgram.levelize(sim, 64, outs, flags=host.LV_STATEFUL)
    assert fp.info["n_state"] == len(outs) and fp.spec_chunk_ptx(1, 0)[1] > 1
    nv, stride, T = 64, 4, 3
    imm = np.stack([random_sliced(64, stride, 40 + t) for t in range(T)])
    want = oracle.run_batch_scalar(sim, len(sim.registers), imm, nv, outs, steps=T, all_steps=True)
    state = np.zeros((fp.info["n_state"], stride), np.uint32)
    for t in range(T):
        got = ptx_interp.run_program(fp, 4, imm[t], state, len(outs), fp.info["lane_slots"])
        assert np.array_equal(mask_tail(got, nv), mask_tail(want[t], nv)), f"step {t}"


def test_cuts_fall_between_flattened_instances_and_kernels_are_shared():
    """Multipliers flattened: every cut lands on an instance boundary, where nothing is live, so no kernel of the
    chain touches the scratch buffer; registers are renumbered and rows made relative per chunk, so all instances
    produce the SAME kernel text (assembled once) and differ only in their row bases."""
    m = 3
    base = circuits.array_multiplier(64)
    circ = circuits.flatten_instances(base, m)
    fp = host.FlatProgram.levelize(circ.sim, circ.n_immediates, circ.out_regs)
    texts = []
    for c in range(m):
        text, n, bases = fp.spec_chunk_ptx(1, c, with_bases=True)
        assert n == m
        assert "%scr;\n  ld.global" not in text and "%scr;\n  st.global" not in text
        assert text.count("lop3.b32") == fp.info["n_luts"] // m
        assert bases == (c * base.n_immediates, c * len(base.out_regs), 0)
        texts.append(text)
    assert texts[0] == texts[1] == texts[2]
    nv, stride = 64, 4
    imm = random_sliced(circ.n_immediates, stride, seed=5)
    want = oracle.run_batch_sliced(circ.sim, len(circ.sim.registers), imm, nv, circ.out_regs)
    got = ptx_interp.run_program(fp, 1, imm, None, len(circ.out_regs), fp.info["lane_slots"])
    assert np.array_equal(mask_tail(got, nv), mask_tail(want, nv))


@st.composite
def programs(draw):
    n_regs = draw(st.integers(1, 12))
    ops = []
    for _ in range(draw(st.integers(0, 60))):
        if draw(st.integers(0, 5)) == 0:
            ops.append(host.Set_op(draw(st.integers(0, n_regs - 1)), draw(st.booleans())))
        else:
            ops.append(host.Nand_op(*[draw(st.integers(0, n_regs - 1)) for _ in range(3)]))
    return ops, n_regs, draw(st.integers(0, n_regs + 1)), draw(st.sampled_from([1, 2, 4])), draw(st.integers(2, 9))


@settings(max_examples=int(os.environ.get("CLG_HYPOTHESIS_EXAMPLES", "150")), deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(programs(), st.integers(0, 2 ** 31))
def test_random_stateful_programs_in_tiny_chunks(prog, seed):
    """Arbitrary programs (multi-writer, read-before-write, carried registers) cut every few instructions
    (CLG_SPEC_CHUNK), two steps: the chain must reproduce the oracle VM register for register."""
    ops, n_regs, n_imm, k, chunk = prog
    sim = host.Simulation(ops=ops, registers=[False] * n_regs)
    nv, stride, steps = 100, 4, 2
    imm = np.stack([random_sliced(n_imm, stride, seed + s) for s in range(steps)]) if n_imm else np.zeros((steps, 0, stride), np.uint32)
    want = oracle.run_batch_scalar(ops, n_regs, imm, nv, list(range(n_regs)), steps=steps, all_steps=True)
    fp = host.FlatProgram.levelize(sim, n_imm, None, flags=host.LV_STATEFUL)
    state = np.zeros((max(1, fp.info["n_state"]), stride), np.uint32)
    old = os.environ.get("CLG_SPEC_CHUNK")
    os.environ["CLG_SPEC_CHUNK"] = str(chunk)
    try:
        for s in range(steps):
            got = ptx_interp.run_program(fp, k, imm[s], state, n_regs, max(1, fp.info["lane_slots"]))
            assert np.array_equal(mask_tail(got, nv), mask_tail(want[s], nv)), f"step {s}"
    finally:
        if old is None:
            del os.environ["CLG_SPEC_CHUNK"]
        else:
            os.environ["CLG_SPEC_CHUNK"] = old


def test_chain_shape_kernels_and_launches():
    """What a host can ask before the first run: how many kernels will be assembled, how many launches a run takes."""
    mul = circuits.flatten_instances(circuits.array_multiplier(64), 5)
    fp = host.FlatProgram.levelize(mul.sim, mul.n_immediates, mul.out_regs)
    assert fp.spec_chain_info(1) == {"chunks": 5, "kernels": 1, "launches": 1}  # equal, independent: one kernel, one launch
    dag = circuits.random_dag(levels=1500, width=48, n_immediates=64, window=2, seed=31)
    fp = host.FlatProgram.levelize(dag.sim, 64, dag.out_regs)
    info = fp.spec_chain_info(1)
    assert info["chunks"] == info["kernels"] == info["launches"] >= 3  # cut mid-circuit: all different, run in sequence
    small = circuits.ripple_adder(32)
    fp = host.FlatProgram.levelize(small.sim, 65, small.out_regs)
    assert fp.spec_chain_info(4) == {"chunks": 1, "kernels": 1, "launches": 1}
    # two independent latch banks that rewrite the SAME carried registers cannot share a launch -- but these do not:
    sim, outs = feedback_dag(levels=40, width=16)
    fp = host.FlatProgram.levelize(sim, 64, outs, flags=host.LV_STATEFUL)
    assert fp.spec_chain_info(1)["chunks"] == 1
    with pytest.raises(host.ClgError):
        fp.spec_chain_info(3)


def test_specialised_text_needs_a_lane_stream():
    c = circuits.ripple_adder(8)
    fp = host.FlatProgram.levelize(c.sim, c.n_immediates, c.out_regs, flags=host.LV_NO_LANE_STREAM)
    assert fp.info["lane_instrs"] == 0
    for call in (lambda: fp.spec_ptx(4), lambda: fp.spec_chunk_ptx(1, 0), lambda: fp.spec_chain_info(1)):
        with pytest.raises(host.ClgError):
            call()
