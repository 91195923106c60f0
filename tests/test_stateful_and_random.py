"""Carried-register (stateful) programs and randomly generated op lists, flat buffer vs oracle, on the CPU.

The reference never clears `registers` between run() calls (simulation.rs:16-30), which is what makes its
latches work (gates.rs:510-571, compile.rs:268-294). In stateful mode the flat program carries exactly the
registers a run can read before writing them; here every lane replays a multi-step immediates sequence and
every register after every step must equal the oracle VM's.
"""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

import complogic_b200 as host
from oracle import oracle
from tests import flat_interp
from tests.helpers import mask_tail, random_sliced, stride_for


def run_steps_flat(fp, imm_steps, n_vectors, stream, seed=0):
    flat = flat_interp.Flat(fp.to_bytes())
    stride = imm_steps.shape[2]
    state = np.zeros((flat.n_state, stride), dtype=np.uint32)  # fresh Simulation: all registers false
    outs = []
    for s in range(imm_steps.shape[0]):
        outs.append(mask_tail(flat_interp.run(flat, stream, imm_steps[s], state, rng=np.random.default_rng(seed + s)),
                              n_vectors))
    return np.stack(outs)


@pytest.fixture(autouse=True)
def _force_row_stream(monkeypatch):
    monkeypatch.setenv("CLG_ROWS_FORCE", "1")


def check_program(ops, n_regs, n_imm, steps, n_vectors, seed, flags=0):
    sim = host.Simulation(ops=ops, registers=[False] * n_regs)
    stride = stride_for(n_vectors)
    imm = np.stack([random_sliced(n_imm, stride, seed + 97 * s) for s in range(steps)]) if n_imm else \
        np.zeros((steps, 0, stride), np.uint32)
    all_regs = list(range(n_regs))
    want = oracle.run_batch_scalar(ops, n_regs, imm, n_vectors, all_regs, steps=steps, all_steps=True)
    want_sliced = oracle.run_batch_sliced(ops, n_regs, imm, n_vectors, all_regs, steps=steps, all_steps=True)
    assert np.array_equal(mask_tail(want, n_vectors), mask_tail(want_sliced, n_vectors))
    fp = host.FlatProgram.levelize(sim, n_imm, None, flags=host.LV_STATEFUL | flags)
    streams = [flat_interp.STREAM_LEVEL, flat_interp.STREAM_LANE]
    if fp.info["rows_instrs"]:  # (CLG_ROWS_FORCE, set for this module: the row stream of the cooperative row kernel)
        streams.append(flat_interp.STREAM_ROWS)
    for stream in streams:
        got = run_steps_flat(fp, imm, n_vectors, stream, seed)
        assert np.array_equal(got, mask_tail(want, n_vectors)), f"stream {stream}"
    if steps == 1:  # stateless build must agree on a fresh simulation
        fp0 = host.FlatProgram.levelize(sim, n_imm, None, flags=flags)
        assert fp0.info["n_state"] == 0
        flat0 = flat_interp.Flat(fp0.to_bytes())
        for stream in (0, 1):
            got = mask_tail(flat_interp.run(flat0, stream, imm[0]), n_vectors)
            assert np.array_equal(got, mask_tail(want[0], n_vectors))
    return fp


@pytest.mark.parametrize("gate", ["RSLatch", "RSLatchTest", "DLatch"])
def test_latches_carry_state_per_lane(gate):
    c = host.Compiler(2)
    g = getattr(host, gate)(0, 1, c.alloc())
    sim = c.compile([g])
    fp = check_program(sim.ops, len(sim.registers), 2, steps=6, n_vectors=200, seed=5)
    assert fp.info["n_state"] > 0  # a latch reads registers before it writes them


def test_rs_latch_truth_sequence_in_every_lane():
    """rs_nor_latch (gates.rs:510-535): FF->0, TF->1, FT->0, TT->0, here with all lanes doing the same thing."""
    c = host.Compiler(2)
    rs = host.RSLatch(0, 1, c.alloc())
    sim = c.compile([rs])
    fp = host.FlatProgram.levelize(sim, 2, [rs.q], flags=host.LV_STATEFUL)
    flat = flat_interp.Flat(fp.to_bytes())
    ones, zeros = np.full(4, 0xFFFFFFFF, np.uint32), np.zeros(4, np.uint32)
    for stream in (0, 1):
        state = np.zeros((flat.n_state, 4), np.uint32)
        seq = []
        for s, r in ((0, 0), (1, 0), (0, 1), (1, 1)):
            imm = np.stack([ones if s else zeros, ones if r else zeros])
            seq.append(int(flat_interp.run(flat, stream, imm, state)[0, 0]))
        assert seq == [0, 0xFFFFFFFF, 0, 0]


def test_acyclic_program_has_no_state():
    c = host.Compiler(3)
    sim = c.compile([host.FullAdder(0, 1, 2, c.alloc(), c.alloc())])
    fp = host.FlatProgram.levelize(sim, 3, None, flags=host.LV_STATEFUL)
    assert fp.info["n_state"] == 0


@st.composite
def programs(draw):
    n_regs = draw(st.integers(1, 12))
    n_ops = draw(st.integers(0, 40))
    ops = []
    for _ in range(n_ops):
        if draw(st.integers(0, 4)) == 0:
            ops.append(host.Set_op(draw(st.integers(0, n_regs - 1)), draw(st.booleans())))
        else:
            ops.append(host.Nand_op(*[draw(st.integers(0, n_regs - 1)) for _ in range(3)]))
    n_imm = draw(st.integers(0, n_regs + 2))  # shorter, equal or longer than the registers a Set can name
    steps = draw(st.integers(1, 3))
    return ops, n_regs, n_imm, steps


@settings(max_examples=int(__import__('os').environ.get('CLG_HYPOTHESIS_EXAMPLES', '150')), deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(programs(), st.integers(0, 2 ** 31))
def test_random_programs_match_the_vm(prog, seed):
    """Hand-built style programs: registers written more than once, read before written, Sets anywhere with
    literals, immediates slice shorter or longer than the Sets need -- every register, every step, bit exact."""
    ops, n_regs, n_imm, steps = prog
    check_program(ops, n_regs, n_imm, steps, n_vectors=70, seed=seed)


@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(programs(), st.integers(0, 2 ** 31))
def test_random_programs_unfused(prog, seed):
    ops, n_regs, n_imm, steps = prog
    check_program(ops, n_regs, n_imm, steps, n_vectors=40, seed=seed, flags=host.LV_NO_FUSE)
