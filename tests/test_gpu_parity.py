"""GPU parity tests (run on a B200 with `-m gpu`): every execution mode of the CUDA path, called through
the C ABI, must reproduce the oracle VM bit for bit on the same seeded inputs.

(The reference's own unit tests also run on the GPU path: tests/test_reference_golden.py, `host` leg.)
"""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

import complogic_b200 as host
from complogic_b200 import circuits
from oracle import oracle
from tests.test_spec_ptx import feedback_dag
from tests.helpers import sampled_word_columns, exhaustive_sliced, mask_tail, random_sliced, stride_for, unslice_int

pytestmark = pytest.mark.gpu

MODES = [host.MODE_SPEC, host.MODE_LANE, host.MODE_COOP]


def oracle_out(circ, imm, nv):
    return mask_tail(oracle.run_batch_sliced(circ.sim, len(circ.sim.registers), imm, nv, circ.out_regs), nv)


def gpu_out(circ, imm, nv, mode, flags=0):
    cs = host.CudaSimulation.from_simulation(circ.sim, circ.n_immediates, circ.out_regs, flags=flags, mode=mode)
    assert cs.mode == mode
    return mask_tail(cs.run_batch(imm, nv), nv)


@pytest.mark.parametrize("mode", MODES)
def test_config1_adder8_exhaustive(mode):
    """BASELINE config 1: 8-bit ripple adder, exhaustive 2^17 vectors: GPU == scalar VM == sliced VM == a+b+cin."""
    circ = circuits.ripple_adder(8)
    n = 1 << 17
    imm = exhaustive_sliced(17)
    got = gpu_out(circ, imm, n, mode)
    scalar = oracle.run_batch_scalar(circ.sim, 169, imm, n, circ.out_regs, threads=8)
    assert np.array_equal(got, mask_tail(scalar, n))
    v = np.arange(n, dtype=np.uint64)
    want = (v & np.uint64(0xFF)) + ((v >> np.uint64(8)) & np.uint64(0xFF)) + (v >> np.uint64(16))
    assert np.array_equal(unslice_int(got, n), want)


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("nv", [1, 31, 33, 127, 4097, 1 << 16])
def test_adder32_ragged_sizes(mode, nv):
    circ = circuits.ripple_adder(32)
    imm = random_sliced(65, stride_for(nv), seed=0xC0FFEE)
    assert np.array_equal(gpu_out(circ, imm, nv, mode), oracle_out(circ, imm, nv))


@pytest.mark.parametrize("mode", MODES)
def test_mul16(mode):
    circ = circuits.array_multiplier(16)
    nv = 1 << 15
    imm = random_sliced(32, stride_for(nv), seed=0xC0FFEF)
    got = gpu_out(circ, imm, nv, mode)
    assert np.array_equal(got, oracle_out(circ, imm, nv))
    a, b = unslice_int(imm[:16], nv), unslice_int(imm[16:], nv)
    assert np.array_equal(unslice_int(got, nv), a * b)


@pytest.mark.parametrize("mode", [host.MODE_COOP, host.MODE_LANE])
@pytest.mark.parametrize("window", [2, 0])
def test_small_dag(mode, window):
    width = 200 if mode == host.MODE_COOP else 60  # the per-thread stack of the lane mode holds ~440 live values
    circ = circuits.random_dag(levels=40, width=width, n_immediates=64, window=window, seed=9)
    nv = 5000
    imm = random_sliced(64, stride_for(nv), seed=21)
    for flags in (0, host.LV_NO_FUSE):
        assert np.array_equal(gpu_out(circ, imm, nv, mode, flags), oracle_out(circ, imm, nv))


def test_dag_1m_gates_prefix():
    """BASELINE config 4's circuit (1 048 576 gates, depth 256, K=2) on a 2^12-vector batch."""
    circ = circuits.random_dag()
    nv = 1 << 12
    imm = random_sliced(1024, stride_for(nv), seed=0xD46)
    got = gpu_out(circ, imm, nv, host.MODE_COOP)
    assert np.array_equal(got, oracle_out(circ, imm, nv))


def test_partitioned_cooperative_launch():
    """Small batches of separable programs run one CTA per (part, column group) -- every batch size from one vector up
    to where the monolithic stream takes over -- and must agree with the oracle; NO_PARTS gives the same bits."""
    base = circuits.array_multiplier(16)
    circ = circuits.flatten_instances(base, 9)
    cs = host.CudaSimulation.from_simulation(circ.sim, circ.n_immediates, circ.out_regs, mode=host.MODE_COOP)
    assert cs.info["n_parts"] == 9
    mono = host.CudaSimulation.from_simulation(circ.sim, circ.n_immediates, circ.out_regs, mode=host.MODE_COOP,
                                               flags=host.LV_NO_PARTS)
    for nv in (1, 33, 128 * 3 + 5, 128 * 147, 128 * 149, 128 * 400):
        imm = random_sliced(circ.n_immediates, stride_for(nv), seed=nv)
        want = oracle_out(circ, imm, nv)
        assert np.array_equal(mask_tail(cs.run_batch(imm, nv), nv), want)
        assert np.array_equal(mask_tail(mono.run_batch(imm, nv), nv), want)


def test_dag_1m_gates_no_window():
    """The K = infinity variant of config 4 (inputs drawn from all earlier levels): live set of ~54k values, one word
    per register in shared memory (K = 1), whole-warp bank packing."""
    circ = circuits.random_dag(window=0)
    nv = 1 << 11
    imm = random_sliced(1024, stride_for(nv), seed=0xD47)
    cs = host.CudaSimulation.from_simulation(circ.sim, 1024, circ.out_regs)
    got = mask_tail(cs.run_batch(imm, nv), nv)
    assert cs.mode == host.MODE_COOP and cs.words_per_thread == 1
    assert np.array_equal(got, oracle_out(circ, imm, nv))


@pytest.mark.parametrize("mode", MODES)
def test_flattened_instances_single_vector(mode):
    """config 5 in miniature: independent multiplier instances flattened, ONE input vector."""
    base = circuits.array_multiplier(8)
    circ = circuits.flatten_instances(base, 6)
    for nv in (1, 40):
        imm = random_sliced(circ.n_immediates, stride_for(nv), seed=77)
        assert np.array_equal(gpu_out(circ, imm, nv, mode), oracle_out(circ, imm, nv))


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("gate", ["RSLatch", "RSLatchTest", "DLatch"])
def test_latch_sequences_per_lane(mode, gate):
    """Carried registers (simulation.rs never clears them): 6 steps x 300 lanes, every register, every step."""
    c = host.Compiler(2)
    sim = c.compile([getattr(host, gate)(0, 1, c.alloc())])
    nv, steps, R = 300, 6, len(sim.registers)
    stride = stride_for(nv)
    imm = np.stack([random_sliced(2, stride, 5 + 97 * s) for s in range(steps)])
    want = oracle.run_batch_scalar(sim, R, imm, nv, list(range(R)), steps=steps, all_steps=True)
    cs = host.CudaSimulation.from_simulation(sim, 2, None, flags=host.LV_STATEFUL, mode=mode)
    state = np.zeros((cs.info["n_state"], stride), np.uint32)
    for s in range(steps):
        got = cs.run_batch(imm[s], nv, state=state)
        assert np.array_equal(mask_tail(got, nv), mask_tail(want[s], nv)), f"step {s}"


def test_chunked_specialised_long_programs():
    """Lane-order programs beyond 16384 instructions run as a chain of specialised kernels that hand their live
    registers over through a device scratch buffer: cuts between instances (multipliers) and in the middle of a
    DAG with a few hundred live values must both give the oracle's bits, at several strides with one executor."""
    mul = circuits.flatten_instances(circuits.array_multiplier(64), 5)
    dag = circuits.random_dag(levels=1500, width=48, n_immediates=64, window=2, seed=31)
    for circ in (mul, dag):
        cs = host.CudaSimulation.from_simulation(circ.sim, circ.n_immediates, circ.out_regs, mode=host.MODE_SPEC)
        assert cs.info["lane_instrs"] > 16384 and cs.mode == host.MODE_SPEC
        # the five multipliers are five equal, independent chunks: ONE launch, blockIdx.y picks the instance's rows
        assert cs.launches_per_run == 1 if circ is mul else cs.launches_per_run > 1
        for nv in (300, 70000, 33):  # the scratch buffer grows with the stride and is kept when it shrinks
            imm = random_sliced(circ.n_immediates, stride_for(nv), seed=nv + 5)
            assert np.array_equal(mask_tail(cs.run_batch(imm, nv), nv), oracle_out(circ, imm, nv)), (circ.name, nv)


def test_chunked_specialised_stateful():
    """Carried registers across a chunked program: the state row is loaded by the first kernel of the chain and stored
    by the last; several steps, against the oracle VM stepping the same ops."""
    sim, outs = feedback_dag()
    R, nv, T = len(sim.registers), 200, 4
    stride = stride_for(nv)
    cs = host.CudaSimulation.from_simulation(sim, 64, outs, flags=host.LV_STATEFUL, mode=host.MODE_SPEC)
    assert cs.info["lane_instrs"] > 16384 and cs.info["n_state"] == len(outs) and cs.launches_per_run > 1
    imm = np.stack([random_sliced(64, stride, 40 + t) for t in range(T)])
    want = oracle.run_batch_scalar(sim, R, imm, nv, outs, steps=T, all_steps=True)
    state = np.zeros((cs.info["n_state"], stride), np.uint32)
    for t in range(T):
        got = cs.run_batch(imm[t], nv, state=state)
        assert np.array_equal(mask_tail(got, nv), mask_tail(want[t], nv)), f"step {t}"
    assert np.array_equal(mask_tail(state, nv), mask_tail(want[T - 1], nv))  # state rows ARE the output registers here


def test_auto_tiers_up_to_the_specialised_chain():
    """A long program whose chain needs more than 4 distinct kernels: AUTO starts in an interpreter and specialises
    once the interpreter has cost as much GPU time as the assembly will (here at once: the knob sets the price of a
    kernel to ~0). Same bits before and after."""
    import os
    circ = circuits.random_dag(levels=3400, width=48, n_immediates=64, window=2, seed=37)
    nv = 4096
    imm = random_sliced(64, stride_for(nv), seed=3)
    want = oracle_out(circ, imm, nv)
    cs = host.CudaSimulation.from_simulation(circ.sim, 64, circ.out_regs)
    assert cs.info["lane_instrs"] > 5 * 16384  # at least 6 kernels, all different
    assert np.array_equal(mask_tail(cs.run_batch(imm, nv), nv), want)
    assert cs.mode in (host.MODE_LANE, host.MODE_COOP)  # one short run has not paid for ~5 s of assembly
    os.environ["CLG_TIER_SECONDS_PER_KERNEL"] = "1e-9"
    try:
        assert np.array_equal(mask_tail(cs.run_batch(imm, nv), nv), want)
    finally:
        del os.environ["CLG_TIER_SECONDS_PER_KERNEL"]
    assert cs.mode == host.MODE_SPEC and cs.launches_per_run >= 5
    assert np.array_equal(mask_tail(cs.run_batch(imm, nv), nv), want)
    assert cs.mode == host.MODE_SPEC  # the chain is built: it stays


def latch_chain(n):
    """n D latches in a row behind one enable: latch i's data input is latch i-1's output."""
    c = host.Compiler(2)
    qs = [c.alloc() for _ in range(n)]
    return c.compile([host.DLatch(0 if i == 0 else qs[i - 1], 1, qs[i]) for i in range(n)]), qs


def test_latch_chain_needs_few_registers_in_lane_order():
    """Wide and shallow sequential logic: thousands of carried registers, every state row rewritten in place. The lane
    order keeps a handful of values live (forced state loads drag their ready readers along), so the chain runs
    specialised -- as a chain of chunk kernels -- and in the lane interpreter; all modes agree with the oracle VM."""
    sim, qs = latch_chain(1500)
    R, nv, T = len(sim.registers), 300, 5
    stride = stride_for(nv)
    imm = np.stack([random_sliced(2, stride, 70 + t) for t in range(T)])
    imm[:, 1] |= np.uint32(0x0F0F0F0F)  # enable high in many lanes so that data actually moves down the chain
    want = oracle.run_batch_scalar(sim, R, imm, nv, qs, steps=T, all_steps=True)
    for mode in MODES:
        cs = host.CudaSimulation.from_simulation(sim, 2, qs, flags=host.LV_STATEFUL, mode=mode)
        assert cs.info["lane_slots"] <= 32
        if mode == host.MODE_SPEC:
            assert cs.launches_per_run > 1
        state = np.zeros((cs.info["n_state"], stride), np.uint32)
        for t in range(T):
            got = cs.run_batch(imm[t], nv, state=state)
            assert np.array_equal(mask_tail(got, nv), mask_tail(want[t], nv)), (host.MODE_NAMES[mode], t)


def test_lane_interpreter_variant_follows_the_table_mix(monkeypatch):
    """The lane interpreter tests for the NAND network's two tables in line (clg_lane_n_k*) only where they are the
    majority of the program's LUTs; arithmetic keeps the plain jump table. Both are bit-exact, and so is the plain
    kernel forced onto the NAND network."""
    dag = circuits.random_dag(24, 96, 64, 2, 9)
    mul = circuits.array_multiplier(8)
    nv = 3000
    for circ, prefix in ((dag, "clg_lane_n_k"), (mul, "clg_lane_k")):
        imm = random_sliced(circ.n_immediates, stride_for(nv), seed=3)
        want = mask_tail(oracle.run_batch_sliced(circ.sim, len(circ.sim.registers), imm, nv, circ.out_regs), nv)
        cs = host.CudaSimulation.from_simulation(circ.sim, circ.n_immediates, circ.out_regs, mode=host.MODE_LANE)
        assert np.array_equal(mask_tail(cs.run_batch(imm, nv), nv), want)
        assert cs.kernel_name.startswith(prefix), cs.kernel_name
    monkeypatch.setenv("CLG_LANE_PLAIN", "1")
    imm = random_sliced(dag.n_immediates, stride_for(nv), seed=3)
    want = mask_tail(oracle.run_batch_sliced(dag.sim, len(dag.sim.registers), imm, nv, dag.out_regs), nv)
    cs = host.CudaSimulation.from_simulation(dag.sim, dag.n_immediates, dag.out_regs, mode=host.MODE_LANE)
    assert np.array_equal(mask_tail(cs.run_batch(imm, nv), nv), want)
    assert cs.kernel_name.startswith("clg_lane_k"), cs.kernel_name


def test_lane_only_program_when_level_order_needs_too_many_registers():
    """20 000 latches side by side need > 65 536 registers in level order: the flat buffer then carries the lane stream
    alone and AUTO runs the lane interpreter."""
    sim, qs = latch_chain(20000)
    outs = qs[::997]
    R, nv, T = len(sim.registers), 64, 2
    stride = stride_for(nv)
    imm = np.stack([random_sliced(2, stride, 80 + t) for t in range(T)])
    imm[:, 1] = np.uint32(0xFFFFFFFF)
    want = oracle.run_batch_scalar(sim, R, imm, nv, outs, steps=T, all_steps=True)
    cs = host.CudaSimulation.from_simulation(sim, 2, outs, flags=host.LV_STATEFUL)
    assert cs.info["level_instrs"] == 0 and cs.info["lane_slots"] <= 128 and cs.mode == host.MODE_LANE
    state = np.zeros((cs.info["n_state"], stride), np.uint32)
    for t in range(T):
        got = cs.run_batch(imm[t], nv, state=state)
        assert np.array_equal(mask_tail(got, nv), mask_tail(want[t], nv)), t
    with pytest.raises(host.ClgError):
        host.CudaSimulation.from_simulation(sim, 2, outs, flags=host.LV_STATEFUL, mode=host.MODE_COOP)


@pytest.mark.parametrize("mode", MODES)
def test_fused_multi_step(mode):
    """clg_exec_run_steps: T run() calls per lane in one call (one launch in the specialised mode, carried registers
    kept in registers), every register after every step equal to the oracle VM's."""
    import torch

    c = host.Compiler(3)
    q1, q2, s_, k_ = c.alloc(), c.alloc(), c.alloc(), c.alloc()
    sim = c.compile([host.DLatch(0, 1, q1), host.RSLatch(q1, 2, q2), host.FullAdder(0, q1, q2, s_, k_)])
    R, nv, T = len(sim.registers), 1000, 7
    stride = stride_for(nv)
    imm = np.stack([random_sliced(3, stride, 900 + t) for t in range(T)])
    want = oracle.run_batch_scalar(sim, R, imm, nv, list(range(R)), steps=T, all_steps=True)
    cs = host.CudaSimulation.from_simulation(sim, 3, None, flags=host.LV_STATEFUL, mode=mode)
    d_imm = torch.from_numpy(imm.view(np.int32)).cuda()
    d_out = torch.zeros((T, R, stride), dtype=torch.int32, device="cuda")
    d_state = torch.zeros((max(1, cs.info["n_state"]), stride), dtype=torch.int32, device="cuda")
    cs.run_steps_device(d_imm.data_ptr(), d_out.data_ptr(), d_state.data_ptr(), nv, stride, T)
    torch.cuda.synchronize()
    got = d_out.cpu().numpy().view(np.uint32)
    for t in range(T):
        assert np.array_equal(mask_tail(got[t], nv), mask_tail(want[t], nv)), f"step {t}"
    # the carried registers left behind equal a step-by-step run
    ref = host.CudaSimulation.from_simulation(sim, 3, None, flags=host.LV_STATEFUL, mode=mode)
    state = np.zeros((cs.info["n_state"], stride), np.uint32)
    for t in range(T):
        ref.run_batch(imm[t], nv, state=state)
    assert np.array_equal(mask_tail(d_state.cpu().numpy().view(np.uint32)[:cs.info["n_state"]], nv), mask_tail(state, nv))


@st.composite
def programs(draw):
    n_regs = draw(st.integers(1, 10))
    ops = []
    for _ in range(draw(st.integers(0, 30))):
        if draw(st.integers(0, 4)) == 0:
            ops.append(host.Set_op(draw(st.integers(0, n_regs - 1)), draw(st.booleans())))
        else:
            ops.append(host.Nand_op(*[draw(st.integers(0, n_regs - 1)) for _ in range(3)]))
    return ops, n_regs, draw(st.integers(0, n_regs + 1)), draw(st.sampled_from(MODES))


@settings(max_examples=int(__import__("os").environ.get("CLG_HYPOTHESIS_EXAMPLES", "300")), deadline=None,
          suppress_health_check=[HealthCheck.too_slow])
@given(programs(), st.integers(0, 2 ** 31))
def test_random_programs(prog, seed):
    ops, n_regs, n_imm, mode = prog
    sim = host.Simulation(ops=ops, registers=[False] * n_regs)
    nv, steps = 70, 2
    stride = stride_for(nv)
    imm = np.stack([random_sliced(n_imm, stride, seed + s) for s in range(steps)]) if n_imm else \
        np.zeros((steps, 0, stride), np.uint32)
    want = oracle.run_batch_scalar(ops, n_regs, imm, nv, list(range(n_regs)), steps=steps, all_steps=True)
    cs = host.CudaSimulation.from_simulation(sim, n_imm, None, flags=host.LV_STATEFUL, mode=mode)
    state = np.zeros((cs.info["n_state"], stride), np.uint32)
    for s in range(steps):
        got = cs.run_batch(imm[s], nv, state=state)
        assert np.array_equal(mask_tail(got, nv), mask_tail(want[s], nv))


@settings(max_examples=int(__import__("os").environ.get("CLG_HYPOTHESIS_EXAMPLES", "300")), deadline=None,
          suppress_health_check=[HealthCheck.too_slow])
@given(programs(), st.integers(0, 2 ** 31), st.integers(2, 9))
def test_random_programs_in_tiny_spec_chunks(prog, seed, chunk):
    """The chunked specialised chain on arbitrary stateful programs, cut every few instructions (CLG_SPEC_CHUNK)."""
    import os
    ops, n_regs, n_imm, _ = prog
    sim = host.Simulation(ops=ops, registers=[False] * n_regs)
    nv, steps = 200, 2
    stride = stride_for(nv)
    imm = np.stack([random_sliced(n_imm, stride, seed + s) for s in range(steps)]) if n_imm else \
        np.zeros((steps, 0, stride), np.uint32)
    want = oracle.run_batch_scalar(ops, n_regs, imm, nv, list(range(n_regs)), steps=steps, all_steps=True)
    old = os.environ.get("CLG_SPEC_CHUNK")
    os.environ["CLG_SPEC_CHUNK"] = str(chunk)
    try:
        cs = host.CudaSimulation.from_simulation(sim, n_imm, None, flags=host.LV_STATEFUL, mode=host.MODE_SPEC)
        state = np.zeros((cs.info["n_state"], stride), np.uint32)
        for s in range(steps):
            got = cs.run_batch(imm[s], nv, state=state)
            assert np.array_equal(mask_tail(got, nv), mask_tail(want[s], nv))
    finally:
        if old is None:
            del os.environ["CLG_SPEC_CHUNK"]
        else:
            os.environ["CLG_SPEC_CHUNK"] = old


@settings(max_examples=int(__import__("os").environ.get("CLG_HYPOTHESIS_EXAMPLES", "300")), deadline=None,
          suppress_health_check=[HealthCheck.too_slow])
@given(programs(), st.integers(0, 2 ** 31))
def test_random_programs_streaming_cooperative_kernel(prog, seed):
    """Small programs normally run the shared-memory-resident cooperative kernel; CLG_COOP_NO_RESIDENT sends them
    through the streaming one (compact 8-byte instructions fetched from global memory) that big programs use."""
    import os
    ops, n_regs, n_imm, _ = prog
    sim = host.Simulation(ops=ops, registers=[False] * n_regs)
    nv, steps = 300, 2
    stride = stride_for(nv)
    imm = np.stack([random_sliced(n_imm, stride, seed + s) for s in range(steps)]) if n_imm else \
        np.zeros((steps, 0, stride), np.uint32)
    want = oracle.run_batch_scalar(ops, n_regs, imm, nv, list(range(n_regs)), steps=steps, all_steps=True)
    os.environ["CLG_COOP_NO_RESIDENT"] = "1"
    try:
        cs = host.CudaSimulation.from_simulation(sim, n_imm, None, flags=host.LV_STATEFUL, mode=host.MODE_COOP)
        state = np.zeros((cs.info["n_state"], stride), np.uint32)
        for s in range(steps):
            got = cs.run_batch(imm[s], nv, state=state)
            assert np.array_equal(mask_tail(got, nv), mask_tail(want[s], nv))
    finally:
        del os.environ["CLG_COOP_NO_RESIDENT"]


def test_streaming_cooperative_kernel_deep_dag():
    """A deep, narrow DAG through the streaming cooperative kernel, from one vector to several column groups per CTA."""
    import os
    circ = circuits.random_dag(levels=300, width=96, n_immediates=64, window=6, seed=17)
    os.environ["CLG_COOP_NO_RESIDENT"] = "1"
    try:
        cs = host.CudaSimulation.from_simulation(circ.sim, 64, circ.out_regs, mode=host.MODE_COOP)
    finally:
        del os.environ["CLG_COOP_NO_RESIDENT"]
    for nv in (1, 129, 40000, 128 * 148 * 3 + 77):
        imm = random_sliced(64, stride_for(nv), seed=nv)
        assert np.array_equal(mask_tail(cs.run_batch(imm, nv), nv), oracle_out(circ, imm, nv))


def test_bools_api_and_transposes():
    """`&[bool]`-shaped entry point (one run()'s immediates per row) through the device transposes."""
    circ = circuits.array_multiplier(8)
    cs = host.CudaSimulation.from_simulation(circ.sim, 16, circ.out_regs)
    rng = np.random.default_rng(3)
    import os
    for nv, chunk in ((1, None), (33, None), (1000, None), (5000, "1024"), (3 * 2048 + 5, "2048"), (7 * 1024, "1024")):
        # (small CLG_BOOLS_CHUNK_VECTORS: the chunked H2D | slice, run, unslice | D2H pipeline wraps its three staging sets)
        if chunk:
            os.environ["CLG_BOOLS_CHUNK_VECTORS"] = chunk
        imm = rng.integers(0, 2, size=(nv, 16), dtype=np.uint8)
        out = cs.run_bools(imm)
        os.environ.pop("CLG_BOOLS_CHUNK_VECTORS", None)
        a = (imm[:, :8].astype(np.uint64) << np.arange(8, dtype=np.uint64)).sum(axis=1)
        b = (imm[:, 8:].astype(np.uint64) << np.arange(8, dtype=np.uint64)).sum(axis=1)
        p = (out.astype(np.uint64) << np.arange(16, dtype=np.uint64)).sum(axis=1)
        assert np.array_equal(p, a * b), (nv, chunk)
    # an odd number of rows on either side (65 in, 33 out: the generic transposes)
    add = circuits.ripple_adder(32)
    ca = host.CudaSimulation.from_simulation(add.sim, 65, add.out_regs)
    os.environ["CLG_BOOLS_CHUNK_VECTORS"] = "4096"
    imm = rng.integers(0, 2, size=(20000, 65), dtype=np.uint8)
    out = ca.run_bools(imm)
    os.environ.pop("CLG_BOOLS_CHUNK_VECTORS", None)
    w = np.uint64(1) << np.arange(32, dtype=np.uint64)
    want = (imm[:, :32] * w).sum(1) + (imm[:, 32:64] * w).sum(1) + imm[:, 64]
    assert np.array_equal((out.astype(np.uint64) * (np.uint64(1) << np.arange(33, dtype=np.uint64))).sum(1), want)


def test_transposes_tiled_and_fallback_paths():
    """clg_slice_bools / clg_unslice_bools: the 128-bit tiled kernels (rows % 16 == 0) and the generic ones, ragged
    vector counts, any non-zero byte counts as true."""
    import torch

    ctx = host.Context.default()
    rng = np.random.default_rng(8)
    import os
    cases = [(1, 16), (33, 16), (300, 48), (5000, 64), (257, 1024), (1000, 17), (64, 33),
             (1025, 80), (40000, 128), (3000, 144), (1024 * 3 + 7, 272), (70000, 32)]  # wide tiles: 32 / 64 / 128 rows, ragged edges
    # default: the TMA-tiled kernels in their default tile shape; the other shapes (CLG_TRANSPOSE_TMA = shape index), the wide
    # LDG/STG ones (CLG_TRANSPOSE_NO_TMA) and the 32-byte-segment ones (CLG_TRANSPOSE_NARROW) stay in the library, keep them honest
    cases += [(256 * 12 + 1, 2048)]
    # any other row count: the staged kernels (one bulk copy of VT x rows contiguous bytes per CTA); rows * 32 > 96 KB and
    # CLG_TRANSPOSE_BYTEWISE: the byte-wise kernels
    odd = [(1, 1), (100, 1), (31, 3), (5000, 65), (777, 33), (300, 3000), (40, 3071), (20, 3073), (1 << 16, 9), (257, 255)]
    cases += odd
    variants = [(a, b, None) for a, b in cases] + [(a, b, "CLG_TRANSPOSE_NO_TMA=1") for a, b in cases if b % 16 == 0] + \
               [(a, b, f"CLG_TRANSPOSE_TMA={t}") for t in range(7) for a, b in ((33, 16), (3000, 144), (1024 * 3 + 7, 272), (70000, 32))] + \
               [(300, 48, "CLG_TRANSPOSE_NARROW=1"), (5000, 64, "CLG_TRANSPOSE_NARROW=1"), (257, 1024, "CLG_TRANSPOSE_NARROW=1")] + \
               [(a, b, "CLG_TRANSPOSE_BYTEWISE=1") for a, b in ((5000, 65), (777, 33), (31, 3))]
    for nv, rows, env in variants:
        for k in ("CLG_TRANSPOSE_NARROW", "CLG_TRANSPOSE_NO_TMA", "CLG_TRANSPOSE_TMA", "CLG_TRANSPOSE_BYTEWISE"):
            os.environ.pop(k, None)
        if env:
            os.environ[env.split("=")[0]] = env.split("=")[1]
        stride = stride_for(nv)
        b = rng.integers(0, 2, size=(nv, rows), dtype=np.uint8) * rng.integers(1, 256, size=(nv, rows)).astype(np.uint8)
        bools = torch.from_numpy(b).cuda()
        sliced = torch.zeros((rows, stride), dtype=torch.int32, device="cuda")
        back = torch.full((nv, rows), 9, dtype=torch.uint8, device="cuda")
        ctx.slice_bools(bools.data_ptr(), sliced.data_ptr(), nv, rows, stride)
        ctx.unslice_bools(sliced.data_ptr(), back.data_ptr(), nv, rows, stride)
        torch.cuda.synchronize()
        want_bits = (b != 0)
        got = sliced.cpu().numpy().view(np.uint32)
        bits = ((got[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(rows, -1)[:, :nv].T.astype(bool)
        assert np.array_equal(bits, want_bits), (nv, rows, env)
        assert np.array_equal(back.cpu().numpy(), want_bits.astype(np.uint8)), (nv, rows, env)
    for k in ("CLG_TRANSPOSE_NARROW", "CLG_TRANSPOSE_NO_TMA", "CLG_TRANSPOSE_TMA", "CLG_TRANSPOSE_BYTEWISE"):
        os.environ.pop(k, None)
    # a vector-major matrix that is not 16-byte aligned (a view one byte into a buffer): the byte-wise kernels
    nv, rows = 3000, 48
    b = rng.integers(0, 2, size=(nv, rows), dtype=np.uint8)
    buf = torch.zeros(nv * rows + 16, dtype=torch.uint8, device="cuda")
    buf[1:1 + nv * rows] = torch.from_numpy(b.reshape(-1)).cuda()
    sliced = torch.zeros((rows, stride_for(nv)), dtype=torch.int32, device="cuda")
    back = torch.full((nv * rows + 16,), 7, dtype=torch.uint8, device="cuda")
    ctx.slice_bools(buf.data_ptr() + 1, sliced.data_ptr(), nv, rows, stride_for(nv))
    ctx.unslice_bools(sliced.data_ptr(), back.data_ptr() + 1, nv, rows, stride_for(nv))
    torch.cuda.synchronize()
    got = back.cpu().numpy()
    assert np.array_equal(got[1:1 + nv * rows].reshape(nv, rows), b) and got[0] == 7 and got[1 + nv * rows] == 7


def test_transposes_random_shapes():
    """200 random (vectors, rows, offset) shapes through whichever transpose kernels the library picks -- TMA-tiled, staged,
    byte-wise (odd offsets) -- against numpy, with guard bytes around the vector-major result."""
    import torch

    ctx = host.Context.default()
    rng = np.random.default_rng(2024)
    for case in range(200):
        rows = int(rng.choice([rng.integers(1, 40), rng.integers(1, 400), 16 * rng.integers(1, 40)]))
        nv = int(rng.choice([rng.integers(1, 70), rng.integers(1, 3000), rng.integers(1, 40000)]))
        off = int(rng.choice([0, 0, 16, 48, 1, 5]))
        stride = stride_for(nv)
        b = (rng.random((nv, rows)) < 0.5).astype(np.uint8) * rng.integers(1, 256, size=(nv, rows)).astype(np.uint8)
        buf = torch.zeros(nv * rows + 128, dtype=torch.uint8, device="cuda")
        buf[off:off + nv * rows] = torch.from_numpy(b.reshape(-1)).cuda()
        sliced = torch.zeros((rows, stride), dtype=torch.int32, device="cuda")
        back = torch.full((nv * rows + 128,), 0xA5, dtype=torch.uint8, device="cuda")
        ctx.slice_bools(buf.data_ptr() + off, sliced.data_ptr(), nv, rows, stride)
        ctx.unslice_bools(sliced.data_ptr(), back.data_ptr() + off, nv, rows, stride)
        torch.cuda.synchronize()
        want = b != 0
        got = sliced.cpu().numpy().view(np.uint32)
        bits = ((got[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(rows, -1)[:, :nv].T.astype(bool)
        assert np.array_equal(bits, want), (case, nv, rows, off)
        out = back.cpu().numpy()
        assert np.array_equal(out[off:off + nv * rows].reshape(nv, rows), want.astype(np.uint8)), (case, nv, rows, off)
        assert (out[:off] == 0xA5).all() and (out[off + nv * rows:] == 0xA5).all(), (case, nv, rows, off)


def test_torch_buffers_and_stream_and_fill():
    """Device pointers from torch, torch's current stream, and the on-device input generator."""
    import torch

    circ = circuits.ripple_adder(32)
    nv = (1 << 18) + 5
    stride = stride_for(nv)
    ctx = host.Context.default()
    cs = host.CudaSimulation.from_simulation(circ.sim, 65, circ.out_regs, ctx=ctx)
    imm = torch.empty((65, stride), dtype=torch.int32, device="cuda:0")
    out = torch.zeros((33, stride), dtype=torch.int32, device="cuda:0")
    s = torch.cuda.current_stream().cuda_stream
    ctx.fill_random(imm.data_ptr(), 65, stride, 0xC0FFEE, s)
    cs.run_device(imm.data_ptr(), out.data_ptr(), nv, stride, stream=s)
    torch.cuda.synchronize()
    imm_h = imm.cpu().numpy().view(np.uint32)
    assert np.array_equal(imm_h, random_sliced(65, stride, 0xC0FFEE))
    got = mask_tail(out.cpu().numpy().view(np.uint32), nv)
    assert np.array_equal(got, oracle_out(circ, imm_h, nv))


def test_full_size_adder32_2p24():
    """BASELINE config 2 at full size (2^24 vectors): a + b + cin for every lane, and a checksum of the words."""
    import torch

    circ = circuits.ripple_adder(32)
    nv = 1 << 24
    stride = stride_for(nv)
    ctx = host.Context.default()
    cs = host.CudaSimulation.from_simulation(circ.sim, 65, circ.out_regs, ctx=ctx)
    imm = torch.empty((65, stride), dtype=torch.int32, device="cuda:0")
    out = torch.zeros((33, stride), dtype=torch.int32, device="cuda:0")
    ctx.fill_random(imm.data_ptr(), 65, stride, 0xC0FFEE, 0)
    cs.run_device(imm.data_ptr(), out.data_ptr(), nv, stride)
    torch.cuda.synchronize()
    imm_h, out_h = imm.cpu().numpy().view(np.uint32), out.cpu().numpy().view(np.uint32)
    want = oracle.run_batch_sliced(circ.sim, len(circ.sim.registers), imm_h, nv, circ.out_regs)
    assert np.array_equal(out_h, want)
    sample = slice(0, 1 << 12)  # 2^17 lanes through the arithmetic identity
    n = 1 << 17
    a, b, cin = (unslice_int(imm_h[:32, sample], n), unslice_int(imm_h[32:64, sample], n), unslice_int(imm_h[64:, sample], n))
    assert np.array_equal(unslice_int(out_h[:, sample], n), a + b + cin)


def _device_random(ctx, n_rows, stride, seed):
    import torch
    t = torch.empty((n_rows, stride), dtype=torch.int32, device="cuda:0")
    ctx.fill_random(t.data_ptr(), n_rows, stride, seed, 0)
    return t


def _check_sampled_columns(circ, imm_dev, out_dev, nv, k=4):
    """Full-size check without periodic inputs: sampled word columns (first, last, every CTA residue, random) of a
    device-resident run against the oracle VM on the same input columns."""
    import torch
    cols = sampled_word_columns((nv + 31) // 32, k=k)
    idx = torch.from_numpy(cols.astype(np.int64)).cuda()
    imm_h = imm_dev.index_select(1, idx).cpu().numpy().view(np.uint32)
    out_h = out_dev.index_select(1, idx).cpu().numpy().view(np.uint32)
    want = oracle.run_batch_sliced(circ.sim, len(circ.sim.registers), np.ascontiguousarray(imm_h), len(cols) * 32, circ.out_regs)
    if nv % 32 and cols[-1] == (nv + 31) // 32 - 1:
        m = np.uint32((1 << (nv % 32)) - 1)
        want[:, -1] &= m
        out_h[:, -1] &= m
    assert np.array_equal(out_h, want)
    return len(cols)


def test_full_size_dag_1m_2p24():
    """BASELINE config 4 at full size: the 1 048 576-gate DAG on 2^24 lanes of NON-periodic random inputs (generated on
    the device, reproducible on the host); ~190 sampled column groups -- the first, the last, one for every residue of
    the persistent grid, random ones -- of all 4096 output rows against the oracle VM. Runs the row kernel."""
    import torch

    circ = circuits.random_dag()
    nv = 1 << 24
    stride = stride_for(nv)
    ctx = host.Context.default()
    cs = host.CudaSimulation.from_simulation(circ.sim, 1024, circ.out_regs, ctx=ctx)
    imm = _device_random(ctx, 1024, stride, 0xD4A)
    out = torch.zeros((4096, stride), dtype=torch.int32, device="cuda:0")
    cs.run_device(imm.data_ptr(), out.data_ptr(), nv, stride)
    torch.cuda.synchronize()
    assert cs.mode == host.MODE_COOP and cs.kernel_name.startswith("clg_coop_rows")
    assert _check_sampled_columns(circ, imm, out, nv) >= 4 * 150
    # the previous generation of the cooperative kernel on the same inputs: identical buffers, word for word
    import os
    os.environ["CLG_COOP_NO_ROWS"] = "1"
    try:
        old = host.CudaSimulation.from_simulation(circ.sim, 1024, circ.out_regs, ctx=ctx, mode=host.MODE_COOP)
        out2 = torch.zeros_like(out)
        old.run_device(imm.data_ptr(), out2.data_ptr(), nv, stride)
        torch.cuda.synchronize()
        assert old.kernel_name.startswith("clg_coop_c")
        assert bool((out == out2).all())
    finally:
        del os.environ["CLG_COOP_NO_ROWS"]


def test_full_size_structured_1m_gate_circuit_2p24(monkeypatch):
    """Twelve flattened 64x64 multipliers (1 009 152 gates) on 2^24 lanes, every instance on its OWN random inputs:
    sampled column groups of all 1 536 output rows against the oracle VM, and two instances' blocks must be a * b on a
    sample of lanes. The halves of the instances run side by side (two launches of twelve), each instance handing its
    registers over in its own part of the scratch buffer -- also when the scratch budget only holds five parts at a time."""
    import torch

    base = circuits.array_multiplier(64)
    circ = circuits.flatten_instances(base, 12)
    nv = 1 << 24
    stride = stride_for(nv)
    ctx = host.Context.default()
    cs = host.CudaSimulation.from_simulation(circ.sim, circ.n_immediates, circ.out_regs, ctx=ctx)
    imm = _device_random(ctx, 12 * 128, stride, 0x12A)
    out = torch.zeros((12 * 128, stride), dtype=torch.int32, device="cuda:0")
    cs.run_device(imm.data_ptr(), out.data_ptr(), nv, stride)
    torch.cuda.synchronize()
    # a batch this big runs the chain whose chunks stay below the instruction-cache cliff: two chunks per instance that
    # hand ~200 registers over through the scratch buffer, two shared kernels, first halves in one launch, second halves in the next
    assert cs.mode == host.MODE_SPEC and cs.launches_per_run == 2 and cs.kernel_name == "clg_spec_chunk"
    assert cs.flat.spec_batch_chain_info(1) == {"chunks": 24, "kernels": 2, "launches": 2, "handed_over": 12 * 203}
    assert _check_sampled_columns(circ, imm, out, nv, k=1) >= 100
    n = 1 << 12
    for inst in (0, 7):
        imm_h = imm[inst * 128:(inst + 1) * 128, :n // 32].contiguous().cpu().numpy().view(np.uint32)
        got = out[inst * 128:(inst + 1) * 128, :n // 32].contiguous().cpu().numpy().view(np.uint32)
        a = [int(x) for x in unslice_int(imm_h[:64], n)]
        b = [int(x) for x in unslice_int(imm_h[64:], n)]
        lo, hi = unslice_int(got[:64], n), unslice_int(got[64:], n)
        assert [int(l) | int(h) << 64 for l, h in zip(lo, hi)] == [x * y for x, y in zip(a, b)]
    # five instances at a time (a part is 222 registers x 2 MB): 5 + 5 + 2, two launches each
    monkeypatch.setenv("CLG_SPEC_SCRATCH_MB", "2400")
    cs5 = host.CudaSimulation.from_simulation(circ.sim, circ.n_immediates, circ.out_regs, ctx=ctx)
    out5 = torch.zeros_like(out)
    cs5.run_device(imm.data_ptr(), out5.data_ptr(), nv, stride)
    torch.cuda.synchronize()
    assert cs5.launches_per_run == 6 and bool((out5 == out).all())
    monkeypatch.delenv("CLG_SPEC_SCRATCH_MB")
    del cs5, out5
    # a small batch on the same executor keeps the instance-parallel single launch; same words
    out2 = torch.zeros((12 * 128, stride), dtype=torch.int32, device="cuda:0")
    cs.run_device(imm.data_ptr(), out2.data_ptr(), n, stride)
    torch.cuda.synchronize()
    assert cs.launches_per_run == 1
    assert bool((out2[:, :n // 32] == out[:, :n // 32]).all())
    # ... and so does the big batch when the second chain is switched off
    import os
    os.environ["CLG_SPEC_NO_TP"] = "1"
    try:
        out3 = torch.zeros_like(out2)
        cs.run_device(imm.data_ptr(), out3.data_ptr(), nv, stride)
        torch.cuda.synchronize()
        assert cs.launches_per_run == 1 and bool((out3 == out).all())
    finally:
        del os.environ["CLG_SPEC_NO_TP"]


def test_throughput_chain_only_where_it_pays():
    """The second specialised chain (chunks below the instruction-cache cliff) is for programs whose default chain hands
    nothing over: a single 64 x 64 multiplier on a big batch runs as two chunks with ~200 registers handed over; a random
    DAG with 225 live values, whose default chain already hands registers over, keeps its two chunks (more cuts make it
    slower). Sampled columns against the oracle VM / a * b, and the two chains against each other over the whole buffer."""
    import os
    import torch

    ctx = host.Context.default()
    nv = 1 << 23
    stride = stride_for(nv)
    for circ, want_launches in ((circuits.array_multiplier(64), 2), (circuits.random_dag(210, 160, 64, 2, 7), 2)):
        is_mul = circ.name.startswith("mul")
        cs = host.CudaSimulation.from_simulation(circ.sim, circ.n_immediates, circ.out_regs, ctx=ctx, mode=host.MODE_SPEC)
        assert cs.info["lane_slots"] >= 160
        if is_mul:
            assert cs.flat.spec_batch_chain_info(1) == {"chunks": 2, "kernels": 2, "launches": 2, "handed_over": 203}
        imm = _device_random(ctx, circ.n_immediates, stride, 0x7C4)
        out = torch.zeros((len(circ.out_regs), stride), dtype=torch.int32, device="cuda:0")
        cs.run_device(imm.data_ptr(), out.data_ptr(), nv, stride)
        torch.cuda.synchronize()
        assert cs.launches_per_run == want_launches and cs.kernel_name.startswith("clg_spec"), (circ.name, cs.launches_per_run)
        assert _check_sampled_columns(circ, imm, out, nv) >= 100
        os.environ["CLG_SPEC_NO_TP"] = "1"
        try:
            out2 = torch.zeros_like(out)
            cs.run_device(imm.data_ptr(), out2.data_ptr(), nv, stride)
            torch.cuda.synchronize()
            assert cs.launches_per_run == (1 if is_mul else 2) and bool((out2 == out).all())
        finally:
            del os.environ["CLG_SPEC_NO_TP"]
        small = 5000
        out3 = torch.zeros_like(out)
        cs.run_device(imm.data_ptr(), out3.data_ptr(), small, stride)
        torch.cuda.synchronize()
        w = (small + 31) // 32
        assert cs.launches_per_run == (1 if is_mul else 2) and bool((out3[:, :w - 1] == out[:, :w - 1]).all())


def test_throughput_chain_one_instance_per_kernel():
    """A flattened program whose instances are short (6 x mul32, 110 live registers): a big batch runs ONE launch of a
    kernel that holds one instance (side by side over blockIdx.y) instead of the default chain's two kernels of three
    instances each -- no register crosses a cut, so the chain is taken whatever the register count. Sampled columns
    against the oracle VM, the two chains against each other over the whole buffer."""
    import os
    import torch

    ctx = host.Context.default()
    nv = 1 << 23
    stride = stride_for(nv)
    circ = circuits.flatten_instances(circuits.array_multiplier(32), 6)
    cs = host.CudaSimulation.from_simulation(circ.sim, circ.n_immediates, circ.out_regs, ctx=ctx, mode=host.MODE_SPEC)
    assert cs.info["lane_slots"] < 160
    assert cs.flat.spec_chain_info(1) == {"chunks": 2, "kernels": 2, "launches": 2}
    assert cs.flat.spec_batch_chain_info(1) == {"chunks": 6, "kernels": 1, "launches": 1, "handed_over": 0}
    imm = _device_random(ctx, circ.n_immediates, stride, 0x7C5)
    out = torch.zeros((len(circ.out_regs), stride), dtype=torch.int32, device="cuda:0")
    cs.run_device(imm.data_ptr(), out.data_ptr(), nv, stride)
    torch.cuda.synchronize()
    assert cs.launches_per_run == 1 and cs.kernel_name == "clg_spec_chunk"
    assert _check_sampled_columns(circ, imm, out, nv) >= 100
    # the chain takes no more launches than the first, so a small batch runs it too (six threads per column group
    # instead of one that runs three instances back to back, twice)
    small = 5000
    out3 = torch.zeros_like(out)
    cs.run_device(imm.data_ptr(), out3.data_ptr(), small, stride)
    torch.cuda.synchronize()
    w = (small + 31) // 32
    assert cs.launches_per_run == 1 and bool((out3[:, :w - 1] == out[:, :w - 1]).all())
    os.environ["CLG_SPEC_NO_TP"] = "1"
    try:
        out2 = torch.zeros_like(out)
        cs.run_device(imm.data_ptr(), out2.data_ptr(), nv, stride)
        torch.cuda.synchronize()
        assert cs.launches_per_run == 2 and bool((out2 == out).all())
    finally:
        del os.environ["CLG_SPEC_NO_TP"]


@pytest.mark.parametrize("k,seed,want", [(1, 11, (12, 2, 2)), (2, 12, (18, 3, 3))])
def test_groups_of_chunks_side_by_side(monkeypatch, k, seed, want):
    """Six flattened instances of a random DAG, every instance cut into two or three chunks that hand ~70 registers over
    (a 1 000-instruction limit): the chunks of all instances run position by position -- as many launches as an instance
    has chunks -- each instance in its own part of the scratch buffer. Every instance on its own random inputs, sampled
    columns against the oracle VM; the same with a scratch budget that holds two parts at a time, and against the
    first chain over the whole buffer."""
    import torch

    monkeypatch.setenv("CLG_SPEC_TP_CHUNK", "1000")
    monkeypatch.setenv("CLG_SPEC_TP_MIN_REGS", "0")
    monkeypatch.setenv("CLG_SPEC_K", str(k))
    ctx = host.Context.default()
    circ = circuits.flatten_instances(circuits.random_dag(40, 64, 32, 2, seed), 6)
    nv = 1 << 24
    stride = stride_for(nv)
    cs = host.CudaSimulation.from_simulation(circ.sim, circ.n_immediates, circ.out_regs, ctx=ctx, mode=host.MODE_SPEC)
    info = cs.flat.spec_batch_chain_info(k)
    assert (info["chunks"], info["kernels"], info["launches"]) == want and info["handed_over"] > 300
    imm = _device_random(ctx, circ.n_immediates, stride, 0x6A0 + seed)
    out = torch.zeros((len(circ.out_regs), stride), dtype=torch.int32, device="cuda:0")
    cs.run_device(imm.data_ptr(), out.data_ptr(), nv, stride)
    torch.cuda.synchronize()
    assert cs.words_per_thread == k and cs.launches_per_run == want[2] and cs.kernel_name == "clg_spec_chunk"
    assert _check_sampled_columns(circ, imm, out, nv) >= 100
    monkeypatch.setenv("CLG_SPEC_SCRATCH_MB", str(2 * cs.info["lane_slots"] * (stride * 4 >> 20) + 1))  # two parts and a bit
    cs2 = host.CudaSimulation.from_simulation(circ.sim, circ.n_immediates, circ.out_regs, ctx=ctx, mode=host.MODE_SPEC)
    out2 = torch.zeros_like(out)
    cs2.run_device(imm.data_ptr(), out2.data_ptr(), nv, stride)
    torch.cuda.synchronize()
    assert cs2.launches_per_run == 3 * want[2] and bool((out2 == out).all())
    monkeypatch.setenv("CLG_SPEC_NO_TP", "1")
    out3 = torch.zeros_like(out)
    cs2.run_device(imm.data_ptr(), out3.data_ptr(), nv, stride)
    torch.cuda.synchronize()
    assert cs2.launches_per_run <= 2 and bool((out3 == out).all())


def test_side_by_side_groups_on_a_full_device():
    """When the device has no room for a scratch part per instance, fewer groups run at a time: same words, more launches."""
    import torch

    ctx = host.Context.default()
    circ = circuits.flatten_instances(circuits.array_multiplier(64), 12)
    nv = 1 << 23
    stride = stride_for(nv)
    imm = _device_random(ctx, circ.n_immediates, stride, 0x51D)
    out = torch.zeros((len(circ.out_regs), stride), dtype=torch.int32, device="cuda:0")
    cs = host.CudaSimulation.from_simulation(circ.sim, circ.n_immediates, circ.out_regs, ctx=ctx, mode=host.MODE_SPEC)
    cs.run_device(imm.data_ptr(), out.data_ptr(), nv, stride)
    torch.cuda.synchronize()
    assert cs.launches_per_run == 2
    part = cs.info["lane_slots"] * stride * 4  # 222 registers x 1 MB
    del cs
    torch.cuda.empty_cache()
    cs = host.CudaSimulation.from_simulation(circ.sim, circ.n_immediates, circ.out_regs, ctx=ctx, mode=host.MODE_SPEC)
    out2 = torch.zeros_like(out)
    free, _ = torch.cuda.mem_get_info()
    filler = torch.empty(free - 5 * part, dtype=torch.uint8, device="cuda:0")  # room for five parts, not for twelve
    try:
        cs.run_device(imm.data_ptr(), out2.data_ptr(), nv, stride)
        torch.cuda.synchronize()
    finally:
        del filler
        torch.cuda.empty_cache()
    assert cs.launches_per_run > 2 and bool((out2 == out).all()), cs.launches_per_run


def test_full_size_mul16_2p26():
    """BASELINE config 3 at full size (2^26 lanes), non-periodic device-generated inputs: sampled column groups (first,
    last, every residue of a 148-CTA grid, random) against a * b and against the oracle VM, and a checksum over the whole
    buffer against a second run in another mode."""
    import torch

    circ = circuits.array_multiplier(16)
    nv = 1 << 26
    stride = stride_for(nv)
    ctx = host.Context.default()
    cs = host.CudaSimulation.from_simulation(circ.sim, 32, circ.out_regs, ctx=ctx)
    imm = _device_random(ctx, 32, stride, 0x316)
    out = torch.zeros((32, stride), dtype=torch.int32, device="cuda:0")
    cs.run_device(imm.data_ptr(), out.data_ptr(), nv, stride)
    torch.cuda.synchronize()
    assert cs.mode == host.MODE_SPEC
    _check_sampled_columns(circ, imm, out, nv, k=cs.words_per_thread)
    cols = sampled_word_columns(nv // 32, k=cs.words_per_thread)
    idx = torch.from_numpy(cols.astype(np.int64)).cuda()
    imm_h = imm.index_select(1, idx).cpu().numpy().view(np.uint32)
    out_h = out.index_select(1, idx).cpu().numpy().view(np.uint32)
    n = len(cols) * 32
    assert np.array_equal(unslice_int(out_h, n), unslice_int(imm_h[:16], n) * unslice_int(imm_h[16:], n))
    lane = host.CudaSimulation.from_simulation(circ.sim, 32, circ.out_regs, ctx=ctx, mode=host.MODE_LANE)
    out2 = torch.zeros_like(out)
    lane.run_device(imm.data_ptr(), out2.data_ptr(), nv, stride)
    torch.cuda.synchronize()
    assert bool((out == out2).all())


def test_modes_agree_at_scale_mul16():
    """Size-independent property at 2^22 vectors: the three execution modes produce identical buffers, and they
    equal a*b on a strided sample."""
    circ = circuits.array_multiplier(16)
    nv = 1 << 22
    imm = random_sliced(32, stride_for(nv), seed=0xC0FFEF)
    outs = [host.CudaSimulation.from_simulation(circ.sim, 32, circ.out_regs, mode=m).run_batch(imm, nv) for m in MODES]
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
    cols = np.arange(0, imm.shape[1], 257)
    n = len(cols) * 32
    a, b = unslice_int(imm[:16][:, cols], n), unslice_int(imm[16:][:, cols], n)
    assert np.array_equal(unslice_int(outs[0][:, cols], n), a * b)


def test_config5_wide_single_vector():
    """BASELINE config 5: 64x64 multiplier x 128 instances flattened (10 764 288 gates), ONE input vector, through
    (i) the cooperative intra-level mode (one CTA per part: the general answer for a wide single-vector circuit) and
    (ii) what AUTO picks here, because the 128 instances are equal, independent chunks: ONE launch of one specialised
    kernel with an instance per blockIdx.y. Neither runs 1.5M instructions on one thread; both must be right."""
    base = circuits.array_multiplier(64)
    circ = circuits.flatten_instances(base, 128)
    assert circ.sim.nand_count == 10764288
    coop = host.CudaSimulation.from_simulation(circ.sim, circ.n_immediates, circ.out_regs, mode=host.MODE_COOP)
    auto = host.CudaSimulation.from_simulation(circ.sim, circ.n_immediates, circ.out_regs)
    for nv in (1, 32):
        imm = random_sliced(circ.n_immediates, stride_for(nv), seed=0x5EED)
        want = oracle_out(circ, imm, nv)
        got = mask_tail(coop.run_batch(imm, nv), nv)
        assert coop.mode == host.MODE_COOP
        for m in (0, 77, 127):
            a, b = unslice_int(imm[128 * m:128 * m + 64], nv), unslice_int(imm[128 * m + 64:128 * m + 128], nv)
            assert list(unslice_int(got[128 * m:128 * m + 128], nv)) == [int(x) * int(y) for x, y in zip(a, b)]
        assert np.array_equal(got, want)
        assert np.array_equal(mask_tail(auto.run_batch(imm, nv), nv), want)
        assert auto.mode == host.MODE_SPEC and auto.launches_per_run == 1


def test_host_path_pipelines_many_chunks(monkeypatch):
    """clg_exec_run_host cuts the batch into column chunks that flow through three streams; force many small,
    ragged chunks and a stateful program, and compare with the oracle."""
    monkeypatch.setenv("CLG_HOST_CHUNK_WORDS", "100")
    circ = circuits.array_multiplier(8)
    nv = 32 * 1000 + 17
    imm = random_sliced(16, stride_for(nv), seed=5)
    for mode in MODES:
        assert np.array_equal(gpu_out(circ, imm, nv, mode), oracle_out(circ, imm, nv))
    c = host.Compiler(2)
    sim = c.compile([host.DLatch(0, 1, c.alloc())])
    R, steps = len(sim.registers), 3
    stride = stride_for(nv)
    imms = np.stack([random_sliced(2, stride, 50 + s) for s in range(steps)])
    want = oracle.run_batch_sliced(sim, R, imms, nv, list(range(R)), steps=steps, all_steps=True)
    cs = host.CudaSimulation.from_simulation(sim, 2, None, flags=host.LV_STATEFUL)
    state = np.zeros((cs.info["n_state"], stride), np.uint32)
    for s_ in range(steps):
        assert np.array_equal(mask_tail(cs.run_batch(imms[s_], nv, state=state), nv), mask_tail(want[s_], nv))


def test_single_process_sharded_driver():
    """clg_run_sharded_host: one host thread + context + stream per device, contiguous 4-word-aligned shards, no
    collective. With fewer GPUs than shards the same device is listed several times (shards still run apart)."""
    import torch

    circ = circuits.array_multiplier(16)
    fp = host.FlatProgram.levelize(circ.sim, 32, circ.out_regs)
    n_dev = torch.cuda.device_count()
    for nv, devices in ((100000, [0, 0, 0]), (33, [0, 0]), (1 << 20, list(range(n_dev)) * (1 if n_dev > 1 else 2))):
        imm = random_sliced(32, stride_for(nv), seed=123)
        out, secs = host.run_sharded(fp, imm, nv, devices)
        assert np.array_equal(mask_tail(out, nv), oracle_out(circ, imm, nv))
        assert len(secs) == len(devices)


def test_persistent_sharded_executor_on_every_visible_device():
    """clg_sharded_*: contexts / executors / staging buffers live across calls, a start barrier per run, every visible
    device takes a shard (on a one-GPU box the same device is listed twice, so the shard arithmetic and the concurrent
    workers are still exercised; the multi-device assertions need more than one GPU). Stateless and stateful programs,
    single and multi-step runs, repeated calls on the same object."""
    import torch

    n_dev = torch.cuda.device_count()
    devices = list(range(n_dev)) if n_dev > 1 else [0, 0]
    circ = circuits.array_multiplier(16)
    fp = host.FlatProgram.levelize(circ.sim, 32, circ.out_regs)
    sh = host.ShardedSimulation(fp, devices)
    for nv in (1 << 20, 100003, 33, 1 << 20):  # the same object again and again, batch sizes up and down
        imm = random_sliced(32, stride_for(nv), seed=nv)
        out = sh.run(imm, nv)
        assert np.array_equal(mask_tail(out, nv), oracle_out(circ, imm, nv))
        assert len(sh.last_seconds) == len(devices)
        if nv >= 1 << 20:
            assert all(t > 0 for t in sh.last_seconds)  # every device really had a shard
    sh.close()
    if n_dev > 1:  # each device's name / distinctness: the shards ran on different GPUs
        assert len({host.Context(d).name + str(d) for d in devices}) == n_dev
    # a sequential circuit: carried registers shard by words like any other row, state travels through the caller's buffer
    latch = circuits.latch_chain(12)
    fpl = host.FlatProgram.levelize(latch.sim, 2, latch.out_regs, flags=host.LV_STATEFUL)
    nv, steps = 4099, 5
    stride = stride_for(nv)
    imms = np.stack([random_sliced(2, stride, seed=70 + s_) for s_ in range(steps)])
    want = oracle.run_batch_sliced(latch.sim, len(latch.sim.registers), imms, nv, latch.out_regs, steps=steps, all_steps=True)
    shl = host.ShardedSimulation(fpl, devices)
    state = np.zeros((fpl.info["n_state"], stride), np.uint32)
    got = shl.run(imms, nv, state=state, steps=steps)  # one call, five steps
    assert np.array_equal(mask_tail(got, nv), mask_tail(want, nv))
    state2 = np.zeros_like(state)
    for s_ in range(steps):  # and step by step on the same object
        assert np.array_equal(mask_tail(shl.run(imms[s_], nv, state=state2), nv), mask_tail(want[s_], nv))
    assert np.array_equal(mask_tail(state, nv), mask_tail(state2, nv))
    shl.close()


def test_prepare_moves_the_first_run_cost_out_of_the_run():
    """clg_exec_prepare: the mode choice, the plan (upload / JIT) and the cooperative CTA-size timing of a batch size,
    ahead of the first run; the results are the oracle's. (No wall-clock assertion: what can be observed without one is
    that the specialised kernel's text exists after prepare and before any run, and that prepare is idempotent.)"""
    circ = circuits.random_dag(64, 1024, 256, 2, 3)
    nv = 1 << 16
    for mode, small in ((host.MODE_AUTO, False), (host.MODE_COOP, False), (host.MODE_COOP, True), (host.MODE_SPEC, False)):
        c2 = circuits.array_multiplier(8) if mode == host.MODE_SPEC else circ
        n2 = 3000 if small else nv
        i2 = random_sliced(c2.n_immediates, stride_for(n2), seed=6)
        cs = host.CudaSimulation.from_simulation(c2.sim, c2.n_immediates, c2.out_regs, mode=mode)
        cs.prepare(n2)
        cs.prepare(n2)
        cs.prepare(0)  # nothing to do
        if mode == host.MODE_SPEC:
            assert "lop3.b32" in cs.ptx
        got = mask_tail(cs.run_batch(i2, n2), n2)
        assert np.array_equal(got, oracle_out(c2, i2, n2)), (mode, small, cs.kernel_name)
    # a big batch of a register-heavy program: prepare builds the chain the run will use (chunks below the instruction-cache cliff)
    mul = circuits.array_multiplier(64)
    big = 148 * 128 * 8 * 32
    cs = host.CudaSimulation.from_simulation(mul.sim, mul.n_immediates, mul.out_regs)
    cs.prepare(big)
    imm = random_sliced(128, stride_for(4096), seed=9)
    got = mask_tail(cs.run_batch(imm, 4096), 4096)  # (a small batch on the same executor: the single kernel)
    assert cs.launches_per_run == 1 and np.array_equal(got, oracle_out(mul, imm, 4096))


def test_row_kernel_on_dags_and_forced_programs(monkeypatch):
    """clg_coop_rows (row stream: lag-2 segments, no CTA barrier): random DAGs of several shapes, and -- with
    CLG_ROWS_FORCE -- structured and sequential programs the pass would otherwise leave to the other modes, on batches
    that give every SM a column group, word for word against the oracle VM."""
    for shape, nv in (((64, 1024, 256, 2, 3), 1 << 16), ((40, 512, 128, 2, 5), 40000), ((30, 300, 64, 0, 9), 19001),
                      ((40, 512, 128, 2, 5), 50000)):
        # the last one through clg_coop_rows_t (instruction stream staged through shared memory with cp.async.bulk)
        if nv == 50000:
            monkeypatch.setenv("CLG_ROWS_TMA", "1")
        circ = circuits.random_dag(*shape)
        imm = random_sliced(circ.n_immediates, stride_for(nv), seed=4)
        cs = host.CudaSimulation.from_simulation(circ.sim, circ.n_immediates, circ.out_regs, mode=host.MODE_COOP)
        got = mask_tail(cs.run_batch(imm, nv), nv)
        assert cs.kernel_name.startswith("clg_coop_rows") and cs.info["rows_segments"] > 0
        assert np.array_equal(got, oracle_out(circ, imm, nv))
        assert (cs.kernel_name == "clg_coop_rows_t") == (nv == 50000)
        small = mask_tail(cs.run_batch(imm[:, :8], 200), 200)  # a batch too small for it goes to the older kernels
        assert not cs.kernel_name.startswith("clg_coop_rows")
        assert np.array_equal(small, oracle_out(circ, np.ascontiguousarray(imm[:, :8]), 200))
    monkeypatch.delenv("CLG_ROWS_TMA")
    monkeypatch.setenv("CLG_ROWS_FORCE", "1")
    nv = 19000
    for circ in (circuits.array_multiplier(8), circuits.ripple_adder(16), circuits.flatten_instances(circuits.array_multiplier(8), 3)):
        imm = random_sliced(circ.n_immediates, stride_for(nv), seed=5)
        cs = host.CudaSimulation.from_simulation(circ.sim, circ.n_immediates, circ.out_regs, mode=host.MODE_COOP)
        got = mask_tail(cs.run_batch(imm, nv), nv)
        assert cs.kernel_name.startswith("clg_coop_rows")
        assert np.array_equal(got, oracle_out(circ, imm, nv))
    # a sequential circuit: carried registers through the row kernel, several steps
    latch = circuits.latch_chain(10)
    steps, stride = 4, stride_for(nv)
    imms = np.stack([random_sliced(2, stride, seed=90 + s_) for s_ in range(steps)])
    want = oracle.run_batch_sliced(latch.sim, len(latch.sim.registers), imms, nv, latch.out_regs, steps=steps, all_steps=True)
    cs = host.CudaSimulation.from_simulation(latch.sim, 2, latch.out_regs, flags=host.LV_STATEFUL, mode=host.MODE_COOP)
    state = np.zeros((cs.info["n_state"], stride), np.uint32)
    for s_ in range(steps):
        got = cs.run_batch(imms[s_], nv, state=state)
        assert cs.kernel_name.startswith("clg_coop_rows")
        assert np.array_equal(mask_tail(got, nv), mask_tail(want[s_], nv))


@settings(max_examples=int(__import__("os").environ.get("CLG_HYPOTHESIS_EXAMPLES", "300")), deadline=None,
          suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])
@given(programs(), st.integers(0, 2 ** 31))
def test_random_programs_row_kernel(prog, seed):
    """Arbitrary stateful programs (multi-writer, read-before-write, short immediates) through the row kernel."""
    import os
    ops, n_regs, n_imm, _ = prog
    sim = host.Simulation(ops=ops, registers=[False] * n_regs)
    nv, steps = 19000, 2
    stride = stride_for(nv)
    imm = np.stack([random_sliced(n_imm, stride, seed + s) for s in range(steps)]) if n_imm else \
        np.zeros((steps, 0, stride), np.uint32)
    want = oracle.run_batch_sliced(ops, n_regs, imm, nv, list(range(n_regs)), steps=steps, all_steps=True)
    os.environ["CLG_ROWS_FORCE"] = "1"
    try:
        cs = host.CudaSimulation.from_simulation(sim, n_imm, None, flags=host.LV_STATEFUL, mode=host.MODE_COOP)
        state = np.zeros((cs.info["n_state"], stride), np.uint32)
        for s in range(steps):
            got = cs.run_batch(imm[s], nv, state=state)
            assert cs.info["rows_instrs"] == 0 or cs.kernel_name.startswith("clg_coop_rows")
            assert np.array_equal(mask_tail(got, nv), mask_tail(want[s], nv))
    finally:
        del os.environ["CLG_ROWS_FORCE"]


def test_json_imported_program_runs_on_the_device():
    """f3: a program that arrives as serde_json text (Simulation, and Compiler + gates) is levelized and evaluated on the
    GPU, bit for bit like the oracle VM on the original."""
    circ = circuits.array_multiplier(8)
    sim2 = host.Simulation.from_json(circ.sim.to_json())
    assert sim2.ops == circ.sim.ops and len(sim2.registers) == len(circ.sim.registers)
    nv = 5000
    imm = random_sliced(16, stride_for(nv), seed=21)
    want = oracle_out(circ, imm, nv)
    for mode in MODES:
        cs = host.CudaSimulation.from_simulation(sim2, 16, circ.out_regs, mode=mode)
        assert np.array_equal(mask_tail(cs.run_batch(imm, nv), nv), want)
    # the front end through JSON as well: compiler state and gate list, then compile + run
    c = host.Compiler(3)
    s_, k_ = c.alloc(), c.alloc()
    gates = [host.FullAdder(0, 1, 2, s_, k_)]
    c2 = host.Compiler.from_json(c.to_json())
    g2 = [host.Gate.from_json(g.to_json()) for g in gates]
    sim3 = c2.compile(g2)
    assert sim3.ops == c.compile(gates).ops
    cs = host.CudaSimulation.from_simulation(sim3, 3, [s_, k_])
    imm = exhaustive_sliced(3)
    got = cs.run_batch(imm, 8)
    v = np.arange(8)
    total = (v & 1) + (v >> 1 & 1) + (v >> 2 & 1)
    assert list(unslice_int(got, 8)) == [int(t & 1) | int(t >> 1) << 1 for t in total]


def test_driver_smoke_entry_point():
    """__graft_entry__.smoke(): what the driver runs on a B200 before the bench -- every mode on a small multiplier, the row
    kernel on a DAG, the reference's and_gate test through the device path, each against the oracle."""
    import __graft_entry__ as entry

    entry.smoke()


def test_errors_are_status_codes():
    circ = circuits.ripple_adder(4)
    cs = host.CudaSimulation.from_simulation(circ.sim, 9, circ.out_regs)
    with pytest.raises(host.ClgError):  # stride not a multiple of 4 words
        cs.run_device(16, 16, 32, 3)
    with pytest.raises(host.ClgError):  # lanes do not fit
        cs.run_device(16, 16, 32 * 5, 4)
    big = circuits.random_dag(levels=6, width=3000, n_immediates=3000, window=0, seed=2)
    with pytest.raises(host.ClgError):  # live set too large for the per-thread modes
        host.CudaSimulation.from_simulation(big.sim, 3000, big.out_regs, mode=host.MODE_LANE)
