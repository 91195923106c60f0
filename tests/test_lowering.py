"""The levelize / LUT-map / register-allocate pass, checked on the CPU: the flat buffer it produces is
interpreted with numpy (tests/flat_interp.py) and compared, bit for bit, with the oracle VM
(scalar tier L0 and bit-sliced tier L1) and with arithmetic ground truth."""
import numpy as np
import pytest

import complogic_b200 as host
from complogic_b200 import circuits
from oracle import oracle
from tests import flat_interp
from tests.helpers import exhaustive_sliced, mask_tail, random_sliced, stride_for, unslice_int


def both_streams(fp, imm, n_vectors, state=None, seed=0):
    flat = flat_interp.Flat(fp.to_bytes())
    outs = []
    for s in (flat_interp.STREAM_LEVEL, flat_interp.STREAM_LANE):
        if flat.streams[s]["n_instr"] == 0:
            continue
        st = None if state is None else state.copy()
        outs.append((mask_tail(flat_interp.run(flat, s, imm, st, rng=np.random.default_rng(seed)), n_vectors), st))
    return outs


def test_adder8_exhaustive_three_ways():
    """BASELINE config 1: 8-bit ripple adder, all 2^17 input vectors; scalar VM == sliced VM == a+b+cin == flat."""
    circ_o = circuits.ripple_adder(8, api=oracle)
    circ_h = circuits.ripple_adder(8, api=host)
    assert circ_o.sim.ops == circ_h.sim.ops and len(circ_o.sim.registers) == len(circ_h.sim.registers) == 169
    assert circ_h.sim.nand_count == 152
    n = 1 << 17
    imm = exhaustive_sliced(17)
    scalar = oracle.run_batch_scalar(circ_o.sim, 169, imm, n, circ_o.out_regs, threads=4)
    sliced = oracle.run_batch_sliced(circ_o.sim, 169, imm, n, circ_o.out_regs)
    assert np.array_equal(scalar, sliced)
    v = np.arange(n, dtype=np.uint64)
    want = (v & np.uint64(0xFF)) + ((v >> np.uint64(8)) & np.uint64(0xFF)) + (v >> np.uint64(16))
    assert np.array_equal(unslice_int(sliced, n), want)
    fp = host.FlatProgram.levelize(circ_h.sim, 17, circ_h.out_regs)
    assert fp.info["ref_nand_count"] == 152 and fp.info["n_luts"] <= 2 * 8 + 2
    for got, _ in both_streams(fp, imm, n):
        assert np.array_equal(got, sliced)


@pytest.mark.parametrize("n", [1, 2, 5, 32])
def test_ripple_adder_random(n):
    circ = circuits.ripple_adder(n, api=host)
    assert circ.sim.nand_count == 19 * n and len(circ.sim.registers) == 21 * n + 1
    nv = 1000
    imm = random_sliced(2 * n + 1, stride_for(nv), seed=0xC0FFEE)
    want = mask_tail(oracle.run_batch_sliced(circ.sim, len(circ.sim.registers), imm, nv, circ.out_regs), nv)
    a, b = unslice_int(imm[:n], nv), unslice_int(imm[n:2 * n], nv)
    cin = unslice_int(imm[2 * n:], nv)
    assert np.array_equal(unslice_int(want, nv), a + b + cin)
    fp = host.FlatProgram.levelize(circ.sim, 2 * n + 1, circ.out_regs)
    for got, _ in both_streams(fp, imm, nv):
        assert np.array_equal(got, want)


def test_mul4_exhaustive():
    circ = circuits.array_multiplier(4, api=host)
    assert circ.sim.nand_count == 216
    n = 256
    imm = exhaustive_sliced(8)
    want = oracle.run_batch_scalar(circ.sim, len(circ.sim.registers), imm, n, circ.out_regs)
    v = np.arange(n, dtype=np.uint64)
    assert np.array_equal(unslice_int(want, n), (v & np.uint64(15)) * (v >> np.uint64(4)))
    fp = host.FlatProgram.levelize(circ.sim, 8, circ.out_regs)
    for got, _ in both_streams(fp, imm, n):
        assert np.array_equal(got, mask_tail(want, n))


def test_mul16_random():
    """BASELINE config 3's circuit: 4 896 NANDs (reference expansions), checked against a*b."""
    circ = circuits.array_multiplier(16, api=host)
    circ_o = circuits.array_multiplier(16, api=oracle)
    assert circ.sim.ops == circ_o.sim.ops
    assert circ.sim.nand_count == 4896 and len(circ.sim.registers) == 4928
    nv = 2048
    imm = random_sliced(32, stride_for(nv), seed=0xC0FFEF)
    want = oracle.run_batch_sliced(circ.sim, 4928, imm, nv, circ.out_regs)
    a, b = unslice_int(imm[:16], nv), unslice_int(imm[16:], nv)
    assert np.array_equal(unslice_int(want, nv), a * b)
    scalar = oracle.run_batch_scalar(circ.sim, 4928, imm, 256, circ.out_regs)
    assert np.array_equal(mask_tail(scalar, 256), mask_tail(want, 256))
    fp = host.FlatProgram.levelize(circ.sim, 32, circ.out_regs)
    assert fp.info["n_luts"] < 1000  # ~6.6 reference NANDs per LOP3
    for got, _ in both_streams(fp, imm, nv):
        assert np.array_equal(got, mask_tail(want, nv))


@pytest.mark.parametrize("window", [2, 0])
def test_random_dag_small(window):
    circ = circuits.random_dag(levels=24, width=96, n_immediates=40, window=window, seed=7, api=host)
    circ_o = circuits.random_dag(levels=24, width=96, n_immediates=40, window=window, seed=7, api=oracle)
    assert np.array_equal(circ.sim.ops_array["a"], circ_o.sim.ops_array["a"])
    nv = 300
    imm = random_sliced(40, stride_for(nv), seed=3)
    R = len(circ.sim.registers)
    want = oracle.run_batch_sliced(circ_o.sim, R, imm, nv, circ.out_regs)
    scalar = oracle.run_batch_scalar(circ_o.sim, R, imm, nv, circ.out_regs)
    assert np.array_equal(mask_tail(want, nv), mask_tail(scalar, nv))
    for flags in (0, host.LV_NO_FUSE):
        fp = host.FlatProgram.levelize(circ.sim, 40, circ.out_regs, flags=flags)
        for got, _ in both_streams(fp, imm, nv):
            assert np.array_equal(got, mask_tail(want, nv))


def test_flatten_instances():
    base = circuits.array_multiplier(4, api=host)
    flat3 = circuits.flatten_instances(base, 3, api=host)
    assert flat3.sim.nand_count == 3 * 216 and flat3.n_immediates == 24
    nv = 64
    imm = random_sliced(24, stride_for(nv), seed=11)
    want = oracle.run_batch_sliced(flat3.sim, len(flat3.sim.registers), imm, nv, flat3.out_regs)
    for m in range(3):
        a, b = unslice_int(imm[8 * m:8 * m + 4], nv), unslice_int(imm[8 * m + 4:8 * m + 8], nv)
        assert np.array_equal(unslice_int(want[8 * m:8 * m + 8], nv), a * b)
    fp = host.FlatProgram.levelize(flat3.sim, 24, flat3.out_regs)
    for got, _ in both_streams(fp, imm, nv):
        assert np.array_equal(got, mask_tail(want, nv))


def test_partitioned_level_stream():
    """Independent components (flattened instances) also come as separately scheduled parts for small batches."""
    base = circuits.array_multiplier(16)
    circ = circuits.flatten_instances(base, 7)
    fp = host.FlatProgram.levelize(circ.sim, circ.n_immediates, circ.out_regs)
    assert fp.info["n_parts"] == 7
    flat = flat_interp.Flat(fp.to_bytes())
    assert sum(p["n_instr"] for p in flat.parts) == flat.streams[0]["n_instr"]
    nv = 100
    imm = random_sliced(circ.n_immediates, stride_for(nv), seed=4)
    want = mask_tail(oracle.run_batch_sliced(circ.sim, len(circ.sim.registers), imm, nv, circ.out_regs), nv)
    got = mask_tail(flat_interp.run(flat, flat_interp.STREAM_PARTS, imm, rng=np.random.default_rng(1)), nv)
    assert np.array_equal(got, want)
    assert host.FlatProgram.levelize(circ.sim, circ.n_immediates, circ.out_regs, flags=host.LV_NO_PARTS).info["n_parts"] == 0
    one = host.FlatProgram.levelize(base.sim, 32, base.out_regs)
    assert one.info["n_parts"] == 0  # a single component has nothing to split
    again = host.FlatProgram.from_bytes(fp.to_bytes())
    assert again.info == fp.info


def test_all_registers_observable_by_default():
    """out_regs=None: every register is read back (Simulation::registers is public, simulation.rs:11)."""
    c = host.Compiler(3)
    sim = c.compile([host.FullAdder(0, 1, 2, c.alloc(), c.alloc())])
    imm = exhaustive_sliced(3)
    want = oracle.run_batch_scalar(sim, 22, imm, 8, list(range(22)))
    fp = host.FlatProgram.levelize(sim, 3)
    assert fp.info["n_outputs"] == 22
    for got, _ in both_streams(fp, imm, 8):
        assert np.array_equal(got, mask_tail(want, 8))


def test_short_immediates_fall_back_to_literals():
    """simulation.rs:26 `immediates.get(id).copied().unwrap_or(val)`; op_set / op_set_no_immediate batch-wise."""
    sim = host.Simulation(ops=[host.Set_op(0, False), host.Set_op(1, True), host.Set_op(2, True), host.Nand_op(1, 2, 3),
                               host.Nand_op(0, 3, 4)], registers=[False] * 5)
    for n_imm in (0, 1, 2, 3):
        imm = exhaustive_sliced(n_imm) if n_imm else np.zeros((0, 4), np.uint32)
        nv = 1 << n_imm
        want = oracle.run_batch_scalar(sim, 5, imm, nv, list(range(5)))
        fp = host.FlatProgram.levelize(sim, n_imm)
        for got, _ in both_streams(fp, imm, nv):
            assert np.array_equal(got, mask_tail(want, nv))


def test_out_of_range_registers_are_errors_not_ub():
    sim = host.Simulation(ops=[host.Nand_op(0, 0, 9)], registers=[False])
    with pytest.raises(host.ReferencePanic):
        host.FlatProgram.levelize(sim, 0)
    sim = host.Simulation(ops=[host.Set_op(0, False)], registers=[False])
    with pytest.raises(host.ReferencePanic):
        host.FlatProgram.levelize(sim, 1, out_regs=[5])


def test_flat_roundtrip_and_validation():
    circ = circuits.ripple_adder(4, api=host)
    fp = host.FlatProgram.levelize(circ.sim, 9, circ.out_regs)
    data = fp.to_bytes()
    again = host.FlatProgram.from_bytes(data)
    assert again.to_bytes() == data and again.info == fp.info
    with pytest.raises(host.ClgError):
        host.FlatProgram.from_bytes(data[:-16])
    flat = flat_interp.Flat(data)
    off = int.from_bytes(data[64 + 16:64 + 20], "little")  # level stream: off_instr
    first_lut = int(np.flatnonzero(flat.streams[0]["instr"]["kind"] == flat_interp.K_LUT)[0])
    bad = bytearray(data)
    bad[off + 16 * first_lut + 2:off + 16 * first_lut + 4] = b"\xff\xff"  # operand slot a -> out of range
    with pytest.raises(host.ClgError):
        host.FlatProgram.from_bytes(bytes(bad))
    bad = bytearray(data)
    bad[off + 16 * first_lut + 9] = 7  # unknown instruction kind
    with pytest.raises(host.ClgError):
        host.FlatProgram.from_bytes(bytes(bad))
    assert flat.n_inputs == 9


def test_json_roundtrip():
    c = host.Compiler(2)
    sim = c.compile([host.And(0, 1, c.alloc())])
    text = sim.to_json()
    assert text == ('{"ops":[{"Set":[0,false]},{"Set":[1,false]},{"Nand":[0,1,3]},{"Nand":[3,3,2]}],'
                    '"registers":[false,false,false,false]}')
    back = host.Simulation.from_json(text)
    assert back.ops == sim.ops and back.registers == sim.registers
    with pytest.raises(host.ClgError):
        host.Simulation.from_json('{"ops":[{"Nor":[0,1,2]}],"registers":[]}')


def test_empty_and_degenerate_programs():
    """compile(vec![]) gives a Simulation with no ops (compile.rs:96-101); constants, pure copies and dead code."""
    c = host.Compiler(2)
    sim = c.compile([])
    assert sim.ops == [] and len(sim.registers) == 2
    fp = host.FlatProgram.levelize(sim, 2)          # every register reads back false: nothing writes them
    flat = flat_interp.Flat(fp.to_bytes())
    imm = exhaustive_sliced(2)
    for s in (0, 1):
        assert not flat_interp.run(flat, s, imm).any()
    # outputs that are plain copies of immediates, inverted copies, and constants
    sim = host.Simulation(ops=[host.Set_op(0, False), host.Set_op(1, True), host.Nand_op(0, 0, 2), host.Set_op(3, True),
                               host.Nand_op(3, 3, 4), host.Nand_op(0, 1, 5)], registers=[False] * 7)
    for n_imm in (2, 1):
        imm = exhaustive_sliced(n_imm)
        want = oracle.run_batch_scalar(sim, 7, imm, 1 << n_imm, list(range(7)))
        fp = host.FlatProgram.levelize(sim, n_imm)
        for got, _ in both_streams(fp, imm, 1 << n_imm):
            assert np.array_equal(got, mask_tail(want, 1 << n_imm))
    # the same register named twice as an output, and an output nobody computes
    fp = host.FlatProgram.levelize(sim, 2, out_regs=[5, 5, 6])
    imm = exhaustive_sliced(2)
    want = oracle.run_batch_scalar(sim, 7, imm, 4, [5, 5, 6])
    for got, _ in both_streams(fp, imm, 4):
        assert np.array_equal(got, mask_tail(want, 4))


def test_latch_chain_lane_order_keeps_few_values_live():
    """Sequential logic rewrites its state rows in place, which forces every row's old value to be loaded before the
    store. The lane order runs the ready readers of such a forced load at once, so a chain of n latches needs a
    handful of registers plus the state loads hoisted for memory-level parallelism (it needed 2n + 2 when the loaded values waited for the depth-first order to reach them);
    when the level order needs more than 65 536 registers the buffer carries the lane stream alone."""
    def chain(n):
        c = host.Compiler(2)
        qs = [c.alloc() for _ in range(n)]
        return c.compile([host.DLatch(0 if i == 0 else qs[i - 1], 1, qs[i]) for i in range(n)]), qs

    sim, qs = chain(400)
    fp = host.FlatProgram.levelize(sim, 2, qs, flags=host.LV_STATEFUL)
    assert fp.info["lane_slots"] <= 32 < fp.info["level_slots"]
    nv, T = 96, 4
    stride = stride_for(nv)
    imm = np.stack([random_sliced(2, stride, 11 + t) for t in range(T)])
    imm[:, 1] |= np.uint32(0x33333333)
    want = oracle.run_batch_scalar(sim, len(sim.registers), imm, nv, qs, steps=T, all_steps=True)
    flat = flat_interp.Flat(fp.to_bytes())
    for stream in (flat_interp.STREAM_LEVEL, flat_interp.STREAM_LANE):
        state = np.zeros((flat.n_state, stride), np.uint32)
        for t in range(T):
            got = flat_interp.run(flat, stream, imm[t], state)
            assert np.array_equal(mask_tail(got, nv), mask_tail(want[t], nv)), (stream, t)
    sim, qs = chain(12000)
    fp = host.FlatProgram.levelize(sim, 2, qs[-3:], flags=host.LV_STATEFUL)
    assert fp.info["level_instrs"] == 0 and fp.info["lane_instrs"] > 0 and fp.info["lane_slots"] <= 32


def test_mutated_flat_buffers_are_rejected_or_safe():
    """clg_flat_from_bytes is the trust boundary of the device code (kernels carry no bounds checks): random byte
    flips in a valid buffer must either be rejected with a status code or yield a buffer that still passes every
    index and hazard check -- and that the text generators can walk -- never a crash."""
    rng = np.random.default_rng(7)
    progs = [circuits.array_multiplier(8), circuits.latch_chain(20), circuits.flatten_instances(circuits.array_multiplier(16), 9)]
    accepted = rejected = 0
    for c in progs:
        fp = host.FlatProgram.levelize(c.sim, c.n_immediates, c.out_regs, flags=host.LV_STATEFUL if c.stateful else 0)
        data = fp.to_bytes()
        for _ in range(250):
            d = bytearray(data)
            for _ in range(int(rng.integers(1, 4))):
                d[int(rng.integers(0, len(d)))] = int(rng.integers(0, 256))
            try:
                g = host.FlatProgram.from_bytes(bytes(d))
            except host.ClgError:
                rejected += 1
                continue
            accepted += 1
            assert g.info["n_inputs"] >= 0
            try:
                g.spec_chain_info(1)
                g.spec_chunk_ptx(1, 0)
            except host.ClgError:
                pass
    assert rejected > 100 and accepted + rejected == 750


def test_unread_lut_operands_are_validated():
    """the encoders and the PTX emitter touch all three operand fields of a LUT, also the ones its table
    does not read -- an out-of-range value there must be rejected like any other slot index."""
    c = circuits.array_multiplier(8)
    fp = host.FlatProgram.levelize(c.sim, c.n_immediates, c.out_regs)
    data = fp.to_bytes()
    flat = flat_interp.Flat(data)
    hit = 0
    for s in (flat_interp.STREAM_LEVEL, flat_interp.STREAM_LANE):
        st = flat.streams[s]
        off_instr = struct_off_instr(data, s)
        for i in range(st["n_instr"]):
            x = st["instr"][i]
            nfan = (int(x["flags"]) >> 8) & 3
            if int(x["kind"]) != 0 or nfan == 3:
                continue
            d = bytearray(data)
            field = 6 if nfan == 2 else 4  # byte offset of c (nfan 2) or b (nfan 1) inside the 16-byte instruction
            d[off_instr + 16 * i + field: off_instr + 16 * i + field + 2] = (0xFFFF).to_bytes(2, "little")
            with pytest.raises(host.ClgError):
                host.FlatProgram.from_bytes(bytes(d))
            hit += 1
            if hit % 7 == 0:
                break
    assert hit >= 2


def struct_off_instr(data: bytes, stream: int) -> int:
    import struct
    return struct.unpack_from("<6I", data, 64 + 32 * stream)[4]


def test_constant_store_over_a_carried_row_stays_with_its_load():
    """in the partitioned form a carried register that is read and then overwritten with a constant
    (`Nand(S,S,X); Set(S,true)`) must have its load and its store in ONE part -- parts run as unordered CTAs."""
    n_comp = 5000  # (the pass partitions programs of >= 4096 LUTs)
    n_imm = 2 * n_comp  # every component reads its own two immediates, so the network really separates
    ops = [host.Set_op(i, False) for i in range(n_imm)]
    S = n_imm
    X = S + 1
    nxt = X + 1
    outs = [X]
    for k in range(n_comp):  # independent XOR-ish components so that the network separates into many parts
        a, b = 2 * k, 2 * k + 1
        t, u = nxt, nxt + 1
        nxt += 2
        ops += [host.Nand_op(a, b, t), host.Nand_op(t, a, u)]
        outs.append(u)
    ops += [host.Nand_op(S, S, X), host.Set_op(S, True)]
    sim = host.Simulation(ops=ops, registers=[False] * nxt)
    fp = host.FlatProgram.levelize(sim, n_imm, outs, flags=host.LV_STATEFUL)
    flat = flat_interp.Flat(fp.to_bytes())
    assert fp.info["n_state"] == 1
    assert flat.n_parts > 100
    if flat.n_parts:
        loads = [q for q, p in enumerate(flat.parts) if ((p["instr"]["kind"] == 1) & (p["instr"]["row"] >= n_imm)).any()]
        stores = [q for q, p in enumerate(flat.parts) if ((p["instr"]["kind"] >= 2) & (p["instr"]["row"] >= len(outs))).any()]
        assert loads == stores and len(loads) == 1
    # and the values: two consecutive runs per lane against the oracle VM
    nv = 64
    imm = random_sliced(n_imm, stride_for(nv), seed=3)
    state = np.zeros((1, stride_for(nv)), dtype=np.uint32)
    for stream in (flat_interp.STREAM_LEVEL, flat_interp.STREAM_PARTS) if flat.n_parts else (flat_interp.STREAM_LEVEL,):
        st = state.copy()
        got1 = mask_tail(flat_interp.run(flat, stream, imm, st, rng=np.random.default_rng(2)), nv)
        got2 = mask_tail(flat_interp.run(flat, stream, imm, st, rng=np.random.default_rng(3)), nv)
        # X = !S: first run sees S = false -> X = 1 in every lane; second run sees S = true -> X = 0
        assert (got1[0, : nv // 32] == 0xFFFFFFFF).all() and (got2[0, : nv // 32] == 0).all()


def _rows_vs_oracle(circ, nv=96):
    """Row stream of the cooperative row kernel, run with the kernel's semantics (a register per lane, warps up to
    lag - 1 segments apart) in all three skews against the oracle VM."""
    fp = host.FlatProgram.levelize(circ.sim, circ.n_immediates, circ.out_regs)
    flat = flat_interp.Flat(fp.to_bytes())
    assert flat.rows is not None, "no row stream"
    imm = random_sliced(circ.n_immediates, stride_for(nv), seed=11)
    want = mask_tail(oracle.run_batch_sliced(circ.sim, len(circ.sim.registers), imm, nv, circ.out_regs), nv)
    for ahead in (0, 1, 2):
        got = mask_tail(flat_interp.run_rows(flat, imm, None, ahead=ahead), nv)
        assert np.array_equal(got, want), f"row stream differs from the VM (skew {ahead})"
    return fp


@pytest.mark.parametrize("shape", [(40, 512, 128, 2, 5), (64, 1024, 256, 2, 3), (30, 300, 64, 0, 9)])
def test_row_stream_random_dag(shape):
    fp = _rows_vs_oracle(circuits.random_dag(*shape))
    info = fp.info
    assert info["rows_segments"] > 0 and info["rows_instrs"] % 32 == 0 and info["rows_slots"] <= 16383


def test_row_stream_forced_on_small_and_structured_programs(monkeypatch):
    """CLG_ROWS_FORCE builds the row stream for programs the pass would leave to the other modes: adders and
    multipliers (narrow, deep), flattened instances."""
    monkeypatch.setenv("CLG_ROWS_FORCE", "1")
    for circ in (circuits.ripple_adder(8), circuits.array_multiplier(8), circuits.flatten_instances(circuits.array_multiplier(8), 5)):
        _rows_vs_oracle(circ)


def test_compiler_and_gate_json_in_serdes_shape():
    """f3: Compiler / Incrementer / Gate as serde_json writes the derives at compile.rs:18,51 and gates.rs:5-117
    (struct = object of its fields in declaration order, Gate = externally tagged enum)."""
    c = host.Compiler(2)
    q = c.alloc()
    text = c.to_json()
    assert text == '{"immediate_count":2,"ops":[],"incrementer":{"val":3}}'
    sim = c.compile([host.And(0, 1, q)])
    text = c.to_json()  # compile() leaves its op list in the public `ops` field (compile.rs:94-108)
    assert text.startswith('{"immediate_count":2,"ops":[{"Set":[0,false]},{"Set":[1,false]},{"Nand":[0,1,3]},{"Nand":[3,3,2]}]')
    again = host.Compiler.from_json(text)
    assert again.immediate_count == 2 and again.incrementer.val == 3 and again.ops == c.ops
    assert again.compile([host.And(0, 1, q)]).ops == sim.ops
    assert host.Compiler.from_json(' { "incrementer" : {"val": 9}, "ops": [], "immediate_count": 4 } ').alloc() == 9  # any field order
    gates = [host.Nand(0, 1, 2), host.Not(3, 4), host.Xor(1, 2, 3), host.RSLatch(0, 1, 2), host.DLatch(5, 6, 7), host.HalfAdder(0, 1, 2, 3),
             host.FullAdder(0, 1, 2, 3, 4), host.FourBitAdder(*range(13))]
    assert gates[6].to_json() == '{"FullAdder":{"a":0,"b":1,"cin":2,"s":3,"cout":4}}'
    assert gates[3].to_json() == '{"RSLatch":{"s":0,"r":1,"q":2}}'
    for g in gates:
        back = host.Gate.from_json(g.to_json())
        assert back == g and type(back) is type(g)
    assert host.Gate.from_json('{"Or":{"out":7,"a":1,"b":2}}') == host.Or(1, 2, 7)
    for bad in ('{"Nope":{"a":1}}', '{"And":{"a":1,"b":2}}', '{"And":{"a":1,"b":2,"out":3,"x":4}}', '{"immediate_count":1}', ''):
        with pytest.raises(host.ClgError):
            (host.Compiler.from_json if "immediate" in bad else host.Gate.from_json)(bad)
