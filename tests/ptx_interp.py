"""TEST INFRASTRUCTURE: numpy interpreter for the straight-line PTX the specialised mode generates
(complogic_b200/csrc/spec.cpp), so the CPU-only suite can check the emitted code -- including the chunked
chain and its register hand-over -- without a GPU. It understands exactly the instruction forms the emitter
uses and raises on anything else; every thread of the launch is evaluated at once (one numpy lane per thread).
"""
import re

import numpy as np

_FULL = np.uint32(0xFFFFFFFF)
_BASES = ("imm", "out", "state", "scr")


def _regs(text):
    return [r.strip() for r in text.strip("{} ").split(",")]


def run_kernel(ptx: str, k: int, groups: int, mem: dict):
    """Execute one generated kernel. mem: {"imm","out","state","scr"} -> uint32 [rows][stride] (modified in place);
    thread g owns words [g*k, (g+1)*k) of every row. Loops (the multi-step form) are not supported."""
    lines = [l.strip() for l in ptx.splitlines()]
    start = lines.index("// body") + 1  # (the prologue computes the thread's column offset and row bases)
    reg = {}
    addr = None
    cols = np.arange(groups)[:, None] * k + np.arange(k)[None, :]  # [groups][k]

    def val(name):
        if name.startswith("0x"):
            return np.full(groups, int(name, 16), np.uint32)
        if name.lstrip("-").isdigit():
            return np.full(groups, int(name) & 0xFFFFFFFF, np.uint32)
        return reg[name]

    scr_nc, scr_stored = set(), set()
    for l in lines[start:]:
        if not l or l.startswith("//") or l.startswith(".reg") or l in ("DONE:", "ret;", "}"):
            continue
        op, _, rest = l.rstrip(";").partition(" ")
        if op == "mad.lo.u64":
            d, sb, n, base = _regs(rest)
            assert d == "%addr" and sb == "%sb" and base[1:] in _BASES, l
            addr = (base[1:], int(n))
        elif op.startswith("ld.global"):
            m = re.fullmatch(r"(\{[^}]*\}|%\w+), \[(%addr|%scr\+\d+)\]", rest)
            assert m, l
            if m.group(2) != "%addr":  # a scratch row: a compile-time offset of 128 k bytes per row from the thread's base
                off = int(m.group(2)[5:])
                assert off % (128 * k) == 0, l
                addr = ("scr", off // (128 * k))
            dst = _regs(m.group(1))
            assert len(dst) == k and (("v%d" % k) in op or k == 1), l
            if addr[0] == "state":
                assert ".nc" not in op, "rows rewritten during the run must use coherent loads: " + l
            if addr[0] == "scr" and ".nc" in op:  # legal only for scratch rows this kernel never writes
                assert addr[1] not in scr_stored, "read-only load of a scratch row this chunk has written: " + l
                scr_nc.add(addr[1])
            row = mem[addr[0]][addr[1]]
            for j, r in enumerate(dst):
                reg[r] = row[cols[:, j]].copy()
        elif op.startswith("st.global"):
            m = re.fullmatch(r"\[(%addr|%scr\+\d+)\], (\{[^}]*\}|%\w+)", rest)
            assert m, l
            if m.group(1) != "%addr":
                off = int(m.group(1)[5:])
                assert off % (128 * k) == 0, l
                addr = ("scr", off // (128 * k))
            src = _regs(m.group(2))
            assert len(src) == k, l
            if addr[0] == "scr":
                assert addr[1] not in scr_nc, "store to a scratch row this chunk reads through the read-only path: " + l
                scr_stored.add(addr[1])
            row = mem[addr[0]][addr[1]]
            for j, r in enumerate(src):
                row[cols[:, j]] = val(r)
        elif op == "lop3.b32":
            d, a, b, c, tt = _regs(rest)
            a, b, c, tt = val(a), val(b), val(c), int(tt, 16)
            res = np.zeros(groups, np.uint32)
            for bit in range(8):
                if tt >> bit & 1:
                    res |= (a if bit & 4 else ~a) & (b if bit & 2 else ~b) & (c if bit & 1 else ~c)
            reg[d] = res
        elif op == "not.b32":
            d, a = _regs(rest)
            reg[d] = ~val(a)
        elif op in ("mov.b32", "mov.u32"):
            d, a = _regs(rest)
            reg[d] = val(a).copy()
        else:
            raise AssertionError("unexpected instruction in generated PTX: " + l)


def run_program(flat_program, k: int, imm: np.ndarray, state: np.ndarray, n_outputs: int, n_slots: int, second: bool = False):
    """Run the whole (possibly chunked) specialised program over imm [n_inputs][stride]; state is updated in place.
    second: the chain big batches run (in chunk order, which is one of the orders its launches allow).
    The scratch buffer starts poisoned so that a register a chunk forgot to hand over shows up as a mismatch."""
    stride = imm.shape[1]
    assert stride % k == 0
    out = np.zeros((max(1, n_outputs), stride), np.uint32)
    scr = np.full((max(1, n_slots), stride), 0xDEADBEEF, np.uint32)
    mem = {"imm": imm, "out": out, "state": state if state is not None else np.zeros((1, stride), np.uint32), "scr": scr}
    _, n = flat_program.spec_chunk_ptx(k, 0, second=second)
    for c in range(n):
        text, _, (bi, bo, bs) = flat_program.spec_chunk_ptx(k, c, with_bases=True, second=second)
        view = {"imm": mem["imm"][bi:], "out": mem["out"][bo:], "state": mem["state"][bs:], "scr": scr}
        run_kernel(text, k, stride // k, view)
    return out[:n_outputs]
