"""Benchmark circuits, built ONLY through the reference's gate API (Compiler + Gate structs) or, for the
synthetic DAG, directly as `Simulation { ops, registers }` (both fields are public in the reference).

Every builder takes the API namespace to build with (the product mirror `complogic_b200` by default;
tests pass the CPU oracle's mirror to build the same circuit on the reference side) and returns a
`Circuit`: the compiled Simulation, the immediates count and the output registers.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np


@dataclass
class Circuit:
    name: str
    sim: object          # api.Simulation
    n_immediates: int
    out_regs: List[int]
    note: str = ""
    stateful: bool = False  # registers carry over from one run() to the next (levelize with LV_STATEFUL)


def _api(api):
    if api is None:
        import complogic_b200 as api
    return api


def ripple_adder(n: int, api=None) -> Circuit:
    """n-bit ripple-carry adder from FullAdders, cin as an immediate (never FourBitAdder: its cin register has
    no writer and compile() panics, gates.rs:421 / compile.rs:137).

    immediates: a_i = reg i, b_i = reg n+i, cin = reg 2n; outputs s_0..s_{n-1}, carry_{n-1}.
    19n NANDs, 21n+1 registers (SURVEY 8 provenance)."""
    api = _api(api)
    c = api.Compiler(2 * n + 1)
    s = [c.alloc() for _ in range(n)]
    k = [c.alloc() for _ in range(n)]
    gates = [api.FullAdder(i, n + i, 2 * n if i == 0 else k[i - 1], s[i], k[i]) for i in range(n)]
    sim = c.compile(gates)
    return Circuit(f"adder{n}", sim, 2 * n + 1, s + [k[n - 1]])


def array_multiplier(n: int, api=None) -> Circuit:
    """n x n row-ripple array multiplier from And / HalfAdder / FullAdder.

    immediates: a_j = reg j, b_i = reg n+i; outputs p_0..p_{2n-1}.
    n^2 Ands + n HalfAdders + n(n-2) FullAdders (n=16: 4 896 NANDs, n=64: 84 096)."""
    api = _api(api)
    c = api.Compiler(2 * n)
    gates = []
    pp = [[0] * n for _ in range(n)]
    for i in range(n):
        for j in range(n):
            pp[i][j] = c.alloc()
            gates.append(api.And(j, n + i, pp[i][j]))
    prod = [pp[0][0]]
    acc = pp[0][1:]          # accumulator bits at weights i+1 .. (n-1 of them before row 1, n afterwards)
    for i in range(1, n):
        carry: Optional[int] = None
        nxt = []
        for j in range(n):
            x = pp[i][j]
            y = acc[j] if j < len(acc) else None
            s, k = c.alloc(), c.alloc()
            if y is None:
                gates.append(api.HalfAdder(x, carry, s, k))      # top bit of row 1: no accumulator bit yet
            elif carry is None:
                gates.append(api.HalfAdder(x, y, s, k))          # column 0: no carry in
            else:
                gates.append(api.FullAdder(x, y, carry, s, k))
            carry = k
            if j == 0:
                prod.append(s)
            else:
                nxt.append(s)
        acc = nxt + [carry]
    prod += acc
    sim = c.compile(gates)
    return Circuit(f"mul{n}", sim, 2 * n, prod)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def random_dag(levels: int = 256, width: int = 4096, n_immediates: int = 1024, window: int = 2, seed: int = 1,
               api=None) -> Circuit:
    """Synthetic random levelized NAND DAG (BASELINE config 4).

    Level 0 is the immediates (Set(i,false), i < n_immediates). The gate g of level l (1-based) is
    Nand(x, y, reg) with x, y drawn uniformly from the registers of levels [l-window, l-1] (window <= 0:
    from all earlier levels), reg = n_immediates + (l-1)*width + g. Draw k of gate (l, g) is
    splitmix64(seed ^ ((l*width + g)*2 + k)). Outputs: the last level."""
    api = _api(api)
    I, W, L = n_immediates, width, levels
    ops = np.zeros(I + L * W, dtype=api.OP_DTYPE)
    ops["tag"][:I] = 1
    ops["a"][:I] = np.arange(I)
    g = np.arange(W, dtype=np.uint64)
    for l in range(1, L + 1):
        lo = 0 if window <= 0 else max(0, l - window)
        # candidate registers: level 0 has I registers, level k>=1 has W starting at I + (k-1)*W
        first = 0 if lo == 0 else I + (lo - 1) * W
        end = I + (l - 1) * W
        span = np.uint64(end - first)
        ctr = (np.uint64(l) * np.uint64(W) + g) * np.uint64(2)
        x = _splitmix64(np.uint64(seed) ^ ctr) % span + np.uint64(first)
        y = _splitmix64(np.uint64(seed) ^ (ctr + np.uint64(1))) % span + np.uint64(first)
        sl = slice(I + (l - 1) * W, I + l * W)
        ops["tag"][sl] = 0
        ops["a"][sl], ops["b"][sl] = x, y
        ops["c"][sl] = np.arange(I + (l - 1) * W, I + l * W)
    n_regs = I + L * W
    sim = _sim_from_array(api, ops, n_regs)
    outs = list(range(I + (L - 1) * W, I + L * W))
    return Circuit(f"dag{L}x{W}k{window}", sim, I, outs, note=f"seed={seed}")


def _sim_from_array(api, ops: np.ndarray, n_regs: int):
    return api.Simulation(registers=[False] * n_regs, _ops_array=ops)


def flatten_instances(base: Circuit, copies: int, api=None) -> Circuit:
    """M independent instances of `base` flattened into one program with distinct registers (BASELINE
    config 5): instance m's immediates are registers [m*I, (m+1)*I), its other registers follow all immediates."""
    api = _api(api)
    src = base.sim.ops_array if hasattr(base.sim, "ops_array") else api.ops_to_array(base.sim.ops)
    I, R = base.n_immediates, len(base.sim.registers)
    sets, nands = src[src["tag"] == 1], src[src["tag"] == 0]
    assert (sets["a"] < I).all(), "flatten_instances expects compile()-shaped programs (all Sets are immediates)"
    M = copies
    out = np.zeros(M * len(src), dtype=src.dtype)

    def remap(r, m):
        r = r.astype(np.int64)
        return np.where(r < I, m * I + r, M * I + m * (R - I) + (r - I)).astype(np.uint64)

    pos = 0
    for m in range(M):
        blk = sets.copy()
        blk["a"] = remap(sets["a"], m)
        out[pos:pos + len(blk)] = blk
        pos += len(blk)
    for m in range(M):
        blk = nands.copy()
        for f in ("a", "b", "c"):
            blk[f] = remap(nands[f], m)
        out[pos:pos + len(blk)] = blk
        pos += len(blk)
    n_regs = M * R
    sim = _sim_from_array(api, out, n_regs)
    outs = []
    for m in range(M):
        outs += [int(x) for x in remap(np.array(base.out_regs), m)]
    return Circuit(f"{base.name}x{M}", sim, M * I, outs)


def latch_chain(n: int, api=None) -> Circuit:
    """A sequential circuit: n D latches in a row behind one enable (immediates: data, enable); latch i's data input
    is latch i-1's output. 16 NANDs and 4 carried registers per latch with the reference's DLatch expansion."""
    api = _api(api)
    c = api.Compiler(2)
    qs = [c.alloc() for _ in range(n)]
    sim = c.compile([api.DLatch(0 if i == 0 else qs[i - 1], 1, qs[i]) for i in range(n)])
    return Circuit(f"latch{n}", sim, 2, qs, note="D-latch chain", stateful=True)
