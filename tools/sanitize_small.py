#!/usr/bin/env python
"""Small, fast exercise of every kernel for compute-sanitizer (memcheck / racecheck / synccheck):
all three execution modes on a multiplier, a small DAG through the cooperative kernel, a stateful latch, the
transposes and the generator. Exits non-zero on any mismatch with the oracle."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import complogic_b200 as host
from complogic_b200 import circuits
from oracle import oracle
from tests.helpers import mask_tail, random_sliced, stride_for


def check(circ, nv, mode, flags=0):
    imm = random_sliced(circ.n_immediates, stride_for(nv), seed=1)
    want = mask_tail(oracle.run_batch_sliced(circ.sim, len(circ.sim.registers), imm, nv, circ.out_regs), nv)
    cs = host.CudaSimulation.from_simulation(circ.sim, circ.n_immediates, circ.out_regs, flags=flags, mode=mode)
    assert np.array_equal(mask_tail(cs.run_batch(imm, nv), nv), want), (circ.name, mode)


mul = circuits.array_multiplier(8)
for mode in (host.MODE_SPEC, host.MODE_LANE, host.MODE_COOP):
    check(mul, 3000, mode)
check(circuits.random_dag(levels=12, width=300, n_immediates=64, window=2, seed=3), 1000, host.MODE_COOP)
check(circuits.ripple_adder(16), 5000, host.MODE_LANE)
c = host.Compiler(2)
sim = c.compile([host.DLatch(0, 1, c.alloc())])
for mode in (host.MODE_SPEC, host.MODE_LANE, host.MODE_COOP):
    cs = host.CudaSimulation.from_simulation(sim, 2, None, flags=host.LV_STATEFUL, mode=mode)
    state = np.zeros((cs.info["n_state"], 4), np.uint32)
    for s in range(3):
        imm = random_sliced(2, 4, 9 + s)
        cs.run_batch(imm, 100, state=state)
cs = host.CudaSimulation.from_simulation(mul.sim, 16, mul.out_regs)
out = cs.run_bools(np.random.default_rng(0).integers(0, 2, size=(77, 16), dtype=np.uint8))
assert out.shape == (77, 16)
print("sanitize_small: ok")
