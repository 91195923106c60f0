"""Small end-to-end run of every kernel family for compute-sanitizer (one tool per gpurun call):

    compute-sanitizer --tool memcheck  python tests/sanitize_gpu.py
    compute-sanitizer --tool racecheck python tests/sanitize_gpu.py

Checks the results against the oracle as well, so a clean sanitizer report is a report about a correct run."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import complogic_b200 as host
from complogic_b200 import circuits
from oracle import oracle
from tests.helpers import mask_tail, random_sliced, stride_for

ctx = host.Context(0)


def check(circ, nv, mode, flags=0, expect=None):
    imm = random_sliced(circ.n_immediates, stride_for(nv), seed=3)
    want = mask_tail(oracle.run_batch_sliced(circ.sim, len(circ.sim.registers), imm, nv, circ.out_regs), nv)
    cs = host.CudaSimulation.from_simulation(circ.sim, circ.n_immediates, circ.out_regs, ctx=ctx, mode=mode, flags=flags)
    got = mask_tail(cs.run_batch(imm, nv), nv)
    assert np.array_equal(got, want), (circ.name, cs.kernel_name)
    assert expect is None or cs.kernel_name.startswith(expect), cs.kernel_name
    print(f"{circ.name:14s} {nv:6d} vectors  {cs.kernel_name:16s} ok", flush=True)


dag = circuits.random_dag(40, 512, 128, 2, 5)
check(dag, 148 * 256, host.MODE_COOP, expect="clg_coop_rows")  # row kernel: mbarrier-ordered shared-memory traffic
check(dag, 300, host.MODE_COOP)                                 # resident cooperative kernel (bulk copy + barrier per level)
os.environ["CLG_COOP_NO_ROWS"] = "1"
os.environ["CLG_COOP_NO_RESIDENT"] = "1"
check(dag, 5000, host.MODE_COOP, expect="clg_coop_c")           # streaming compact cooperative kernel
mul = circuits.array_multiplier(8)
check(mul, 5000, host.MODE_LANE, expect="clg_lane")             # lane interpreter (bulk-copy instruction ring)
check(mul, 5000, host.MODE_SPEC, expect="clg_spec")             # generated straight-line kernel
print("sanitize run ok")
