"""Assembly (driver JIT) time of a chunked specialised program that has no two equal chunks:
CUDA_CACHE_DISABLE=1 python tools/bench_jit.py
(An earlier version assembled on CLG_JIT_THREADS host threads: 12.7 s on one, 14.2 s on sixteen -- the driver
serialises the calls, so the pool was removed; profiles/r01_bench_jit_threads.txt.)"""
import os
import sys
import time

sys.path.insert(0, ".")
import complogic_b200 as host
from complogic_b200 import circuits

circ = circuits.random_dag(levels=6000, width=48, n_immediates=64, window=2, seed=5)
t0 = time.time()
cs = host.CudaSimulation.from_simulation(circ.sim, 64, circ.out_regs, mode=host.MODE_SPEC)
print(f"lane_instrs={cs.info['lane_instrs']} kernels={cs.launches_per_run} "
      f"lower+assemble={time.time() - t0:.2f}s", flush=True)
