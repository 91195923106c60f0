#!/usr/bin/env python
"""The second specialised chain (one instance of a flattened program per kernel, instances side by side) against the
first (as many instances per kernel as fit 16 384 instructions) from one vector to 2^22: python tools/bench_second_chain_batches.py
CLG_SPEC_TP_BIG_ONLY=1 restricts the second chain to big batches, which is what the first column measures."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import complogic_b200 as host
from complogic_b200 import circuits

ctx = host.Context(0)
for n, m in ((16, 16), (32, 12), (8, 256), (24, 16)):
    circ = circuits.flatten_instances(circuits.array_multiplier(n), m)
    for lg in (0, 10, 14, 17, 20, 22):
        nv = 1 << lg
        stride = host.round_stride(nv)
        imm = torch.empty((circ.n_immediates, stride), dtype=torch.int32, device="cuda")
        out = torch.zeros((len(circ.out_regs), stride), dtype=torch.int32, device="cuda")
        ctx.fill_random(imm.data_ptr(), circ.n_immediates, stride, 5, 0)
        res = []
        for big_only in (True, False):
            os.environ.pop("CLG_SPEC_TP_BIG_ONLY", None)
            if big_only:
                os.environ["CLG_SPEC_TP_BIG_ONLY"] = "1"
            cs = host.CudaSimulation.from_simulation(circ.sim, circ.n_immediates, circ.out_regs, ctx=ctx, mode=host.MODE_SPEC)
            for _ in range(3):
                cs.run_device(imm.data_ptr(), out.data_ptr(), nv, stride)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                cs.run_device(imm.data_ptr(), out.data_ptr(), nv, stride)
            e1.record()
            torch.cuda.synchronize()
            res.append((e0.elapsed_time(e1) / 20, cs.launches_per_run, cs.kernel_name))
        print(f"mul{n} x {m} 2^{lg}: first chain {res[0][0]*1e3:8.1f} us ({res[0][1]} launches {res[0][2]})  second chain {res[1][0]*1e3:8.1f} us ({res[1][1]} launches)", flush=True)
