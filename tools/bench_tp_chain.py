#!/usr/bin/env python
"""The second specialised chain (chunks below the instruction-cache cliff) against the default one on register-heavy
programs at a big batch: python tools/bench_tp_chain.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import complogic_b200 as host
from complogic_b200 import circuits

ctx = host.Context(0)
for name, circ, lg in (("dag210x160 (225 live, 19 k instr)", circuits.random_dag(210, 160, 64, 2, 7), 23),
                       ("dag180x180 (248 live, 18 k instr)", circuits.random_dag(180, 180, 64, 2, 7), 23),
                       ("mul64 (222 live, 12 k instr)", circuits.array_multiplier(64), 24),
                       ("mul64 x 4", circuits.flatten_instances(circuits.array_multiplier(64), 4), 24)):
    nv = 1 << lg
    stride = host.round_stride(nv)
    imm = torch.empty((circ.n_immediates, stride), dtype=torch.int32, device="cuda")
    out = torch.zeros((len(circ.out_regs), stride), dtype=torch.int32, device="cuda")
    ctx.fill_random(imm.data_ptr(), circ.n_immediates, stride, 5, 0)
    for env in (None, "CLG_SPEC_NO_TP"):
        if env:
            os.environ[env] = "1"
        cs = host.CudaSimulation.from_simulation(circ.sim, circ.n_immediates, circ.out_regs, ctx=ctx, mode=host.MODE_SPEC)
        for _ in range(3):
            cs.run_device(imm.data_ptr(), out.data_ptr(), nv, stride)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            cs.run_device(imm.data_ptr(), out.data_ptr(), nv, stride)
        e1.record()
        torch.cuda.synchronize()
        print(f"{name:36s} 2^{lg}  {'default chain' if env else 'second chain ':14s} launches {cs.launches_per_run:3d}  {e0.elapsed_time(e1) / 10:8.3f} ms", flush=True)
        if env:
            del os.environ[env]
