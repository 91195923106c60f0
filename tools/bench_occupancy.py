#!/usr/bin/env python
"""How much of the LOP3 pipe a specialised straight-line kernel reaches as a function of warps per scheduler: array
multipliers of growing width (live registers grow with the width, so registers per thread and with them the resident
warps change) flattened to about the same 1M gates, words per thread forced: python tools/bench_occupancy.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import complogic_b200 as host
from complogic_b200 import circuits

ctx = host.Context(0)
clock_hz = 1.965e9
peak = ctx.sm_count * 64 * clock_hz  # LOP3 / s
for n, m, lg in ((16, 16, 24), (24, 16, 24), (32, 12, 24), (40, 8, 24), (48, 8, 23), (64, 4, 23)):
    circ = circuits.flatten_instances(circuits.array_multiplier(n), m)
    nv = 1 << lg
    stride = host.round_stride(nv)
    imm = torch.empty((circ.n_immediates, stride), dtype=torch.int32, device="cuda")
    out = torch.zeros((len(circ.out_regs), stride), dtype=torch.int32, device="cuda")
    ctx.fill_random(imm.data_ptr(), circ.n_immediates, stride, 5, 0)
    for k in (1, 2):
        os.environ["CLG_SPEC_K"] = str(k)
        try:
            cs = host.CudaSimulation.from_simulation(circ.sim, circ.n_immediates, circ.out_regs, ctx=ctx, mode=host.MODE_SPEC)
            for _ in range(3):
                cs.run_device(imm.data_ptr(), out.data_ptr(), nv, stride)
        except host.ClgError as e:
            print(f"mul{n} x {m} K={k}: {e}")
            continue
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            cs.run_device(imm.data_ptr(), out.data_ptr(), nv, stride)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        lop3 = cs.info["n_luts"] * (nv / 32) / (ms * 1e-3)  # a LOP3 = one thread, one 32-bit word
        print(f"mul{n} x {m:2d}  2^{lg}  K={k}  live {cs.info['lane_slots']:3d} ({cs.info['lane_slots'] * k:3d} registers)  launches {cs.launches_per_run:2d}  "
              f"{ms:7.3f} ms  LOP3 issued {lop3 / peak * 100:5.1f} % of peak  kernel {cs.kernel_name}", flush=True)
