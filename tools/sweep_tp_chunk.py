#!/usr/bin/env python
"""Chunk length of the second specialised chain (CLG_SPEC_TP_CHUNK) and the register threshold from which a program
takes it (CLG_SPEC_TP_MIN_REGS) on flattened array multipliers: python tools/sweep_tp_chunk.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import complogic_b200 as host
from complogic_b200 import circuits

ctx = host.Context(0)
peak = ctx.sm_count * 64 * 1.965e9  # LOP3 per second (one thread, one 32-bit word)
for n, m, lg, chunks in ((64, 12, 24, (0, 3200, 4200, 5000, 6200, -1)), (32, 12, 24, (0, 1600, -1)), (40, 8, 24, (0, 2500, -1)),
                         (48, 8, 23, (0, 2400, -1)), (24, 16, 24, (0, -1)), (16, 16, 24, (0, -1)), (8, 256, 23, (0, -1))):
    circ = circuits.flatten_instances(circuits.array_multiplier(n), m)
    nv = 1 << lg
    stride = host.round_stride(nv)
    imm = torch.empty((circ.n_immediates, stride), dtype=torch.int32, device="cuda")
    out = torch.zeros((len(circ.out_regs), stride), dtype=torch.int32, device="cuda")
    ctx.fill_random(imm.data_ptr(), circ.n_immediates, stride, 5, 0)
    for ch in chunks:
        os.environ.pop("CLG_SPEC_NO_TP", None)
        os.environ.pop("CLG_SPEC_TP_CHUNK", None), os.environ.pop("CLG_SPEC_TP_MIN_REGS", None)
        if ch > 0:
            os.environ["CLG_SPEC_TP_CHUNK"], os.environ["CLG_SPEC_TP_MIN_REGS"] = str(ch), "0"
        elif ch == 0:
            os.environ["CLG_SPEC_NO_TP"] = "1"
        try:
            cs = host.CudaSimulation.from_simulation(circ.sim, circ.n_immediates, circ.out_regs, ctx=ctx, mode=host.MODE_SPEC)
            for _ in range(3):
                cs.run_device(imm.data_ptr(), out.data_ptr(), nv, stride)
        except host.ClgError as e:
            print(f"mul{n} x {m} chunk {ch}: {e}")
            continue
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            cs.run_device(imm.data_ptr(), out.data_ptr(), nv, stride)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        lop3 = cs.info["n_luts"] * (nv / 32) / (ms * 1e-3)
        print(f"mul{n} x {m:2d}  2^{lg}  {'chunk %5d' % ch if ch > 0 else 'library    ' if ch else 'first chain'}  K={cs.words_per_thread}  live {cs.info['lane_slots']:3d}  launches {cs.launches_per_run:2d}  "
              f"{ms:7.3f} ms  LOP3 issued {lop3 / peak * 100:5.1f} % of peak  {cs.kernel_name}", flush=True)
