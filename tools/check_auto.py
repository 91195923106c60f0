#!/usr/bin/env python
"""Does CLG_MODE_AUTO pick the fastest mode? Times every mode and AUTO over a grid of programs x batch sizes and flags
the cases where AUTO is more than 20 % behind the best pinned mode.   python tools/check_auto.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import complogic_b200 as host
from complogic_b200 import circuits

PROGRAMS = [
    ("adder32", lambda: circuits.ripple_adder(32)),
    ("mul16", lambda: circuits.array_multiplier(16)),
    ("mul64", lambda: circuits.array_multiplier(64)),
    ("mul64x8", lambda: circuits.flatten_instances(circuits.array_multiplier(64), 8)),
    ("mul16x64", lambda: circuits.flatten_instances(circuits.array_multiplier(16), 64)),
    ("dag64k", lambda: circuits.random_dag(levels=64, width=1024, n_immediates=256, window=2, seed=3)),
    ("dag_thin", lambda: circuits.random_dag(levels=1500, width=48, n_immediates=64, window=2, seed=31)),
    ("latch1500", lambda: circuits.latch_chain(1500)),
]
ctx = host.Context(0)
bad = 0
for name, make in PROGRAMS:
    circ = make()
    flags = host.LV_STATEFUL if circ.stateful else 0
    sims = {}
    for mode in (host.MODE_AUTO, host.MODE_SPEC, host.MODE_LANE, host.MODE_COOP):
        try:
            sims[mode] = host.CudaSimulation.from_simulation(circ.sim, circ.n_immediates, circ.out_regs, ctx=ctx, mode=mode, flags=flags)
        except host.ClgError:
            pass
    info = sims[host.MODE_AUTO].info
    for lg in (0, 10, 16, 20, 23):
        nv = 1 << lg
        stride = host.round_stride(nv)
        if (info["n_inputs"] + info["n_outputs"] + info["n_state"]) * stride * 4 > 40e9:
            continue
        imm = torch.empty((max(1, info["n_inputs"]), stride), dtype=torch.int32, device="cuda")
        ctx.fill_random(imm.data_ptr(), info["n_inputs"], stride, 5)
        out = torch.zeros((max(1, info["n_outputs"]), stride), dtype=torch.int32, device="cuda")
        state = torch.zeros((max(1, info["n_state"]), stride), dtype=torch.int32, device="cuda")
        ms, ran = {}, {}
        for mode, cs in sims.items():
            for _ in range(2):
                cs.run_device(imm.data_ptr(), out.data_ptr(), nv, stride, state_dptr=state.data_ptr())
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5
            e0.record()
            for _ in range(reps):
                cs.run_device(imm.data_ptr(), out.data_ptr(), nv, stride, state_dptr=state.data_ptr())
            e1.record()
            torch.cuda.synchronize()
            ms[mode], ran[mode] = e0.elapsed_time(e1) / reps, cs.mode_name
        best = min((m for m in ms if m != host.MODE_AUTO), key=lambda m: ms[m])
        flag = "" if ms[host.MODE_AUTO] <= 1.2 * ms[best] + 0.005 else "   <-- AUTO is behind"
        bad += bool(flag)
        print(f"{name:10s} 2^{lg:<2d} auto->{ran[host.MODE_AUTO]:4s} {ms[host.MODE_AUTO]:9.4f} ms | " +
              " ".join(f"{host.MODE_NAMES[m]}={ms[m]:.4f}" for m in ms if m != host.MODE_AUTO) + flag, flush=True)
        del imm, out, state
print("cases where AUTO is behind:", bad)
