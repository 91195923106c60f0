#!/usr/bin/env python
"""A wide sequential circuit (n D latches in a row, ~10 instructions and 4 carried registers per latch) on a big
lane batch, one step per call, in every execution mode: python tools/bench_latch_chain.py [n] [log2 lanes]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import complogic_b200 as host

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
lg = int(sys.argv[2]) if len(sys.argv) > 2 else 20
c = host.Compiler(2)
qs = [c.alloc() for _ in range(n)]
sim = c.compile([host.DLatch(0 if i == 0 else qs[i - 1], 1, qs[i]) for i in range(n)])
ctx = host.Context(0)
nv = 1 << lg
stride = host.round_stride(nv)
imm = torch.empty((2, stride), dtype=torch.int32, device="cuda")
ctx.fill_random(imm.data_ptr(), 2, stride, 7)
ref = None
for mode in (host.MODE_SPEC, host.MODE_COOP, host.MODE_LANE):
    try:
        cs = host.CudaSimulation.from_simulation(sim, 2, qs, flags=host.LV_STATEFUL, ctx=ctx, mode=mode)
    except host.ClgError as e:
        print(host.MODE_NAMES[mode], "unavailable:", e)
        continue
    out = torch.zeros((n, stride), dtype=torch.int32, device="cuda")
    state = torch.zeros((cs.info["n_state"], stride), dtype=torch.int32, device="cuda")
    for _ in range(3):
        cs.run_device(imm.data_ptr(), out.data_ptr(), nv, stride, state_dptr=state.data_ptr())
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    ref = got if ref is None else ref
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        cs.run_device(imm.data_ptr(), out.data_ptr(), nv, stride, state_dptr=state.data_ptr())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    rows = 2 + n + 2 * cs.info["n_state"]
    print(f"latch_chain n={n} lanes=2^{lg} mode={cs.mode_name} K={cs.words_per_thread} launches={cs.launches_per_run} "
          f"slots(lane/level)={cs.info['lane_slots']}/{cs.info['level_slots']} ms={ms:.3f} "
          f"gate-evals/s={cs.info['ref_nand_count'] * nv / (ms * 1e-3):.3e} HBM={rows * stride * 4 / (ms * 1e-3) / 1e9:.0f} GB/s "
          f"same_bits={np.array_equal(got, ref)}", flush=True)
