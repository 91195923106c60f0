#!/usr/bin/env python
"""Fused multi-step (clg_exec_run_steps, one launch, carried registers in registers) vs one launch per step."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import complogic_b200 as host

# a small sequential circuit: 8 D-latches fed by a ripple adder of the latch outputs and 8 immediates
n = 8
c = host.Compiler(n + 1)
q = [c.alloc() for _ in range(n)]
s = [c.alloc() for _ in range(n)]
k = [c.alloc() for _ in range(n)]
gates = [host.FullAdder(i, q[i], n if i == 0 else k[i - 1], s[i], k[i]) for i in range(n)]
gates += [host.DLatch(s[i], n, q[i]) for i in range(n)]
sim = c.compile(gates)
ctx = host.Context(0)
cs = host.CudaSimulation.from_simulation(sim, n + 1, q, flags=host.LV_STATEFUL, ctx=ctx)
print("program:", cs.info, "mode", cs.mode_name)
nv, T = 1 << 22, 64
stride = host.round_stride(nv)
imm = torch.empty((T, n + 1, stride), dtype=torch.int32, device="cuda")
ctx.fill_random(imm.data_ptr(), T * (n + 1), stride, 7)
out = torch.zeros((T, n, stride), dtype=torch.int32, device="cuda")
state = torch.zeros((cs.info["n_state"], stride), dtype=torch.int32, device="cuda")


def fused():
    cs.run_steps_device(imm.data_ptr(), out.data_ptr(), state.data_ptr(), nv, stride, T)


def looped():
    for t in range(T):
        cs.run_device(imm[t].data_ptr(), out[t].data_ptr(), nv, stride, state_dptr=state.data_ptr())


res = {}
for name, fn in (("looped", looped), ("fused", fused)):
    state.zero_()
    fn()
    torch.cuda.synchronize()
    res[name] = out.clone()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{name:7s}: {ms:8.3f} ms for {T} steps x 2^22 lanes  ({cs.info['ref_nand_count'] * nv * T / (ms * 1e-3):.3e} gate-evals/s)")
assert torch.equal(res["looped"], res["fused"])
print("ok: identical outputs")
