"""Build (assemble) time and run time of the chunked specialised mode against the interpreters, on flattened
64x64 multipliers: python tools/bench_chunked.py [instances ...]"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import complogic_b200 as host
from complogic_b200 import circuits

LOG2 = 18
for m in [int(a) for a in sys.argv[1:]] or [8, 32]:
    circ = circuits.flatten_instances(circuits.array_multiplier(64), m)
    nv = 1 << LOG2
    stride = nv // 32
    d_imm = torch.randint(-2 ** 31, 2 ** 31 - 1, (circ.n_immediates, stride), dtype=torch.int32, device="cuda")
    d_out = torch.zeros((len(circ.out_regs), stride), dtype=torch.int32, device="cuda")
    ref = None
    for mode in (host.MODE_SPEC, host.MODE_AUTO, host.MODE_COOP, host.MODE_LANE):
        t0 = time.time()
        cs = host.CudaSimulation.from_simulation(circ.sim, circ.n_immediates, circ.out_regs, mode=mode)
        t1 = time.time()
        cs.run_device(d_imm.data_ptr(), d_out.data_ptr(), nv, stride)
        torch.cuda.synchronize()
        t2 = time.time()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            cs.run_device(d_imm.data_ptr(), d_out.data_ptr(), nv, stride)
        e1.record()
        torch.cuda.synchronize()
        got = d_out.cpu().numpy()
        ref = got if ref is None else ref
        print(f"mul64x{m} 2^{LOG2} requested={host.MODE_NAMES[mode]} ran={host.MODE_NAMES[cs.mode]} K={cs.words_per_thread} "
              f"launches={cs.launches_per_run} lower={t1 - t0:.2f}s first_run(build)={t2 - t1:.2f}s "
              f"ms={e0.elapsed_time(e1) / 3:.3f} same_bits={np.array_equal(got, ref)}", flush=True)
        del cs
