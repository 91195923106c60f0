#!/usr/bin/env python
"""Bandwidth of the vector-major <-> bit-sliced transposes (clg_slice_bools / clg_unslice_bools)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import complogic_b200 as host

ctx = host.Context(0)
for nv, rows in ((1 << 24, 32), (1 << 22, 64), (1 << 22, 128), (1 << 20, 1024), (1 << 18, 4096), (1 << 24, 17), (1 << 22, 33)):
    stride = host.round_stride(nv)
    bools = torch.randint(0, 2, (nv, rows), dtype=torch.uint8, device="cuda")
    sliced = torch.zeros((rows, stride), dtype=torch.int32, device="cuda")
    back = torch.zeros_like(bools)
    s = torch.cuda.current_stream().cuda_stream
    for name, fn in (("slice", lambda: ctx.slice_bools(bools.data_ptr(), sliced.data_ptr(), nv, rows, stride, s)),
                     ("unslice", lambda: ctx.unslice_bools(sliced.data_ptr(), back.data_ptr(), nv, rows, stride, s))):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        byt = nv * rows + rows * stride * 4
        print(f"{name:8s} nv=2^{nv.bit_length() - 1} rows={rows:5d}: {ms:8.3f} ms  {byt / ms / 1e6:8.1f} GB/s")
    assert torch.equal(back, bools), "round trip mismatch"
print("ok")
