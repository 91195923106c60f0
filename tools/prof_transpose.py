#!/usr/bin/env python
"""One shape of the transposes, a few launches (for ncu): python tools/prof_transpose.py [log2 vectors] [rows]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import complogic_b200 as host

ctx = host.Context(0)
nv, rows = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 20), int(sys.argv[2]) if len(sys.argv) > 2 else 1024
stride = host.round_stride(nv)
bools = torch.randint(0, 2, (nv, rows), dtype=torch.uint8, device="cuda")
sliced = torch.zeros((rows, stride), dtype=torch.int32, device="cuda")
for _ in range(3):
    ctx.slice_bools(bools.data_ptr(), sliced.data_ptr(), nv, rows, stride)
    ctx.unslice_bools(sliced.data_ptr(), bools.data_ptr(), nv, rows, stride)
torch.cuda.synchronize()
print("ok")
