#!/usr/bin/env python
"""Dump the shared-memory slot indices of a row stream for tools/microbench_replay.cu:
uint16 [n_rows][4 (a, b, c, dst)][32 lanes], 0xFFFF = the lane does not access shared memory for that operand."""
import sys
import numpy as np
sys.path.insert(0, ".")
from tests import flat_interp as fi
import complogic_b200 as host
from complogic_b200 import circuits

circ = circuits.random_dag(256, 4096, 1024, 2, 1)
fp = host.FlatProgram.levelize(circ.sim, circ.n_immediates, circ.out_regs)
ins = fi.Flat(fp.to_bytes()).rows["instr"].reshape(-1, 32)
lut = ins["kind"] == fi.K_LUT
nfan = (ins["flags"] >> 8) & 3
act = {"a": lut & ((ins["flags"] & fi.FLAG_FROM_REG) == 0), "b": lut & (nfan >= 2), "c": lut & (nfan >= 3), "dst": lut & ((ins["flags"] & fi.FLAG_NO_STORE) == 0)}
out = np.full((len(ins), 4, 32), 0xFFFF, np.uint16)
for k, name in enumerate(("a", "b", "c", "dst")):
    out[:, k, :] = np.where(act[name], ins[name], 0xFFFF)
limit = int(sys.argv[2]) if len(sys.argv) > 2 else len(out)
out[:limit].tofile(sys.argv[1] if len(sys.argv) > 1 else "tools/_data/dag1m_rows.bin")
print(len(out), "rows,", limit, "written")
