#!/usr/bin/env python
"""Static profile of a flat buffer's row stream (clg_coop_rows): LSU wavefronts per row (instruction fetch, loads,
stores, bank-conflict replays per quarter-warp), lanes without an instruction, rows per warp and segment."""
import sys
import numpy as np
sys.path.insert(0, ".")
from tests import flat_interp as fi


def stats(flat: "fi.Flat"):
    r = flat.rows
    ins = r["instr"].reshape(-1, 32)
    lut = ins["kind"] == fi.K_LUT
    nfan = (ins["flags"] >> 8) & 3
    from_reg = (ins["flags"] & fi.FLAG_FROM_REG) != 0
    no_store = (ins["flags"] & fi.FLAG_NO_STORE) != 0
    n_rows = len(ins)
    out = {"rows": n_rows, "luts": int(lut.sum()), "empty_lane_frac": 1 - lut.sum() / (n_rows * 32.0)}
    wf = {"fetch": 2 * n_rows}
    ideal = {}
    for name, act in (("a", lut & ~from_reg), ("b", lut & (nfan >= 2)), ("c", lut & (nfan >= 3)), ("dst", lut & ~no_store)):
        v = ins[name] % 8
        tot = idl = 0
        for q in range(4):
            a = act[:, 8 * q:8 * q + 8]
            vv = v[:, 8 * q:8 * q + 8]
            cnt = np.zeros((n_rows, 8), np.int32)
            for b in range(8):
                cnt[:, b] = ((vv == b) & a).sum(axis=1)
            tot += int(cnt.max(axis=1).sum())
            idl += int(a.any(axis=1).sum())
        wf[name], ideal[name] = tot, idl
    out["wavefronts"] = wf
    out["ideal"] = ideal
    out["wavefronts_per_row"] = sum(wf.values()) / n_rows
    out["conflict_ratio"] = sum(wf[k] for k in ideal) / max(1, sum(ideal.values()))
    rt = r["rowtab"].astype(np.int64)
    per = (rt[1:] - rt[:-1]).reshape(r["n_segs"], r["n_warps"]) // 32
    out["rows_per_warp_total"] = [int(per[:, w].sum()) for w in (0, r["n_warps"] // 2, r["n_warps"] - 1)]
    out["max_rows_per_warp_segment"] = int(per.max())
    out["distinct_tt_rows"] = {int(k): int(v) for k, v in zip(*np.unique([len(set(row["tt"][row["kind"] == 0])) for row in ins[::7]], return_counts=True))}
    return out


if __name__ == "__main__":
    import complogic_b200 as host
    from complogic_b200 import circuits
    circ = circuits.random_dag()
    fp = host.FlatProgram.levelize(circ.sim, circ.n_immediates, circ.out_regs)
    print(fp.info)
    for k, v in stats(fi.Flat(fp.to_bytes())).items():
        print(k, v)
