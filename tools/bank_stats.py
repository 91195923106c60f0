#!/usr/bin/env python
"""Static shared-memory bank-conflict estimate of a level-ordered stream for the cooperative kernel at K=4:
for every quarter-warp (8 consecutive instructions of a level) the number of wavefronts each LDS.128/STS.128
needs = the largest number of its addresses that share slot % 8."""
import sys
import numpy as np
sys.path.insert(0, ".")
from tests import flat_interp as fi


def stats(flat: "fi.Flat"):
    st = flat.streams[fi.STREAM_LEVEL]
    ins, lv = st["instr"], st["levels"]
    tot = {"a": 0, "b": 0, "c": 0, "dst": 0}
    ideal = {"a": 0, "b": 0, "c": 0, "dst": 0}
    for l in range(st["n_levels"]):
        blk = ins[lv[l]:lv[l + 1]]
        n = len(blk)
        pad = (-n) % 8
        kind = np.concatenate([blk["kind"], np.full(pad, 255, np.uint8)]).reshape(-1, 8)
        nfan = np.concatenate([(blk["flags"] >> 8) & 3, np.zeros(pad, np.uint16)]).reshape(-1, 8)
        for name, need in (("a", lambda k, f: ((k == 0) & (f >= 1)) | (k == 2)), ("b", lambda k, f: (k == 0) & (f >= 2)),
                           ("c", lambda k, f: (k == 0) & (f >= 3)), ("dst", lambda k, f: (k == 0) | (k == 1))):
            v = np.concatenate([blk[name], np.zeros(pad, np.uint16)]).reshape(-1, 8) % 8
            act = need(kind, nfan)
            cnt = np.zeros((v.shape[0], 8), np.int32)
            for b in range(8):
                cnt[:, b] = ((v == b) & act).sum(axis=1)
            tot[name] += int(cnt.max(axis=1).sum())
            ideal[name] += int(act.any(axis=1).sum())
    return tot, ideal


if __name__ == "__main__":
    import complogic_b200 as host
    from complogic_b200 import circuits
    circ = circuits.random_dag()
    fp = host.FlatProgram.levelize(circ.sim, circ.n_immediates, circ.out_regs)
    print(fp.info)
    tot, ideal = stats(fi.Flat(fp.to_bytes()))
    for k in tot:
        print(f"{k:4s} wavefronts {tot[k]:9d}  ideal {ideal[k]:9d}  ratio {tot[k] / max(1, ideal[k]):.3f}")
    print("total", sum(tot.values()), "ideal", sum(ideal.values()), "ratio", sum(tot.values()) / sum(ideal.values()))
