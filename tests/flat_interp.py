"""TEST INFRASTRUCTURE: numpy interpreter of the flat device buffer (complogic_b200/csrc/flat.hpp).

Lets the CPU-only suite check the levelize/register-allocate pass end to end (flat buffer vs the
oracle VM) without a GPU. The level stream is evaluated with the semantics the cooperative kernel
has: all instructions of a level read the register stack as it was when the level started.
"""
import struct

import numpy as np

INSTR = np.dtype([("dst", "<u2"), ("a", "<u2"), ("b", "<u2"), ("c", "<u2"), ("tt", "u1"), ("kind", "u1"),
                  ("flags", "<u2"), ("row", "<u4")])
assert INSTR.itemsize == 16
K_LUT, K_LDIN, K_STOUT, K_STCONST, K_NOP = range(5)
FLAG_FROM_REG, FLAG_NO_STORE = 2, 4
STREAM_LEVEL, STREAM_LANE = 0, 1


class Flat:
    def __init__(self, data: bytes):
        self.data = data
        (self.magic, self.version, self.header_bytes, self.total_bytes, self.n_inputs, self.n_outputs, self.n_state,
         self.n_luts) = struct.unpack_from("<8I", data, 0)
        self.ref_nand_count, self.ref_set_count = struct.unpack_from("<2Q", data, 32)
        self.flags, self.n_registers, self.off_state_regs, self.off_out_regs = struct.unpack_from("<4I", data, 48)
        assert self.magic == 0x46474C43 and self.total_bytes == len(data)
        self.streams = []
        for s in range(2):
            n_slots, n_instr, n_levels, off_levels, off_instr, max_w = struct.unpack_from("<6I", data, 64 + 32 * s)
            levels = np.frombuffer(data, dtype="<u4", count=n_levels + 1, offset=off_levels) if n_instr else None
            instr = np.frombuffer(data, dtype=INSTR, count=n_instr, offset=off_instr) if n_instr else None
            self.streams.append(dict(n_slots=n_slots, n_instr=n_instr, n_levels=n_levels, levels=levels, instr=instr,
                                     max_level_width=max_w))
        self.n_parts, self.off_parts = struct.unpack_from("<2I", data, 128)
        self.parts = []
        for q in range(self.n_parts):
            n_slots, n_instr, n_levels, off_levels, off_instr, max_w = struct.unpack_from("<6I", data, self.off_parts + 32 * q)
            self.parts.append(dict(n_slots=n_slots, n_instr=n_instr, n_levels=n_levels, max_level_width=max_w,
                                   levels=np.frombuffer(data, dtype="<u4", count=n_levels + 1, offset=off_levels),
                                   instr=np.frombuffer(data, dtype=INSTR, count=n_instr, offset=off_instr)))
        # row stream (FlatRows, right behind the part table fields)
        (n_slots, n_instr, n_segs, n_warps, lag, off_rowtab, off_instr, n_side, off_sidetab, off_side, n_from_reg,
         n_no_store) = struct.unpack_from("<12I", data, 144)
        self.rows = None
        if n_instr:
            self.rows = dict(n_slots=n_slots, n_instr=n_instr, n_segs=n_segs, n_warps=n_warps, lag=lag, n_side=n_side,
                             n_from_reg=n_from_reg, n_no_store=n_no_store,
                             rowtab=np.frombuffer(data, dtype="<u4", count=n_segs * n_warps + 1, offset=off_rowtab),
                             instr=np.frombuffer(data, dtype=INSTR, count=n_instr, offset=off_instr),
                             sidetab=np.frombuffer(data, dtype="<u4", count=n_segs + 1, offset=off_sidetab),
                             side=np.frombuffer(data, dtype=INSTR, count=n_side, offset=off_side))
        self.state_regs = np.frombuffer(data, dtype="<u8", count=self.n_state, offset=self.off_state_regs)
        self.out_regs = np.frombuffer(data, dtype="<u8", count=self.n_outputs, offset=self.off_out_regs)


def _lut(tt, a, b, c):
    """tt: [n] uint8, a/b/c: [n][W] uint32."""
    res = np.zeros_like(a)
    full = np.uint32(0xFFFFFFFF)
    for k in range(8):
        sel = (((tt >> k) & 1).astype(np.uint32) * full)[:, None]
        t = (a if k & 4 else ~a) & (b if k & 2 else ~b) & (c if k & 1 else ~c)
        res |= t & sel
    return res


STREAM_PARTS = 2  # pseudo stream id: run every part of the partitioned level stream, one after the other
STREAM_ROWS = 3   # the row stream of the cooperative row kernel


def run_rows(flat: Flat, imm: np.ndarray, state: np.ndarray = None, ahead: int = 0, rng=None):
    """The row stream with the kernel's semantics: every lane keeps its last result in a register (FLAG_FROM_REG reads
    it, FLAG_NO_STORE results exist nowhere else), a warp's rows run one after the other, and warps are only loosely
    synchronised: a warp may be `lag - 1` segments ahead of another. `ahead` picks which warps run ahead (warps with
    index % 2 == ahead - 1; 0: nobody), so a schedule that relies on more ordering than the kernel gives fails here."""
    r = flat.rows
    assert r is not None, "row stream absent"
    W = imm.shape[1]
    nw, lag = r["n_warps"], r["lag"]
    slots = np.zeros((r["n_slots"], W), dtype=np.uint32)
    regs = np.zeros((nw * 32, W), dtype=np.uint32)
    out = np.zeros((flat.n_outputs, W), dtype=np.uint32)
    ins, rowtab, side, sidetab = r["instr"], r["rowtab"], r["side"], r["sidetab"]
    full = np.uint32(0xFFFFFFFF)

    def in_row(q):
        return imm[q] if q < flat.n_inputs else state[q - flat.n_inputs]

    def st_row(q, v):
        if q < flat.n_outputs:
            out[q] = v
        else:
            state[q - flat.n_outputs] = v

    def run_warp_segment(w, sgm):
        b, e = int(rowtab[sgm * nw + w]), int(rowtab[sgm * nw + w + 1])
        for r0 in range(b, e, 32):
            row = ins[r0:r0 + 32]
            act = np.nonzero(row["kind"] == K_LUT)[0]
            if len(act) == 0:
                continue
            x = row[act]
            lanes = w * 32 + act
            from_reg = (x["flags"] & FLAG_FROM_REG) != 0
            a = np.where(from_reg[:, None], regs[lanes], slots[x["a"]])
            res = _lut(x["tt"], a, slots[x["b"]], slots[x["c"]])
            regs[w * 32:(w + 1) * 32] = np.uint32(0xDEADBEEF)  # the row's LOP3s run in every lane: idle lanes lose their register
            regs[lanes] = res
            st = (x["flags"] & FLAG_NO_STORE) == 0
            slots[x["dst"][st]] = res[st]
        # this warp's share of the segment's side list: entry i belongs to thread i % (32 * nw)
        for i in range(int(sidetab[sgm]), int(sidetab[sgm + 1])):
            if ((i - int(sidetab[sgm])) % (32 * nw)) // 32 != w:
                continue
            y = side[i]
            k = y["kind"]
            if k == K_LUT:
                v = _lut(np.array([y["tt"]], np.uint8), slots[y["a"]][None], slots[y["b"]][None], slots[y["c"]][None])[0]
                slots[y["dst"]] = v
            elif k == K_LDIN:
                slots[y["dst"]] = in_row(y["row"])
            elif k == K_STOUT:
                st_row(y["row"], slots[y["a"]] ^ (full if y["flags"] & 1 else np.uint32(0)))
            else:
                st_row(y["row"], np.full(W, 0xFFFFFFFF if y["flags"] & 1 else 0, np.uint32))

    fast = [w for w in range(nw) if ahead and w % 2 == ahead - 1]
    slow = [w for w in range(nw) if w not in fast]
    skew = lag - 1 if fast else 0
    for t in range(r["n_segs"] + skew):
        order = []
        if t < r["n_segs"]:
            order += [(w, t) for w in fast]
        if t - skew >= 0:
            order += [(w, t - skew) for w in slow]
        if rng is not None:  # warps of one time step run in any order -- except that the fast ones go first
            pass
        for w, sgm in order:
            run_warp_segment(w, sgm)
    return out


def run(flat: Flat, stream: int, imm: np.ndarray, state: np.ndarray = None, rng=None, _part=None, _out=None):
    """imm [n_inputs][W]; state [n_state][W] is updated in place. Returns out [n_outputs][W]."""
    W = imm.shape[1]
    if stream == STREAM_ROWS:
        return run_rows(flat, imm, state, ahead=0 if rng is None else int(rng.integers(0, 3)), rng=rng)
    if stream == STREAM_PARTS:
        assert flat.n_parts, "no partitioned stream"
        out = np.zeros((flat.n_outputs, W), dtype=np.uint32)
        order = list(range(flat.n_parts))
        if rng is not None:
            order = [int(x) for x in rng.permutation(flat.n_parts)]
        for q in order:  # parts are independent: any order (the device runs them concurrently)
            run(flat, STREAM_LEVEL, imm, state, rng, _part=flat.parts[q], _out=out)
        return out
    st = flat.streams[stream] if _part is None else _part
    assert st["n_instr"], "stream absent"
    slots = np.zeros((st["n_slots"], W), dtype=np.uint32)
    out = np.zeros((flat.n_outputs, W), dtype=np.uint32) if _out is None else _out
    ins, lv = st["instr"], st["levels"]

    def in_row(r):
        return imm[r] if r < flat.n_inputs else state[r - flat.n_inputs]

    def st_row(r, v):
        if r < flat.n_outputs:
            out[r] = v
        else:
            state[r - flat.n_outputs] = v

    if stream == STREAM_LANE:
        for x in ins:
            k = x["kind"]
            if k == K_LUT:
                slots[x["dst"]] = _lut(np.array([x["tt"]], np.uint8), slots[x["a"]][None], slots[x["b"]][None],
                                       slots[x["c"]][None])[0]
            elif k == K_LDIN:
                slots[x["dst"]] = in_row(x["row"])
            elif k == K_STOUT:
                st_row(x["row"], slots[x["a"]] ^ np.uint32(0xFFFFFFFF if x["flags"] & 1 else 0))
            else:
                st_row(x["row"], np.full(W, 0xFFFFFFFF if x["flags"] & 1 else 0, np.uint32))
        return out
    for l in range(st["n_levels"]):
        blk = ins[lv[l]:lv[l + 1]]
        if rng is not None:
            blk = blk[rng.permutation(len(blk))]
        lut = blk[blk["kind"] == K_LUT]
        new_vals = _lut(lut["tt"], slots[lut["a"]], slots[lut["b"]], slots[lut["c"]]) if len(lut) else None
        loads = blk[blk["kind"] == K_LDIN]
        load_vals = [in_row(r).copy() for r in loads["row"]]
        for x in blk[blk["kind"] == K_STOUT]:
            st_row(x["row"], slots[x["a"]] ^ np.uint32(0xFFFFFFFF if x["flags"] & 1 else 0))
        for x in blk[blk["kind"] == K_STCONST]:
            st_row(x["row"], np.full(W, 0xFFFFFFFF if x["flags"] & 1 else 0, np.uint32))
        if len(lut):
            slots[lut["dst"]] = new_vals
        for x, v in zip(loads, load_vals):
            slots[x["dst"]] = v
    return out
