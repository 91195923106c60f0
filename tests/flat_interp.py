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
K_LUT, K_LDIN, K_STOUT, K_STCONST = range(4)
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


def run(flat: Flat, stream: int, imm: np.ndarray, state: np.ndarray = None, rng=None, _part=None, _out=None):
    """imm [n_inputs][W]; state [n_state][W] is updated in place. Returns out [n_outputs][W]."""
    W = imm.shape[1]
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
