// Flat device-buffer format produced by the levelize/register-allocate pass and consumed by the
// kernels. Everything is little-endian; offsets are in bytes from the start of the buffer and are
// 16-byte aligned. Documented for users in DESIGN.md ("Flat program format").
#pragma once
#include <cstdint>

namespace clg {

constexpr uint32_t FLAT_MAGIC = 0x46474C43u;  // "CLGF"
constexpr uint32_t FLAT_VERSION = 5u;

enum InstrKind : uint8_t {
  K_LUT = 0,      // slot[dst] = LUT3(tt; slot[a], slot[b], slot[c])      (one LOP3 per 32 lanes)
  K_LDIN = 1,     // slot[dst] = input row `row`   (immediates first, then carried state rows)
  K_STOUT = 2,    // output row `row` = slot[a] (inverted if flags & 1)   (named outputs, then state)
  K_STCONST = 3,  // output row `row` = all-zeros / all-ones (flags & 1)
  K_NOP = 4,      // row stream only: a lane of a row without an instruction
};

// 16 bytes, one 128-bit load. Slots are physical registers of the bit-sliced register stack.
struct FlatInstr {
  uint16_t dst, a, b, c;
  uint8_t tt;    // LOP3 immediate: bit ((va << 2) | (vb << 1) | vc) of tt is the result
  uint8_t kind;  // InstrKind
  uint16_t flags;  // bit 0: invert (STOUT) / value (STCONST); bits 8-9: number of operands a LUT really reads (1..3)
  uint32_t row;
};
constexpr uint16_t FLAG_INV = 1;
constexpr uint16_t FLAG_FROM_REG = 2;  // row stream, K_LUT: operand a is the result this lane computed last (field a is not read)
constexpr uint16_t FLAG_NO_STORE = 4;  // row stream, K_LUT: the result is only ever read that way and gets no slot (field dst is not read)
constexpr int NFAN_SHIFT = 8;
inline int lut_nfan(const FlatInstr &x) { return (x.flags >> NFAN_SHIFT) & 3; }
static_assert(sizeof(FlatInstr) == 16, "FlatInstr must be 16 bytes");

constexpr uint32_t STREAM_LEVEL = 0;  // level-ordered: instructions of one level are independent
constexpr uint32_t STREAM_LANE = 1;   // one sequential order chosen for a small live set

struct FlatStream {
  uint32_t n_slots;
  uint32_t n_instr;
  uint32_t n_levels;         // entries-1 of the level offset table (lane stream: 1)
  uint32_t off_levels;       // uint32[n_levels + 1] instruction indices
  uint32_t off_instr;        // FlatInstr[n_instr]
  uint32_t max_level_width;  // widest level, in instructions
  uint32_t reserved[2];
};

// The row stream (cooperative row kernel, clg_coop_rows). The program is cut into SEGMENTS; in every segment each of
// the n_warps warps of the CTA executes a few ROWS of 32 lane instructions (K_LUT or K_NOP), one after the other, and
// then its share of the segment's side list (loads, stores, LUTs that do not fit a row). Warps are NOT barrier-
// synchronised segment by segment: a warp may start segment s as soon as every warp has finished segment s - lag, so
//   * a slot written in segment s may be read in segments >= s + lag only, and written again only in segments >= (last
//     segment that reads it) + lag -- except that
//   * a lane may take operand a from its own previous result (FLAG_FROM_REG), whatever segment that was computed in.
struct FlatRows {
  uint32_t n_slots;
  uint32_t n_instr;      // lane instructions = 32 * rows, NOP lanes included
  uint32_t n_segs;
  uint32_t n_warps;      // warps of the CTA this stream was scheduled for
  uint32_t lag;          // visibility distance in segments (2)
  uint32_t off_rowtab;   // uint32[n_segs * n_warps + 1]: first lane instruction of (segment, warp); every entry a multiple of 32
  uint32_t off_instr;    // FlatInstr[n_instr]
  uint32_t n_side;
  uint32_t off_sidetab;  // uint32[n_segs + 1]: first side instruction of each segment
  uint32_t off_side;     // FlatInstr[n_side]
  uint32_t n_from_reg, n_no_store;  // statistics: LUTs that take an operand from a register / whose result is never stored
};

struct FlatHeader {
  uint32_t magic, version, header_bytes, total_bytes;
  uint32_t n_inputs, n_outputs, n_state, n_luts;
  uint64_t ref_nand_count, ref_set_count;
  uint32_t flags;          // CLG_LV_* the buffer was built with
  uint32_t n_registers;    // registers of the source Simulation
  uint32_t off_state_regs; // uint64[n_state]  register behind each state row
  uint32_t off_out_regs;   // uint64[n_outputs] register behind each output row
  FlatStream stream[2];    // [STREAM_LEVEL], [STREAM_LANE] (n_instr == 0 when absent)
  // Optional partitioned form of the level-ordered stream: the LUT network's connected components, bin-packed
  // into n_parts independent programs (own levels, own slots; rows stay global). A batch too small to give every
  // SM a column group runs one CTA per (part, group) instead. n_parts == 0 when the network is one component.
  uint32_t n_parts;
  uint32_t off_parts;      // FlatStream[n_parts]
  uint32_t reserved[2];
  FlatRows rows;           // n_instr == 0 when absent
};
static_assert(sizeof(FlatHeader) % 16 == 0, "header must keep 16-byte alignment");

}  // namespace clg
