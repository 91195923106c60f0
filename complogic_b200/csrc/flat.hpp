// Flat device-buffer format produced by the levelize/register-allocate pass and consumed by the
// kernels. Everything is little-endian; offsets are in bytes from the start of the buffer and are
// 16-byte aligned. Documented for users in DESIGN.md ("Flat program format").
#pragma once
#include <cstdint>

namespace clg {

constexpr uint32_t FLAT_MAGIC = 0x46474C43u;  // "CLGF"
constexpr uint32_t FLAT_VERSION = 4u;

enum InstrKind : uint8_t {
  K_LUT = 0,      // slot[dst] = LUT3(tt; slot[a], slot[b], slot[c])      (one LOP3 per 32 lanes)
  K_LDIN = 1,     // slot[dst] = input row `row`   (immediates first, then carried state rows)
  K_STOUT = 2,    // output row `row` = slot[a] (inverted if flags & 1)   (named outputs, then state)
  K_STCONST = 3,  // output row `row` = all-zeros / all-ones (flags & 1)
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
};
static_assert(sizeof(FlatHeader) % 16 == 0, "header must keep 16-byte alignment");

}  // namespace clg
