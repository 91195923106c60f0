// Device side of the backend: thin dynamic binding to the CUDA driver API (libcuda.so.1 is
// dlopen'ed on first use, cudarc-style; the library itself has no link-time CUDA dependency, so it
// loads -- and the host-side passes work -- on machines without a driver), contexts, resident
// programs and launches.
#pragma once
#include <cstdlib>
#include <memory>
#include <string>
#include <vector>

#include "host.hpp"

namespace clg {

std::string emit_spec_ptx(const Flat &f, int K, bool multi_step = false);
std::vector<uint32_t> spec_live_slots(const Flat &f, uint32_t at);
uint32_t spec_chunk_limit(int K);                        // longest chunk, in lane-order instructions
std::vector<uint32_t> spec_chunk_cuts(const Flat &f, int K, uint32_t limit = 0, bool fine = false);  // n_chunks + 1 instruction indices
// One kernel of the chunked specialised form, in chunk-local terms (spec.cpp)
constexpr uint16_t FLAG_STATE_ROW = 0x10;  // chunk-local code only: `row` indexes the state buffer
struct SpecChunk {
  std::vector<FlatInstr> code;              // rows relative to the bases below
  std::vector<uint32_t> live_in, live_out;  // registers handed over through the scratch buffer (chunk-local names)
  std::vector<uint32_t> scr_row;            // chunk-local register -> its row of the scratch buffer (empty: the same number)
  uint32_t n_slots = 0;
  uint32_t imm_base = 0, out_base = 0, state_base = 0;  // rows the kernel's pointers are advanced by at launch
  bool crosses_in = false, crosses_out = false;         // some register is live across the cut before / after the chunk (whether or not the chunk touches it)
  bool same_kernel(const SpecChunk &o) const;           // equal code: the two chunks can share one assembled kernel
};
std::vector<SpecChunk> spec_chunks(const Flat &f, int K, uint32_t limit = 0, bool fine = false);  // limit: longest chunk in lane-order instructions (0: spec_chunk_limit(K))
struct SpecChain {
  std::vector<SpecChunk> chunks;
  std::vector<uint32_t> unique;     // chunks that get assembled (one per distinct kernel)
  std::vector<uint32_t> kernel_of;  // chunk -> index into `unique`
  struct Launch {
    // chunks first, first + stride, ... (count of them) run side by side in one launch (grid.y = count). stride 1: equal
    // chunks that hand nothing over. stride G > 1: the launch is one of G consecutive launches that together run `count`
    // equal groups of G chunks (a group hands registers over inside but not across its ends: the instances of a flattened
    // program whose instance is longer than a chunk) -- position by position, every group on its own part of the scratch buffer.
    uint32_t first, count, stride;
  };
  std::vector<Launch> launches;
};
SpecChain spec_chain(const Flat &f, int K, uint32_t limit = 0, bool fine = false);  // fine: cut at the FIRST point no register crosses (one instance of a flattened program per kernel)
std::string emit_spec_chunk_ptx(const Flat &f, int K, const SpecChunk &c);

// mirrors kernels.cu
struct RunParams {
  const void *instr;
  const uint32_t *levels;
  const uint32_t *imm;
  uint32_t *out;
  uint32_t *state;
  uint32_t n_inputs, n_outputs;
  uint32_t n_instr, n_levels, n_slots;
  uint32_t stride;
  uint32_t n_groups;
  const void *parts;
  uint32_t n_parts;
  uint32_t code_off;
  const void *cinstr;
  const uint32_t *clevels;
  const void *rinstr;  // row form (clg_coop_rows)
  const uint32_t *rwarp;
  const uint32_t *rside_off;
  uint32_t ring_off;  // clg_coop_rows_t: byte offset of the per-warp instruction rings in shared memory
};

constexpr int LANE_MIN_THREADS = 64, LANE_MAX_THREADS = 992;  // consumer threads of a lane-kernel CTA (+ 32 producer)
constexpr int RING_STAGES = 4;
constexpr int RING_CHUNK = 256;
constexpr size_t LANE_RING_BYTES = RING_STAGES * RING_CHUNK * 16 + 128;
constexpr size_t SMEM_MAX = 227 * 1024;
constexpr uint32_t SPEC_CHUNK = 16384;  // most instructions per kernel of a chunked specialised program ...
inline uint32_t spec_chunk_limit() {    // ... unless CLG_SPEC_CHUNK says otherwise (tuning and tests; read at every plan build)
  const char *e = std::getenv("CLG_SPEC_CHUNK");
  const long n = e ? std::atol(e) : 0;
  return n >= 2 ? (uint32_t)n : SPEC_CHUNK;
}

// The second chain, for big batches (backend.cpp, wants_tp_chain): chunks short enough to stay in the instruction cache.
inline uint32_t spec_tp_chunk_limit() {  // lane-order instructions at K = 1 (6 800: ~110 KB of SASS with the hand-over code)
  const char *e = std::getenv("CLG_SPEC_TP_CHUNK");
  const long n = e ? std::atol(e) : 0;
  return n >= 2 ? (uint32_t)n : 6800u;
}

}  // namespace clg

// C ABI objects
struct clg_simulation {
  clg::Simulation sim;
};
struct clg_compiler {
  clg::Compiler c;
  explicit clg_compiler(uint64_t ic) : c(ic) {}
};
struct clg_flat {
  clg::Flat f;
};
