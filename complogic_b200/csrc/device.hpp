// Device side of the backend: thin dynamic binding to the CUDA driver API (libcuda.so.1 is
// dlopen'ed on first use, cudarc-style; the library itself has no link-time CUDA dependency, so it
// loads -- and the host-side passes work -- on machines without a driver), contexts, resident
// programs and launches.
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "host.hpp"

namespace clg {

std::string emit_spec_ptx(const Flat &f, int K, bool multi_step = false);
std::vector<uint32_t> spec_live_slots(const Flat &f, uint32_t at);
std::string emit_spec_chunk_ptx(const Flat &f, int K, uint32_t begin, uint32_t end, const std::vector<uint32_t> &live_in,
                                const std::vector<uint32_t> &live_out);

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
};

constexpr int LANE_MIN_THREADS = 64, LANE_MAX_THREADS = 992;  // consumer threads of a lane-kernel CTA (+ 32 producer)
constexpr int RING_STAGES = 4;
constexpr int RING_CHUNK = 256;
constexpr size_t LANE_RING_BYTES = RING_STAGES * RING_CHUNK * 16 + 128;
constexpr size_t SMEM_MAX = 227 * 1024;
constexpr uint32_t SPEC_CHUNK = 16384;  // most instructions per kernel of a chunked specialised program
inline uint32_t spec_n_chunks(uint32_t n_instr) { return n_instr ? (n_instr + SPEC_CHUNK - 1) / SPEC_CHUNK : 1; }
// first instruction of chunk c (c == n_chunks: the end); chunks are equal-sized
inline uint32_t spec_chunk_begin(uint32_t n_instr, uint32_t c) {
  const uint32_t n = spec_n_chunks(n_instr);
  return (uint32_t)(((uint64_t)n_instr * c + n - 1) / n);
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
