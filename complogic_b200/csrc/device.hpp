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
