// A C++ host driving the backend through the C ABI only (no Python, no torch): builds an n-bit ripple adder
// with the reference's gate API (Compiler + FullAdder, as complogic's tests do), compiles it to the NAND program,
// levelizes it and evaluates a batch of random vectors on GPU 0, checking a + b + cin on the host.
//
//   g++ -std=c++17 -Iinclude examples/adder_demo.cpp -Lcomplogic_b200 -lcomplogic_cuda -Wl,-rpath,$PWD/complogic_b200 -o adder_demo
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "complogic_cuda.h"

#define CHECK(call)                                                        \
  do {                                                                     \
    int rc_ = (call);                                                      \
    if (rc_ != CLG_OK) {                                                   \
      std::fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, clg_last_error()); \
      return 1;                                                            \
    }                                                                      \
  } while (0)

int main(int argc, char **argv) {
  const unsigned n = argc > 1 ? (unsigned)std::atoi(argv[1]) : 32;  // adder width, <= 63 for the host check
  const size_t n_vectors = argc > 2 ? (size_t)std::atoll(argv[2]) : 100000;

  // Compiler::new(2n+1); a_i = reg i, b_i = reg n+i, cin = reg 2n           (compile.rs:65)
  clg_compiler *c = nullptr;
  CHECK(clg_compiler_new(2 * n + 1, &c));
  std::vector<uint64_t> s(n), k(n);
  for (unsigned i = 0; i < n; i++) s[i] = clg_compiler_alloc(c);          // compile.rs:88
  for (unsigned i = 0; i < n; i++) k[i] = clg_compiler_alloc(c);
  std::vector<clg_gate> gates(n);
  for (unsigned i = 0; i < n; i++) {                                      // FullAdder{a,b,cin,s,cout}, gates.rs:76
    gates[i] = clg_gate{CLG_GATE_FULL_ADDER, 0, {i, n + i, i == 0 ? 2 * n : k[i - 1], s[i], k[i]}};
  }
  clg_simulation *sim = nullptr;
  CHECK(clg_compiler_compile(c, gates.data(), gates.size(), &sim));       // Compiler::compile, compile.rs:93
  size_t n_ops = 0, n_regs = 0;
  clg_simulation_ops(sim, &n_ops);
  clg_simulation_registers(sim, &n_regs);
  std::printf("adder%u: %zu ops (%llu NAND), %zu registers\n", n, n_ops, (unsigned long long)clg_simulation_nand_count(sim), n_regs);

  std::vector<uint64_t> outs(s);
  outs.push_back(k[n - 1]);
  clg_flat *flat = nullptr;
  CHECK(clg_levelize(sim, 2 * n + 1, outs.data(), outs.size(), 0, &flat));
  clg_flat_info info;
  CHECK(clg_flat_get_info(flat, &info));
  std::printf("levelized: %u LOP3 for %llu NAND, %u levels, %u lane registers\n", info.n_luts,
              (unsigned long long)info.ref_nand_count, info.n_levels, info.lane_slots);

  clg_ctx *ctx = nullptr;
  CHECK(clg_ctx_create(0, &ctx));
  clg_exec *ex = nullptr;
  CHECK(clg_exec_create(ctx, flat, CLG_MODE_AUTO, &ex));

  const size_t stride = ((n_vectors + 31) / 32 + 3) / 4 * 4;
  std::vector<uint32_t> imm((2 * n + 1) * stride), out((n + 1) * stride);
  std::mt19937 rng(1);
  for (auto &w : imm) w = rng();
  CHECK(clg_exec_run_host(ex, imm.data(), out.data(), nullptr, n_vectors, stride));  // Simulation::run x n_vectors

  size_t bad = 0;
  for (size_t v = 0; v < n_vectors; v++) {
    auto bit = [&](const std::vector<uint32_t> &buf, size_t row) { return (uint64_t)((buf[row * stride + v / 32] >> (v % 32)) & 1u); };
    uint64_t a = 0, b = 0, sum = 0;
    for (unsigned i = 0; i < n; i++) a |= bit(imm, i) << i, b |= bit(imm, n + i) << i;
    for (unsigned i = 0; i <= n; i++) sum |= bit(out, i) << i;
    bad += (sum != a + b + bit(imm, 2 * n));
  }
  std::printf("%zu vectors on the GPU (mode %d): %zu mismatches against a + b + cin\n", n_vectors, clg_exec_mode(ex), bad);

  clg_exec_destroy(ex);
  clg_ctx_destroy(ctx);
  clg_flat_destroy(flat);
  clg_simulation_destroy(sim);
  clg_compiler_destroy(c);
  return bad ? 2 : 0;
}
