// Host-side core of the backend: the mirror of complogic's Op / Simulation / Compiler / Gate
// (the API surface the hot path sits behind) and the levelize + register-allocate pass.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/complogic_cuda.h"
#include "flat.hpp"

namespace clg {

struct Error {
  int code;
  std::string msg;
};
[[noreturn]] void fail(int code, const std::string &msg);

// compile.rs:7-14
using Op = clg_op;
inline Op nand_op(uint64_t a, uint64_t b, uint64_t out) { return Op{CLG_OP_NAND, 0, a, b, out}; }
inline Op set_op(uint64_t id, bool val) { return Op{CLG_OP_SET, 0, id, val ? 1u : 0u, 0}; }

// simulation.rs:5-12
struct Simulation {
  std::vector<Op> ops;
  std::vector<uint8_t> registers;
};

// gates.rs:193-456
void gate_create(const clg_gate &g, uint64_t &incrementer, std::vector<Op> &out);

// compile.rs:51-203
struct Compiler {
  uint64_t immediate_count = 0;
  std::vector<Op> ops;
  uint64_t incrementer = 0;
  std::string graph_dot;  // text the reference would have written to ./graph.dot

  explicit Compiler(uint64_t ic) : immediate_count(ic), incrementer(ic) {}
  void reset_ops();
  void reset_incrementer() { incrementer = immediate_count; }
  uint64_t alloc() { return incrementer++; }
  Simulation compile(const clg_gate *gates, size_t n_gates);
};

// The flat buffer (flat.hpp) plus decoded views.
struct Flat {
  std::vector<uint8_t> bytes;
  const FlatHeader &header() const { return *reinterpret_cast<const FlatHeader *>(bytes.data()); }
  const FlatInstr *instrs(int s) const {
    return reinterpret_cast<const FlatInstr *>(bytes.data() + header().stream[s].off_instr);
  }
  const uint32_t *levels(int s) const {
    return reinterpret_cast<const uint32_t *>(bytes.data() + header().stream[s].off_levels);
  }
  const FlatStream *parts() const { return reinterpret_cast<const FlatStream *>(bytes.data() + header().off_parts); }
  const FlatInstr *part_instrs(uint32_t p) const { return reinterpret_cast<const FlatInstr *>(bytes.data() + parts()[p].off_instr); }
  const uint32_t *part_levels(uint32_t p) const { return reinterpret_cast<const uint32_t *>(bytes.data() + parts()[p].off_levels); }
  const FlatRows &rows() const { return header().rows; }
  const FlatInstr *row_instrs() const { return reinterpret_cast<const FlatInstr *>(bytes.data() + rows().off_instr); }
  const uint32_t *row_tab() const { return reinterpret_cast<const uint32_t *>(bytes.data() + rows().off_rowtab); }
  const FlatInstr *row_side() const { return reinterpret_cast<const FlatInstr *>(bytes.data() + rows().off_side); }
  const uint32_t *row_sidetab() const { return reinterpret_cast<const uint32_t *>(bytes.data() + rows().off_sidetab); }
  const uint64_t *state_regs() const {
    return reinterpret_cast<const uint64_t *>(bytes.data() + header().off_state_regs);
  }
};

// lower.cpp
Flat levelize(const Simulation &sim, size_t n_immediates, const uint64_t *out_regs, size_t n_out, uint32_t flags);
void validate_flat(const Flat &f);  // throws Error(CLG_E_FORMAT)

// json.cpp
std::string simulation_to_json(const Simulation &sim);
Simulation simulation_from_json(const char *json);
std::string compiler_to_json(const Compiler &c);
Compiler compiler_from_json(const char *json);
std::string gate_to_json(const clg_gate &g);
clg_gate gate_from_json(const char *json);

}  // namespace clg
