// Host mirror of complogic's gate library and Compiler (the API surface in front of the hot path).
//   Gate::create            complogic/src/gates.rs:192-456
//   Compiler::{reset_ops, compile}  complogic/src/compile.rs:74-80, 93-202
// Built from scratch around a small emitter; register numbering and the emitted op order are the
// reference's (tests compare op for op against the oracle and the reference's golden vectors).
#include <algorithm>
#include <cstdio>

#include "host.hpp"

namespace clg {

namespace {

// Emits NANDs while drawing scratch registers from the incrementer in the reference's order:
// every composite gate reserves the outputs of its children first (struct construction in
// gates.rs happens before any child's create()), then expands the children left to right.
struct Lowerer {
  uint64_t &inc;
  std::vector<Op> &ops;

  uint64_t fresh() { return inc++; }
  void nand(uint64_t a, uint64_t b, uint64_t out) { ops.push_back(nand_op(a, b, out)); }  // gates.rs:195
  void inv(uint64_t a, uint64_t out) { nand(a, a, out); }                                  // gates.rs:198
  void and_(uint64_t a, uint64_t b, uint64_t out) {                                        // gates.rs:201
    uint64_t t = fresh();
    nand(a, b, t);
    inv(t, out);
  }
  void or_(uint64_t a, uint64_t b, uint64_t out) {  // gates.rs:218
    uint64_t na = fresh(), nb = fresh();
    nand(a, a, na);
    nand(b, b, nb);
    nand(na, nb, out);
  }
  void nor(uint64_t a, uint64_t b, uint64_t out) {  // gates.rs:242
    uint64_t t = fresh();
    or_(a, b, t);
    inv(t, out);
  }
  void xor_(uint64_t a, uint64_t b, uint64_t out) {  // gates.rs:259
    uint64_t t_or = fresh(), t_nand = fresh();
    or_(a, b, t_or);
    nand(a, b, t_nand);
    and_(t_or, t_nand, out);
  }
  void rs_latch(uint64_t s, uint64_t r, uint64_t q, bool reversed) {  // gates.rs:283 / :312
    uint64_t q_patch = fresh(), qn_patch = fresh();
    if (!reversed) {
      nor(s, qn_patch, q_patch);
      nor(r, q_patch, qn_patch);
      or_(qn_patch, qn_patch, q);
    } else {
      or_(qn_patch, qn_patch, q);
      nor(r, q_patch, qn_patch);
      nor(s, qn_patch, q_patch);
    }
  }
  void d_latch(uint64_t d, uint64_t e, uint64_t q) {  // gates.rs:340
    uint64_t nd = fresh(), r_in = fresh(), s_in = fresh();
    inv(d, nd);
    and_(nd, e, r_in);
    and_(e, d, s_in);
    rs_latch(s_in, r_in, q, false);  // s/r swapped on purpose, gates.rs:357-363
  }
  void half_adder(uint64_t a, uint64_t b, uint64_t s, uint64_t c) {  // gates.rs:373
    xor_(a, b, s);
    and_(a, b, c);
  }
  void full_adder(uint64_t a, uint64_t b, uint64_t cin, uint64_t s, uint64_t cout) {  // gates.rs:391
    uint64_t s1 = fresh(), c1 = fresh(), c2 = fresh();
    half_adder(a, b, s1, c1);
    half_adder(s1, cin, s, c2);
    or_(c1, c2, cout);
  }
  void four_bit_adder(const uint64_t *f) {  // gates.rs:417; f = a1..a4 b1..b4 s1..s4 cout
    uint64_t cin = fresh();  // never written by any op (gates.rs:421)
    uint64_t k1 = fresh(), k2 = fresh(), k3 = fresh();
    full_adder(f[0], f[4], cin, f[8], k1);
    full_adder(f[1], f[5], k1, f[9], k2);
    full_adder(f[2], f[6], k2, f[10], k3);
    full_adder(f[3], f[7], k3, f[11], f[12]);
  }
};

std::string op_debug(const Op &op) {  // Rust `{:?}` of Op
  char buf[96];
  if (op.tag == CLG_OP_NAND)
    std::snprintf(buf, sizeof buf, "Nand(%llu, %llu, %llu)", (unsigned long long)op.a, (unsigned long long)op.b,
                  (unsigned long long)op.c);
  else
    std::snprintf(buf, sizeof buf, "Set(%llu, %s)", (unsigned long long)op.a, op.b ? "true" : "false");
  return buf;
}

}  // namespace

void gate_create(const clg_gate &g, uint64_t &incrementer, std::vector<Op> &out) {
  Lowerer L{incrementer, out};
  const uint64_t *f = g.f;
  switch (g.kind) {
    case CLG_GATE_NAND: L.nand(f[0], f[1], f[2]); break;
    case CLG_GATE_NOT: L.inv(f[0], f[1]); break;
    case CLG_GATE_AND: L.and_(f[0], f[1], f[2]); break;
    case CLG_GATE_OR: L.or_(f[0], f[1], f[2]); break;
    case CLG_GATE_NOR: L.nor(f[0], f[1], f[2]); break;
    case CLG_GATE_XOR: L.xor_(f[0], f[1], f[2]); break;
    case CLG_GATE_RSLATCH: L.rs_latch(f[0], f[1], f[2], false); break;
    case CLG_GATE_RSLATCH_TEST: L.rs_latch(f[0], f[1], f[2], true); break;
    case CLG_GATE_DLATCH: L.d_latch(f[0], f[1], f[2]); break;
    case CLG_GATE_HALF_ADDER: L.half_adder(f[0], f[1], f[2], f[3]); break;
    case CLG_GATE_FULL_ADDER: L.full_adder(f[0], f[1], f[2], f[3], f[4]); break;
    case CLG_GATE_FOUR_BIT_ADDER: L.four_bit_adder(f); break;
    default: fail(CLG_E_INVALID, "unknown gate kind " + std::to_string(g.kind));
  }
}

void Compiler::reset_ops() {
  ops.clear();
  for (uint64_t i = 0; i < immediate_count; i++) ops.push_back(set_op(i, false));
}

// compile.rs:93-202. The reference keeps the ops in a petgraph DiGraph whose node index doubles
// as the register id; here the same ordering is computed on plain arrays:
//   owner[r]   the op whose result register r holds (node weight after the relabel, :133-148)
//   preds/succs  CSR of the register->register edges the reference adds for every Nand whose three
//                ids are all < #ops (:115-125)
// and the pass loop (:155-196) is reproduced with its quirks: readiness is re-evaluated inside a
// pass, queues keep duplicates, and a pass that makes no progress force-emits its queue.
Simulation Compiler::compile(const clg_gate *gates, size_t n_gates) {
  reset_ops();
  graph_dot.clear();
  Simulation sim;
  if (n_gates == 0) {
    sim.registers.assign(immediate_count, 0);
    return sim;
  }
  uint64_t scratch = incrementer;  // the caller's incrementer is not advanced (:105)
  for (size_t i = 0; i < n_gates; i++) gate_create(gates[i], scratch, ops);

  const size_t n = ops.size();
  std::vector<Op> owner(ops);
  std::vector<uint64_t> esrc, edst;
  for (const Op &op : ops)
    if (op.tag == CLG_OP_NAND && op.a < n && op.b < n && op.c < n) {
      esrc.push_back(op.a), edst.push_back(op.c);
      esrc.push_back(op.b), edst.push_back(op.c);
    }

  std::vector<uint8_t> pending(n, 0);
  std::vector<uint64_t> queue, next;
  for (const Op &op : ops) {
    uint64_t reg = op.tag == CLG_OP_NAND ? op.c : op.a;
    if (reg >= n)
      fail(CLG_E_BAD_REGISTER, "compile: register " + std::to_string(reg) + " has no graph node (" +
                                   std::to_string(n) + " ops); the reference panics at compile.rs:137");
    owner[reg] = op;
    pending[reg] = 1;
    if (op.tag == CLG_OP_SET) queue.push_back(reg);
  }

  {  // Dot::with_config(&graph, &[]) as `{:?}` prints it (:150), kept in memory
    std::string &d = graph_dot;
    d = "digraph {\n";
    for (size_t i = 0; i < n; i++) d += "    " + std::to_string(i) + " [ label = \"" + op_debug(owner[i]) + "\" ]\n";
    for (size_t e = 0; e < esrc.size(); e++)
      d += "    " + std::to_string(esrc[e]) + " -> " + std::to_string(edst[e]) + " [ label = \"()\" ]\n";
    d += "}\n";
  }

  std::vector<uint32_t> pred_off(n + 1, 0), succ_off(n + 1, 0);
  for (size_t e = 0; e < esrc.size(); e++) pred_off[edst[e] + 1]++, succ_off[esrc[e] + 1]++;
  for (size_t i = 0; i < n; i++) pred_off[i + 1] += pred_off[i], succ_off[i + 1] += succ_off[i];
  std::vector<uint64_t> preds(esrc.size()), succs(esrc.size());
  {
    std::vector<uint32_t> pc(pred_off.begin(), pred_off.end() - 1), sc(succ_off.begin(), succ_off.end() - 1);
    for (size_t e = 0; e < esrc.size(); e++) preds[pc[edst[e]]++] = esrc[e], succs[sc[esrc[e]]++] = edst[e];
  }

  bool stuck = false;  // `recursion_flag`
  while (!queue.empty()) {
    for (uint64_t reg : queue) {
      if (!pending[reg]) continue;
      bool blocked = false;
      if (!stuck)
        for (uint32_t k = pred_off[reg]; k < pred_off[reg + 1] && !blocked; k++) blocked = pending[preds[k]];
      if (blocked) {
        next.push_back(reg);
        continue;
      }
      sim.ops.push_back(owner[reg]);
      pending[reg] = 0;
      next.insert(next.end(), succs.begin() + succ_off[reg], succs.begin() + succ_off[reg + 1]);
    }
    std::sort(queue.begin(), queue.end());
    std::sort(next.begin(), next.end());
    stuck = (queue == next);
    queue.swap(next);
    next.clear();
  }
  sim.registers.assign(scratch, 0);
  return sim;
}

}  // namespace clg
