// serde_json-shaped interchange of `Simulation { ops, registers }` -- what the derives at
// complogic/src/simulation.rs:5 and compile.rs:7 produce with serde's default (externally tagged)
// enum representation:
//   {"ops":[{"Nand":[0,1,2]},{"Set":[0,false]}],"registers":[false,false,true]}
#include <cctype>
#include <cstdlib>
#include <cstring>

#include "host.hpp"

namespace clg {

std::string simulation_to_json(const Simulation &sim) {
  std::string s = "{\"ops\":[";
  for (size_t i = 0; i < sim.ops.size(); i++) {
    const Op &op = sim.ops[i];
    if (i) s += ',';
    if (op.tag == CLG_OP_NAND)
      s += "{\"Nand\":[" + std::to_string(op.a) + "," + std::to_string(op.b) + "," + std::to_string(op.c) + "]}";
    else
      s += "{\"Set\":[" + std::to_string(op.a) + "," + (op.b ? "true" : "false") + "]}";
  }
  s += "],\"registers\":[";
  for (size_t i = 0; i < sim.registers.size(); i++) {
    if (i) s += ',';
    s += sim.registers[i] ? "true" : "false";
  }
  s += "]}";
  return s;
}

namespace {
struct Parser {
  const char *p;
  [[noreturn]] void bad(const char *what) { fail(CLG_E_FORMAT, std::string("Simulation JSON: expected ") + what); }
  void ws() {
    while (*p && std::isspace((unsigned char)*p)) p++;
  }
  bool eat(char c) {
    ws();
    if (*p == c) {
      p++;
      return true;
    }
    return false;
  }
  void need(char c, const char *what) {
    if (!eat(c)) bad(what);
  }
  std::string str() {
    need('"', "string");
    const char *b = p;
    while (*p && *p != '"') p++;
    if (!*p) bad("closing quote");
    return std::string(b, p++);
  }
  uint64_t num() {
    ws();
    if (!std::isdigit((unsigned char)*p)) bad("unsigned integer");
    char *e;
    uint64_t v = std::strtoull(p, &e, 10);
    p = e;
    return v;
  }
  bool boolean() {
    ws();
    if (!std::strncmp(p, "true", 4)) {
      p += 4;
      return true;
    }
    if (!std::strncmp(p, "false", 5)) {
      p += 5;
      return false;
    }
    bad("true/false");
  }
};
}  // namespace

Simulation simulation_from_json(const char *json) {
  Parser P{json};
  Simulation sim;
  bool have_ops = false, have_regs = false;
  P.need('{', "{");
  if (!P.eat('}')) {
    do {
      std::string key = P.str();
      P.need(':', ":");
      P.need('[', "[");
      if (key == "ops") {
        have_ops = true;
        if (!P.eat(']')) {
          do {
            P.need('{', "{");
            std::string tag = P.str();
            P.need(':', ":");
            P.need('[', "[");
            if (tag == "Nand") {
              uint64_t a = P.num();
              P.need(',', ",");
              uint64_t b = P.num();
              P.need(',', ",");
              uint64_t c = P.num();
              sim.ops.push_back(nand_op(a, b, c));
            } else if (tag == "Set") {
              uint64_t id = P.num();
              P.need(',', ",");
              bool v = P.boolean();
              sim.ops.push_back(set_op(id, v));
            } else
              P.bad("Nand or Set");
            P.need(']', "]");
            P.need('}', "}");
          } while (P.eat(','));
          P.need(']', "]");
        }
      } else if (key == "registers") {
        have_regs = true;
        if (!P.eat(']')) {
          do sim.registers.push_back(P.boolean() ? 1 : 0);
          while (P.eat(','));
          P.need(']', "]");
        }
      } else
        P.bad("\"ops\" or \"registers\"");
    } while (P.eat(','));
    P.need('}', "}");
  }
  if (!have_ops || !have_regs) P.bad("both \"ops\" and \"registers\"");
  return sim;
}

// ---- Compiler, Incrementer and the gate structs (compile.rs:18-61, gates.rs:5-117) ------------------------------
// serde's derives: a struct is an object of its fields in declaration order, `Gate` is externally tagged with the
// variant struct as its content:
//   Compiler     {"immediate_count":2,"ops":[{"Set":[0,false]},...],"incrementer":{"val":3}}
//   Incrementer  {"val":3}
//   Gate         {"FullAdder":{"a":0,"b":1,"cin":2,"s":3,"cout":4}}
namespace {
struct GateShape {
  const char *name;
  int n;
  const char *field[13];
};
const GateShape GATE_SHAPES[12] = {
    {"Nand", 3, {"a", "b", "out"}},
    {"Not", 2, {"a", "out"}},
    {"And", 3, {"a", "b", "out"}},
    {"Or", 3, {"a", "b", "out"}},
    {"Nor", 3, {"a", "b", "out"}},
    {"Xor", 3, {"a", "b", "out"}},
    {"RSLatch", 3, {"s", "r", "q"}},
    {"RSLatchTest", 3, {"s", "r", "q"}},
    {"DLatch", 3, {"d", "e", "q"}},
    {"HalfAdder", 4, {"a", "b", "s", "c"}},
    {"FullAdder", 5, {"a", "b", "cin", "s", "cout"}},
    {"FourBitAdder", 13, {"a1", "a2", "a3", "a4", "b1", "b2", "b3", "b4", "s1", "s2", "s3", "s4", "cout"}},
};

void ops_to_json(const std::vector<Op> &ops, std::string &s) {
  s += '[';
  for (size_t i = 0; i < ops.size(); i++) {
    const Op &op = ops[i];
    if (i) s += ',';
    if (op.tag == CLG_OP_NAND)
      s += "{\"Nand\":[" + std::to_string(op.a) + "," + std::to_string(op.b) + "," + std::to_string(op.c) + "]}";
    else
      s += "{\"Set\":[" + std::to_string(op.a) + "," + (op.b ? "true" : "false") + "]}";
  }
  s += ']';
}

void ops_from_json(Parser &P, std::vector<Op> &ops) {
  P.need('[', "[");
  if (P.eat(']')) return;
  do {
    P.need('{', "{");
    std::string tag = P.str();
    P.need(':', ":");
    P.need('[', "[");
    if (tag == "Nand") {
      uint64_t a = P.num();
      P.need(',', ",");
      uint64_t b = P.num();
      P.need(',', ",");
      uint64_t c = P.num();
      ops.push_back(nand_op(a, b, c));
    } else if (tag == "Set") {
      uint64_t id = P.num();
      P.need(',', ",");
      bool v = P.boolean();
      ops.push_back(set_op(id, v));
    } else
      P.bad("Nand or Set");
    P.need(']', "]");
    P.need('}', "}");
  } while (P.eat(','));
  P.need(']', "]");
}
}  // namespace

std::string gate_to_json(const clg_gate &g) {
  if (g.kind >= 12) fail(CLG_E_INVALID, "unknown gate kind");
  const GateShape &sh = GATE_SHAPES[g.kind];
  std::string s = std::string("{\"") + sh.name + "\":{";
  for (int i = 0; i < sh.n; i++) s += std::string(i ? "," : "") + "\"" + sh.field[i] + "\":" + std::to_string(g.f[i]);
  return s + "}}";
}

clg_gate gate_from_json(const char *json) {
  Parser P{json};
  clg_gate g{};
  P.need('{', "{");
  const std::string tag = P.str();
  int kind = -1;
  for (int k = 0; k < 12; k++)
    if (tag == GATE_SHAPES[k].name) kind = k;
  if (kind < 0) P.bad("a gate name (Nand, Not, And, Or, Nor, Xor, RSLatch, DLatch, HalfAdder, FullAdder, FourBitAdder)");
  const GateShape &sh = GATE_SHAPES[kind];
  g.kind = (uint32_t)kind;
  P.need(':', ":");
  P.need('{', "{");
  uint32_t seen = 0;
  if (!P.eat('}')) {
    do {  // serde accepts the fields of a struct in any order
      const std::string key = P.str();
      P.need(':', ":");
      int f = -1;
      for (int i = 0; i < sh.n; i++)
        if (key == sh.field[i]) f = i;
      if (f < 0) P.bad("a field of that gate");
      if (seen >> f & 1u) P.bad("every field once");
      seen |= 1u << f, g.f[f] = P.num();
    } while (P.eat(','));
    P.need('}', "}");
  }
  if (seen != (1u << sh.n) - 1) P.bad("all fields of the gate");
  P.need('}', "}");
  return g;
}

std::string compiler_to_json(const Compiler &c) {
  std::string s = "{\"immediate_count\":" + std::to_string(c.immediate_count) + ",\"ops\":";
  ops_to_json(c.ops, s);
  return s + ",\"incrementer\":{\"val\":" + std::to_string(c.incrementer) + "}}";
}

Compiler compiler_from_json(const char *json) {
  Parser P{json};
  Compiler c(0);
  bool have[3] = {false, false, false};
  P.need('{', "{");
  if (!P.eat('}')) {
    do {
      const std::string key = P.str();
      P.need(':', ":");
      if (key == "immediate_count") have[0] = true, c.immediate_count = P.num();
      else if (key == "ops") have[1] = true, ops_from_json(P, c.ops);
      else if (key == "incrementer") {
        have[2] = true;
        P.need('{', "{");
        if (P.str() != "val") P.bad("\"val\"");
        P.need(':', ":");
        c.incrementer = P.num();
        P.need('}', "}");
      } else
        P.bad("\"immediate_count\", \"ops\" or \"incrementer\"");
    } while (P.eat(','));
    P.need('}', "}");
  }
  if (!have[0] || !have[1] || !have[2]) P.bad("\"immediate_count\", \"ops\" and \"incrementer\"");
  return c;
}

}  // namespace clg
