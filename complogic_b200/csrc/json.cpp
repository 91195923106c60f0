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

}  // namespace clg
