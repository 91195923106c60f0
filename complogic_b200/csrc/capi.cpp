// extern "C" surface (include/complogic_cuda.h). Exceptions never cross it: every entry point
// catches clg::Error and turns it into a status code + thread-local message.
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <new>
#include <condition_variable>
#include <mutex>
#include <thread>

#include "device.hpp"

namespace clg {
struct Ctx;
struct Exec;
Ctx *ctx_create(int device);
void ctx_destroy(Ctx *);
const std::string &ctx_name(Ctx *);
int ctx_sm_count(Ctx *);
void ctx_sync(Ctx *);
void *dev_alloc(Ctx *, size_t);
void dev_free(Ctx *, void *);
void dev_memset(Ctx *, void *, int, size_t, void *);
void h2d(Ctx *, void *, const void *, size_t, void *);
void d2h(Ctx *, void *, const void *, size_t, void *);
void *host_alloc(Ctx *, size_t);
void host_free(Ctx *, void *);
void stream_sync(Ctx *, void *);
Exec *exec_create(Ctx *, const Flat &, int);
void exec_destroy(Exec *);
int exec_mode(const Exec *);
int exec_k(const Exec *);
int exec_launches(const Exec *);
const char *exec_kernel_name(const Exec *);
const std::string &exec_ptx(const Exec *);
const Flat &exec_flat(const Exec *);
Ctx *exec_ctx(const Exec *);
void exec_run(Exec *, const uint32_t *, uint32_t *, uint32_t *, size_t, size_t, void *);
void exec_prepare(Exec *, size_t);
void exec_run_bools_host(Exec *, const uint8_t *, uint8_t *, size_t);
void exec_run_host(Exec *, const uint32_t *, uint32_t *, uint32_t *, size_t, size_t);
void exec_run_steps(Exec *, const uint32_t *, uint32_t *, uint32_t *, size_t, size_t, size_t, void *);
void launch_slice(Ctx *, bool, const void *, void *, size_t, size_t, size_t, void *);
void launch_fill(Ctx *, uint32_t *, size_t, size_t, uint64_t, void *);
double measure_lop3_peak(Ctx *, int);

static thread_local std::string g_err;
void fail(int code, const std::string &msg) { throw Error{code, msg}; }
}  // namespace clg

using namespace clg;

struct clg_ctx {
  Ctx *c;
};
struct clg_exec {
  Exec *e;
};

#define CLG_TRY try {
#define CLG_CATCH                                   \
  }                                                 \
  catch (const clg::Error &e) {                     \
    clg::g_err = e.msg;                             \
    return e.code;                                  \
  }                                                 \
  catch (const std::bad_alloc &) {                  \
    clg::g_err = "out of memory";                   \
    return CLG_E_NOMEM;                             \
  }                                                 \
  catch (const std::exception &e) {                 \
    clg::g_err = std::string("internal: ") + e.what(); \
    return CLG_E_INVALID;                           \
  }                                                 \
  return CLG_OK;
#define NEED(p) ((p) ? (void)0 : fail(CLG_E_INVALID, "null argument: " #p))

extern "C" {

const char *clg_last_error(void) { return g_err.c_str(); }
uint32_t clg_abi_version(void) { return CLG_ABI_VERSION; }
void clg_free(void *p) { std::free(p); }

int clg_gate_create(const clg_gate *gate, uint64_t *incrementer_val, clg_op **ops_out, size_t *n_ops_out) {
  CLG_TRY
  NEED(gate), NEED(incrementer_val), NEED(ops_out), NEED(n_ops_out);
  std::vector<Op> ops;
  gate_create(*gate, *incrementer_val, ops);
  clg_op *o = (clg_op *)std::malloc(std::max<size_t>(1, ops.size()) * sizeof(clg_op));
  if (!o) throw std::bad_alloc();
  std::memcpy(o, ops.data(), ops.size() * sizeof(clg_op));
  *ops_out = o, *n_ops_out = ops.size();
  CLG_CATCH
}

// ---- Simulation ----------------------------------------------------------------------------------------
int clg_simulation_new(const clg_op *ops, size_t n_ops, const uint8_t *registers, size_t n_registers, clg_simulation **out) {
  CLG_TRY
  NEED(out);
  if (n_ops) NEED(ops);
  auto *s = new clg_simulation();
  s->sim.ops.assign(ops, ops + n_ops);
  for (Op &op : s->sim.ops) op.reserved = 0;
  if (registers) s->sim.registers.assign(registers, registers + n_registers);
  else s->sim.registers.assign(n_registers, 0);
  *out = s;
  CLG_CATCH
}
void clg_simulation_destroy(clg_simulation *sim) { delete sim; }
const clg_op *clg_simulation_ops(const clg_simulation *sim, size_t *n_ops) {
  if (n_ops) *n_ops = sim->sim.ops.size();
  return sim->sim.ops.data();
}
uint8_t *clg_simulation_registers(clg_simulation *sim, size_t *n_registers) {
  if (n_registers) *n_registers = sim->sim.registers.size();
  return sim->sim.registers.data();
}
int clg_simulation_register(const clg_simulation *sim, size_t id, int *value_out) {
  CLG_TRY
  NEED(sim), NEED(value_out);
  if (id >= sim->sim.registers.size()) fail(CLG_E_BAD_REGISTER, "register index out of range (the reference panics at simulation.rs:34)");
  *value_out = sim->sim.registers[id] != 0;
  CLG_CATCH
}
uint64_t clg_simulation_nand_count(const clg_simulation *sim) {
  uint64_t n = 0;
  for (const Op &op : sim->sim.ops) n += op.tag == CLG_OP_NAND;
  return n;
}

// ---- Compiler --------------------------------------------------------------------------------------------
int clg_compiler_new(uint64_t immediate_count, clg_compiler **out) {
  CLG_TRY
  NEED(out);
  *out = new clg_compiler(immediate_count);
  CLG_CATCH
}
void clg_compiler_destroy(clg_compiler *c) { delete c; }
void clg_compiler_reset_ops(clg_compiler *c) { c->c.reset_ops(); }
void clg_compiler_reset_incrementer(clg_compiler *c) { c->c.reset_incrementer(); }
uint64_t clg_compiler_alloc(clg_compiler *c) { return c->c.alloc(); }
uint64_t clg_compiler_immediate_count(const clg_compiler *c) { return c->c.immediate_count; }
void clg_compiler_set_immediate_count(clg_compiler *c, uint64_t v) { c->c.immediate_count = v; }
uint64_t clg_compiler_incrementer(const clg_compiler *c) { return c->c.incrementer; }
void clg_compiler_set_incrementer(clg_compiler *c, uint64_t v) { c->c.incrementer = v; }
const clg_op *clg_compiler_ops(const clg_compiler *c, size_t *n_ops) {
  if (n_ops) *n_ops = c->c.ops.size();
  return c->c.ops.data();
}
int clg_compiler_compile(clg_compiler *c, const clg_gate *gates, size_t n_gates, clg_simulation **out) {
  CLG_TRY
  NEED(c), NEED(out);
  if (n_gates) NEED(gates);
  auto *s = new clg_simulation();
  try {
    s->sim = c->c.compile(gates, n_gates);
  } catch (...) {
    delete s;
    throw;
  }
  *out = s;
  CLG_CATCH
}
const char *clg_compiler_graph_dot(const clg_compiler *c) { return c->c.graph_dot.c_str(); }

// ---- levelize / flat ---------------------------------------------------------------------------------------
int clg_levelize(const clg_simulation *sim, size_t n_immediates, const uint64_t *out_regs, size_t n_out, uint32_t flags,
                 clg_flat **out) {
  CLG_TRY
  NEED(sim), NEED(out);
  if (n_out) NEED(out_regs);  // (NULL with n_out == 0 means "every register")
  auto *f = new clg_flat();
  try {
    f->f = levelize(sim->sim, n_immediates, out_regs, n_out, flags);
    validate_flat(f->f);
  } catch (...) {
    delete f;
    throw;
  }
  *out = f;
  CLG_CATCH
}
void clg_flat_destroy(clg_flat *f) { delete f; }
int clg_flat_bytes(const clg_flat *f, const void **data, size_t *len) {
  CLG_TRY
  NEED(f), NEED(data), NEED(len);
  *data = f->f.bytes.data(), *len = f->f.bytes.size();
  CLG_CATCH
}
int clg_flat_from_bytes(const void *data, size_t len, clg_flat **out) {
  CLG_TRY
  NEED(data), NEED(out);
  auto *f = new clg_flat();
  try {
    f->f.bytes.assign((const uint8_t *)data, (const uint8_t *)data + len);
    validate_flat(f->f);
  } catch (...) {
    delete f;
    throw;
  }
  *out = f;
  CLG_CATCH
}
int clg_flat_get_info(const clg_flat *f, clg_flat_info *info) {
  CLG_TRY
  NEED(f), NEED(info);
  const FlatHeader &h = f->f.header();
  info->ref_nand_count = h.ref_nand_count, info->ref_set_count = h.ref_set_count;
  info->n_inputs = h.n_inputs, info->n_outputs = h.n_outputs, info->n_state = h.n_state, info->n_luts = h.n_luts;
  info->n_levels = h.stream[STREAM_LEVEL].n_levels, info->max_level_width = h.stream[STREAM_LEVEL].max_level_width;
  info->level_slots = h.stream[STREAM_LEVEL].n_slots, info->lane_slots = h.stream[STREAM_LANE].n_slots;
  info->level_instrs = h.stream[STREAM_LEVEL].n_instr, info->lane_instrs = h.stream[STREAM_LANE].n_instr;
  info->n_parts = h.n_parts, info->reserved = 0;
  info->rows_segments = h.rows.n_segs, info->rows_slots = h.rows.n_instr ? h.rows.n_slots : 0, info->rows_instrs = h.rows.n_instr;
  info->rows_from_reg = h.rows.n_from_reg, info->rows_no_store = h.rows.n_no_store, info->rows_warps = h.rows.n_instr ? h.rows.n_warps : 0;
  CLG_CATCH
}
int clg_flat_state_regs(const clg_flat *f, const uint64_t **regs, size_t *n) {
  CLG_TRY
  NEED(f), NEED(regs), NEED(n);
  *regs = f->f.state_regs(), *n = f->f.header().n_state;
  CLG_CATCH
}
static int spec_ptx_text(const clg_flat *f, int k, bool steps, char **ptx_out);
int clg_flat_spec_ptx(const clg_flat *f, int k, char **ptx_out) { return spec_ptx_text(f, k, false, ptx_out); }
int clg_flat_spec_steps_ptx(const clg_flat *f, int k, char **ptx_out) { return spec_ptx_text(f, k, true, ptx_out); }
static int spec_ptx_text(const clg_flat *f, int k, bool steps, char **ptx_out) {
  CLG_TRY
  NEED(f), NEED(ptx_out);
  std::string s = emit_spec_ptx(f->f, k, steps);
  char *p = (char *)std::malloc(s.size() + 1);
  if (!p) throw std::bad_alloc();
  std::memcpy(p, s.c_str(), s.size() + 1);
  *ptx_out = p;
  CLG_CATCH
}
static void chunk_ptx(const clg_flat *f, int k, bool second, uint32_t chunk, uint32_t *n_chunks, uint32_t *row_bases, char **ptx_out) {
  NEED(f), NEED(ptx_out);
  if (k != 1 && k != 2 && k != 4) fail(CLG_E_INVALID, "K must be 1, 2 or 4");
  if (f->f.header().stream[STREAM_LANE].n_instr == 0) fail(CLG_E_UNSUPPORTED, "no lane-ordered stream in this flat buffer");
  const std::vector<SpecChunk> chunks = second ? spec_chunks(f->f, k, std::max(2u, spec_tp_chunk_limit() / (uint32_t)k), true) : spec_chunks(f->f, k);
  if (n_chunks) *n_chunks = (uint32_t)chunks.size();
  if (chunk >= chunks.size()) fail(CLG_E_INVALID, "chunk index out of range");
  const SpecChunk &c = chunks[chunk];
  if (row_bases) row_bases[0] = c.imm_base, row_bases[1] = c.out_base, row_bases[2] = c.state_base;
  std::string s = emit_spec_chunk_ptx(f->f, k, c);
  char *p = (char *)std::malloc(s.size() + 1);
  if (!p) throw std::bad_alloc();
  std::memcpy(p, s.c_str(), s.size() + 1);
  *ptx_out = p;
}
int clg_flat_spec_chunk_ptx(const clg_flat *f, int k, uint32_t chunk, uint32_t *n_chunks, uint32_t *row_bases, char **ptx_out) {
  CLG_TRY
  chunk_ptx(f, k, false, chunk, n_chunks, row_bases, ptx_out);
  CLG_CATCH
}
int clg_flat_spec_batch_chunk_ptx(const clg_flat *f, int k, uint32_t chunk, uint32_t *n_chunks, uint32_t *row_bases, char **ptx_out) {
  CLG_TRY
  chunk_ptx(f, k, true, chunk, n_chunks, row_bases, ptx_out);
  CLG_CATCH
}
int clg_flat_spec_chain_info(const clg_flat *f, int k, uint32_t *n_chunks, uint32_t *n_kernels, uint32_t *n_launches) {
  CLG_TRY
  NEED(f);
  if (k != 1 && k != 2 && k != 4) fail(CLG_E_INVALID, "K must be 1, 2 or 4");
  if (f->f.header().stream[STREAM_LANE].n_instr == 0) fail(CLG_E_UNSUPPORTED, "no lane-ordered stream in this flat buffer");
  const SpecChain ch = spec_chain(f->f, k);
  if (n_chunks) *n_chunks = (uint32_t)ch.chunks.size();
  if (n_kernels) *n_kernels = (uint32_t)ch.unique.size();
  if (n_launches) *n_launches = (uint32_t)ch.launches.size();
  CLG_CATCH
}
int clg_flat_spec_batch_chain_info(const clg_flat *f, int k, uint32_t *n_chunks, uint32_t *n_kernels, uint32_t *n_launches, uint32_t *n_handed_over) {
  CLG_TRY
  NEED(f);
  if (k != 1 && k != 2 && k != 4) fail(CLG_E_INVALID, "K must be 1, 2 or 4");
  if (f->f.header().stream[STREAM_LANE].n_instr == 0) fail(CLG_E_UNSUPPORTED, "no lane-ordered stream in this flat buffer");
  const SpecChain ch = spec_chain(f->f, k, std::max(2u, spec_tp_chunk_limit() / (uint32_t)k), true);
  if (n_chunks) *n_chunks = (uint32_t)ch.chunks.size();
  if (n_kernels) *n_kernels = (uint32_t)ch.unique.size();
  if (n_launches) *n_launches = (uint32_t)ch.launches.size();
  if (n_handed_over) {
    *n_handed_over = 0;
    for (const SpecChunk &c : ch.chunks) *n_handed_over += (uint32_t)c.live_out.size();
  }
  CLG_CATCH
}

// ---- device -------------------------------------------------------------------------------------------------
int clg_ctx_create(int device, clg_ctx **out) {
  CLG_TRY
  NEED(out);
  *out = new clg_ctx{ctx_create(device)};
  CLG_CATCH
}
void clg_ctx_destroy(clg_ctx *ctx) {
  if (!ctx) return;
  try {
    ctx_destroy(ctx->c);
  } catch (...) {
  }
  delete ctx;
}
int clg_ctx_device_name(clg_ctx *ctx, char *buf, size_t cap) {
  CLG_TRY
  NEED(ctx), NEED(buf);
  std::snprintf(buf, cap, "%s", ctx_name(ctx->c).c_str());
  CLG_CATCH
}
int clg_ctx_sm_count(clg_ctx *ctx) { return ctx ? ctx_sm_count(ctx->c) : 0; }
int clg_ctx_synchronize(clg_ctx *ctx) {
  CLG_TRY
  NEED(ctx);
  ctx_sync(ctx->c);
  CLG_CATCH
}
int clg_dev_alloc(clg_ctx *ctx, size_t bytes, void **dptr) {
  CLG_TRY
  NEED(ctx), NEED(dptr);
  *dptr = dev_alloc(ctx->c, bytes);
  CLG_CATCH
}
int clg_dev_free(clg_ctx *ctx, void *dptr) {
  CLG_TRY
  NEED(ctx);
  if (dptr) dev_free(ctx->c, dptr);
  CLG_CATCH
}
int clg_dev_memset(clg_ctx *ctx, void *dptr, int byte, size_t bytes, void *stream) {
  CLG_TRY
  NEED(ctx), NEED(dptr);
  dev_memset(ctx->c, dptr, byte, bytes, stream);
  CLG_CATCH
}
int clg_memcpy_h2d(clg_ctx *ctx, void *dst, const void *src, size_t bytes, void *stream) {
  CLG_TRY
  NEED(ctx), NEED(dst), NEED(src);
  h2d(ctx->c, dst, src, bytes, stream);
  CLG_CATCH
}
int clg_memcpy_d2h(clg_ctx *ctx, void *dst, const void *src, size_t bytes, void *stream) {
  CLG_TRY
  NEED(ctx), NEED(dst), NEED(src);
  d2h(ctx->c, dst, src, bytes, stream);
  CLG_CATCH
}
int clg_host_alloc_pinned(clg_ctx *ctx, size_t bytes, void **hptr) {
  CLG_TRY
  NEED(ctx), NEED(hptr);
  *hptr = host_alloc(ctx->c, bytes);
  CLG_CATCH
}
int clg_host_free_pinned(clg_ctx *ctx, void *hptr) {
  CLG_TRY
  NEED(ctx);
  if (hptr) host_free(ctx->c, hptr);
  CLG_CATCH
}

int clg_exec_create(clg_ctx *ctx, const clg_flat *flat, int mode, clg_exec **out) {
  CLG_TRY
  NEED(ctx), NEED(flat), NEED(out);
  *out = new clg_exec{exec_create(ctx->c, flat->f, mode)};
  CLG_CATCH
}
void clg_exec_destroy(clg_exec *ex) {
  if (!ex) return;
  try {
    exec_destroy(ex->e);
  } catch (...) {
  }
  delete ex;
}
int clg_exec_mode(const clg_exec *ex) { return ex ? exec_mode(ex->e) : 0; }
int clg_exec_words_per_thread(const clg_exec *ex) { return ex ? exec_k(ex->e) : 0; }
int clg_exec_launches_per_run(const clg_exec *ex) { return ex ? exec_launches(ex->e) : 0; }
const char *clg_exec_kernel_name(const clg_exec *ex) { return ex ? exec_kernel_name(ex->e) : ""; }
const char *clg_exec_ptx(const clg_exec *ex) { return ex ? exec_ptx(ex->e).c_str() : ""; }

int clg_exec_prepare(clg_exec *ex, size_t n_vectors) {
  CLG_TRY
  NEED(ex);
  exec_prepare(ex->e, n_vectors);
  CLG_CATCH
}

int clg_exec_run(clg_exec *ex, const uint32_t *imm, uint32_t *out, uint32_t *state, size_t n_vectors, size_t stride, void *stream) {
  CLG_TRY
  NEED(ex);
  exec_run(ex->e, imm, out, state, n_vectors, stride, stream);
  CLG_CATCH
}

int clg_exec_run_steps(clg_exec *ex, const uint32_t *imm, uint32_t *out, uint32_t *state, size_t n_vectors, size_t stride,
                       size_t n_steps, void *stream) {
  CLG_TRY
  NEED(ex);
  exec_run_steps(ex->e, imm, out, state, n_vectors, stride, n_steps, stream);
  CLG_CATCH
}

namespace {
struct DevBuf {
  Ctx *c;
  void *p = nullptr;
  DevBuf(Ctx *c_, size_t bytes) : c(c_) { p = dev_alloc(c, bytes); }
  ~DevBuf() {
    try {
      if (p) dev_free(c, p);
    } catch (...) {
    }
  }
};
}  // namespace

int clg_exec_run_host(clg_exec *ex, const uint32_t *imm, uint32_t *out, uint32_t *state, size_t n_vectors, size_t stride) {
  CLG_TRY
  NEED(ex);
  exec_run_host(ex->e, imm, out, state, n_vectors, stride);
  CLG_CATCH
}

int clg_exec_run_bools_host(clg_exec *ex, const uint8_t *imm, uint8_t *out, size_t n_vectors) {
  CLG_TRY
  NEED(ex);
  exec_run_bools_host(ex->e, imm, out, n_vectors);
  CLG_CATCH
}

int clg_slice_bools(clg_ctx *ctx, const uint8_t *bools, uint32_t *sliced, size_t n_vectors, size_t n_rows, size_t stride, void *stream) {
  CLG_TRY
  NEED(ctx), NEED(bools), NEED(sliced);
  launch_slice(ctx->c, false, bools, sliced, n_vectors, n_rows, stride, stream);
  CLG_CATCH
}
int clg_unslice_bools(clg_ctx *ctx, const uint32_t *sliced, uint8_t *bools, size_t n_vectors, size_t n_rows, size_t stride, void *stream) {
  CLG_TRY
  NEED(ctx), NEED(bools), NEED(sliced);
  launch_slice(ctx->c, true, sliced, bools, n_vectors, n_rows, stride, stream);
  CLG_CATCH
}
int clg_measure_lop3_peak(clg_ctx *ctx, double *lane_ops_per_second) {
  CLG_TRY
  NEED(ctx), NEED(lane_ops_per_second);
  *lane_ops_per_second = measure_lop3_peak(ctx->c, 5);
  CLG_CATCH
}
int clg_fill_random(clg_ctx *ctx, uint32_t *sliced, size_t n_rows, size_t stride, uint64_t seed, void *stream) {
  CLG_TRY
  NEED(ctx), NEED(sliced);
  launch_fill(ctx->c, sliced, n_rows, stride, seed, stream);
  CLG_CATCH
}

// ---- Simulation::run for one vector, carried registers, on the device -----------------------------------------
struct clg_sim_runner {
  Ctx *c;
  clg_simulation *sim;
  Exec *e = nullptr;
  size_t n_imm;
  void *d_imm = nullptr, *d_out = nullptr, *d_state = nullptr;
  std::vector<uint32_t> h_imm, h_out, h_state;
  clg_sim_runner(Ctx *c_, clg_simulation *s_, size_t n) : c(c_), sim(s_), n_imm(n) {}
  ~clg_sim_runner() {
    try {
      if (e) exec_destroy(e);
      if (d_imm) dev_free(c, d_imm);
      if (d_out) dev_free(c, d_out);
      if (d_state) dev_free(c, d_state);
    } catch (...) {
    }
  }
};

int clg_sim_runner_create(clg_ctx *ctx, clg_simulation *sim, size_t n_immediates, int mode, clg_sim_runner **out) {
  CLG_TRY
  NEED(ctx), NEED(sim), NEED(out);
  Flat f = levelize(sim->sim, n_immediates, nullptr, 0, CLG_LV_STATEFUL);
  validate_flat(f);
  auto *r = new clg_sim_runner(ctx->c, sim, n_immediates);
  try {
    r->e = exec_create(ctx->c, f, mode);
    const FlatHeader &h = f.header();
    r->h_imm.assign((size_t)h.n_inputs * 4, 0), r->h_out.assign((size_t)h.n_outputs * 4, 0), r->h_state.assign((size_t)h.n_state * 4, 0);
    r->d_imm = dev_alloc(ctx->c, r->h_imm.size() * 4), r->d_out = dev_alloc(ctx->c, r->h_out.size() * 4);
    r->d_state = dev_alloc(ctx->c, r->h_state.size() * 4);
  } catch (...) {
    delete r;
    throw;
  }
  *out = r;
  CLG_CATCH
}
void clg_sim_runner_destroy(clg_sim_runner *r) { delete r; }

int clg_sim_runner_run(clg_sim_runner *r, const uint8_t *immediates, size_t n_immediates) {
  CLG_TRY
  NEED(r);
  if (n_immediates != r->n_imm) fail(CLG_E_INVALID, "runner was built for a different immediates length");
  if (n_immediates) NEED(immediates);
  const Flat &f = exec_flat(r->e);
  const FlatHeader &h = f.header();
  std::vector<uint8_t> &regs = r->sim->sim.registers;
  if (regs.size() != h.n_registers) fail(CLG_E_INVALID, "Simulation::registers was resized after the runner was created");
  for (size_t i = 0; i < h.n_inputs; i++) r->h_imm[i * 4] = immediates[i] ? 1u : 0u;
  const uint64_t *sr = f.state_regs();
  for (size_t k = 0; k < h.n_state; k++) r->h_state[k * 4] = regs[sr[k]] ? 1u : 0u;  // registers are public and writable
  if (h.n_inputs) h2d(r->c, r->d_imm, r->h_imm.data(), r->h_imm.size() * 4, nullptr);
  if (h.n_state) h2d(r->c, r->d_state, r->h_state.data(), r->h_state.size() * 4, nullptr);
  exec_run(r->e, (const uint32_t *)r->d_imm, (uint32_t *)r->d_out, h.n_state ? (uint32_t *)r->d_state : nullptr, 1, 4, nullptr);
  if (h.n_outputs) d2h(r->c, r->h_out.data(), r->d_out, r->h_out.size() * 4, nullptr);
  stream_sync(r->c, nullptr);
  for (size_t i = 0; i < h.n_outputs; i++) regs[i] = (uint8_t)(r->h_out[i * 4] & 1u);  // outputs == every register, in order
  CLG_CATCH
}

// ---- single-process multi-GPU sharding --------------------------------------------------------------------------
static void shard_range(size_t n_vectors, int n_shards, int shard, size_t &w0, size_t &w1) {
  const size_t words = (n_vectors + 31) / 32, quads = (words + 3) / 4;
  w0 = std::min(words, quads * (size_t)shard / (size_t)n_shards * 4);
  w1 = std::min(words, quads * (size_t)(shard + 1) / (size_t)n_shards * 4);
}

int clg_shard_range(size_t n_vectors, int n_shards, int shard, size_t *w0, size_t *w1) {
  CLG_TRY
  NEED(w0), NEED(w1);
  if (n_shards < 1 || shard < 0 || shard >= n_shards) fail(CLG_E_INVALID, "shard index out of range");
  shard_range(n_vectors, n_shards, shard, *w0, *w1);
  CLG_CATCH
}

int clg_run_sharded_host(const clg_flat *flat, int mode, const int *devices, int n_devices, const uint32_t *imm, uint32_t *out,
                         size_t n_vectors, size_t stride, double *seconds_out) {
  CLG_TRY
  NEED(flat), NEED(devices);
  if (n_devices < 1) fail(CLG_E_INVALID, "need at least one device");
  const FlatHeader &h = flat->f.header();
  if (h.n_state) fail(CLG_E_UNSUPPORTED, "sharded runs are for stateless programs");
  if (stride % 4) fail(CLG_E_INVALID, "stride must be a multiple of 4 words");
  std::vector<Error> errs(n_devices, Error{CLG_OK, ""});
  std::vector<std::thread> th;
  for (int d = 0; d < n_devices; d++)
    th.emplace_back([&, d] {
      try {
        size_t w0, w1;  // contiguous ranges of 4-word quads: rows stay 16-byte aligned on every shard
        shard_range(n_vectors, n_devices, d, w0, w1);
        if (seconds_out) seconds_out[d] = 0;
        if (w1 <= w0) return;
        const size_t nv = std::min(n_vectors, w1 * 32) - w0 * 32;
        std::unique_ptr<Ctx, void (*)(Ctx *)> c(ctx_create(devices[d]), ctx_destroy);
        std::unique_ptr<Exec, void (*)(Exec *)> e(exec_create(c.get(), flat->f, mode), exec_destroy);
        // this device's shard is words [w0, w1) of every row: same row stride, pointers moved by w0; the host-buffer
        // path pipelines H2D | kernel | D2H chunks on this device's own streams
        const auto t0 = std::chrono::steady_clock::now();
        exec_run_host(e.get(), imm ? imm + w0 : nullptr, out ? out + w0 : nullptr, nullptr, nv, stride);
        if (seconds_out) seconds_out[d] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      } catch (const Error &e) {
        errs[d] = e;
      } catch (const std::exception &e) {
        errs[d] = Error{CLG_E_INVALID, e.what()};
      }
    });
  for (auto &t : th) t.join();
  for (const Error &e : errs)
    if (e.code != CLG_OK) throw e;
  CLG_CATCH
}

// ---- persistent single-process multi-GPU executor -----------------------------------------------------------------
// One worker thread per device, each with its own context and executor created ONCE: resident programs, assembled
// kernels, tuned launch shapes and the staging buffers of the host path live across calls. A run posts one job; the
// workers meet at a start barrier (so the per-device times are comparable and no device idles while another one is still
// setting up), run their shard of words through the pipelined host-buffer path, and report back.
struct clg_sharded {
  Flat flat;
  int mode = CLG_MODE_AUTO;
  struct Worker {
    int device = 0;
    Ctx *c = nullptr;
    Exec *e = nullptr;
    std::thread th;
  };
  std::vector<Worker> w;
  std::mutex m;
  std::condition_variable cv_job, cv_done, cv_start;
  uint64_t gen = 0;
  int pending = 0, at_start = 0;
  bool quit = false;
  // the job
  const uint32_t *imm = nullptr;
  uint32_t *out = nullptr, *state = nullptr;
  size_t nv = 0, stride = 0, steps = 1;
  std::vector<Error> errs;
  std::vector<double> secs;

  void worker(int d) {
    uint64_t seen = 0;
    try {
      w[d].c = ctx_create(w[d].device);
      w[d].e = exec_create(w[d].c, flat, mode);
    } catch (const Error &e) {
      errs[d] = e;
    } catch (const std::exception &e) {
      errs[d] = Error{CLG_E_INVALID, e.what()};
    }
    {
      std::lock_guard<std::mutex> lk(m);
      pending--;
    }
    cv_done.notify_all();
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(m);
        cv_job.wait(lk, [&] { return quit || gen != seen; });
        if (quit) break;
        seen = gen;
        // start barrier: nobody begins before every worker has picked the job up
        if (++at_start == (int)w.size()) cv_start.notify_all();
        else cv_start.wait(lk, [&] { return at_start == (int)w.size(); });
      }
      try {
        size_t w0, w1;  // contiguous ranges of 4-word quads: rows stay 16-byte aligned on every shard
        shard_range(nv, (int)w.size(), d, w0, w1);
        secs[d] = 0;
        if (w1 > w0) {
          const FlatHeader &h = flat.header();
          const size_t n = std::min(nv, w1 * 32) - w0 * 32;
          const size_t imm_step = (size_t)h.n_inputs * stride, out_step = (size_t)h.n_outputs * stride;
          const auto t0 = std::chrono::steady_clock::now();
          // this device's shard is words [w0, w1) of every row (immediates, outputs and carried state alike): same row
          // stride, pointers moved by w0; consecutive steps carry the state through the caller's buffer
          for (size_t st = 0; st < steps; st++)
            exec_run_host(w[d].e, imm ? imm + st * imm_step + w0 : nullptr, out ? out + st * out_step + w0 : nullptr, state ? state + w0 : nullptr, n, stride);
          secs[d] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        }
      } catch (const Error &e) {
        errs[d] = e;
      } catch (const std::exception &e) {
        errs[d] = Error{CLG_E_INVALID, e.what()};
      }
      {
        std::lock_guard<std::mutex> lk(m);
        pending--;
      }
      cv_done.notify_all();
    }
    if (w[d].e) exec_destroy(w[d].e);
    if (w[d].c) ctx_destroy(w[d].c);
  }
  void wait_all() {
    std::unique_lock<std::mutex> lk(m);
    cv_done.wait(lk, [&] { return pending == 0; });
  }
  void throw_first_error() {
    for (Error &e : errs)
      if (e.code != CLG_OK) {
        Error copy = e;
        e = Error{CLG_OK, ""};
        throw copy;
      }
  }
  ~clg_sharded() {
    {
      std::lock_guard<std::mutex> lk(m);
      quit = true;
    }
    cv_job.notify_all();
    for (Worker &x : w)
      if (x.th.joinable()) x.th.join();
  }
};

int clg_sharded_create(const clg_flat *flat, int mode, const int *devices, int n_devices, clg_sharded **out) {
  CLG_TRY
  NEED(flat), NEED(devices), NEED(out);
  if (n_devices < 1 || n_devices > 64) fail(CLG_E_INVALID, "need between 1 and 64 devices");
  std::unique_ptr<clg_sharded> s(new clg_sharded());
  s->flat = flat->f, s->mode = mode;
  s->w.resize(n_devices), s->errs.assign(n_devices, Error{CLG_OK, ""}), s->secs.assign(n_devices, 0.0);
  s->pending = n_devices;
  for (int d = 0; d < n_devices; d++) s->w[d].device = devices[d];
  for (int d = 0; d < n_devices; d++) s->w[d].th = std::thread([p = s.get(), d] { p->worker(d); });
  s->wait_all();  // contexts and executors exist (or an error is recorded)
  s->throw_first_error();
  *out = s.release();
  CLG_CATCH
}
void clg_sharded_destroy(clg_sharded *s) { delete s; }
int clg_sharded_n_devices(const clg_sharded *s) { return s ? (int)s->w.size() : 0; }

static void sharded_run(clg_sharded *s, const uint32_t *imm, uint32_t *out, uint32_t *state, size_t n_vectors, size_t stride, size_t n_steps,
                        double *seconds_out) {
  const FlatHeader &h = s->flat.header();
  if (stride % 4) fail(CLG_E_INVALID, "stride must be a multiple of 4 words");
  if ((n_vectors + 31) / 32 > stride) fail(CLG_E_INVALID, "n_vectors does not fit in stride words");
  if (h.n_inputs && !imm) fail(CLG_E_INVALID, "imm is null");
  if (h.n_outputs && !out) fail(CLG_E_INVALID, "out is null");
  if (h.n_state && !state) fail(CLG_E_INVALID, "state is null for a stateful program");
  {
    std::lock_guard<std::mutex> lk(s->m);
    s->imm = imm, s->out = out, s->state = state, s->nv = n_vectors, s->stride = stride, s->steps = n_steps;
    s->pending = (int)s->w.size(), s->at_start = 0;
    s->gen++;
  }
  s->cv_job.notify_all();
  s->wait_all();
  if (seconds_out)
    for (size_t d = 0; d < s->w.size(); d++) seconds_out[d] = s->secs[d];
  s->throw_first_error();
}
int clg_sharded_run_host(clg_sharded *s, const uint32_t *imm, uint32_t *out, uint32_t *state, size_t n_vectors, size_t stride, double *seconds_out) {
  CLG_TRY
  NEED(s);
  if (n_vectors) sharded_run(s, imm, out, state, n_vectors, stride, 1, seconds_out);
  CLG_CATCH
}
int clg_sharded_run_steps_host(clg_sharded *s, const uint32_t *imm, uint32_t *out, uint32_t *state, size_t n_vectors, size_t stride, size_t n_steps,
                               double *seconds_out) {
  CLG_TRY
  NEED(s);
  if (n_vectors && n_steps) sharded_run(s, imm, out, state, n_vectors, stride, n_steps, seconds_out);
  CLG_CATCH
}

// ---- JSON ---------------------------------------------------------------------------------------------------------
int clg_simulation_to_json(const clg_simulation *sim, char **json_out) {
  CLG_TRY
  NEED(sim), NEED(json_out);
  std::string s = simulation_to_json(sim->sim);
  char *p = (char *)std::malloc(s.size() + 1);
  if (!p) throw std::bad_alloc();
  std::memcpy(p, s.c_str(), s.size() + 1);
  *json_out = p;
  CLG_CATCH
}
int clg_simulation_from_json(const char *json, clg_simulation **out) {
  CLG_TRY
  NEED(json), NEED(out);
  auto *s = new clg_simulation();
  try {
    s->sim = simulation_from_json(json);
  } catch (...) {
    delete s;
    throw;
  }
  *out = s;
  CLG_CATCH
}

static char *dup_string(const std::string &s) {
  char *p = (char *)std::malloc(s.size() + 1);
  if (!p) throw std::bad_alloc();
  std::memcpy(p, s.c_str(), s.size() + 1);
  return p;
}
int clg_compiler_to_json(const clg_compiler *c, char **json_out) {
  CLG_TRY
  NEED(c), NEED(json_out);
  *json_out = dup_string(compiler_to_json(c->c));
  CLG_CATCH
}
int clg_compiler_from_json(const char *json, clg_compiler **out) {
  CLG_TRY
  NEED(json), NEED(out);
  Compiler parsed = compiler_from_json(json);
  auto *c = new clg_compiler(parsed.immediate_count);
  c->c = std::move(parsed);
  *out = c;
  CLG_CATCH
}
int clg_gate_to_json(const clg_gate *gate, char **json_out) {
  CLG_TRY
  NEED(gate), NEED(json_out);
  *json_out = dup_string(gate_to_json(*gate));
  CLG_CATCH
}
int clg_gate_from_json(const char *json, clg_gate *out) {
  CLG_TRY
  NEED(json), NEED(out);
  *out = gate_from_json(json);
  CLG_CATCH
}

}  // extern "C"
