/*
 * complogic_cuda.h -- C ABI of the B200-native backend for complogic's NAND-VM hot path.
 *
 * This is the drop-in boundary: exactly what a Rust `complogic-cuda` crate (or any FFI host)
 * binds. Plain pointers and sizes only; no C++/torch types; nothing unwinds across it.
 * Every entry point cites the reference interface (Vandesm14/complogic, file:line) it replaces
 * or sits beside. The reference has no FFI of its own; its boundary is the Rust public API
 * re-exported at complogic/src/lib.rs:5-7.
 *
 * Conventions
 *   - every function returning `int` returns a clg_status; clg_last_error() gives the text
 *     (thread-local). Where the reference would panic (index out of bounds at
 *     simulation.rs:20-26, compile.rs:137) the call returns CLG_E_BAD_REGISTER instead.
 *   - opaque handles are created by *_new/_create/_from_* and released by the matching
 *     *_destroy; arrays handed back through `T**` are released with clg_free().
 *   - bit-sliced layout (the device ABI): a batch of N vectors is `rows x stride` uint32 words,
 *     row r = immediate (or output) r, vector v <-> bit (v % 32) of word (v / 32) of each row.
 *     `stride` (words between rows) must be a multiple of 4 so rows stay 16-byte aligned and the
 *     kernels can use 128-bit accesses; lanes past N in the last words are don't-care on input
 *     and unspecified on output.
 *   - a clg_ctx is bound to one CUDA device and used by one host thread at a time; multi-GPU is
 *     one ctx (and one host thread or process) per device. Programs and flat buffers are
 *     immutable after creation and may be shared between contexts.
 *   - there is NO CPU execution path in this library: without a CUDA driver and a device,
 *     clg_ctx_create fails with CLG_E_NO_DRIVER and nothing can be run.
 */
#ifndef COMPLOGIC_CUDA_H
#define COMPLOGIC_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden */
#endif

#define CLG_ABI_VERSION 1u

typedef enum clg_status {
  CLG_OK = 0,
  CLG_E_INVALID = 1,      /* bad argument (null pointer, misaligned stride, ...) */
  CLG_E_BAD_REGISTER = 2, /* register index out of range: where the reference panics */
  CLG_E_NOMEM = 3,
  CLG_E_CUDA = 4,        /* a CUDA driver call failed; see clg_last_error() */
  CLG_E_NO_DRIVER = 5,   /* libcuda.so.1 / a device is not available: nothing can run */
  CLG_E_UNSUPPORTED = 6, /* program does not fit the requested execution mode */
  CLG_E_PTX = 7,         /* the specialised-kernel assembler rejected generated code */
  CLG_E_FORMAT = 8       /* malformed flat buffer / JSON */
} clg_status;

const char *clg_last_error(void);
uint32_t clg_abi_version(void);
void clg_free(void *p);

/* ---------------------------------------------------------------------------------------------
 * Instruction set -- `enum Op { Nand(usize,usize,usize), Set(usize,bool) }`, compile.rs:7-14.
 * 32 bytes like the Rust enum. tag 0: Nand(a, b, out = c).  tag 1: Set(id = a, val = b != 0).
 * ------------------------------------------------------------------------------------------- */
#define CLG_OP_NAND 0u
#define CLG_OP_SET 1u
typedef struct clg_op {
  uint32_t tag;
  uint32_t reserved;
  uint64_t a, b, c;
} clg_op;

/* ---------------------------------------------------------------------------------------------
 * Gate library -- the 11 gate structs + test-only RSLatchTest, gates.rs:5-117.
 * `f` holds the struct's fields in declaration order:
 *   Nand{a,b,out} Not{a,out} And/Or/Nor/Xor{a,b,out} RSLatch{s,r,q} DLatch{d,e,q}
 *   HalfAdder{a,b,s,c} FullAdder{a,b,cin,s,cout} FourBitAdder{a1..a4,b1..b4,s1..s4,cout}
 * ------------------------------------------------------------------------------------------- */
typedef enum clg_gate_kind {
  CLG_GATE_NAND = 0,
  CLG_GATE_NOT = 1,
  CLG_GATE_AND = 2,
  CLG_GATE_OR = 3,
  CLG_GATE_NOR = 4,
  CLG_GATE_XOR = 5,
  CLG_GATE_RSLATCH = 6,
  CLG_GATE_RSLATCH_TEST = 7,
  CLG_GATE_DLATCH = 8,
  CLG_GATE_HALF_ADDER = 9,
  CLG_GATE_FULL_ADDER = 10,
  CLG_GATE_FOUR_BIT_ADDER = 11
} clg_gate_kind;

typedef struct clg_gate {
  uint32_t kind;
  uint32_t reserved;
  uint64_t f[13];
} clg_gate;

/* Gate::create(&self, &mut Incrementer) -> Ops, gates.rs:193. *incrementer_val is Incrementer.val
 * (compile.rs:19-49) and is advanced exactly as the reference advances it. */
int clg_gate_create(const clg_gate *gate, uint64_t *incrementer_val, clg_op **ops_out, size_t *n_ops_out);

/* ---------------------------------------------------------------------------------------------
 * Simulation -- `struct Simulation { pub ops: Vec<Op>, pub registers: Vec<bool> }`,
 * simulation.rs:5-12. Fields are public in the reference, so both are exposed as views; the
 * registers view is writable. (Running it is a device operation: see clg_sim_* below.)
 * ------------------------------------------------------------------------------------------- */
typedef struct clg_simulation clg_simulation;
int clg_simulation_new(const clg_op *ops, size_t n_ops, const uint8_t *registers, size_t n_registers,
                       clg_simulation **out);
void clg_simulation_destroy(clg_simulation *sim);
const clg_op *clg_simulation_ops(const clg_simulation *sim, size_t *n_ops);
uint8_t *clg_simulation_registers(clg_simulation *sim, size_t *n_registers);
/* Simulation::register(&self, id) -> bool, simulation.rs:33-35; *value_out = 0/1. */
int clg_simulation_register(const clg_simulation *sim, size_t id, int *value_out);
/* number of Op::Nand in `ops` -- the numerator unit of the gate-evals/s metric */
uint64_t clg_simulation_nand_count(const clg_simulation *sim);

/* ---------------------------------------------------------------------------------------------
 * Compiler -- `struct Compiler { pub immediate_count, pub ops, pub incrementer }`,
 * compile.rs:51-203.  The three public fields are reachable through get/set because the
 * reference's only in-tree caller mutates them directly (complogic-gui/src/app.rs:410,425).
 * ------------------------------------------------------------------------------------------- */
typedef struct clg_compiler clg_compiler;
int clg_compiler_new(uint64_t immediate_count, clg_compiler **out); /* Compiler::new, compile.rs:65 */
void clg_compiler_destroy(clg_compiler *c);
void clg_compiler_reset_ops(clg_compiler *c);         /* compile.rs:74-80 */
void clg_compiler_reset_incrementer(clg_compiler *c); /* compile.rs:83-85 */
uint64_t clg_compiler_alloc(clg_compiler *c);         /* compile.rs:88-90 */
uint64_t clg_compiler_immediate_count(const clg_compiler *c);
void clg_compiler_set_immediate_count(clg_compiler *c, uint64_t v);
uint64_t clg_compiler_incrementer(const clg_compiler *c);
void clg_compiler_set_incrementer(clg_compiler *c, uint64_t v); /* Incrementer::set / skip */
const clg_op *clg_compiler_ops(const clg_compiler *c, size_t *n_ops);
/* Compiler::compile(&mut self, Vec<&Gate>) -> Simulation, compile.rs:93-202: lower every gate,
 * then emit ops in the reference's dependency order (same op sequence, op for op).
 * Deviation, on purpose: the reference writes ./graph.dot on every compile (compile.rs:150);
 * this keeps the text in memory instead -- clg_compiler_graph_dot(). */
int clg_compiler_compile(clg_compiler *c, const clg_gate *gates, size_t n_gates, clg_simulation **out);
const char *clg_compiler_graph_dot(const clg_compiler *c);

/* ---------------------------------------------------------------------------------------------
 * The new compiler pass (north-star): SSA-rename the in-order NAND program, fold constants,
 * map NAND cones onto 3-input LUTs (one LOP3 each), levelize, register-allocate, and serialise
 * a flat, level-ordered device buffer.
 *
 *   n_immediates  length of the immediates slice every run() of the batch will pass: a Set(id,_)
 *                 with id < n_immediates reads input row id, any other Set writes its literal
 *                 (simulation.rs:26, `immediates.get(id).copied().unwrap_or(val)`).
 *   out_regs      registers the caller wants to read back (Simulation::registers is fully
 *                 public in the reference; at 10^6 registers x 2^24 vectors that cannot be
 *                 materialised, so the observable set is named up front). NULL + n_out == 0
 *                 means "every register".
 *   flags         CLG_LV_STATEFUL: registers that a run() can read before it writes them are
 *                 carried between runs per lane (simulation.rs never clears `registers`; this
 *                 is what makes latches work, gates.rs:510-571). Without it every lane starts
 *                 from an all-false Simulation and runs once.
 * ------------------------------------------------------------------------------------------- */
#define CLG_LV_STATEFUL 1u
#define CLG_LV_NO_FUSE 2u      /* keep one LUT per surviving NAND (no cone mapping) */
#define CLG_LV_NO_LANE_STREAM 4u /* do not build the lane-ordered stream */
#define CLG_LV_NO_PARTS 8u      /* do not build the partitioned (per connected component) level streams */
#define CLG_LV_NO_ROWS 16u      /* do not build the row stream of the cooperative row kernel */

typedef struct clg_flat clg_flat;
int clg_levelize(const clg_simulation *sim, size_t n_immediates, const uint64_t *out_regs, size_t n_out,
                 uint32_t flags, clg_flat **out);
void clg_flat_destroy(clg_flat *f);
/* the serialised buffer (versioned header, documented in DESIGN.md); valid while `f` lives */
int clg_flat_bytes(const clg_flat *f, const void **data, size_t *len);
/* re-open a serialised buffer (validates every index and the level hazards) */
int clg_flat_from_bytes(const void *data, size_t len, clg_flat **out);

typedef struct clg_flat_info {
  uint64_t ref_nand_count; /* Op::Nand in the source program */
  uint64_t ref_set_count;
  uint32_t n_inputs;  /* immediate rows actually read (= n_immediates) */
  uint32_t n_outputs; /* named output rows */
  uint32_t n_state;   /* carried state rows (stateful) */
  uint32_t n_luts;    /* LUT3 instructions after fusion */
  uint32_t n_levels;
  uint32_t max_level_width;
  uint32_t level_slots; /* physical registers of the level-ordered stream */
  uint32_t lane_slots;  /* physical registers of the lane-ordered stream (0 if absent) */
  uint32_t level_instrs;
  uint32_t lane_instrs;
  uint32_t n_parts; /* independent parts of the partitioned level stream (0: one component) */
  uint32_t reserved;
  /* row stream (cooperative row kernel; all 0 if absent) */
  uint32_t rows_segments;  /* segments (a warp may run lag segments ahead of the slowest one) */
  uint32_t rows_slots;     /* physical registers */
  uint32_t rows_instrs;    /* lane instructions, empty lanes included (32 per row) */
  uint32_t rows_from_reg;  /* LUTs that take one operand from the lane's registers instead of shared memory */
  uint32_t rows_no_store;  /* LUTs whose result is never written to shared memory */
  uint32_t rows_warps;     /* warps of the CTA the stream was scheduled for */
} clg_flat_info;
int clg_flat_get_info(const clg_flat *f, clg_flat_info *info);
/* registers behind the state rows, in row order (stateful programs) */
int clg_flat_state_regs(const clg_flat *f, const uint64_t **regs, size_t *n);
/* PTX text of the specialised kernel for this program at k words per thread (1, 2 or 4); needs no
 * device, so it can be assembled and inspected offline (ptxas -arch=sm_100a). Free with clg_free. */
int clg_flat_spec_ptx(const clg_flat *f, int k, char **ptx_out);
/* same for the fused multi-step form (clg_exec_run_steps) */
int clg_flat_spec_steps_ptx(const clg_flat *f, int k, char **ptx_out);
/* Programs longer than 16384 lane-order instructions (fewer at k > 1) are specialised as a chain of chunk kernels
 * (entry clg_spec_chunk), cut where few registers are live; registers live across a cut are handed over through a
 * device scratch buffer. A chunk with nothing live at either end addresses its rows relative to the lowest row it
 * touches, so that the instances of a flattened program share one kernel: row_bases (optional, [3]) receives the
 * rows by which the imm / out / state pointers are advanced for this chunk. Text of chunk `chunk`; *n_chunks
 * (optional) receives the chain length (1 for a short program). Free with clg_free. */
int clg_flat_spec_chunk_ptx(const clg_flat *f, int k, uint32_t chunk, uint32_t *n_chunks, uint32_t *row_bases, char **ptx_out);
/* Shape of that chain without generating any code: chunks, distinct kernels that would have to be assembled (about a
 * second each on the first run), and launches per run (equal independent chunks share one launch). Any may be NULL. */
int clg_flat_spec_chain_info(const clg_flat *f, int k, uint32_t *n_chunks, uint32_t *n_kernels, uint32_t *n_launches);
/* The same for the SECOND chain, the one big batches run (a launch that fills the machine many times over): chunks short
 * enough to stay in the SM's instruction cache (CLG_SPEC_TP_CHUNK lane-order instructions, 6 800 / k by default), cut at
 * the first boundary no register crosses -- one instance of a flattened program per kernel, all instances side by side
 * in one launch -- and, where an instance is longer than that, into equal parts. *n_handed_over (optional): registers
 * that cross a cut, summed over the chain (0: the chain is taken whatever the program's register count). */
int clg_flat_spec_batch_chain_info(const clg_flat *f, int k, uint32_t *n_chunks, uint32_t *n_kernels, uint32_t *n_launches, uint32_t *n_handed_over);
/* Text of chunk `chunk` of that second chain (arguments as for clg_flat_spec_chunk_ptx). */
int clg_flat_spec_batch_chunk_ptx(const clg_flat *f, int k, uint32_t chunk, uint32_t *n_chunks, uint32_t *row_bases, char **ptx_out);

/* ---------------------------------------------------------------------------------------------
 * Device side. Raw CUDA driver API underneath (libcuda.so.1 is dlopen'ed when the first context
 * is created); the context is the device's primary context, so pointers from cudaMalloc / torch
 * are usable directly.
 * ------------------------------------------------------------------------------------------- */
typedef struct clg_ctx clg_ctx;
int clg_ctx_create(int device, clg_ctx **out);
void clg_ctx_destroy(clg_ctx *ctx);
int clg_ctx_device_name(clg_ctx *ctx, char *buf, size_t cap);
int clg_ctx_sm_count(clg_ctx *ctx);
int clg_ctx_synchronize(clg_ctx *ctx);

int clg_dev_alloc(clg_ctx *ctx, size_t bytes, void **dptr);
int clg_dev_free(clg_ctx *ctx, void *dptr);
int clg_dev_memset(clg_ctx *ctx, void *dptr, int byte, size_t bytes, void *stream);
int clg_memcpy_h2d(clg_ctx *ctx, void *dst, const void *src, size_t bytes, void *stream);
int clg_memcpy_d2h(clg_ctx *ctx, void *dst, const void *src, size_t bytes, void *stream);
int clg_host_alloc_pinned(clg_ctx *ctx, size_t bytes, void **hptr);
int clg_host_free_pinned(clg_ctx *ctx, void *hptr);

/* execution modes */
typedef enum clg_mode {
  CLG_MODE_AUTO = 0,
  CLG_MODE_LANE = 1, /* lane-parallel interpreter: one thread = K words, register stack in smem */
  CLG_MODE_COOP = 2, /* cooperative intra-level interpreter: one CTA shares the stack, threads split a level */
  CLG_MODE_SPEC = 3  /* lane-parallel specialised kernel: straight-line LOP3 code generated per program */
} clg_mode;

/* a program made resident on one device (uploaded streams, specialised module if any) */
typedef struct clg_exec clg_exec;
int clg_exec_create(clg_ctx *ctx, const clg_flat *flat, int mode, clg_exec **out);
void clg_exec_destroy(clg_exec *ex);
int clg_exec_mode(const clg_exec *ex); /* the mode actually chosen */
int clg_exec_words_per_thread(const clg_exec *ex); /* K: 32-bit words of each row one thread (lane/spec) or CTA (coop) owns */
/* kernel launches one clg_exec_run issues in the mode last used (1, or the chain length of a chunked specialised program) */
int clg_exec_launches_per_run(const clg_exec *ex);
/* name of the kernel the most recent run launched ("clg_coop_rows", "clg_coop_c_k4", "clg_spec", ...; "" before any run);
 * static storage, do not free */
const char *clg_exec_kernel_name(const clg_exec *ex);
/* specialised mode only: generated PTX text (NUL-terminated) for inspection */
const char *clg_exec_ptx(const clg_exec *ex);
/* Everything a first clg_exec_run of `n_vectors` lanes would otherwise do before its launch, done now: the mode is
 * chosen for that batch size (CLG_MODE_AUTO), its streams are encoded and uploaded, a specialised kernel or chain is
 * assembled, and the cooperative CTA size is timed on scratch buffers (a few tens of ms, once per plan). Synchronous. A
 * later run of a batch that selects another plan prepares that one on its own first call, as before. */
int clg_exec_prepare(clg_exec *ex, size_t n_vectors);

/* The hot path: `Simulation::run(&mut self, immediates: &[bool])`, simulation.rs:16-30, for
 * n_vectors independent lanes at once. All pointers are DEVICE pointers.
 *   imm    [n_inputs][stride]  bit-sliced immediates
 *   out    [n_outputs][stride] bit-sliced values of the named output registers after the run
 *   state  [n_state][stride]   carried registers, updated in place (NULL unless stateful)
 * Asynchronous on `stream` (a CUstream / cudaStream_t; NULL = the legacy default stream). Runs of ONE executor must not
 * overlap on the device: issue them on one stream, or order the streams with events -- a chunked specialised program
 * hands registers from kernel to kernel through a scratch buffer that belongs to the executor (growing it synchronises
 * the context), and the first big cooperative run times CTA sizes on the legacy stream unless clg_exec_prepare was called. */
int clg_exec_run(clg_exec *ex, const uint32_t *imm, uint32_t *out, uint32_t *state, size_t n_vectors,
                 size_t stride, void *stream);

/* n_steps consecutive run() calls per lane in one call (clocked / sequential circuits: what the reference does by
 * calling run() repeatedly on one Simulation, e.g. gates.rs:523-534). DEVICE pointers:
 *   imm [n_steps][n_inputs][stride]   out [n_steps][n_outputs][stride]   state [n_state][stride] (in place)
 * In the specialised mode this is ONE launch and the carried registers never leave the register file between
 * steps; the interpreter modes issue one launch per step. */
int clg_exec_run_steps(clg_exec *ex, const uint32_t *imm, uint32_t *out, uint32_t *state, size_t n_vectors, size_t stride,
                       size_t n_steps, void *stream);

/* Host-buffer entry point. Same layouts, HOST pointers (pinned memory -- clg_host_alloc_pinned -- gives full
 * PCIe speed). The batch is cut into column chunks that are pipelined over three streams, so H2D copies,
 * kernels and D2H copies overlap; returns when the results are in `out` (and `state`). */
int clg_exec_run_host(clg_exec *ex, const uint32_t *imm, uint32_t *out, uint32_t *state, size_t n_vectors,
                      size_t stride);

/* `&[bool]`-shaped host API: imm is [n_vectors][n_inputs] bytes (0/1, one run()'s immediates per
 * row, as built at complogic-gui/src/app.rs:543-552), out is [n_vectors][n_outputs] bytes.
 * Transposes to and from the bit-sliced layout on the device. */
int clg_exec_run_bools_host(clg_exec *ex, const uint8_t *imm, uint8_t *out, size_t n_vectors);

/* vector-major <-> bit-sliced transposes (device pointers). bools: [n_vectors][n_rows] bytes. */
int clg_slice_bools(clg_ctx *ctx, const uint8_t *bools, uint32_t *sliced, size_t n_vectors, size_t n_rows,
                    size_t stride, void *stream);
int clg_unslice_bools(clg_ctx *ctx, const uint32_t *sliced, uint8_t *bools, size_t n_vectors, size_t n_rows,
                      size_t stride, void *stream);

/* counter-based Bernoulli(1/2) generator used by the benchmarks: word (row, w) =
 * hi32(splitmix64(seed ^ (row << 40) ^ w)); fills rows x stride words on the device. */
int clg_fill_random(clg_ctx *ctx, uint32_t *sliced, size_t n_rows, size_t stride, uint64_t seed, void *stream);

/* Measured denominator of the integer-ALU roofline: a kernel of nothing but dependent LOP3 chains; returns
 * 32-bit LOP3 lane-operations per second on this device (nominal: 148 SM x 64 lanes/clk x SM clock). */
int clg_measure_lop3_peak(clg_ctx *ctx, double *lane_ops_per_second);

/* Simulation::run for ONE vector, on the device, with the reference's carried-register
 * semantics: binds `sim` to a 1-lane stateful program (all registers observable), and every
 * call evaluates one run(immediates) and writes sim.registers back. This is what lets the
 * reference's own unit tests execute through the CUDA path unchanged. */
typedef struct clg_sim_runner clg_sim_runner;
int clg_sim_runner_create(clg_ctx *ctx, clg_simulation *sim, size_t n_immediates, int mode, clg_sim_runner **out);
void clg_sim_runner_destroy(clg_sim_runner *r);
int clg_sim_runner_run(clg_sim_runner *r, const uint8_t *immediates, size_t n_immediates);

/* Shard arithmetic used by every multi-GPU driver: the words [*w0, *w1) of each row that shard `shard` of
 * `n_shards` owns. Ranges are contiguous, cover ceil(n_vectors/32) words exactly once, and start on multiples
 * of 4 words so that a shard's rows stay 16-byte aligned. Pure host logic (no device needed). */
int clg_shard_range(size_t n_vectors, int n_shards, int shard, size_t *w0, size_t *w1);

/* Single-process multi-GPU driver (north-star: the vector batch shards independently across the
 * GPUs of one box, per-GPU device-side result buffers copied back by the host; no collective).
 * HOST pointers; words are split into contiguous multiples-of-4 ranges, one host thread, context
 * and streams per device; every shard runs through the pipelined host-buffer path. seconds_out (optional, n_devices
 * entries) = wall time of each device's shard, copies included. */
int clg_run_sharded_host(const clg_flat *flat, int mode, const int *devices, int n_devices,
                         const uint32_t *imm, uint32_t *out, size_t n_vectors, size_t stride,
                         double *seconds_out);

/* The same driver, persistent: contexts, executors (resident programs, assembled kernels, tuned launch shapes) and
 * staging buffers are created once, one worker thread per device, and live across calls; every run starts the devices
 * together behind a barrier. Stateful programs are allowed: a carried-state row shards by words like any other row.
 * clg_sharded_run_steps_host runs n_steps consecutive run() calls per lane (imm is [n_steps][n_inputs][stride], out is
 * [n_steps][n_outputs][stride], state is carried from step to step through the caller's buffer). One caller thread at a
 * time per clg_sharded. */
typedef struct clg_sharded clg_sharded;
int clg_sharded_create(const clg_flat *flat, int mode, const int *devices, int n_devices, clg_sharded **out);
void clg_sharded_destroy(clg_sharded *s);
int clg_sharded_n_devices(const clg_sharded *s);
int clg_sharded_run_host(clg_sharded *s, const uint32_t *imm, uint32_t *out, uint32_t *state, size_t n_vectors,
                         size_t stride, double *seconds_out);
int clg_sharded_run_steps_host(clg_sharded *s, const uint32_t *imm, uint32_t *out, uint32_t *state, size_t n_vectors,
                               size_t stride, size_t n_steps, double *seconds_out);

/* serde_json-shaped interchange of `Simulation` (the derives at simulation.rs:5, compile.rs:7):
 * {"ops":[{"Nand":[a,b,out]},{"Set":[id,val]},...],"registers":[false,...]} */
int clg_simulation_to_json(const clg_simulation *sim, char **json_out);
int clg_simulation_from_json(const char *json, clg_simulation **out);
/* ... of `Compiler` (compile.rs:51-61; its Incrementer, compile.rs:18-21, as {"val":n}):
 * {"immediate_count":2,"ops":[...],"incrementer":{"val":3}}
 * and of `Gate` (gates.rs:5-117), serde's externally tagged enum around the variant's struct:
 * {"FullAdder":{"a":0,"b":1,"cin":2,"s":3,"cout":4}}. JSON text is returned malloc'ed: free with clg_free. */
int clg_compiler_to_json(const clg_compiler *c, char **json_out);
int clg_compiler_from_json(const char *json, clg_compiler **out);
int clg_gate_to_json(const clg_gate *gate, char **json_out);
int clg_gate_from_json(const char *json, clg_gate *out);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* COMPLOGIC_CUDA_H */
