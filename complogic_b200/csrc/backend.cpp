// CUDA driver binding, contexts, resident programs, launches. See device.hpp.
#include <cuda.h>
#include <dlfcn.h>

#include <algorithm>
#include <iterator>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "device.hpp"

#include "kernels_cubin.inc"  // static const unsigned char clg_cubin[] (nvcc -cubin of kernels.cu, sm_100a)

namespace clg {

namespace {

// ---- libcuda.so.1, bound at run time ---------------------------------------------------------------
#define CU_FNS(X)                                                                                                      \
  X(cuInit) X(cuDeviceGet) X(cuDeviceGetCount) X(cuDeviceGetName) X(cuDeviceGetAttribute) X(cuDevicePrimaryCtxRetain)  \
  X(cuDevicePrimaryCtxRelease) X(cuCtxSetCurrent) X(cuCtxSynchronize) X(cuModuleLoadData) X(cuModuleLoadDataEx)        \
  X(cuModuleGetFunction) X(cuModuleUnload) X(cuFuncSetAttribute) X(cuLaunchKernel) X(cuMemAlloc) X(cuMemFree)          \
  X(cuMemcpyHtoDAsync) X(cuMemcpyDtoHAsync) X(cuMemcpy2DAsync) X(cuStreamWaitEvent) X(cuMemsetD8Async) X(cuMemAllocHost) X(cuMemFreeHost) X(cuStreamSynchronize) \
  X(cuStreamCreate) X(cuStreamDestroy) X(cuEventCreate) X(cuEventRecord) X(cuEventSynchronize) X(cuEventElapsedTime)   \
  X(cuEventDestroy) X(cuGetErrorString) X(cuTensorMapEncodeTiled)

struct Driver {
#define X(name) decltype(&::name) name = nullptr;
  CU_FNS(X)
#undef X
  bool ok = false;
  std::string why;
};

Driver &driver() {
  static Driver d;
  static std::once_flag once;
  std::call_once(once, [] {
    void *h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libcuda.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
      d.why = std::string("cannot load the CUDA driver (libcuda.so.1): ") + dlerror();
      return;
    }
    // cuda.h maps the unsuffixed names onto the current ABI (_v2 ...); stringify after expansion
#define STR2(x) #x
#define STR(x) STR2(x)
#define X(name)                                                     \
  d.name = reinterpret_cast<decltype(d.name)>(dlsym(h, STR(name))); \
  if (!d.name) {                                                    \
    d.why = "CUDA driver lacks " STR(name);                         \
    return;                                                         \
  }
    CU_FNS(X)
#undef X
    CUresult r = d.cuInit(0);
    if (r != CUDA_SUCCESS) {
      d.why = "cuInit failed (" + std::to_string((int)r) + "): no usable CUDA device";
      return;
    }
    d.ok = true;
  });
  return d;
}

void cu_check(CUresult r, const char *what) {
  if (r == CUDA_SUCCESS) return;
  const char *s = nullptr;
  driver().cuGetErrorString(r, &s);
  fail(CLG_E_CUDA, std::string(what) + ": " + (s ? s : "unknown CUDA error") + " (" + std::to_string((int)r) + ")");
}
#define CU(call) cu_check(driver().call, #call)

}  // namespace

// ---- context ---------------------------------------------------------------------------------------
struct Ctx {
  CUdevice dev = 0;
  CUcontext cu = nullptr;
  CUmodule mod = nullptr;
  int sm_count = 0;
  CUfunction lane[3] = {}, lane_n[3] = {}, coop[3] = {}, coop_res[3] = {}, coop_c[3] = {}, coop_rows = nullptr, coop_rows_t = nullptr, slice = nullptr, unslice = nullptr, slice_v4 = nullptr, unslice_v4 = nullptr, slice_w[3] = {}, unslice_w[3] = {}, unslice_s64 = nullptr, slice_tma[7] = {}, unslice_tma[7] = {}, slice_gen = nullptr, unslice_gen = nullptr, fill = nullptr, lop3_peak = nullptr;  // index: K = 1,2,4 -> 0,1,2
  std::string name;

  static constexpr uint32_t GEN_MAX_BYTES = 96 * 1024;  // staged transposes for any row count (kernels.cu, clg_slice_gen)
  static constexpr uint32_t TMA_RT[7] = {32, 64, 64, 128, 64, 128, 32}, TMA_VT[7] = {1024, 1024, 512, 512, 256, 256, 512};  // tile shapes of the TMA-tiled transposes (kernels.cu)
  static size_t tma_smem(int i) { return (size_t)TMA_VT[i] * TMA_RT[i] + (size_t)TMA_RT[i] * (TMA_VT[i] / 32 + 1) * 4 + 16 + 1024; }  // (+ slack to align the base to 1024)

  explicit Ctx(int device) {
    Driver &d = driver();
    if (!d.ok) fail(CLG_E_NO_DRIVER, d.why + " -- this backend has no CPU execution path");
    int n = 0;
    CU(cuDeviceGetCount(&n));
    if (device < 0 || device >= n) fail(CLG_E_INVALID, "no CUDA device " + std::to_string(device));
    CU(cuDeviceGet(&dev, device));
    char nm[128] = {0};
    CU(cuDeviceGetName(nm, sizeof nm, dev));
    name = nm;
    int major = 0, minor = 0;
    CU(cuDeviceGetAttribute(&major, CU_DEVICE_ATTRIBUTE_COMPUTE_CAPABILITY_MAJOR, dev));
    CU(cuDeviceGetAttribute(&minor, CU_DEVICE_ATTRIBUTE_COMPUTE_CAPABILITY_MINOR, dev));
    if (major != 10 || minor != 0)
      fail(CLG_E_UNSUPPORTED, "device " + name + " is sm_" + std::to_string(major) + std::to_string(minor) +
                                  "; this backend is built for sm_100a (B200) only");
    CU(cuDeviceGetAttribute(&sm_count, CU_DEVICE_ATTRIBUTE_MULTIPROCESSOR_COUNT, dev));
    CU(cuDevicePrimaryCtxRetain(&cu, dev));
    bind();
    CU(cuModuleLoadData(&mod, clg_cubin));
    static const char *ks[3] = {"1", "2", "4"};
    for (int i = 0; i < 3; i++) {
      CU(cuModuleGetFunction(&lane[i], mod, (std::string("clg_lane_k") + ks[i]).c_str()));
      CU(cuModuleGetFunction(&coop[i], mod, (std::string("clg_coop_k") + ks[i]).c_str()));
      CU(cuFuncSetAttribute(lane[i], CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)SMEM_MAX));
      CU(cuModuleGetFunction(&lane_n[i], mod, (std::string("clg_lane_n_k") + ks[i]).c_str()));
      CU(cuFuncSetAttribute(lane_n[i], CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)SMEM_MAX));
      CU(cuFuncSetAttribute(coop[i], CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)SMEM_MAX));
      CU(cuModuleGetFunction(&coop_c[i], mod, (std::string("clg_coop_c_k") + ks[i]).c_str()));
      CU(cuFuncSetAttribute(coop_c[i], CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)SMEM_MAX));
      CU(cuModuleGetFunction(&coop_res[i], mod, (std::string("clg_coop_res_k") + ks[i]).c_str()));
      CU(cuFuncSetAttribute(coop_res[i], CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)SMEM_MAX));
    }
    CU(cuModuleGetFunction(&coop_rows, mod, "clg_coop_rows"));
    CU(cuFuncSetAttribute(coop_rows, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)SMEM_MAX));
    CU(cuModuleGetFunction(&coop_rows_t, mod, "clg_coop_rows_t"));
    CU(cuFuncSetAttribute(coop_rows_t, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)SMEM_MAX));
    CU(cuModuleGetFunction(&slice, mod, "clg_slice"));
    CU(cuModuleGetFunction(&unslice, mod, "clg_unslice"));
    CU(cuModuleGetFunction(&slice_v4, mod, "clg_slice_v4"));
    CU(cuModuleGetFunction(&unslice_v4, mod, "clg_unslice_v4"));
    CU(cuModuleGetFunction(&unslice_s64, mod, "clg_unslice_s64"));
    static const char *rt[3] = {"32", "64", "128"};
    for (int i = 0; i < 3; i++) {
      CU(cuModuleGetFunction(&slice_w[i], mod, (std::string("clg_slice_w") + rt[i]).c_str()));
      CU(cuModuleGetFunction(&unslice_w[i], mod, (std::string("clg_unslice_w") + rt[i]).c_str()));
    }
    for (int i = 0; i < 7; i++) {  // TMA-tiled transposes: VT vectors x RT bytes staged per CTA + the word tile + one mbarrier
      const std::string sfx = std::to_string(TMA_RT[i]) + "_" + std::to_string(TMA_VT[i]);
      CU(cuModuleGetFunction(&slice_tma[i], mod, ("clg_slice_tma" + sfx).c_str()));
      CU(cuModuleGetFunction(&unslice_tma[i], mod, ("clg_unslice_tma" + sfx).c_str()));
      CU(cuFuncSetAttribute(slice_tma[i], CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)tma_smem(i)));
      CU(cuFuncSetAttribute(unslice_tma[i], CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)tma_smem(i)));
    }
    CU(cuModuleGetFunction(&slice_gen, mod, "clg_slice_gen"));
    CU(cuModuleGetFunction(&unslice_gen, mod, "clg_unslice_gen"));
    CU(cuFuncSetAttribute(slice_gen, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)(GEN_MAX_BYTES + 256)));
    CU(cuFuncSetAttribute(unslice_gen, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)(GEN_MAX_BYTES + 256)));
    CU(cuModuleGetFunction(&fill, mod, "clg_fill_random"));
    CU(cuModuleGetFunction(&lop3_peak, mod, "clg_lop3_peak"));
  }
  ~Ctx() {
    Driver &d = driver();
    if (cu) {
      d.cuCtxSetCurrent(cu);
      if (mod) d.cuModuleUnload(mod);
      d.cuDevicePrimaryCtxRelease(dev);
    }
  }
  void bind() const { CU(cuCtxSetCurrent(cu)); }
};

static int kidx(int K) { return K == 4 ? 2 : K == 2 ? 1 : 0; }

// ---- resident program --------------------------------------------------------------------------------
// Device encoding of a flat instruction stream (kernels.cu): byte offsets pre-scaled by the register-stack stride.
static void encode_stream(const FlatInstr *src, uint32_t n, uint32_t slot_bytes, uint32_t *enc) {
  for (uint32_t i = 0; i < n; i++) {
    const FlatInstr &x = src[i];
    uint32_t *e = enc + (size_t)i * 4;
    switch (x.kind) {
      case K_LUT: {
        const int nf = lut_nfan(x);
        const uint32_t kind = nf >= 3 ? 0u : nf == 2 ? 1u : 2u;  // D_LUT3 / D_LUT2 / D_LUT1
        e[0] = x.a * slot_bytes | (uint32_t)x.tt << 24;
        e[1] = x.b * slot_bytes | kind << 28;
        e[2] = x.c * slot_bytes;
        e[3] = x.dst * slot_bytes;
        break;
      }
      case K_LDIN: e[0] = 0, e[1] = 3u << 28, e[2] = x.row, e[3] = x.dst * slot_bytes; break;
      case K_STOUT: e[0] = x.a * slot_bytes, e[1] = 4u << 28 | (uint32_t)(x.flags & 1) << 24, e[2] = x.row, e[3] = 0; break;
      default: e[0] = 0, e[1] = 5u << 28 | (uint32_t)(x.flags & 1) << 24, e[2] = x.row, e[3] = 0; break;
    }
  }
}

// One execution mode of a program made resident on the device.
struct Plan {
  bool ready = false;
  int K = 0;
  CUdeviceptr d_instr = 0, d_levels = 0, d_parts = 0, d_cinstr = 0, d_clevels = 0;  // (compact cooperative form)
  uint32_t n_instr = 0, n_levels = 0, n_slots = 0, coop_threads = 0, n_parts = 0, code_off = 0, ring_off = 0;
  bool nandish = false;   // lane mode: most LUTs are the NAND network's two tables (clg_lane_n_k*)
  bool rows_tma = false;  // row form: instruction stream staged through shared memory (clg_coop_rows_t)
  size_t smem = 0;
  uint32_t tuned_threads = 0;  // cooperative, big batches: CTA size found by timing candidates on this device (0: not tuned yet)
  bool resident = false;  // cooperative: the (largest part's) instruction stream is staged in shared memory
  // cooperative, row form (clg_coop_rows): the row stream of the flat buffer, encoded for the device
  CUdeviceptr d_rinstr = 0, d_rwarp = 0, d_rside = 0, d_rside_off = 0;
  uint32_t rows_nt = 0;
  // long programs: the lane stream cut into SPEC_CHUNK-instruction kernels; live registers cross the cuts through `scratch`
  struct ChunkLaunch {
    CUfunction fn;
    uint32_t first, count;  // entries of the row-base table: `count` equal, independent chunks run side by side (grid.y)
    uint32_t group;         // > 1: this and the next group - 1 launches run `count` groups of `group` chunks each, position by position
  };
  std::vector<CUmodule> chunk_mods;  // one per DISTINCT chunk
  std::vector<ChunkLaunch> chunks;   // the chain, in order
  CUdeviceptr d_bases = 0;           // uint32[chunks of the program][4]: imm / out / state row base of each chunk
  CUdeviceptr scratch = 0;
  size_t scratch_bytes = 0;
  uint32_t scratch_parts_max = UINT32_MAX;  // lowered when the device had no room for as many parts as the budget allows
  CUmodule spec_mod = nullptr, steps_mod = nullptr;
  CUfunction spec_fn = nullptr, steps_fn = nullptr;  // steps_fn: the fused multi-step form, assembled on first use
  std::string ptx;
};

// Assemble generated PTX with the driver's JIT (it serialises concurrent calls internally: 13 kernels took 12.7 s on
// one host thread and 14.2 s on sixteen, so there is no thread pool here). CLG_SPEC_MAXREG caps registers per thread.
static void assemble(const std::string &text, const char *entry, const std::string &what, CUmodule *mod, CUfunction *fn) {
  char log[8192] = {0};
  CUjit_option opt[3] = {CU_JIT_ERROR_LOG_BUFFER, CU_JIT_ERROR_LOG_BUFFER_SIZE_BYTES, CU_JIT_MAX_REGISTERS};
  void *val[3] = {log, (void *)(uintptr_t)sizeof log, nullptr};
  unsigned n_opt = 2;
  if (const char *env = std::getenv("CLG_SPEC_MAXREG")) val[2] = (void *)(uintptr_t)std::max(16, std::atoi(env)), n_opt = 3;
  if (driver().cuModuleLoadDataEx(mod, text.c_str(), n_opt, opt, val) != CUDA_SUCCESS)
    fail(CLG_E_PTX, "assembling " + what + " failed: " + log);
  CU(cuModuleGetFunction(fn, *mod, entry));
}

void launch_slice(Ctx *c, bool unslice, const void *src, void *dst, size_t n_vectors, size_t n_rows, size_t stride, void *stream);

struct Exec {
  Ctx *ctx;
  Flat flat;  // own copy: the caller may destroy its clg_flat
  int want;   // CLG_MODE_AUTO or the mode the caller pinned
  int mode;   // mode of the most recent (or, before any run, the preferred) plan
  int last_k = 0, last_parts = 0, last_launches = 0;  // (last_launches: 0 before any run)
  bool last_rows = false;  // the most recent cooperative run used the row kernel
  const char *last_kernel = "";
  // shape of the chunked specialised chain (computed on demand, for the mode choice): distinct kernels to assemble, and
  // per launch {instructions, instances side by side}
  struct ChainShape {
    int k = 0, distinct = 0;
    std::vector<std::pair<uint32_t, uint32_t>> launches;
  };
  mutable ChainShape shape_[2];  // [0] the first chain, [1] the second (wants_tp_chain)
  const ChainShape &chain_shape(int k, bool second = false) const {
    ChainShape &sh = shape_[second];
    if (sh.k != k) {
      const SpecChain ch = second ? spec_chain(flat, k, std::max(2u, tp_chunk() / (uint32_t)k), true) : spec_chain(flat, k);
      sh = ChainShape{k, (int)ch.unique.size(), {}};
      for (const SpecChain::Launch &l : ch.launches) sh.launches.push_back({(uint32_t)ch.chunks[l.first].code.size(), l.count});
    }
    return sh;
  }
  Plan plan[7];  // indexed by clg_mode; [PLAN_PARTS] = cooperative mode over the partitioned level stream, [PLAN_ROWS] = over the row
                 // stream, [PLAN_SPEC_TP] = the specialised chain in chunks below the instruction-cache cliff (big batches)
  static constexpr int PLAN_PARTS = 4, PLAN_ROWS = 5, PLAN_SPEC_TP = 6;

  const FlatStream &level_stream() const { return flat.header().stream[STREAM_LEVEL]; }
  const FlatStream &lane_stream() const { return flat.header().stream[STREAM_LANE]; }

  // K (words per thread / per CTA) a mode would run with; 0 = the program does not fit that mode
  int fit_k(int m) const {
    const FlatStream &lv = level_stream(), &ln = lane_stream();
    switch (m) {
      case CLG_MODE_SPEC:
        // Assembling straight-line code is super-linear in its length (12k instructions: 2 s, 50k: 9 s, 200k: 75 s),
        // so long programs are cut into chunk kernels of about a second each, up to 4M instructions. (When AUTO
        // takes such a chain is decided in choose(): at once if few distinct kernels have to be assembled, else after
        // the interpreters have cost as much GPU time as the assembly will.)
        if (ln.n_instr == 0 || ln.n_instr > 4000000u) return 0;
        {
          int k = 0;
          if (const char *env = std::getenv("CLG_SPEC_K")) k = std::atoi(env);  // tuning knob
          // Wide words (K = 4, 2) make every load and store a 128/64-bit access, which is what an I/O-bound program
          // wants; a long compute-bound one runs at the instruction-fetch rate and wants resident warps instead, i.e.
          // few registers per thread (a 39k-instruction DAG with 91 live values: K = 1 0.81 ms, K = 2 1.21 ms,
          // K = 4 2.27 ms; the 784-instruction mul16: K = 2 0.114 ms, K = 1 0.138 ms).
          const FlatHeader &h = flat.header();
          const uint64_t io_rows = (uint64_t)h.n_inputs + h.n_outputs + 2ull * h.n_state;
          const bool long_compute_bound = ln.n_instr > 4096 && io_rows * 4 < ln.n_instr;
          for (int c : {4, 2, 1})
            if (!k && (uint64_t)ln.n_slots * c <= 200 && (c == 1 || !long_compute_bound)) k = c;
          // one word per thread tolerates a live set somewhat beyond the register file: ptxas spills the coldest
          // values to L1-backed local memory (mul64: 315 live values -> 255 registers + 488 B of spills)
          // (the assembler's time explodes with the spill count: 315 live values 2 s, 681 live values 54 s per
          // kernel -- so AUTO stops at 400, a caller who pins the mode may go to 512)
          if (!k && ln.n_slots <= (want == CLG_MODE_SPEC ? 512u : 400u)) k = 1;
          return k;
        }
      case CLG_MODE_LANE:
        if (ln.n_instr == 0) return 0;
        if (const char *env = std::getenv("CLG_LANE_K")) return std::atoi(env);  // tuning knob
        // The kernel executes ~39 warp instructions per interpreted instruction whatever K is, and each is a long
        // dependent chain (fetch, three LDS, table dispatch, STS): with many warps it is bound by instruction issue
        // (K = 1, 31 warps: 81 % of the issue slots), with few by the chain's latency. Shared memory fixes warps x K
        // (stack bytes per thread = slots * K * 4). Prefer the widest vector that still leaves >= 8 consumer warps
        // (tools/sweep_lane_k.sh: mul16 K = 1 / 2 / 4: 0.63 / 0.49 / 0.47 ms at 2^24 vectors, a 77-slot DAG 5.8 / 4.7 / 4.7,
        // adder32 0.17 / 0.12 / 0.08; below 8 warps it turns around: mul64, 222 slots, 1.48 / 2.79).
        for (int k : {4, 2, 1})
          if (lane_resident(k) >= 256u) return k;
        {  // big stacks: whatever keeps the most words in flight per SM
          int best = 0;
          uint32_t best_words = 0;
          for (int k : {4, 2, 1})
            if (lane_resident(k) * (uint32_t)k > best_words) best_words = lane_resident(k) * (uint32_t)k, best = k;
          return best;
        }
      case CLG_MODE_COOP:
        if (lv.n_instr == 0) return 0;
        if (const char *env = std::getenv("CLG_COOP_K")) return std::atoi(env);  // tuning knob
        for (int k : {4, 2, 1})
          if ((size_t)lv.n_slots * k * 4 <= SMEM_MAX) return k;
        return 0;
    }
    return 0;
  }
  // consumer threads of one lane-kernel CTA: as many as one SM's shared memory holds stacks for (one CTA per SM
  // when the stack is big, several when it is small)
  uint32_t lane_consumers(int k) const {
    const size_t per_thread = std::max<size_t>(4, (size_t)lane_stream().n_slots * k * 4);
    const size_t fit = (SMEM_MAX - LANE_RING_BYTES) / per_thread / 32 * 32;
    if (fit < (size_t)LANE_MIN_THREADS) return 0;
    if (fit <= (size_t)LANE_MAX_THREADS) return (uint32_t)fit;
    // small stacks: several CTAs per SM; 224 consumers + 1 producer warp = 8 warps per CTA
    return 224;
  }
  uint32_t lane_resident(int k) const {  // consumer threads resident on one SM
    const uint32_t nc = lane_consumers(k);
    if (!nc) return 0;
    const size_t smem = LANE_RING_BYTES + (size_t)lane_stream().n_slots * nc * k * 4;
    return nc * (uint32_t)std::min<size_t>(2048 / (nc + 32), SMEM_MAX / smem);
  }
  static uint32_t coop_threads_for(const FlatStream &lv) {
    if (const char *env = std::getenv("CLG_COOP_THREADS")) return std::max(32, std::atoi(env)) / 32 * 32;  // tuning knob
    return std::min<uint32_t>(1024u, std::max<uint32_t>(64u, (lv.max_level_width + 31u) / 32u * 32u));
  }

  // Rough cycle models, only used to pick between the two interpreters for a given batch: the lane mode wins on
  // big batches of small-live-set programs, the cooperative mode on small batches of wide programs.
  double lane_cycles(size_t words) const {
    const int k = fit_k(CLG_MODE_LANE);
    const FlatStream &ln = lane_stream();
    const uint32_t nc = lane_consumers(k);
    const size_t smem = LANE_RING_BYTES + (size_t)ln.n_slots * nc * k * 4;
    const double warps_per_sm = (nc / 32.0) * std::min<double>(std::floor(2048.0 / (nc + 32)), std::floor((double)SMEM_MAX / (double)smem));
    const double groups = std::ceil((double)words / k), per_wave = warps_per_sm * 32.0 * ctx->sm_count;
    return std::ceil(groups / per_wave) * ln.n_instr * std::max(48.0, 17.0 * std::min(warps_per_sm, std::ceil(groups / 32.0 / ctx->sm_count)));
  }
  double coop_cycles(size_t words) const {
    const int k = fit_k(CLG_MODE_COOP);
    const FlatStream &lv = level_stream();
    const uint32_t nt = coop_threads_for(lv), *lvl = flat.levels(STREAM_LEVEL);
    double per_group = 0;
    for (uint32_t l = 0; l < lv.n_levels; l++) {
      const uint32_t n = lvl[l + 1] - lvl[l];
      per_group += 60.0 + (n / nt) * std::max(300.0, 0.75 * nt) + ((n % nt) ? std::max(300.0, 0.75 * (n % nt)) : 0.0);
    }
    return std::ceil(std::ceil((double)words / k) / ctx->sm_count) * per_group;
  }
  // the chunked specialised form: a chain of launches of dependent straight-line code. One warp alone is bound by
  // the LOP3 latency (~4 cycles an instruction), a full SM by issue (2 cycles per warp instruction per scheduler;
  // measured on mul64 x 8: 7 cycles per instruction at 8 resident warps).
  double spec_cycles(size_t words) const {
    const int k = fit_k(CLG_MODE_SPEC);
    const FlatStream &ln = lane_stream();
    const double regs = std::min(255.0, (double)ln.n_slots * k + 24.0);
    const double resident = std::min(2048.0, std::floor(65536.0 / regs / 128.0) * 128.0);
    const double groups = std::ceil((double)words / k);
    auto launch = [&](double n_instr, double instances) {
      const double threads = groups * instances, warps = std::min(resident, std::ceil(threads / ctx->sm_count / 32.0) * 32.0) / 32.0;
      return std::ceil(threads / (resident * ctx->sm_count)) * n_instr * k * std::max(4.0, 0.85 * warps) + 5000.0;
    };
    const bool second = wants_tp_chain(words);
    if (!second && ln.n_instr <= spec_chunk_limit(k)) return launch(ln.n_instr, 1);
    double cycles = 0;
    for (const auto &l : chain_shape(k, second).launches) cycles += launch(l.first, l.second);
    return cycles;
  }
  // CTA size of the specialised kernels: 128 threads, narrower when the batch would otherwise leave SMs without a CTA
  // (2^20 lanes at K = 4 are 8192 threads: 64 CTAs of 128 for 148 SMs)
  uint32_t spec_threads(uint32_t groups) const {
    const uint32_t sms = (uint32_t)ctx->sm_count;
    return groups >= 2 * sms * 128 ? 128 : groups >= 2 * sms * 64 ? 64 : 32;
  }
  mutable double interp_seconds = 0;  // estimated GPU time spent interpreting a program that could be specialised
  mutable bool tier_pending = false;
  static double tier_seconds_per_kernel() {
    if (const char *env = std::getenv("CLG_TIER_SECONDS_PER_KERNEL")) return std::atof(env);  // knob (tests: tier up at once)
    return 1.0;
  }
  int choose(size_t words) const {
    tier_pending = false;
    if (want != CLG_MODE_AUTO) return want;
    const int spec_k = fit_k(CLG_MODE_SPEC);
    const bool lane_ok = fit_k(CLG_MODE_LANE) != 0, coop_ok = fit_k(CLG_MODE_COOP) != 0;
    if (spec_k) {
      // a program that is ONE kernel always runs specialised; a chain has to beat the interpreters at this batch size
      // (one vector through 128 flattened multipliers is 128 dependent launches: the cooperative mode's case)
      if (lane_stream().n_instr <= spec_chunk_limit(spec_k)) return CLG_MODE_SPEC;
      const double sc = spec_cycles(words);
      if ((!lane_ok || sc <= lane_cycles(words)) && (!coop_ok || sc <= coop_cycles(words))) {
        // Tiered specialisation: every distinct kernel of the chain costs about a second of assembly. Up to 4 of them
        // are paid at once; beyond that the program runs in an interpreter until that has cost as much GPU time as
        // the assembly will (run() keeps the tally), and is specialised then.
        const bool second = wants_tp_chain(words);
        const int distinct = chain_shape(spec_k, second).distinct;
        if (distinct <= 4 || plan[second ? PLAN_SPEC_TP : CLG_MODE_SPEC].ready || interp_seconds >= distinct * tier_seconds_per_kernel()) return CLG_MODE_SPEC;
        tier_pending = lane_ok || coop_ok;
        if (!tier_pending) return CLG_MODE_SPEC;  // nothing else can run it
      }
    }
    if (lane_ok && coop_ok) return lane_cycles(words) <= coop_cycles(words) ? CLG_MODE_LANE : CLG_MODE_COOP;
    return lane_ok ? CLG_MODE_LANE : CLG_MODE_COOP;
  }

  Exec(Ctx *c, const Flat &f, int want_) : ctx(c), flat(f), want(want_) {
    if (want < CLG_MODE_AUTO || want > CLG_MODE_SPEC) fail(CLG_E_INVALID, "unknown execution mode");
    mode = choose((size_t)1 << 20);  // the plan a large batch would use; small batches may pick another (auto)
    last_k = (mode == CLG_MODE_SPEC && wants_tp_chain(1) ? get_spec_tp() : get(mode)).K;  // (a program that runs its second chain at every batch size never builds the first)
  }

  // K the partitioned cooperative plan would use (stack sized for the largest part); 0 = no partitioned form
  int parts_k() const {
    const FlatHeader &h = flat.header();
    if (h.n_parts < 2) return 0;
    uint32_t slots = 0;
    for (uint32_t q = 0; q < h.n_parts; q++) slots = std::max(slots, flat.parts()[q].n_slots);
    for (int k : {4, 2, 1})
      if ((size_t)slots * k * 4 <= SMEM_MAX) return k;
    return 0;
  }
  // keep the instruction stream in shared memory behind the stack when it fits: [stack | mbarrier | instructions]
  static void make_resident(Plan &P, uint32_t n_instr, uint32_t n_levels) {
    const size_t code_off = (P.smem + 127) / 128 * 128;
    const size_t need = code_off + 16 + ((size_t)n_levels + 1 + 3) / 4 * 16 + (size_t)n_instr * 16;
    if (need <= SMEM_MAX) P.resident = true, P.code_off = (uint32_t)code_off, P.smem = need;
  }
  Plan &get_parts() {
    Plan &P = plan[PLAN_PARTS];
    if (P.ready) return P;
    ctx->bind();
    const FlatHeader &h = flat.header();
    P.K = parts_k();
    if (!P.K) fail(CLG_E_UNSUPPORTED, "no partitioned level stream");
    P.n_parts = h.n_parts;
    uint32_t width = 0;
    size_t total_instr = 0, total_levels = 0;
    for (uint32_t q = 0; q < h.n_parts; q++) {
      const FlatStream &st = flat.parts()[q];
      P.n_slots = std::max(P.n_slots, st.n_slots), width = std::max(width, st.max_level_width);
      total_instr += st.n_instr, total_levels += (st.n_levels + 1 + 3) / 4 * 4;
    }
    P.smem = std::max<size_t>((size_t)P.n_slots * P.K * 4, 16);
    P.coop_threads = std::min<uint32_t>(1024u, std::max<uint32_t>(64u, (width + 31u) / 32u * 32u));
    uint32_t biggest = 0;
    for (uint32_t q = 0; q < h.n_parts; q++) biggest = std::max(biggest, flat.parts()[q].n_instr);
    uint32_t deepest = 0;
    for (uint32_t q = 0; q < h.n_parts; q++) deepest = std::max(deepest, flat.parts()[q].n_levels);
    make_resident(P, biggest, deepest);
    std::vector<uint32_t> enc(total_instr * 4), lv(total_levels), desc((size_t)h.n_parts * 4);
    size_t io = 0, lo = 0;
    for (uint32_t q = 0; q < h.n_parts; q++) {
      const FlatStream &st = flat.parts()[q];
      encode_stream(flat.part_instrs(q), st.n_instr, (uint32_t)P.K * 4, enc.data() + io * 4);
      std::memcpy(lv.data() + lo, flat.part_levels(q), ((size_t)st.n_levels + 1) * 4);
      desc[q * 4 + 0] = (uint32_t)io, desc[q * 4 + 1] = (uint32_t)lo, desc[q * 4 + 2] = st.n_levels, desc[q * 4 + 3] = 0;
      io += st.n_instr, lo += (st.n_levels + 1 + 3) / 4 * 4;
    }
    P.n_instr = (uint32_t)total_instr;
    CU(cuMemAlloc(&P.d_instr, std::max<size_t>(16, enc.size() * 4)));
    CU(cuMemcpyHtoDAsync(P.d_instr, enc.data(), enc.size() * 4, nullptr));
    CU(cuMemAlloc(&P.d_levels, lv.size() * 4));
    CU(cuMemcpyHtoDAsync(P.d_levels, lv.data(), lv.size() * 4, nullptr));
    CU(cuMemAlloc(&P.d_parts, desc.size() * 4));
    CU(cuMemcpyHtoDAsync(P.d_parts, desc.data(), desc.size() * 4, nullptr));
    CU(cuStreamSynchronize(nullptr));
    P.ready = true;
    return P;
  }

  // chunked specialisation: cut where few registers are live; equal chunks (the instances of a flattened program)
  // share one assembled kernel and differ only in the rows their pointers are advanced by
  void build_chain(Plan &P, uint32_t limit, bool fine = false) {
    const SpecChain chain = spec_chain(flat, P.K, limit, fine);
    const std::vector<SpecChunk> &chunks = chain.chunks;
    const std::vector<uint32_t> &unique = chain.unique;
    P.chunk_mods.assign(unique.size(), nullptr);
    std::vector<CUfunction> fns(unique.size(), nullptr);
    try {
      for (uint32_t u = 0; u < unique.size(); u++) {
        const std::string text = emit_spec_chunk_ptx(flat, P.K, chunks[unique[u]]);
        if (u == 0) P.ptx = text;
        assemble(text, "clg_spec_chunk", "chunk " + std::to_string(unique[u]) + " of the specialised program", &P.chunk_mods[u], &fns[u]);
      }
    } catch (...) {  // leave no half-built chain behind: a later run() would build it again from scratch
      for (CUmodule m : P.chunk_mods)
        if (m) driver().cuModuleUnload(m);
      P.chunk_mods.clear();
      throw;
    }
    // the row-base table in launch order: {imm, out, state row, part of the scratch buffer} of every member of every launch
    std::vector<uint32_t> bases;
    for (const SpecChain::Launch &l : chain.launches) {
      P.chunks.push_back({fns[chain.kernel_of[l.first]], (uint32_t)(bases.size() / 4), l.count, l.stride});
      for (uint32_t i = 0; i < l.count; i++) {
        const SpecChunk &c = chunks[l.first + (size_t)i * l.stride];
        bases.insert(bases.end(), {c.imm_base, c.out_base, c.state_base, l.stride > 1 ? i : 0u});
      }
    }
    CU(cuMemAlloc(&P.d_bases, bases.size() * 4));
    CU(cuMemcpyHtoDAsync(P.d_bases, bases.data(), bases.size() * 4, nullptr));
    CU(cuStreamSynchronize(nullptr));
  }

  // ---- the specialised chain for big batches of register-heavy programs ------------------------------------------
  // Straight-line code is executed once per warp, so what it costs to fetch depends on whether the kernel still sits in
  // the SM's instruction cache when the next wave of CTAs starts: up to 128 KB of SASS it does (89-92 % of LOP3 peak at two
  // warps per scheduler), beyond it does not (65 %; tools/microbench_ifetch.cu). Programs with few live registers run four
  // or more warps per scheduler and do not care; a program that needs the whole register file (mul64: 222 live values)
  // does. For those -- if their default chain hands nothing over -- a batch big enough that every launch fills the machine
  // several times over runs a second chain whose chunks stay below the cliff: more launches and a register hand-over
  // through the scratch buffer, 7-10 % less time (12 x mul64 x 2^24 vectors: 7.3 -> 6.8 ms; mul64 x 2^24: 0.61 -> 0.57 ms). Small batches (config 5's single vector: one launch for all instances)
  // keep the chain with the fewest launches.
  static uint32_t tp_chunk() { return spec_tp_chunk_limit(); }
  static size_t scratch_budget() {  // bytes of hand-over scratch a plan may hold (side-by-side groups share it)
    const char *e = std::getenv("CLG_SPEC_SCRATCH_MB");
    return (size_t)(e ? std::max(1L, std::atol(e)) : 8192L) << 20;
  }
  static uint32_t tp_min_regs() {  // live registers x words per thread from which a program takes the second chain
    const char *e = std::getenv("CLG_SPEC_TP_MIN_REGS");
    return e ? (uint32_t)std::atol(e) : 160u;
  }
  mutable uint32_t tp_narrowest_ = 1;         // fewest instances side by side in a launch of the second chain
  mutable int tp_state_ = 0;                  // 0: not decided, 1: the program has a second chain for big batches, 2: for every batch, -1: it has none
  bool wants_tp_chain(size_t words) const {
    if (std::getenv("CLG_SPEC_CHUNK") || std::getenv("CLG_SPEC_NO_TP")) return false;
    const int k = fit_k(CLG_MODE_SPEC);
    const FlatStream &ln = lane_stream();
    if (!k || (uint64_t)ln.n_instr * k <= 8192 || ln.n_instr * (uint64_t)k > 4000000u) return false;
    if (tp_state_ == 0) {
      // Only for programs whose default chain hands nothing over (one kernel, or the isolated instances of a flattened
      // program): there the cliff is paid in full and one extra cut per instance is cheap next to it. A chain that
      // already hands ~200 registers over at every cut gets slower with more cuts (a 19 k-instruction random DAG with
      // 225 live values: 0.86 ms in four chunks against 0.52 ms in two; tools/bench_tp_chain.py). And at most four kernels
      // to assemble, else the default chain's tiering rules would have to apply.
      // A second chain that itself hands nothing over (every cut falls between two instances) is taken whatever the
      // register count: 12 x mul32 (110 live) 2.12 -> 1.24 ms, 8 x mul40 2.20 -> 1.44 ms, 16 x mul16 0.55 -> 0.48 ms --
      // and, if it does not take more launches than the first, whatever the batch: its kernels are shorter and its
      // launches wider (16 x mul16: one thread per column group runs all sixteen instances in the first chain, eight
      // threads two each in the second). One that cuts inside the instances pays for the hand-over and is only for
      // programs that run two warps per scheduler anyway (160 registers or more), on big batches.
      bool isolated = true, tp_isolated = true;
      const SpecChain first = spec_chain(flat, k);
      for (const SpecChunk &c : first.chunks) isolated &= !c.crosses_in && !c.crosses_out;
      const SpecChain ch = spec_chain(flat, k, std::max(2u, tp_chunk() / (uint32_t)k), true);
      for (const SpecChunk &c : ch.chunks) tp_isolated &= !c.crosses_in && !c.crosses_out;
      tp_narrowest_ = UINT32_MAX;
      for (const SpecChain::Launch &l : ch.launches) tp_narrowest_ = std::min(tp_narrowest_, l.count);
      if (!isolated || ch.unique.size() > 4 || ch.chunks.size() <= 1) tp_state_ = -1;
      else if (tp_isolated) tp_state_ = ch.launches.size() <= first.launches.size() && !std::getenv("CLG_SPEC_TP_BIG_ONLY") ? 2 : 1;
      else tp_state_ = ln.n_slots * (uint32_t)k >= tp_min_regs() ? 1 : -1;
    }
    if (tp_state_ == 2) return true;
    // every launch (column groups x the instances it runs side by side): eight waves of 128-thread CTAs or more
    return tp_state_ == 1 && (words + k - 1) / k * tp_narrowest_ >= (size_t)ctx->sm_count * 128 * 8;
  }
  Plan &get_spec_tp() {
    Plan &P = plan[PLAN_SPEC_TP];
    if (P.ready) return P;
    ctx->bind();
    P.K = fit_k(CLG_MODE_SPEC);
    build_chain(P, std::max(2u, tp_chunk() / (uint32_t)P.K), true);
    P.ready = true;
    return P;
  }

  Plan &get(int m) {
    Plan &P = plan[m];
    if (P.ready) return P;
    ctx->bind();
    const FlatStream &lv = level_stream(), &ln = lane_stream();
    const FlatStream *st = nullptr;
    int sidx = STREAM_LANE;
    P.K = fit_k(m);
    switch (m) {
      case CLG_MODE_SPEC:
        if (!P.K) fail(CLG_E_UNSUPPORTED, "program does not fit the specialised mode (no lane stream, too long, or too many live registers)");
        break;
      case CLG_MODE_LANE:
        if (!P.K) fail(CLG_E_UNSUPPORTED, "live register set does not fit a per-thread shared-memory stack; use the cooperative mode");
        st = &ln;
        P.coop_threads = lane_consumers(P.K);  // (consumer threads of the lane kernel)
        P.smem = LANE_RING_BYTES + (size_t)ln.n_slots * P.coop_threads * P.K * 4;
        {
          const FlatInstr *src = flat.instrs(STREAM_LANE);
          size_t luts = 0, hot = 0;
          for (uint32_t i = 0; i < ln.n_instr; i++)
            if (src[i].kind == K_LUT) luts++, hot += src[i].tt == 0x0E || src[i].tt == 0x03;
          P.nandish = !std::getenv("CLG_LANE_PLAIN") && hot * 2 > luts;  // (knob: A/B runs)
        }
        break;
      case CLG_MODE_COOP:
        if (!P.K) fail(CLG_E_UNSUPPORTED, "live register set exceeds one SM's shared memory even at one word per register");
        st = &lv, sidx = STREAM_LEVEL;
        P.smem = std::max<size_t>((size_t)lv.n_slots * P.K * 4, 16);
        P.coop_threads = coop_threads_for(lv);
        if (!std::getenv("CLG_COOP_NO_RESIDENT")) make_resident(P, lv.n_instr, lv.n_levels);  // (knob: tests drive the streaming kernels with small programs)
        break;
      default: fail(CLG_E_INVALID, "unknown execution mode");
    }
    if (st) {
      P.n_instr = st->n_instr, P.n_levels = st->n_levels, P.n_slots = st->n_slots;
      // device encoding (kernels.cu): byte offsets pre-scaled for this mode's register-stack layout
      const uint32_t slot_bytes = (m == CLG_MODE_LANE ? P.coop_threads : 1) * P.K * 4;
      if ((uint64_t)P.n_slots * slot_bytes >= (1u << 24)) fail(CLG_E_UNSUPPORTED, "register stack exceeds the 24-bit offset field");
      std::vector<uint32_t> enc((size_t)P.n_instr * 4);
      encode_stream(flat.instrs(sidx), P.n_instr, slot_bytes, enc.data());
      if (m == CLG_MODE_COOP && !P.resident && P.n_slots <= 16383 && !std::getenv("CLG_COOP_WIDE")) {
        // compact form: LUTs as 8 bytes (slot indices), everything else stays 16 bytes in a per-level side list
        const FlatInstr *src = flat.instrs(sidx);
        const uint32_t *lvl = flat.levels(sidx);
        std::vector<uint32_t> cenc, menc, clv(2 * ((size_t)P.n_levels + 1));
        cenc.reserve((size_t)P.n_instr * 2);
        for (uint32_t l = 0; l < P.n_levels; l++) {
          clv[l] = (uint32_t)(cenc.size() / 2), clv[P.n_levels + 1 + l] = (uint32_t)(menc.size() / 4);
          for (uint32_t i = lvl[l]; i < lvl[l + 1]; i++) {
            const FlatInstr &x = src[i];
            if (x.kind == K_LUT) {
              const int nf = lut_nfan(x);
              const uint32_t b = nf >= 2 ? x.b : 0x3FFFu, c = nf >= 3 ? x.c : 0x3FFFu;
              cenc.push_back(x.a | b << 14 | (uint32_t)(x.tt & 15) << 28);
              cenc.push_back(c | (uint32_t)x.dst << 14 | (uint32_t)(x.tt >> 4) << 28);
            } else {
              menc.insert(menc.end(), &enc[(size_t)i * 4], &enc[(size_t)i * 4 + 4]);
            }
          }
        }
        clv[P.n_levels] = (uint32_t)(cenc.size() / 2), clv[2 * (size_t)P.n_levels + 1] = (uint32_t)(menc.size() / 4);
        CU(cuMemAlloc(&P.d_cinstr, std::max<size_t>(16, cenc.size() * 4)));
        CU(cuMemcpyHtoDAsync(P.d_cinstr, cenc.data(), cenc.size() * 4, nullptr));
        CU(cuMemAlloc(&P.d_clevels, clv.size() * 4));
        CU(cuMemcpyHtoDAsync(P.d_clevels, clv.data(), clv.size() * 4, nullptr));
        CU(cuStreamSynchronize(nullptr));
        enc.swap(menc);  // `instr` now holds only the side list
        enc.resize(std::max<size_t>(enc.size(), 4));
      }
      CU(cuMemAlloc(&P.d_instr, std::max<size_t>(16, enc.size() * 4)));
      CU(cuMemcpyHtoDAsync(P.d_instr, enc.data(), enc.size() * 4, nullptr));
      CU(cuMemAlloc(&P.d_levels, ((size_t)P.n_levels + 1 + 3) / 4 * 16));  // padded: the resident kernel bulk-copies it
      CU(cuMemcpyHtoDAsync(P.d_levels, flat.levels(sidx), ((size_t)P.n_levels + 1) * 4, nullptr));
      CU(cuStreamSynchronize(nullptr));
    } else if (ln.n_instr > spec_chunk_limit(P.K)) {
      build_chain(P, 0);
    } else {
      P.ptx = emit_spec_ptx(flat, P.K);
      assemble(P.ptx, "clg_spec", "the specialised kernel", &P.spec_mod, &P.spec_fn);
    }
    P.ready = true;
    return P;
  }
  ~Exec() {
    Driver &d = driver();
    d.cuCtxSetCurrent(ctx->cu);
    for (Plan &P : plan) {
      if (P.d_instr) d.cuMemFree(P.d_instr);
      if (P.d_levels) d.cuMemFree(P.d_levels);
      if (P.d_parts) d.cuMemFree(P.d_parts);
      if (P.d_cinstr) d.cuMemFree(P.d_cinstr);
      if (P.d_clevels) d.cuMemFree(P.d_clevels);
      if (P.d_rinstr) d.cuMemFree(P.d_rinstr);
      if (P.d_rwarp) d.cuMemFree(P.d_rwarp);
      if (P.d_rside_off) d.cuMemFree(P.d_rside_off);
      if (P.d_rside) d.cuMemFree(P.d_rside);
      if (P.spec_mod) d.cuModuleUnload(P.spec_mod);
      if (P.steps_mod) d.cuModuleUnload(P.steps_mod);
      for (CUmodule m : P.chunk_mods)
        if (m) d.cuModuleUnload(m);
      if (P.scratch) d.cuMemFree(P.scratch);
      if (P.d_bases) d.cuMemFree(P.d_bases);
    }
    for (int b = 0; b < NBUF; b++)
      for (CUdeviceptr q : bstage[b])
        if (q) d.cuMemFree(q);
    for (int b = 0; b < NBUF; b++) {
      if (stage_in[b]) d.cuMemFree(stage_in[b]);
      if (stage_out[b]) d.cuMemFree(stage_out[b]);
      if (stage_state[b]) d.cuMemFree(stage_state[b]);
      if (ev_in[b]) d.cuEventDestroy(ev_in[b]);
      if (ev_run[b]) d.cuEventDestroy(ev_run[b]);
      if (ev_out[b]) d.cuEventDestroy(ev_out[b]);
    }
    if (s_in) d.cuStreamDestroy(s_in);
    if (s_run) d.cuStreamDestroy(s_run);
    if (s_out) d.cuStreamDestroy(s_out);
  }

  // ---- host-buffer path: column chunks pipelined over three streams (H2D | kernel | D2H) -----------------
  static constexpr int NBUF = 3;
  CUstream s_in = nullptr, s_run = nullptr, s_out = nullptr;
  CUevent ev_in[NBUF] = {}, ev_run[NBUF] = {}, ev_out[NBUF] = {};
  CUdeviceptr stage_in[NBUF] = {}, stage_out[NBUF] = {}, stage_state[NBUF] = {};
  size_t stage_words = 0;
  CUdeviceptr bstage[NBUF][4] = {};  // run_bools_host: per staging set {bools in, sliced in, sliced out, bools out}
  size_t bstage_vectors = 0;

  void copy2d(bool to_device, CUdeviceptr dev, size_t dev_pitch, const void *host, size_t host_pitch, size_t width, size_t rows,
              CUstream st) {
    if (!width || !rows) return;
    CUDA_MEMCPY2D c{};
    if (to_device) {
      c.srcMemoryType = CU_MEMORYTYPE_HOST, c.srcHost = host, c.srcPitch = host_pitch;
      c.dstMemoryType = CU_MEMORYTYPE_DEVICE, c.dstDevice = dev, c.dstPitch = dev_pitch;
    } else {
      c.srcMemoryType = CU_MEMORYTYPE_DEVICE, c.srcDevice = dev, c.srcPitch = dev_pitch;
      c.dstMemoryType = CU_MEMORYTYPE_HOST, c.dstHost = const_cast<void *>(host), c.dstPitch = host_pitch;
    }
    c.WidthInBytes = width, c.Height = rows;
    CU(cuMemcpy2DAsync(&c, st));
  }

  void ensure_streams() {
    if (s_in) return;
    CU(cuStreamCreate(&s_in, CU_STREAM_NON_BLOCKING));
    CU(cuStreamCreate(&s_run, CU_STREAM_NON_BLOCKING));
    CU(cuStreamCreate(&s_out, CU_STREAM_NON_BLOCKING));
    for (int b = 0; b < NBUF; b++) {
      CU(cuEventCreate(&ev_in[b], CU_EVENT_DISABLE_TIMING));
      CU(cuEventCreate(&ev_run[b], CU_EVENT_DISABLE_TIMING));
      CU(cuEventCreate(&ev_out[b], CU_EVENT_DISABLE_TIMING));
    }
  }

  // `&[bool]`-shaped batch on HOST buffers: imm is [n_vectors][n_inputs] bytes (one run()'s immediates per vector), out is
  // [n_vectors][n_outputs] bytes. Chunks of vectors go through H2D | slice, kernel, unslice | D2H on three streams with
  // staging buffers that live with the executor, so PCIe (one byte per bool each way) is what a big batch waits for.
  void run_bools_host(const uint8_t *imm, uint8_t *out, size_t n_vectors) {
    const FlatHeader &h = flat.header();
    if (h.n_state) fail(CLG_E_UNSUPPORTED, "run_bools_host is for stateless programs; use clg_sim_runner for carried state");
    if (n_vectors == 0) return;
    if (h.n_inputs && !imm) fail(CLG_E_INVALID, "imm is null");
    if (h.n_outputs && !out) fail(CLG_E_INVALID, "out is null");
    ctx->bind();
    ensure_streams();
    const size_t per_vector = std::max<size_t>(1, (size_t)h.n_inputs + h.n_outputs);
    size_t cv = ((size_t)64 << 20) / per_vector;  // ~64 MB of host bytes per chunk
    if (const char *env = std::getenv("CLG_BOOLS_CHUNK_VECTORS")) cv = std::strtoull(env, nullptr, 10);  // tuning / tests
    cv = std::max<size_t>(1024, cv / 1024 * 1024);
    cv = std::min(cv, (n_vectors + 1023) / 1024 * 1024);
    if (cv > bstage_vectors) {
      CU(cuCtxSynchronize());
      const size_t sizes[4] = {cv * h.n_inputs, (size_t)h.n_inputs * (cv / 32) * 4, (size_t)h.n_outputs * (cv / 32) * 4, cv * h.n_outputs};
      for (int b = 0; b < NBUF; b++)
        for (int k = 0; k < 4; k++) {
          if (bstage[b][k]) driver().cuMemFree(bstage[b][k]);
          bstage[b][k] = 0;
          CU(cuMemAlloc(&bstage[b][k], std::max<size_t>(16, sizes[k])));
          if (k == 1) CU(cuMemsetD8Async(bstage[b][k], 0, std::max<size_t>(16, sizes[k]), nullptr));  // (the words behind a short last chunk)
        }
      CU(cuCtxSynchronize());
      bstage_vectors = cv;
    }
    const size_t cstride = bstage_vectors / 32, n_chunks = (n_vectors + cv - 1) / cv;
    for (size_t c = 0; c < n_chunks; c++) {
      const int b = (int)(c % NBUF);
      const size_t v0 = c * cv, nv = std::min(n_vectors, v0 + cv) - v0;
      if (c >= NBUF) CU(cuStreamWaitEvent(s_in, ev_run[b], 0));  // the bools of chunk c - NBUF have been sliced
      if (h.n_inputs) CU(cuMemcpyHtoDAsync(bstage[b][0], imm + v0 * h.n_inputs, nv * h.n_inputs, s_in));
      CU(cuEventRecord(ev_in[b], s_in));
      CU(cuStreamWaitEvent(s_run, ev_in[b], 0));
      if (c >= NBUF) CU(cuStreamWaitEvent(s_run, ev_out[b], 0));  // ... and its results have been copied out
      if (h.n_inputs) launch_slice(ctx, false, (const void *)bstage[b][0], (void *)bstage[b][1], nv, h.n_inputs, cstride, s_run);
      run((const uint32_t *)bstage[b][1], (uint32_t *)bstage[b][2], nullptr, nv, cstride, s_run);
      if (h.n_outputs) launch_slice(ctx, true, (const void *)bstage[b][2], (void *)bstage[b][3], nv, h.n_outputs, cstride, s_run);
      CU(cuEventRecord(ev_run[b], s_run));
      CU(cuStreamWaitEvent(s_out, ev_run[b], 0));
      if (h.n_outputs) CU(cuMemcpyDtoHAsync(out + v0 * h.n_outputs, bstage[b][3], nv * h.n_outputs, s_out));
      CU(cuEventRecord(ev_out[b], s_out));
    }
    CU(cuStreamSynchronize(s_in));
    CU(cuStreamSynchronize(s_run));
    CU(cuStreamSynchronize(s_out));
  }

  void run_host(const uint32_t *imm, uint32_t *out, uint32_t *state, size_t n_vectors, size_t stride) {
    const FlatHeader &h = flat.header();
    if (stride % 4 != 0) fail(CLG_E_INVALID, "stride must be a multiple of 4 words (16-byte aligned rows)");
    const size_t words = (n_vectors + 31) / 32;
    if (words > stride) fail(CLG_E_INVALID, "n_vectors does not fit in stride words");
    if (h.n_inputs && !imm) fail(CLG_E_INVALID, "imm is null");
    if (h.n_outputs && !out) fail(CLG_E_INVALID, "out is null");
    if (h.n_state && !state) fail(CLG_E_INVALID, "state is null for a stateful program");
    if (n_vectors == 0) return;
    ctx->bind();
    // chunk = a contiguous range of words of every row, ~128 MB of I/O, so copies and kernels overlap
    const size_t rows = (size_t)h.n_inputs + h.n_outputs + 2 * (size_t)h.n_state;
    size_t cw = ((size_t)128 << 20) / (std::max<size_t>(rows, 1) * 4);
    if (const char *env = std::getenv("CLG_HOST_CHUNK_WORDS")) cw = std::strtoull(env, nullptr, 10);  // tuning / tests
    cw = std::max<size_t>(4, cw) / 4 * 4;
    // whole rounds of the persistent cooperative grid (one 4-word column group per CTA and round): a chunk of 16.6 groups
    // per SM costs 17 rounds (1M-gate DAG x 2^24 vectors, same box: 193 ms per call with 9 828-word chunks, 174-182 ms with
    // any whole number of rounds between 2 and 32)
    const size_t round_words = 4 * (size_t)ctx->sm_count;
    if (cw >= 4 * round_words && !std::getenv("CLG_HOST_CHUNK_WORDS")) cw = cw / round_words * round_words;
    const size_t padded = (words + 3) / 4 * 4;
    cw = std::min(cw, padded);
    ensure_streams();
    if (cw > stage_words) {
      CU(cuCtxSynchronize());
      for (int b = 0; b < NBUF; b++) {
        if (stage_in[b]) driver().cuMemFree(stage_in[b]);
        if (stage_out[b]) driver().cuMemFree(stage_out[b]);
        if (stage_state[b]) driver().cuMemFree(stage_state[b]);
        CU(cuMemAlloc(&stage_in[b], std::max<size_t>(16, (size_t)h.n_inputs * cw * 4)));
        CU(cuMemAlloc(&stage_out[b], std::max<size_t>(16, (size_t)h.n_outputs * cw * 4)));
        CU(cuMemAlloc(&stage_state[b], std::max<size_t>(16, (size_t)h.n_state * cw * 4)));
      }
      stage_words = cw;
    }
    const size_t n_chunks = (words + cw - 1) / cw;
    for (size_t c = 0; c < n_chunks; c++) {
      const int b = (int)(c % NBUF);
      const size_t w0 = c * cw, w1 = std::min(words, w0 + cw), nw = w1 - w0;
      const size_t nv = std::min(n_vectors, w1 * 32) - w0 * 32;
      if (c >= NBUF) {  // staging set b is free again once chunk c - NBUF has been run (inputs) and copied out (outputs)
        CU(cuStreamWaitEvent(s_in, ev_run[b], 0));
        CU(cuStreamWaitEvent(s_in, ev_out[b], 0));
      }
      copy2d(true, stage_in[b], cw * 4, imm + w0, stride * 4, nw * 4, h.n_inputs, s_in);
      copy2d(true, stage_state[b], cw * 4, state + w0, stride * 4, nw * 4, h.n_state, s_in);
      CU(cuEventRecord(ev_in[b], s_in));
      CU(cuStreamWaitEvent(s_run, ev_in[b], 0));
      if (c >= NBUF) CU(cuStreamWaitEvent(s_run, ev_out[b], 0));
      run((const uint32_t *)stage_in[b], (uint32_t *)stage_out[b], h.n_state ? (uint32_t *)stage_state[b] : nullptr, nv, cw, s_run);
      CU(cuEventRecord(ev_run[b], s_run));
      CU(cuStreamWaitEvent(s_out, ev_run[b], 0));
      copy2d(false, stage_out[b], cw * 4, out + w0, stride * 4, nw * 4, h.n_outputs, s_out);
      copy2d(false, stage_state[b], cw * 4, state + w0, stride * 4, nw * 4, h.n_state, s_out);
      CU(cuEventRecord(ev_out[b], s_out));
    }
    CU(cuStreamSynchronize(s_in));
    CU(cuStreamSynchronize(s_run));
    CU(cuStreamSynchronize(s_out));
  }

  // n_steps consecutive run() calls per lane: imm is [n_steps][n_inputs][stride], out is [n_steps][n_outputs][stride],
  // state is carried from step to step. The specialised mode does it in ONE launch with the carried registers held in
  // registers across the steps; the interpreters issue one launch per step.
  void run_steps(const uint32_t *imm, uint32_t *out, uint32_t *state, size_t n_vectors, size_t stride, size_t n_steps, CUstream stream) {
    const FlatHeader &h = flat.header();
    if (n_steps == 0 || n_vectors == 0) return;
    const size_t words = (n_vectors + 31) / 32;
    const size_t imm_step = (size_t)h.n_inputs * stride, out_step = (size_t)h.n_outputs * stride;
    // the fused form is ONE kernel of n_instr * K instructions: a program that needs a chunk chain runs step by step
    // (building the fused kernel would first assemble the whole unused chain and then minutes of straight-line code)
    // (... and so does one whose second chain runs this batch: an instance per thread beats one thread for all instances)
    if (choose(words) != CLG_MODE_SPEC || n_steps == 1 || lane_stream().n_instr > spec_chunk_limit(fit_k(CLG_MODE_SPEC)) || wants_tp_chain(words)) {
      for (size_t s_ = 0; s_ < n_steps; s_++) run(imm ? imm + s_ * imm_step : nullptr, out ? out + s_ * out_step : nullptr, state, n_vectors, stride, stream);
      return;
    }
    if (stride % 4 != 0) fail(CLG_E_INVALID, "stride must be a multiple of 4 words (16-byte aligned rows)");
    if (words > stride) fail(CLG_E_INVALID, "n_vectors does not fit in stride words");
    if (h.n_inputs && !imm) fail(CLG_E_INVALID, "imm is null");
    if (h.n_outputs && !out) fail(CLG_E_INVALID, "out is null");
    if (h.n_state && !state) fail(CLG_E_INVALID, "state is null for a stateful program");
    if (((uintptr_t)imm | (uintptr_t)out | (uintptr_t)state) & 15) fail(CLG_E_INVALID, "buffers must be 16-byte aligned");
    if (n_steps > 0xFFFFFFFFull) fail(CLG_E_INVALID, "too many steps");
    ctx->bind();
    mode = CLG_MODE_SPEC;
    Plan &P = get(CLG_MODE_SPEC);
    if (!P.steps_fn) {
      const std::string text = emit_spec_ptx(flat, P.K, true);
      assemble(text, "clg_spec_steps", "the multi-step specialised kernel", &P.steps_mod, &P.steps_fn);
    }
    last_k = P.K, last_parts = 0;
    uint32_t stride32 = (uint32_t)stride, g = (uint32_t)((words + P.K - 1) / P.K), steps32 = (uint32_t)n_steps;
    uint64_t imm_b = imm_step * 4, out_b = out_step * 4;
    void *args[] = {&imm, &out, &state, &stride32, &g, &steps32, &imm_b, &out_b};
    const uint32_t nt = spec_threads(g);
    last_kernel = "clg_spec_steps";
    CU(cuLaunchKernel(P.steps_fn, (g + nt - 1) / nt, 1, 1, nt, 1, 1, 0, stream, args, nullptr));
  }

  void run(const uint32_t *imm, uint32_t *out, uint32_t *state, size_t n_vectors, size_t stride, CUstream stream) {
    const FlatHeader &h = flat.header();
    if (stride % 4 != 0) fail(CLG_E_INVALID, "stride must be a multiple of 4 words (16-byte aligned rows)");
    const size_t words = (n_vectors + 31) / 32;
    if (words > stride) fail(CLG_E_INVALID, "n_vectors does not fit in stride words");
    if (stride > 0xFFFFFFFFull) fail(CLG_E_INVALID, "stride too large");
    if (h.n_inputs && !imm) fail(CLG_E_INVALID, "imm is null");
    if (h.n_outputs && !out) fail(CLG_E_INVALID, "out is null");
    if (h.n_state && !state) fail(CLG_E_INVALID, "state is null for a stateful program");
    if (((uintptr_t)imm | (uintptr_t)out | (uintptr_t)state) & 15) fail(CLG_E_INVALID, "buffers must be 16-byte aligned");
    if (n_vectors == 0) return;
    ctx->bind();
    mode = choose(words);
    if (tier_pending) interp_seconds += (mode == CLG_MODE_LANE ? lane_cycles(words) : coop_cycles(words)) / 1.9e9;
    // a cooperative run that cannot give every SM its own column group uses the partitioned form if there is one
    const bool use_parts = mode == CLG_MODE_COOP && parts_k() && (words + parts_k() - 1) / parts_k() < (size_t)ctx->sm_count;
    // a batch that gives every SM a column group runs the row stream when the pass built one
    if (mode == CLG_MODE_COOP && !use_parts && has_rows() && (words + 3) / 4 >= (size_t)ctx->sm_count) {
      Plan &R = get_rows();
      last_k = 4, last_parts = 0, last_rows = true, last_launches = 1, last_kernel = R.rows_tma ? "clg_coop_rows_t" : "clg_coop_rows";
      launch_rows(R, params(R, imm, out, state, stride, (uint32_t)((words + 3) / 4)), stream);
      return;
    }
    last_rows = false;
    Plan &P = use_parts ? get_parts() : mode == CLG_MODE_SPEC && wants_tp_chain(words) ? get_spec_tp() : get(mode);
    last_launches = mode == CLG_MODE_SPEC && !P.chunks.empty() ? (int)P.chunks.size() : 1;
    const int K = P.K;
    last_k = K, last_parts = (int)P.n_parts;
    const uint32_t groups = (uint32_t)((words + K - 1) / K);
    if (mode == CLG_MODE_SPEC) {
      uint32_t stride32 = (uint32_t)stride, g = groups;
      const uint32_t nt = spec_threads(groups);
      if (!P.chunks.empty()) {
        // registers that are live across chunk cuts: [part][column group / 32][slot][32 lanes][K words] (spec.cpp). Groups
        // of chunks that run side by side hand over in a part each: as many at a time as the scratch budget holds.
        const size_t part_bytes = (size_t)lane_stream().n_slots * (((size_t)groups + 31) / 32 * 32) * K * 4;
        uint32_t widest = 1;
        for (const Plan::ChunkLaunch &c : P.chunks)
          if (c.group > 1) widest = std::max(widest, c.count);
        uint32_t at_once = (uint32_t)std::min<size_t>(std::min(widest, P.scratch_parts_max), std::max<size_t>(1, scratch_budget() / std::max<size_t>(part_bytes, 1)));
        if (part_bytes * at_once > P.scratch_bytes) {
          CU(cuCtxSynchronize());
          if (P.scratch) driver().cuMemFree(P.scratch);
          P.scratch = 0, P.scratch_bytes = 0;
          // a device too full for that many parts runs fewer groups at a time (one part is what the chain needs at least)
          while (at_once > 1 && driver().cuMemAlloc(&P.scratch, part_bytes * at_once) != CUDA_SUCCESS) P.scratch = 0, at_once /= 2, P.scratch_parts_max = at_once;
          if (!P.scratch) CU(cuMemAlloc(&P.scratch, std::max<size_t>(16, part_bytes * at_once)));
          P.scratch_bytes = part_bytes * at_once;
        }
        last_kernel = "clg_spec_chunk";
        last_launches = 0;
        auto launch = [&](const Plan::ChunkLaunch &c, uint32_t first, uint32_t count) {
          const uint32_t ntc = spec_threads((uint32_t)std::min<uint64_t>((uint64_t)groups * count, 0xFFFFFFFFull));
          CUdeviceptr table = P.d_bases + ((size_t)c.first + first) * 16;
          CUdeviceptr scratch = P.scratch - (size_t)first * part_bytes;  // (the kernel adds the member's number x part_bytes)
          void *cargs[] = {&imm, &out, &state, &stride32, &g, &scratch, &table};
          CU(cuLaunchKernel(c.fn, (groups + ntc - 1) / ntc, count, 1, ntc, 1, 1, 0, stream, cargs, nullptr));
          last_launches++;
        };
        for (size_t i = 0; i < P.chunks.size();) {
          const Plan::ChunkLaunch &c = P.chunks[i];
          if (c.group <= 1) {
            launch(c, 0, c.count);
            i++;
            continue;
          }
          for (uint32_t first = 0; first < c.count; first += at_once)
            for (uint32_t p = 0; p < c.group; p++) launch(P.chunks[i + p], first, std::min(at_once, c.count - first));
          i += c.group;
        }
        return;
      }
      void *args[] = {&imm, &out, &state, &stride32, &g};
      last_kernel = "clg_spec";
      CU(cuLaunchKernel(P.spec_fn, (groups + nt - 1) / nt, 1, 1, nt, 1, 1, 0, stream, args, nullptr));
      return;
    }
    if (mode == CLG_MODE_LANE) {
      RunParams p = params(P, imm, out, state, stride, groups);
      void *args[] = {&p};
      const uint32_t nc = P.coop_threads, grid = (groups + nc - 1) / nc;
      static const char *lane_names[2][3] = {{"clg_lane_k1", "clg_lane_k2", "clg_lane_k4"}, {"clg_lane_n_k1", "clg_lane_n_k2", "clg_lane_n_k4"}};
      last_kernel = lane_names[P.nandish][kidx(K)];
      CU(cuLaunchKernel((P.nandish ? ctx->lane_n : ctx->lane)[kidx(K)], grid, 1, 1, nc + 32, 1, 1, (unsigned)P.smem, stream, args, nullptr));
      return;
    }
    // Cooperative CTA size. One wave of column groups or less is a latency problem: as many threads as the widest
    // level has instructions. More than that is a throughput problem whose optimum depends on how many CTAs fit an
    // SM and how wide the levels are (a 65k-gate DAG with ~570 instructions a level: 15.0 ms with 1024 threads,
    // 7.1 ms with 256), so it is measured once per plan on this device.
    uint32_t nt = P.coop_threads;
    if (!P.n_parts && groups > (uint32_t)ctx->sm_count * coop_ctas_per_sm(P, nt)) {
      if (!P.tuned_threads) P.tuned_threads = tune_coop_threads(P);
      nt = P.tuned_threads;
    }
    launch_coop(P, params(P, imm, out, state, stride, groups), nt, stream);
  }

  // what run() does before its launch for a batch of this size: mode choice, plan, CTA-size timing (clg_exec_prepare)
  void prepare(size_t n_vectors) {
    if (n_vectors == 0) return;
    const size_t words = (n_vectors + 31) / 32;
    ctx->bind();
    mode = choose(words);
    const bool use_parts = mode == CLG_MODE_COOP && parts_k() && (words + parts_k() - 1) / parts_k() < (size_t)ctx->sm_count;
    if (mode == CLG_MODE_COOP && !use_parts && has_rows() && (words + 3) / 4 >= (size_t)ctx->sm_count) {
      get_rows();
    } else {
      Plan &P = use_parts ? get_parts() : mode == CLG_MODE_SPEC && wants_tp_chain(words) ? get_spec_tp() : get(mode);
      const uint32_t groups = (uint32_t)std::min<size_t>((words + P.K - 1) / P.K, 0xFFFFFFFFu);
      if (mode == CLG_MODE_COOP && !P.n_parts && groups > (uint32_t)ctx->sm_count * coop_ctas_per_sm(P, P.coop_threads) && !P.tuned_threads)
        P.tuned_threads = tune_coop_threads(P);
    }
    CU(cuCtxSynchronize());
  }

  // ---- row form of the cooperative mode (clg_coop_rows; device layout in kernels.cu, RunParams) -----------------
  bool has_rows() const { return flat.rows().n_instr != 0 && !std::getenv("CLG_COOP_NO_ROWS"); }
  Plan &get_rows() {
    Plan &P = plan[PLAN_ROWS];
    if (P.ready) return P;
    ctx->bind();
    const FlatRows &r = flat.rows();
    const uint32_t nw = r.n_warps, S0 = r.n_segs, S = std::max<uint32_t>(4, (S0 + 3) / 4 * 4);  // padded with empty segments (kernels.cu, rows_step)
    if (r.lag != 2) fail(CLG_E_UNSUPPORTED, "row stream: the kernel synchronises two segments back (lag 2)");
    const FlatInstr *src = flat.row_instrs();
    const uint32_t *rt = flat.row_tab(), *stab = flat.row_sidetab();
    P.K = 4, P.n_slots = r.n_slots, P.rows_nt = nw * 32;
    P.code_off = (uint32_t)(((size_t)r.n_slots * 16 + 15) / 16 * 16);
    P.smem = P.code_off + 16;  // register stack + two mbarriers
    if (P.smem > SMEM_MAX) fail(CLG_E_UNSUPPORTED, "row stream: register stack exceeds one SM's shared memory");
    // clg_coop_rows_t: + per warp a ring of two stages of four rows (2 KB) and two mbarriers, if there is room
    P.ring_off = (uint32_t)((P.smem + 127) / 128 * 128);
    const size_t with_ring = (size_t)P.ring_off + (size_t)nw * (2 * 4 * 256 + 2 * 8);
    // (measured slower: dag1m x 2^22 52.4 ms against 42.4 ms with the ld.global fetch -- 32 warps x one 1 KB bulk copy every
    // four rows is more requests than the copy engine turns around, and bigger stages do not fit next to the stack. Kept
    // as an option, CLG_ROWS_TMA=1, and in the parity tests.)
    P.rows_tma = with_ring <= SMEM_MAX && std::getenv("CLG_ROWS_TMA");
    if (P.rows_tma) P.smem = with_ring;
    std::vector<uint32_t> words, wbase(2 * (size_t)nw);  // 2 words per lane instruction; wbase: nw offsets, then nw row counts
    words.reserve(((size_t)r.n_instr + (size_t)(S + 3) * nw * 32) * 2);
    constexpr uint32_t EMPTY_X = 0x40000000u, EMPTY_Y = 0x80000000u;  // no load of a, no store
    for (uint32_t w = 0; w < nw; w++) {
      wbase[w] = (uint32_t)(words.size() / 2);
      const size_t first_word = words.size();
      for (uint32_t sgm = 0; sgm < S; sgm++) {
        const uint32_t b = sgm < S0 ? rt[(size_t)sgm * nw + w] : 0, e = sgm < S0 ? rt[(size_t)sgm * nw + w + 1] : 0;
        const uint32_t n_rows = std::max<uint32_t>(1, (e - b) / 32);  // at least one (empty) row: it carries the end-of-segment mark
        for (uint32_t q = 0; q < n_rows; q++) {
          uint32_t hdr = 0;
          uint32_t x[32], y[32];
          for (uint32_t i = 0; i < 32; i++) x[i] = EMPTY_X, y[i] = EMPTY_Y;
          if (b + q * 32 < e) {
            const FlatInstr *row = src + b + q * 32;
            uint8_t tts[3] = {0, 0, 0};
            uint32_t n_tt = 0;
            bool reads_c = false;
            for (uint32_t i = 0; i < 32; i++) {
              const FlatInstr &in = row[i];
              if (in.kind != K_LUT) continue;
              uint32_t k = 0;
              while (k < n_tt && tts[k] != in.tt) k++;
              if (k == n_tt) {
                if (n_tt == 3) fail(CLG_E_FORMAT, "row stream: more than three truth tables in one row");
                tts[n_tt++] = in.tt;
              }
              const int nf = lut_nfan(in);
              const bool from_reg = in.flags & FLAG_FROM_REG, no_store = in.flags & FLAG_NO_STORE;
              const uint32_t a = from_reg ? 0u : in.a, bb = nf >= 2 ? in.b : a, c = nf >= 3 ? in.c : bb;  // operands it does not read alias one it does
              reads_c |= nf >= 3;
              x[i] = a | bb << 14 | k << 28 | (from_reg ? 1u << 30 : 0u);
              y[i] = c | (no_store ? 0u : (uint32_t)in.dst << 14) | (no_store ? 1u << 31 : 0u);
            }
            if (n_tt) hdr = tts[0] | (uint32_t)tts[1] << 8 | (uint32_t)tts[2] << 16 | n_tt << 24 | (reads_c ? 1u << 26 : 0u);
            if (n_tt == 1 && ((tts[0] == 0x0E && reads_c) || (tts[0] == 0x03 && !reads_c))) hdr |= 1u << 31;  // the kernel's in-line tables
            // The row's loads run in every lane. A lane without an instruction repeats the b and c addresses of a lane
            // that has one, in its own quarter-warp if possible: equal addresses are served by the same wavefront, so
            // the idle lane costs nothing (reading any fixed slot instead would collide with the lane that really
            // reads that bank group).
            for (uint32_t q0 = 0; q0 < 32 && n_tt; q0 += 8) {
              int donor = -1;
              for (uint32_t i = q0; i < q0 + 8 && donor < 0; i++)
                if (row[i].kind == K_LUT) donor = (int)i;
              for (uint32_t i = 0; i < 32 && donor < 0; i++)
                if (row[i].kind == K_LUT) donor = (int)i;
              for (uint32_t i = q0; i < q0 + 8; i++)
                if (row[i].kind != K_LUT) x[i] = EMPTY_X | (x[donor] & (0x3FFFu << 14)), y[i] = EMPTY_Y | (y[donor] & 0x3FFFu);
            }
          }
          if (q == 0 && sgm >= 2) hdr |= 1u << 30;
          if (q + 1 == n_rows) {
            hdr |= 1u << 27;
            if (sgm < S0 && stab[sgm + 1] > stab[sgm]) hdr |= 1u << 28;
            if (sgm + 1 == S) hdr |= 1u << 29;
          }
          for (uint32_t i = 0; i < 32; i++) words.push_back(x[i] | ((hdr >> i & 1u) << 31)), words.push_back(y[i]);
        }
      }
      // a multiple of four rows per warp (clg_coop_rows_t stages them four at a time): empty rows before the last one
      {
        const size_t n_rows_w = (words.size() - first_word) / 64, pad = (4 - n_rows_w % 4) % 4;
        std::vector<uint32_t> last(words.end() - 64, words.end());
        words.resize(words.size() - 64);
        for (size_t i = 0; i < pad * 32; i++) words.push_back(EMPTY_X), words.push_back(EMPTY_Y);
        words.insert(words.end(), last.begin(), last.end());
        wbase[nw + w] = (uint32_t)(n_rows_w + pad);
      }
      for (uint32_t i = 0; i < 8 * 32; i++) words.push_back(EMPTY_X), words.push_back(EMPTY_Y);  // clg_coop_rows fetches up to seven rows ahead
    }
    if (words.size() / 2 >= (1ull << 32)) fail(CLG_E_UNSUPPORTED, "row stream too long");
    std::vector<uint32_t> side((size_t)std::max<uint32_t>(r.n_side, 1) * 4, 0);
    encode_stream(flat.row_side(), r.n_side, 16, side.data());
    CU(cuMemAlloc(&P.d_rinstr, words.size() * 4));
    CU(cuMemcpyHtoDAsync(P.d_rinstr, words.data(), words.size() * 4, nullptr));
    CU(cuMemAlloc(&P.d_rwarp, wbase.size() * 4));
    CU(cuMemcpyHtoDAsync(P.d_rwarp, wbase.data(), wbase.size() * 4, nullptr));
    CU(cuMemAlloc(&P.d_rside, side.size() * 4));
    CU(cuMemcpyHtoDAsync(P.d_rside, side.data(), side.size() * 4, nullptr));
    std::vector<uint32_t> soff(stab, stab + S0 + 1);
    soff.resize((size_t)S + 1, stab[S0]);  // the padding segments have no side entries
    CU(cuMemAlloc(&P.d_rside_off, soff.size() * 4));
    CU(cuMemcpyHtoDAsync(P.d_rside_off, soff.data(), soff.size() * 4, nullptr));
    CU(cuStreamSynchronize(nullptr));
    P.ready = true;
    return P;
  }
  void launch_rows(Plan &P, RunParams p, CUstream stream) {
    p.instr = (const void *)P.d_rside, p.rinstr = (const void *)P.d_rinstr, p.rwarp = (const uint32_t *)P.d_rwarp;
    p.rside_off = (const uint32_t *)P.d_rside_off, p.code_off = P.code_off, p.ring_off = P.ring_off;
    const uint32_t per_sm = coop_ctas_per_sm(P, P.rows_nt);
    const uint32_t grid = std::min<uint32_t>(p.n_groups, (uint32_t)ctx->sm_count * per_sm);
    void *args[] = {&p};
    CU(cuLaunchKernel(P.rows_tma ? ctx->coop_rows_t : ctx->coop_rows, grid, 1, 1, P.rows_nt, 1, 1, (unsigned)P.smem, stream, args, nullptr));
  }

  RunParams params(const Plan &P, const uint32_t *imm, uint32_t *out, uint32_t *state, size_t stride, uint32_t groups) const {
    const FlatHeader &h = flat.header();
    RunParams p{};
    p.instr = (const void *)P.d_instr, p.levels = (const uint32_t *)P.d_levels;
    p.imm = imm, p.out = out, p.state = state;
    p.n_inputs = h.n_inputs, p.n_outputs = h.n_outputs;
    p.n_instr = P.n_instr, p.n_levels = P.n_levels, p.n_slots = P.n_slots;
    p.stride = (uint32_t)stride, p.n_groups = groups;
    p.parts = (const void *)P.d_parts, p.n_parts = P.n_parts, p.code_off = P.code_off;
    p.cinstr = (const void *)P.d_cinstr, p.clevels = (const uint32_t *)P.d_clevels;
    return p;
  }
  uint32_t coop_ctas_per_sm(const Plan &P, uint32_t nt) const {
    return (uint32_t)std::max<size_t>(1, std::min<size_t>(SMEM_MAX / (P.smem + 1024), 2048 / nt));
  }
  void launch_coop(const Plan &P, RunParams p, uint32_t nt, CUstream stream) {
    const uint32_t per_sm = coop_ctas_per_sm(P, nt), groups = p.n_groups;
    uint32_t grid = std::min<uint32_t>(groups, (uint32_t)ctx->sm_count * per_sm);
    if (P.n_parts) {  // one CTA per (part, group slot); at least one group slot, about one wave of CTAs
      const uint32_t slots_g = std::max<uint32_t>(1u, std::min<uint32_t>(groups, ((uint32_t)ctx->sm_count * per_sm + P.n_parts - 1) / P.n_parts));
      grid = P.n_parts * slots_g;
    }
    void *args[] = {&p};
    const int k = kidx(P.K);
    static const char *names[3][3] = {{"clg_coop_res_k1", "clg_coop_res_k2", "clg_coop_res_k4"}, {"clg_coop_c_k1", "clg_coop_c_k2", "clg_coop_c_k4"}, {"clg_coop_k1", "clg_coop_k2", "clg_coop_k4"}};
    last_kernel = names[P.resident ? 0 : P.d_cinstr ? 1 : 2][k];
    CU(cuLaunchKernel(P.resident ? ctx->coop_res[k] : P.d_cinstr ? ctx->coop_c[k] : ctx->coop[k], grid, 1, 1, nt, 1, 1, (unsigned)P.smem, stream, args, nullptr));
  }
  // Time the cooperative kernel of this plan on scratch buffers (zero inputs: the work does not depend on the data)
  // for a few CTA sizes, each on enough column groups for 8 passes per CTA, and keep the fastest per group.
  uint32_t tune_coop_threads(const Plan &P) {
    const uint32_t wide = P.coop_threads;
    if (std::getenv("CLG_COOP_THREADS") || std::getenv("CLG_COOP_NO_TUNE") || wide <= 64) return wide;
    const FlatHeader &h = flat.header();
    const uint32_t groups = (uint32_t)ctx->sm_count * 16;
    const size_t stride = ((size_t)groups * P.K + 3) / 4 * 4;
    const size_t rows[3] = {h.n_inputs, h.n_outputs, h.n_state};
    if ((rows[0] + rows[1] + rows[2]) * stride * 4 > ((size_t)1 << 30)) return wide;  // (not worth a gigabyte of scratch)
    CUdeviceptr buf[3] = {0, 0, 0};
    CUevent e0 = nullptr, e1 = nullptr;
    uint32_t best = wide;
    bool ok = true;
    for (int i = 0; i < 3 && ok; i++) {
      const size_t bytes = std::max<size_t>(16, rows[i] * stride * 4);
      ok = driver().cuMemAlloc(&buf[i], bytes) == CUDA_SUCCESS && driver().cuMemsetD8Async(buf[i], 0, bytes, nullptr) == CUDA_SUCCESS;
    }
    ok = ok && driver().cuEventCreate(&e0, CU_EVENT_DEFAULT) == CUDA_SUCCESS && driver().cuEventCreate(&e1, CU_EVENT_DEFAULT) == CUDA_SUCCESS;
    if (ok) {
      const RunParams p = params(P, (const uint32_t *)buf[0], (uint32_t *)buf[1], (uint32_t *)buf[2], stride, groups);
      double best_ms = 1e30;
      for (uint32_t nt : {1024u, 768u, 512u, 384u, 256u, 128u}) {
        if (nt > wide && nt != 1024u) continue;
        nt = std::min(nt, wide);
        float ms = 1e30f;
        launch_coop(P, p, nt, nullptr);  // warm-up
        for (int rep = 0; rep < 2; rep++) {
          float t = 0;
          CU(cuEventRecord(e0, nullptr));
          launch_coop(P, p, nt, nullptr);
          CU(cuEventRecord(e1, nullptr));
          CU(cuEventSynchronize(e1));
          CU(cuEventElapsedTime(&t, e0, e1));
          ms = std::min(ms, t);
        }
        if (ms < best_ms * 0.97) best_ms = ms, best = nt;  // (a smaller CTA has to win by 3 %)
      }
    }
    driver().cuStreamSynchronize(nullptr);
    if (e0) driver().cuEventDestroy(e0);
    if (e1) driver().cuEventDestroy(e1);
    for (CUdeviceptr b : buf)
      if (b) driver().cuMemFree(b);
    return best;
  }
};

// ---- helpers used by the C ABI -------------------------------------------------------------------------
Ctx *ctx_create(int device) { return new Ctx(device); }
void ctx_destroy(Ctx *c) { delete c; }
const std::string &ctx_name(Ctx *c) { return c->name; }
int ctx_sm_count(Ctx *c) { return c->sm_count; }
void ctx_sync(Ctx *c) {
  c->bind();
  CU(cuCtxSynchronize());
}
void *dev_alloc(Ctx *c, size_t bytes) {
  c->bind();
  CUdeviceptr p = 0;
  CU(cuMemAlloc(&p, std::max<size_t>(bytes, 16)));
  return (void *)p;
}
void dev_free(Ctx *c, void *p) {
  c->bind();
  CU(cuMemFree((CUdeviceptr)p));
}
void dev_memset(Ctx *c, void *p, int byte, size_t bytes, void *stream) {
  c->bind();
  CU(cuMemsetD8Async((CUdeviceptr)p, (unsigned char)byte, bytes, (CUstream)stream));
}
void h2d(Ctx *c, void *dst, const void *src, size_t bytes, void *stream) {
  c->bind();
  CU(cuMemcpyHtoDAsync((CUdeviceptr)dst, src, bytes, (CUstream)stream));
}
void d2h(Ctx *c, void *dst, const void *src, size_t bytes, void *stream) {
  c->bind();
  CU(cuMemcpyDtoHAsync(dst, (CUdeviceptr)src, bytes, (CUstream)stream));
}
void *host_alloc(Ctx *c, size_t bytes) {
  c->bind();
  void *p = nullptr;
  CU(cuMemAllocHost(&p, std::max<size_t>(bytes, 16)));
  return p;
}
void host_free(Ctx *c, void *p) {
  c->bind();
  CU(cuMemFreeHost(p));
}
void stream_sync(Ctx *c, void *stream) {
  c->bind();
  CU(cuStreamSynchronize((CUstream)stream));
}

Exec *exec_create(Ctx *c, const Flat &f, int mode) { return new Exec(c, f, mode); }
void exec_destroy(Exec *e) { delete e; }
int exec_mode(const Exec *e) { return e->mode; }
int exec_k(const Exec *e) { return e->last_k; }
int exec_launches(const Exec *e) {
  if (e->last_launches) return e->last_launches;  // what the most recent run issued
  if (e->mode != CLG_MODE_SPEC) return 1;
  const Plan &P = e->plan[e->plan[CLG_MODE_SPEC].ready ? CLG_MODE_SPEC : Exec::PLAN_SPEC_TP];
  return P.chunks.empty() ? 1 : (int)P.chunks.size();
}
const char *exec_kernel_name(const Exec *e) { return e->last_kernel; }
const std::string &exec_ptx(const Exec *e) { return e->plan[e->plan[CLG_MODE_SPEC].ptx.empty() ? Exec::PLAN_SPEC_TP : CLG_MODE_SPEC].ptx; }
const Flat &exec_flat(const Exec *e) { return e->flat; }
Ctx *exec_ctx(const Exec *e) { return e->ctx; }
void exec_run_steps(Exec *e, const uint32_t *imm, uint32_t *out, uint32_t *state, size_t n_vectors, size_t stride, size_t n_steps,
                    void *stream) {
  e->run_steps(imm, out, state, n_vectors, stride, n_steps, (CUstream)stream);
}
void exec_run_host(Exec *e, const uint32_t *imm, uint32_t *out, uint32_t *state, size_t n_vectors, size_t stride) {
  e->run_host(imm, out, state, n_vectors, stride);
}
void exec_prepare(Exec *e, size_t n_vectors) { e->prepare(n_vectors); }
void exec_run_bools_host(Exec *e, const uint8_t *imm, uint8_t *out, size_t n_vectors) { e->run_bools_host(imm, out, n_vectors); }
void exec_run(Exec *e, const uint32_t *imm, uint32_t *out, uint32_t *state, size_t n_vectors, size_t stride, void *stream) {
  e->run(imm, out, state, n_vectors, stride, (CUstream)stream);
}

void launch_slice(Ctx *c, bool unslice, const void *src, void *dst, size_t n_vectors, size_t n_rows, size_t stride, void *stream) {
  if (n_vectors == 0 || n_rows == 0) return;
  if (n_vectors > 0xFFFFFFFFull || n_rows > 0xFFFFFFFFull || stride > 0xFFFFFFFFull) fail(CLG_E_INVALID, "transpose dimensions too large");
  if ((n_vectors + 31) / 32 > stride) fail(CLG_E_INVALID, "n_vectors does not fit in stride words");
  c->bind();
  uint32_t nv = (uint32_t)n_vectors, nr = (uint32_t)n_rows, st = (uint32_t)stride;
  void *args[] = {&src, &dst, &nv, &nr, &st};
  const void *vec_side = unslice ? dst : src;
  if (n_rows % 16 == 0 && ((uintptr_t)vec_side & 15) == 0 && !std::getenv("CLG_TRANSPOSE_NARROW") && !std::getenv("CLG_TRANSPOSE_NO_TMA")) {
    // TMA-tiled kernels: the vector-major side as a 2-D byte tensor [n_vectors][n_rows], boxes of 256 vectors x 32 / 64
    // bytes, swizzled in shared memory by the same span (kernels.cu)
    int i = n_rows >= 48 ? 2 : 0;  // 64 x 512 tiles; 32 x 1024 for narrow matrices
    if (const char *env = std::getenv("CLG_TRANSPOSE_TMA")) i = std::max(0, std::min(6, std::atoi(env)));  // tuning knob: tile shape by index
    const uint32_t rt = Ctx::TMA_RT[i], vt = Ctx::TMA_VT[i];
    CUtensorMap tm;
    const cuuint64_t dims[2] = {(cuuint64_t)n_rows, (cuuint64_t)n_vectors}, strides[1] = {(cuuint64_t)n_rows};
    const cuuint32_t box[2] = {rt, 256}, estr[2] = {1, 1};
    CU(cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(vec_side), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              rt == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : rt == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE));
    const void *other = unslice ? src : dst;
    void *targs[] = {&tm, &other, &nv, &nr, &st};
    const uint64_t gx = (n_vectors + vt - 1) / vt, gy = (n_rows + rt - 1) / rt;
    if (gx * gy > 0x7FFFFFFFull) fail(CLG_E_INVALID, "transpose too large for one launch");
    CU(cuLaunchKernel(unslice ? c->unslice_tma[i] : c->slice_tma[i], (unsigned)(gx * gy), 1, 1, 256, 1, 1, (unsigned)Ctx::tma_smem(i), (CUstream)stream, targs, nullptr));
    return;
  }
  if (n_rows % 16 == 0 && ((uintptr_t)vec_side & 15) == 0 && !std::getenv("CLG_TRANSPOSE_NARROW")) {
    // wide tiled kernels: a CTA owns 1024 vectors x 32 / 64 / 128 rows; up to 128 contiguous bytes per lane on the
    // vector-major side, whole 128-byte lines on the sliced side
    if (unslice && n_rows >= 64 && !std::getenv("CLG_TRANSPOSE_RT")) {  // staged 64-row tiles: 64 contiguous bytes per vector and store instruction
      const uint64_t gx = ((n_vectors + 31) / 32 + 31) / 32, gy = (n_rows + 63) / 64;
      if (gx * gy > 0x7FFFFFFFull) fail(CLG_E_INVALID, "transpose too large for one launch");
      CU(cuLaunchKernel(c->unslice_s64, (unsigned)(gx * gy), 1, 1, 256, 1, 1, 0, (CUstream)stream, args, nullptr));
      return;
    }
    int i = n_rows >= 64 ? 1 : 0;  // 64-row tiles (64 contiguous bytes per lane), 32-row tiles for narrow matrices
    if (const char *env = std::getenv("CLG_TRANSPOSE_RT")) i = std::atoi(env) >= 128 ? 2 : std::atoi(env) >= 64 ? 1 : 0;  // tuning knob
    const uint64_t rt = 32u << i, gx = ((n_vectors + 31) / 32 + 31) / 32, gy = (n_rows + rt - 1) / rt;
    if (gx * gy > 0x7FFFFFFFull) fail(CLG_E_INVALID, "transpose too large for one launch");
    CU(cuLaunchKernel(unslice ? c->unslice_w[i] : c->slice_w[i], (unsigned)(gx * gy), 1, 1, 256, 1, 1, 0, (CUstream)stream, args, nullptr));
    return;
  }
  if (n_rows % 16 == 0 && ((uintptr_t)vec_side & 15) == 0) {
    // tiled kernels: 128-bit accesses on the vector-major side, 32-byte segments on the sliced side
    const uint64_t gx = ((n_vectors + 31) / 32 + 7) / 8, gy = (n_rows + 31) / 32;
    if (gx * gy > 0x7FFFFFFFull) fail(CLG_E_INVALID, "transpose too large for one launch");
    CU(cuLaunchKernel(unslice ? c->unslice_v4 : c->slice_v4, (unsigned)(gx * gy), 1, 1, 256, 1, 1, 0, (CUstream)stream, args, nullptr));
    return;
  }
  if (((uintptr_t)vec_side & 15) == 0 && n_rows * 32 <= Ctx::GEN_MAX_BYTES && !std::getenv("CLG_TRANSPOSE_NARROW") && !std::getenv("CLG_TRANSPOSE_BYTEWISE")) {
    // any row count: VT consecutive vectors are one contiguous, 32-byte aligned run of VT * n_rows bytes, staged in shared
    // memory by a 1-D bulk copy (kernels.cu, clg_slice_gen)
    // vectors per CTA: up to 2048 (eight groups of 256, so that a narrow matrix still has a unit of work for every warp),
    // in whole groups while more than one fits, else a multiple of 32
    size_t fit = Ctx::GEN_MAX_BYTES / n_rows;
    uint32_t vt = (uint32_t)(fit >= 512 ? std::min<size_t>(2048, fit / 256 * 256) : std::min<size_t>(256, fit / 32 * 32));
    vt = (uint32_t)std::min<size_t>(vt, (n_vectors + 31) / 32 * 32);
    const uint64_t grid = (n_vectors + vt - 1) / vt;
    if (grid > 0x7FFFFFFFull) fail(CLG_E_INVALID, "transpose too large for one launch");
    const void *other = unslice ? src : dst, *vec = vec_side;
    void *gargs[] = {unslice ? (void *)&other : (void *)&vec, unslice ? (void *)&vec : (void *)&other, &nv, &nr, &st, &vt};
    CU(cuLaunchKernel(unslice ? c->unslice_gen : c->slice_gen, (unsigned)grid, 1, 1, 256, 1, 1, (unsigned)((size_t)vt * n_rows + 256), (CUstream)stream, gargs, nullptr));
    return;
  }
  uint64_t warps = (uint64_t)((n_rows + 31) / 32) * ((n_vectors + 31) / 32);
  uint64_t blocks = (warps * 32 + 255) / 256;
  if (blocks > 0x7FFFFFFFull) fail(CLG_E_INVALID, "transpose too large for one launch");
  CU(cuLaunchKernel(unslice ? c->unslice : c->slice, (unsigned)blocks, 1, 1, 256, 1, 1, 0, (CUstream)stream, args, nullptr));
}

void launch_fill(Ctx *c, uint32_t *dst, size_t n_rows, size_t stride, uint64_t seed, void *stream) {
  if (n_rows == 0 || stride == 0) return;
  if (n_rows > 0xFFFFFFFFull || stride > 0xFFFFFFFFull) fail(CLG_E_INVALID, "fill dimensions too large");
  c->bind();
  uint32_t nr = (uint32_t)n_rows, st = (uint32_t)stride;
  void *args[] = {&dst, &nr, &st, &seed};
  CU(cuLaunchKernel(c->fill, (unsigned)(c->sm_count * 8), 1, 1, 256, 1, 1, 0, (CUstream)stream, args, nullptr));
}

// pure-LOP3 microbenchmark: returns lane-ops (32-bit LOP3 results) per second
double measure_lop3_peak(Ctx *c, int reps) {
  c->bind();
  CUdeviceptr seed = 0, sink = 0;
  const uint32_t threads = 512, blocks = (uint32_t)c->sm_count * 4, iters = 64;
  CU(cuMemAlloc(&seed, 4096));
  CU(cuMemAlloc(&sink, (size_t)threads * blocks * 4));
  CU(cuMemsetD8Async(seed, 0x5A, 4096, nullptr));
  CUevent a, b;
  CU(cuEventCreate(&a, CU_EVENT_DEFAULT));
  CU(cuEventCreate(&b, CU_EVENT_DEFAULT));
  uint32_t it = iters;
  void *args[] = {&seed, &sink, &it};
  double best = 0;
  for (int r = 0; r < reps + 2; r++) {
    CU(cuEventRecord(a, nullptr));
    CU(cuLaunchKernel(c->lop3_peak, blocks, 1, 1, threads, 1, 1, 0, nullptr, args, nullptr));
    CU(cuEventRecord(b, nullptr));
    CU(cuEventSynchronize(b));
    float ms = 0;
    CU(cuEventElapsedTime(&ms, a, b));
    const double ops = (double)threads * blocks * 8.0 * 64.0 * iters;
    if (r >= 2) best = std::max(best, ops / (ms * 1e-3));
  }
  driver().cuEventDestroy(a), driver().cuEventDestroy(b);
  driver().cuMemFree(seed), driver().cuMemFree(sink);
  return best;
}

}  // namespace clg
