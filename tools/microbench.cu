// On-chip data-path microbenchmarks that bound the cooperative interpreter (clg_coop*) on B200.
// Built and run on the GPU box:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mb tools/microbench.cu && /tmp/mb
//
//   gather    the 1M-gate DAG's shared-memory access pattern with NO instruction fetch, decode or dispatch:
//             every thread does `loads` LDS.128 from pseudo-random slots + 4 LOP3 + `stores` STS.128 per step, in a
//             persistent CTA per SM with a 10.5k-slot stack, bank-conflict-free per quarter-warp or fully random.
//             This is the ceiling any interpreter that keeps the register stack in shared memory can reach.
//   shfl      SHFL.IDX alone, LDS.128 alone, both interleaved: is the shuffle network a second data path next to the
//             shared-memory banks?
//   tmem      tcgen05.st / tcgen05.ld 32x32b.x4 with a run-time column: the [slot][thread] stack shape of clg_lane.
//   barrier   bar.sync cost for 1024 threads with a little work between barriers.
// Every line printed is "name  params  ->  measured rate" with the SM clock taken from clock64 deltas and the
// wall time from CUDA events.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess) {                                                               \
      std::fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
      std::exit(1);                                                                        \
    }                                                                                      \
  } while (0)

__device__ __forceinline__ uint32_t lop3_0e(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, 0x0E;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ---- gather -------------------------------------------------------------------------------------------
// offs: [steps][threads] uint4 of byte offsets {a, b, c, dst} into the stack, precomputed on the host and read
// through the read-only path one step ahead (so the measurement includes a realistic 16 B/step fetch from L2
// when FETCH=1, or nothing at all when FETCH=0: the offsets are then derived in registers from a counter).
template <int LOADS, bool STORE, bool FETCH, bool BARRIER>
__global__ void __launch_bounds__(1024) k_gather(const uint4 *offs, uint32_t steps, uint32_t slots, uint32_t per_level, uint32_t *sink,
                                                 unsigned long long *cycles) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t tid = threadIdx.x, nt = blockDim.x;
  uint32_t sb = (uint32_t)__cvta_generic_to_shared(smem);
  for (uint32_t i = tid; i < slots * 4; i += nt) reinterpret_cast<uint32_t *>(smem)[i] = i * 2654435761u;
  __syncthreads();
  uint4 acc = make_uint4(tid, tid * 3, tid * 5, tid * 7);
  const unsigned long long t0 = clock64();
  // register-only address stream: conflict-free by construction. lane l of a quarter reads bank group (l + r) % 8
  uint32_t r = tid * 2654435761u, rq = (tid >> 3) * 2246822519u + 7u;
  const uint32_t rows = slots / 8;
  uint4 nxt = make_uint4(0, 0, 0, 0);
  if (FETCH) nxt = __ldg(offs + tid);
  for (uint32_t s = 0; s < steps; s++) {
    uint4 o;
    if (FETCH) {
      o = nxt;
      if (s + 1 < steps) nxt = __ldg(offs + (size_t)(s + 1) * nt + tid);
    } else {
      r = r * 1664525u + 1013904223u;
      rq = rq * 1664525u + 1013904223u;  // same sequence in the 8 lanes of a quarter-warp: one rotation per quarter
      const uint32_t g = tid & 7;
      o.x = (((r >> 8) % rows) * 8 + ((g + (rq >> 29)) & 7)) * 16;
      o.y = (((r >> 12) % rows) * 8 + ((g + ((rq >> 26) & 7)) & 7)) * 16;
      o.z = (((r >> 16) % rows) * 8 + ((g + ((rq >> 23) & 7)) & 7)) * 16;
      o.w = (((r >> 4) % rows) * 8 + g) * 16;
    }
    uint4 a = lds128(sb + o.x), b = acc, c = acc;
    if (LOADS >= 2) b = lds128(sb + o.y);
    if (LOADS >= 3) c = lds128(sb + o.z);
    acc.x ^= lop3_0e(a.x, b.x, c.x), acc.y ^= lop3_0e(a.y, b.y, c.y), acc.z ^= lop3_0e(a.z, b.z, c.z), acc.w ^= lop3_0e(a.w, b.w, c.w);
    if (STORE) sts128(sb + o.w, acc);
    if (BARRIER && (s + 1) % per_level == 0) __syncthreads();
  }
  const unsigned long long t1 = clock64();
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
  if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x12345678u) sink[blockIdx.x * nt + tid] = acc.x;
}

// ---- gather, clean form ---------------------------------------------------------------------------------
// No modulo, no 16-byte fetch: each thread keeps four precomputed conflict-free address quads in registers and
// walks them through the stack in 128-byte steps (which keeps every bank group), so the loop is LDS/STS/LOP3 plus
// two cheap integer instructions per address. FETCH8 adds the real kernel's 8 bytes per LUT from an L2-resident
// stream (ld.global.nc.v2, one step ahead). This is the number the interpreter is compared with.
template <int LOADS, int STORES, bool FETCH8, int BAR_EVERY>
__global__ void __launch_bounds__(1024) k_gather2(const uint2 *stream, uint32_t steps, uint32_t *sink, unsigned long long *cycles) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t tid = threadIdx.x, nt = blockDim.x;
  uint32_t sb = (uint32_t)__cvta_generic_to_shared(smem);
  for (uint32_t i = tid; i < 8192 * 4; i += nt) reinterpret_cast<uint32_t *>(smem)[i] = i * 2654435761u;
  __syncthreads();
  uint32_t base[4][4];
  uint32_t r = tid * 2654435761u + 1u, rq = (tid >> 3) * 2246822519u + 7u;
#pragma unroll
  for (int u = 0; u < 4; u++)
#pragma unroll
    for (int k = 0; k < 4; k++) {
      r = r * 1664525u + 1013904223u, rq = rq * 1664525u + 1013904223u;
      base[u][k] = (((r >> 10) & 1023u) * 8u + (((tid & 7u) + (k == 3 ? 0u : rq >> 29)) & 7u)) * 16u;
    }
  uint4 acc = make_uint4(tid, tid * 3, tid * 5, tid * 7);
  uint2 nxt = make_uint2(0, 0);
  if (FETCH8) nxt = __ldg(stream + tid);
  uint32_t rot = 0;
  const unsigned long long t0 = clock64();
  for (uint32_t s = 0; s < steps; s += 4) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
      uint32_t x = 0;
      if (FETCH8) {
        x = nxt.x ^ nxt.y;  // (the stream holds zeros: the addresses do not change, the dependence is real)
        nxt = __ldg(stream + (size_t)(s + u + 1) * nt + tid);
      }
      rot += 128u * 37u;
      const uint32_t oa = ((base[u][0] + rot) & 0x1FFFFu) | x, ob = ((base[u][1] + rot) & 0x1FFFFu) | x, oc = ((base[u][2] + rot) & 0x1FFFFu) | x,
                     od = ((base[u][3] + rot) & 0x1FFFFu) | x;
      uint4 a = lds128(sb + oa), b = acc, c = acc;
      if (LOADS >= 2) b = lds128(sb + ob);
      if (LOADS >= 3) c = lds128(sb + oc);
      acc.x ^= lop3_0e(a.x, b.x, c.x), acc.y ^= lop3_0e(a.y, b.y, c.y), acc.z ^= lop3_0e(a.z, b.z, c.z), acc.w ^= lop3_0e(a.w, b.w, c.w);
      if (STORES >= 1) sts128(sb + od, acc);
      if (STORES >= 2) sts128(sb + oa, acc);
      if (BAR_EVERY && (s + u + 1) % BAR_EVERY == 0) __syncthreads();
    }
  }
  const unsigned long long t1 = clock64();
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
  if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x12345678u) sink[blockIdx.x * nt + tid] = acc.x;
}

// ---- shfl vs lds --------------------------------------------------------------------------------------
template <int MODE>  // 0: shfl only, 1: lds only, 2: both
__global__ void __launch_bounds__(1024) k_shfl(uint32_t steps, uint32_t *sink, unsigned long long *cycles) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t tid = threadIdx.x;
  uint32_t sb = (uint32_t)__cvta_generic_to_shared(smem);
  for (uint32_t i = tid; i < 8192 * 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = i;
  __syncthreads();
  uint4 acc = make_uint4(tid, tid + 1, tid + 2, tid + 3);
  uint32_t r = tid * 2654435761u;
  const unsigned long long t0 = clock64();
  for (uint32_t s = 0; s < steps; s++) {
    r = r * 1664525u + 1013904223u;
    if (MODE != 1) {
      const uint32_t src = r >> 27;
      acc.x ^= __shfl_sync(0xFFFFFFFFu, acc.y, src);
      acc.y ^= __shfl_sync(0xFFFFFFFFu, acc.z, src);
      acc.z ^= __shfl_sync(0xFFFFFFFFu, acc.w, src);
      acc.w ^= __shfl_sync(0xFFFFFFFFu, acc.x, src);
    }
    if (MODE != 0) {
      const uint32_t off = ((((r >> 8) & 1023) * 8) + (tid & 7)) * 16;
      const uint4 v = lds128(sb + off);
      acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
    }
  }
  const unsigned long long t1 = clock64();
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
  if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x12345678u) sink[blockIdx.x * blockDim.x + tid] = acc.x;
}

// ---- tensor memory as a per-thread register stack -----------------------------------------------------
// 128 lanes x 512 columns of 32 bits per SM; a warp reaches lanes 32*(warp%4)..+31, the column comes from a
// register. MODE 0: st only, 1: ld only, 2: ld,ld,ld,st per step (one interpreted LUT at K = 4).
template <int MODE>
__global__ void __launch_bounds__(1024) k_tmem(uint32_t steps, uint32_t *sink, unsigned long long *cycles) {
  __shared__ uint32_t base_s;
  const uint32_t tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = base_s + (((warp & 3) * 32) << 16);  // lane field in bits 31:16
  uint32_t a0 = tid, a1 = tid + 1, a2 = tid + 2, a3 = tid + 3;
  // fill this warp's columns (each warp of a lane quarter takes its own column range so that they do not race)
  const uint32_t wq = warp >> 2, nq = (blockDim.x >> 5) / 4, span = 512 / nq, c0 = wq * span;
  for (uint32_t c = 0; c < span; c += 4)
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(base + c0 + c), "r"(a0), "r"(a1), "r"(a2), "r"(a3));
  asm volatile("tcgen05.wait::st.sync.aligned;");
  uint32_t r = warp * 2654435761u + 12345u;
  const unsigned long long t0 = clock64();
  for (uint32_t s = 0; s < steps; s++) {
    r = r * 1664525u + 1013904223u;
    const uint32_t ca = c0 + ((r >> 8) % (span / 4)) * 4, cb = c0 + ((r >> 14) % (span / 4)) * 4, cc = c0 + ((r >> 20) % (span / 4)) * 4,
                   cd = c0 + ((r >> 4) % (span / 4)) * 4;
    if (MODE == 0) {
      asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(base + cd), "r"(a0), "r"(a1), "r"(a2), "r"(a3));
      a0 += s;
    } else if (MODE == 1) {
      uint32_t x0, x1, x2, x3;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3) : "r"(base + ca));
      asm volatile("tcgen05.wait::ld.sync.aligned;");
      a0 ^= x0, a1 ^= x1, a2 ^= x2, a3 ^= x3;
    } else {
      uint32_t x0, x1, x2, x3, y0, y1, y2, y3, z0, z1, z2, z3;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3) : "r"(base + ca));
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(y0), "=r"(y1), "=r"(y2), "=r"(y3) : "r"(base + cb));
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(z0), "=r"(z1), "=r"(z2), "=r"(z3) : "r"(base + cc));
      asm volatile("tcgen05.wait::ld.sync.aligned;");
      a0 = lop3_0e(x0, y0, z0), a1 = lop3_0e(x1, y1, z1), a2 = lop3_0e(x2, y2, z2), a3 = lop3_0e(x3, y3, z3);
      asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(base + cd), "r"(a0), "r"(a1), "r"(a2), "r"(a3));
    }
  }
  asm volatile("tcgen05.wait::st.sync.aligned;");
  const unsigned long long t1 = clock64();
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
  if ((a0 ^ a1 ^ a2 ^ a3) == 0x12345678u) sink[blockIdx.x * blockDim.x + tid] = a0;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base_s));
}

// ---- barrier -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_barrier(uint32_t steps, uint32_t *sink, unsigned long long *cycles) {
  uint32_t acc = threadIdx.x;
  const unsigned long long t0 = clock64();
  for (uint32_t s = 0; s < steps; s++) {
    acc = lop3_0e(acc, s, acc >> 3);
    __syncthreads();
  }
  const unsigned long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[threadIdx.x] = acc;
}

static double max_cycles(unsigned long long *d, int n) {
  std::vector<unsigned long long> h(n);
  CK(cudaMemcpy(h.data(), d, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  unsigned long long m = 0;
  for (auto x : h) m = x > m ? x : m;
  return (double)m;
}

int main() {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  uint32_t *sink;
  unsigned long long *cyc;
  CK(cudaMalloc(&sink, (size_t)sms * 1024 * 4));
  CK(cudaMalloc(&cyc, sms * sizeof(unsigned long long)));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  std::printf("device SMs %d\n", sms);

  // ---- gather
  const uint32_t slots = 10576, steps = 6000, nt = 1024, per_level = 3;
  const size_t smem = (size_t)slots * 16;
  // host-built offset stream for the FETCH variants: conflict-free quarters or fully random banks
  for (int random_banks = 0; random_banks < 2; random_banks++) {
    std::vector<uint4> offs((size_t)steps * nt);
    uint64_t x = 88172645463325252ull;
    auto rnd = [&]() {
      x ^= x << 13, x ^= x >> 7, x ^= x << 17;
      return (uint32_t)(x >> 16);
    };
    const uint32_t rows = slots / 8;
    uint32_t qrot[3] = {0, 0, 0};
    for (size_t i = 0; i < offs.size(); i++) {
      const uint32_t g = (uint32_t)(i & 7);
      if (g == 0)
        for (uint32_t &q : qrot) q = rnd() & 7;  // one rotation per quarter-warp and operand: its 8 lanes hit 8 bank groups
      int which = 0;
      auto pick = [&](bool dst) {
        const uint32_t row = rnd() % rows, rot = dst ? 0 : qrot[which++];
        const uint32_t bank = random_banks ? rnd() & 7 : (g + rot) & 7;
        return (row * 8 + bank) * 16;
      };
      offs[i] = make_uint4(pick(false), pick(false), pick(false), pick(true));
    }
    uint4 *d_offs;
    CK(cudaMalloc(&d_offs, offs.size() * 16));
    CK(cudaMemcpy(d_offs, offs.data(), offs.size() * 16, cudaMemcpyHostToDevice));
    auto run = [&](const char *name, auto kern, int loads, int stores, int fetch) {
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      for (int rep = 0; rep < 2; rep++) {
        CK(cudaEventRecord(e0));
        kern<<<sms, nt, smem>>>(d_offs, steps, slots, per_level, sink, cyc);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
      }
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      const double c = max_cycles(cyc, sms), luts = (double)steps * nt;
      const double bytes = luts * (loads + stores) * 16 + (fetch ? luts * 16 : 0);
      std::printf("gather %-28s banks=%s  %.1f cyc/1024-LUT step  %.2f word-LUT/clk/SM  %.1f smem B/clk/SM (stack only %.1f)  %.3f ms  clk %.0f MHz\n", name,
                  random_banks ? "random " : "conflict-free", c / steps, luts * 4 / c, bytes / c, luts * (loads + stores) * 16 / c, ms, c / ms / 1e3);
    };
    if (!random_banks) {
      run("3ld+1st regs-only", k_gather<3, true, false, false>, 3, 1, 0);
      run("2ld+1st regs-only", k_gather<2, true, false, false>, 2, 1, 0);
      run("1ld+1st regs-only", k_gather<1, true, false, false>, 1, 1, 0);
      run("3ld+0st regs-only", k_gather<3, false, false, false>, 3, 0, 0);
      run("3ld+1st regs-only +bar/3", k_gather<3, true, false, true>, 3, 1, 0);
    }
    run("3ld+1st fetch16B", k_gather<3, true, true, false>, 3, 1, 1);
    run("3ld+1st fetch16B +bar/3", k_gather<3, true, true, true>, 3, 1, 1);
    run("2ld+1st fetch16B", k_gather<2, true, true, false>, 2, 1, 1);
    CK(cudaFree(d_offs));
  }

  // ---- gather, clean form
  {
    const uint32_t st = 8000;
    uint2 *d_stream;
    CK(cudaMalloc(&d_stream, (size_t)(st + 8) * 1024 * 8));
    CK(cudaMemset(d_stream, 0, (size_t)(st + 8) * 1024 * 8));
    auto run = [&](const char *name, auto kern, int loads, int stores, int fetch8) {
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 16));
      for (int rep = 0; rep < 2; rep++) kern<<<sms, 1024, 8192 * 16>>>(d_stream, st, sink, cyc);
      CK(cudaDeviceSynchronize());
      const double c = max_cycles(cyc, sms), luts = (double)st * 1024;
      const double wf = luts / 32 * ((loads + stores) * 4 + (fetch8 ? 2 : 0));
      std::printf("gather2 %-26s %.1f cyc/1024-LUT step  %.2f word-LUT/clk/SM  stack %.1f B/clk/SM  ideal wavefronts/cycles %.3f\n", name, c / st, luts * 4 / c,
                  luts * (loads + stores) * 16 / c, wf / c);
    };
    run("1ld+0st", k_gather2<1, 0, false, 0>, 1, 0, 0);
    run("3ld+0st", k_gather2<3, 0, false, 0>, 3, 0, 0);
    run("1ld+1st", k_gather2<1, 1, false, 0>, 1, 1, 0);
    run("1ld+2st", k_gather2<1, 2, false, 0>, 1, 2, 0);
    run("2ld+1st", k_gather2<2, 1, false, 0>, 2, 1, 0);
    run("3ld+1st", k_gather2<3, 1, false, 0>, 3, 1, 0);
    run("3ld+1st bar/3", k_gather2<3, 1, false, 3>, 3, 1, 0);
    run("3ld+1st fetch8", k_gather2<3, 1, true, 0>, 3, 1, 1);
    run("3ld+1st fetch8 bar/3", k_gather2<3, 1, true, 3>, 3, 1, 1);
    run("2ld+1st fetch8", k_gather2<2, 1, true, 0>, 2, 1, 1);
    run("2ld+1st fetch8 bar/3", k_gather2<2, 1, true, 3>, 2, 1, 1);
    run("2ld+1st fetch8 bar/6", k_gather2<2, 1, true, 6>, 2, 1, 1);
    CK(cudaFree(d_stream));
  }

  // ---- shfl
  {
    auto run = [&](const char *name, auto kern, double shfl_lanes, double lds_bytes) {
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 16));
      const uint32_t st = 20000;
      for (int rep = 0; rep < 2; rep++) kern<<<sms, 1024, 8192 * 16>>>(st, sink, cyc);
      CK(cudaDeviceSynchronize());
      const double c = max_cycles(cyc, sms);
      std::printf("shfl   %-12s %.2f cyc/step  shfl %.1f lanes/clk/SM  lds %.1f B/clk/SM\n", name, c / st, shfl_lanes * st * 1024 / c, lds_bytes * st * 1024 / c);
    };
    run("shfl x4", k_shfl<0>, 4, 0);
    run("lds.128", k_shfl<1>, 0, 16);
    run("both", k_shfl<2>, 4, 16);
  }

  // ---- tmem
  for (int threads : {128, 256, 512, 1024}) {
    auto run = [&](const char *name, auto kern, double bytes_per_thread_step) {
      const uint32_t st = 20000;
      for (int rep = 0; rep < 2; rep++) kern<<<sms, threads>>>(st, sink, cyc);
      CK(cudaDeviceSynchronize());
      const double c = max_cycles(cyc, sms);
      std::printf("tmem   %-14s threads=%4d  %.2f cyc/step  %.1f B/clk/SM\n", name, threads, c / st, bytes_per_thread_step * st * threads / c);
    };
    run("st.x4", k_tmem<0>, 16);
    run("ld.x4", k_tmem<1>, 16);
    run("3ld+lop3+st", k_tmem<2>, 64);
  }

  // ---- barrier
  for (int threads : {256, 512, 1024}) {
    const uint32_t st = 20000;
    for (int rep = 0; rep < 2; rep++) k_barrier<<<sms, threads>>>(st, sink, cyc);
    CK(cudaDeviceSynchronize());
    std::printf("barrier threads=%4d  %.1f cyc per bar.sync\n", threads, max_cycles(cyc, sms) / st);
  }
  return 0;
}
