// Does the shared-memory pipe of B200 merge lanes that read the SAME 16-byte slot with one LDS.128 (or the same
// 8 bytes with LDS.64) when they sit in different quarter-warps?  The row kernel (clg_coop_rows) pays 4 wavefronts per
// LDS.128 of a row, and every value of the 1M-gate DAG has ~2.6 readers: if duplicates merged, rows built from LUTs
// that share operands would move fewer wavefronts.  Built and run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mbd tools/microbench_dup.cu && /tmp/mbd
// Prints cycles per warp instruction per SM sub-partition stream (1024 threads per SM, all 148 SMs); 4.0 is the
// conflict-free cost of 32 distinct slots (512 bytes through a 128 B/clk pipe).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess) {                                                               \
      std::fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
      std::exit(1);                                                                        \
    }                                                                                      \
  } while (0)

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

// slot[t]: the 16-byte slot thread t reads (pattern built on the host); every step adds 128 bytes (keeps the bank group)
template <int WIDTH>
__global__ void __launch_bounds__(1024) k_dup(const uint32_t *slot, uint32_t steps, uint32_t *sink, unsigned long long *cycles) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t tid = threadIdx.x;
  uint32_t sb = (uint32_t)__cvta_generic_to_shared(smem);
  for (uint32_t i = tid; i < 8192 * 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = i * 2654435761u;
  __syncthreads();
  const uint32_t base = slot[tid] * 16u;
  uint32_t rot = 0, acc = tid;
  const unsigned long long t0 = clock64();
  for (uint32_t s = 0; s < steps; s += 4) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
      rot += 128u * 37u;
      const uint32_t o = sb + ((base + rot) & 0x1FFFFu);
      if (WIDTH == 16) {
        const uint4 a = lds128(o);
        acc ^= a.x ^ a.y ^ a.z ^ a.w;
      } else if (WIDTH == 8) {
        const uint2 a = lds64(o);
        acc ^= a.x ^ a.y;
      } else {
        acc ^= lds32(o);
      }
    }
  }
  const unsigned long long t1 = clock64();
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[blockIdx.x * blockDim.x + tid] = acc;
}

int main() {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  uint32_t *sink, *d_slot;
  unsigned long long *cyc;
  CK(cudaMalloc(&sink, (size_t)sms * 1024 * 4));
  CK(cudaMalloc(&d_slot, 1024 * 4));
  CK(cudaMalloc(&cyc, sms * sizeof(unsigned long long)));
  uint64_t x = 88172645463325252ull;
  auto rnd = [&]() {
    x ^= x << 13, x ^= x >> 7, x ^= x << 17;
    return (uint32_t)(x >> 16);
  };
  struct Pat {
    const char *name;
    int distinct;  // distinct slots per warp
    // lane -> (which distinct slot, bank group of that slot)
    uint32_t (*which)(uint32_t lane);
    uint32_t (*group)(uint32_t which);
  };
  const Pat pats[] = {
      {"32 distinct, conflict-free quarters", 32, [](uint32_t l) { return l; }, [](uint32_t w) { return w & 7; }},
      {"32 distinct, 2-way conflict per quarter", 32, [](uint32_t l) { return l; }, [](uint32_t w) { return w & 3; }},
      {"16 distinct: lanes l, l+16 share", 16, [](uint32_t l) { return l & 15; }, [](uint32_t w) { return w & 7; }},
      {"16 distinct: lanes l, l+8 share", 16, [](uint32_t l) { return (l & 7) | (l >> 4) << 3; }, [](uint32_t w) { return w & 7; }},
      {"16 distinct: lanes 2i, 2i+1 share", 16, [](uint32_t l) { return l >> 1; }, [](uint32_t w) { return w & 7; }},
      {"16 distinct: l, l^1 share, groups 2/quarter", 16, [](uint32_t l) { return l >> 1; }, [](uint32_t w) { return (w & 3) * 2 + (w >> 3); }},
      {"8 distinct: l, l+8, l+16, l+24 share", 8, [](uint32_t l) { return l & 7; }, [](uint32_t w) { return w & 7; }},
      {"8 distinct: 4 adjacent lanes share", 8, [](uint32_t l) { return l >> 2; }, [](uint32_t w) { return w & 7; }},
      {"1 distinct: all lanes share", 1, [](uint32_t) { return 0u; }, [](uint32_t w) { return w & 7; }},
      {"16 distinct: shuffled lanes, 2 per group", 16, [](uint32_t l) { return (l * 5 + (l >> 3)) & 15; }, [](uint32_t w) { return w & 7; }},
  };
  const uint32_t st = 40000;
  for (int width : {16, 8, 4}) {
    for (const Pat &p : pats) {
      std::vector<uint32_t> slot(1024);
      for (uint32_t w = 0; w < 32; w++) {
        uint32_t rows[32];
        for (auto &r : rows) r = rnd() % 1000;
        for (uint32_t l = 0; l < 32; l++) {
          const uint32_t d = p.which(l);
          slot[w * 32 + l] = rows[d] * 8 + p.group(d);
        }
      }
      CK(cudaMemcpy(d_slot, slot.data(), 4096, cudaMemcpyHostToDevice));
      auto launch = [&](auto kern) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 16));
        for (int rep = 0; rep < 2; rep++) kern<<<sms, 1024, 8192 * 16>>>(d_slot, st, sink, cyc);
        CK(cudaDeviceSynchronize());
      };
      if (width == 16) launch(k_dup<16>);
      else if (width == 8) launch(k_dup<8>);
      else launch(k_dup<4>);
      std::vector<unsigned long long> h(sms);
      CK(cudaMemcpy(h.data(), cyc, sms * 8, cudaMemcpyDeviceToHost));
      const double c = (double)*std::max_element(h.begin(), h.end());
      std::printf("dup  LDS.%-3d %-46s %.2f cycles per warp instruction (ideal if merged %.2f)\n", width * 8, p.name, c / ((double)st * 32),
                  std::max(1.0, p.distinct * width / 128.0));
    }
  }
  return 0;
}
