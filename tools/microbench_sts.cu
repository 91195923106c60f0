// What does one STS.128 cost, by how the 32 lanes' bank groups (16-byte units mod 8) are arranged?  ncu shows 4.8
// wavefronts per STS.128 in clg_coop_rows although every quarter-warp writes eight different bank groups (the static
// model says 4.0), and tools/microbench.cu measures exactly 4.0 when lane l writes bank group l % 8.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mbs tools/microbench_sts.cu && /tmp/mbs
#include <algorithm>
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

template <bool LOAD>
__global__ void __launch_bounds__(1024) k_sts(const uint32_t *slot, uint32_t steps, uint32_t *sink, unsigned long long *cycles) {
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
      if (LOAD) {
        uint4 v;
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(o));
        acc ^= v.x ^ v.y ^ v.z ^ v.w;
      } else {
        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(o), "r"(acc), "r"(rot), "r"(tid), "r"(s) : "memory");
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
    int kind;
  };
  const Pat pats[] = {{"contiguous: slot = 32 w + lane", 0},
                      {"random rows, group = lane % 8", 1},
                      {"random rows, random permutation of groups per quarter", 2},
                      {"random rows, one random permutation for all quarters", 3},
                      {"random rows, group = 7 - lane % 8", 4},
                      {"random rows, group = (lane % 8) ^ 1", 5},
                      {"random rows, group = (lane % 8) ^ 2", 6},
                      {"random rows, group = (lane % 8) ^ 4", 7},
                      {"random rows, group = (lane % 8 + quarter) % 8", 8},
                      {"random rows, group = (lane % 8 + 2 quarter) % 8", 9},
                      {"same row per quarter, random permutation of groups", 10},
                      {"same row per quarter, group = lane % 8", 11}};
  const uint32_t st = 40000;
  for (int load = 0; load < 2; load++)
    for (const Pat &p : pats) {
      std::vector<uint32_t> slot(1024);
      for (uint32_t w = 0; w < 32; w++) {
        uint32_t perm_all[8] = {0, 1, 2, 3, 4, 5, 6, 7};
        for (int i = 7; i > 0; i--) std::swap(perm_all[i], perm_all[rnd() % (i + 1)]);
        for (uint32_t q = 0; q < 4; q++) {
          uint32_t perm[8] = {0, 1, 2, 3, 4, 5, 6, 7};
          for (int i = 7; i > 0; i--) std::swap(perm[i], perm[rnd() % (i + 1)]);
          const uint32_t qrow = rnd() % 1000;
          for (uint32_t i = 0; i < 8; i++) {
            const uint32_t l = q * 8 + i, row = rnd() % 1000;
            uint32_t s = 0;
            switch (p.kind) {
              case 0: s = w * 32 + l; break;
              case 1: s = row * 8 + i; break;
              case 2: s = row * 8 + perm[i]; break;
              case 3: s = row * 8 + perm_all[i]; break;
              case 4: s = row * 8 + 7 - i; break;
              case 5: s = row * 8 + (i ^ 1); break;
              case 6: s = row * 8 + (i ^ 2); break;
              case 7: s = row * 8 + (i ^ 4); break;
              case 8: s = row * 8 + (i + q) % 8; break;
              case 9: s = row * 8 + (i + 2 * q) % 8; break;
              case 10: s = qrow * 8 + perm[i]; break;
              case 11: s = qrow * 8 + i; break;
            }
            slot[l + w * 32] = s;
          }
        }
      }
      CK(cudaMemcpy(d_slot, slot.data(), 4096, cudaMemcpyHostToDevice));
      auto launch = [&](auto kern) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 16));
        for (int rep = 0; rep < 2; rep++) kern<<<sms, 1024, 8192 * 16>>>(d_slot, st, sink, cyc);
        CK(cudaDeviceSynchronize());
      };
      if (load) launch(k_sts<true>);
      else launch(k_sts<false>);
      std::vector<unsigned long long> h(sms);
      CK(cudaMemcpy(h.data(), cyc, sms * 8, cudaMemcpyDeviceToHost));
      const double c = (double)*std::max_element(h.begin(), h.end());
      std::printf("%s  %-58s %.2f cycles per warp instruction\n", load ? "LDS.128" : "STS.128", p.name, c / ((double)st * 32));
    }
  return 0;
}
