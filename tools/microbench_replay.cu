// Replays the shared-memory accesses of a row stream (tools/dump_rows.py) one row and one operand at a time, with all
// 32 warps of a CTA issuing the SAME LDS.128 / STS.128 pattern back to back, so cycles / (iterations * 32) is the
// number of wavefronts the hardware really spends on that pattern. Compared with the static model of
// tools/rows_stats.py (per quarter-warp: the largest number of distinct slots in one 16-byte bank group).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mbr tools/microbench_replay.cu && /tmp/mbr tools/_data/dag1m_rows.bin
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <set>
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

__global__ void __launch_bounds__(1024) k_replay(const uint16_t *tab, uint32_t n_rows, uint32_t iters, uint32_t *cyc, uint32_t *sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t tid = threadIdx.x, lane = tid & 31;
  uint32_t sb = (uint32_t)__cvta_generic_to_shared(smem);
  for (uint32_t i = tid; i < 8192 * 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = i * 2654435761u;
  uint32_t acc = tid;
  for (uint32_t row = blockIdx.x; row < n_rows; row += gridDim.x) {
    for (uint32_t op = 0; op < 4; op++) {
      const uint32_t idx = tab[(row * 4 + op) * 32 + lane];
      const bool active = idx != 0xFFFFu;
      const uint32_t base = idx * 16u;
      uint32_t rot = 0;
      __syncthreads();
      const unsigned long long t0 = clock64();
      const uint32_t on = active ? 1u : 0u;
      if (op < 3) {
        for (uint32_t it = 0; it < iters; it += 4) {
          uint4 v[4];
#pragma unroll
          for (int u = 0; u < 4; u++) {
            rot += 128u * 37u;
            const uint32_t o = sb + ((base + rot) & 0x1FFFFu);
            v[u] = make_uint4(0, 0, 0, 0);
            asm volatile("{\n.reg .pred p;\nsetp.ne.u32 p, %5, 0;\n@p ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];\n}\n"
                         : "+r"(v[u].x), "+r"(v[u].y), "+r"(v[u].z), "+r"(v[u].w)
                         : "r"(o), "r"(on));
          }
#pragma unroll
          for (int u = 0; u < 4; u++) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
        }
      } else {
        for (uint32_t it = 0; it < iters; it += 4) {
#pragma unroll
          for (int u = 0; u < 4; u++) {
            rot += 128u * 37u;
            const uint32_t o = sb + ((base + rot) & 0x1FFFFu);
            asm volatile("{\n.reg .pred p;\nsetp.ne.u32 p, %5, 0;\n@p st.shared.v4.u32 [%0], {%1,%2,%3,%4};\n}\n" ::"r"(o), "r"(acc), "r"(rot), "r"(tid), "r"(it),
                         "r"(on)
                         : "memory");
          }
        }
      }
      __syncthreads();
      const unsigned long long t1 = clock64();
      if (tid == 0) cyc[row * 4 + op] = (uint32_t)(t1 - t0);
    }
  }
  if (acc == 0x12345678u) sink[blockIdx.x * blockDim.x + tid] = acc;
}

// static model. pair_merge: lanes 2i, 2i+1 with the same slot count once AND, when every active pair of a half-warp
// is such a pair, the half-warp is one unit of 8 slots (what tools/microbench_dup.cu measured)
static int predict(const uint16_t *s, bool quarter_only) {
  int total = 0;
  if (!quarter_only) {
    // try half-warp units first: all lanes of the half paired, 8 distinct bank groups
    for (int h = 0; h < 2; h++) {
      bool all_pairs = true, any = false;
      int cnt[8] = {};
      std::set<uint16_t> seen;
      for (int i = 0; i < 8; i++) {
        const uint16_t x = s[16 * h + 2 * i], y = s[16 * h + 2 * i + 1];
        if (x != y) all_pairs = false;
        if (x != 0xFFFF) any = true;
      }
      if (all_pairs && any) {
        int mx = 0;
        for (int i = 0; i < 8; i++) {
          const uint16_t x = s[16 * h + 2 * i];
          if (x != 0xFFFF && seen.insert(x).second) mx = std::max(mx, ++cnt[x & 7]);
        }
        total += mx;
      } else {
        for (int q = 2 * h; q < 2 * h + 2; q++) {
          int c2[8] = {}, mx = 0;
          std::set<uint16_t> sn;
          for (int i = 0; i < 8; i++) {
            const uint16_t x = s[8 * q + i];
            if (x != 0xFFFF && sn.insert(x).second) mx = std::max(mx, ++c2[x & 7]);
          }
          total += mx;
        }
      }
    }
    return total;
  }
  for (int q = 0; q < 4; q++) {
    int cnt[8] = {}, mx = 0;
    std::set<uint16_t> seen;
    for (int i = 0; i < 8; i++) {
      const uint16_t x = s[8 * q + i];
      if (x != 0xFFFF && seen.insert(x).second) mx = std::max(mx, ++cnt[x & 7]);
    }
    total += mx;
  }
  return total;
}

int main(int argc, char **argv) {
  const char *path = argc > 1 ? argv[1] : "tools/_data/dag1m_rows.bin";
  FILE *f = std::fopen(path, "rb");
  if (!f) return std::fprintf(stderr, "cannot open %s\n", path), 1;
  std::vector<uint16_t> tab;
  {
    uint16_t buf[4096];
    size_t n;
    while ((n = std::fread(buf, 2, 4096, f)) > 0) tab.insert(tab.end(), buf, buf + n);
    std::fclose(f);
  }
  uint32_t n_rows = (uint32_t)(tab.size() / 128);
  // synthetic probes appended behind the real rows: lane predication and partial pair merging
  auto probe = [&](auto fn) {
    for (int op = 0; op < 4; op++)
      for (int l = 0; l < 32; l++) tab.push_back(fn(l));
    return n_rows++;
  };
  const uint32_t p0 = probe([](int l) { return (uint16_t)(l & 1 ? 0xFFFF : 100 * 8 + l * 8 + (l / 2) % 8); });        // odd lanes off, 16 distinct
  const uint32_t p1 = probe([](int l) { return (uint16_t)(l < 16 ? 0xFFFF : 200 * 8 + l * 8 + l % 8); });            // lanes 0-15 off
  const uint32_t p2 = probe([](int l) { return (uint16_t)(l < 16 ? 300 * 8 + (l / 2) * 8 + (l / 2) % 8 : 400 * 8 + l * 8 + l % 8); });  // half 0 paired, half 1 distinct
  const uint32_t p3 = probe([](int l) { return (uint16_t)(l < 8 ? 500 * 8 + (l / 2) * 8 + (l / 2) % 8 : 600 * 8 + l * 8 + l % 8); });   // quarter 0 paired only
  const uint32_t p4 = probe([](int l) { return (uint16_t)(l % 8 < 4 ? 0xFFFF : 700 * 8 + l * 8 + l % 8); });         // 4 lanes of every quarter off
  const uint32_t p5 = probe([](int l) { return (uint16_t)(l >= 8 ? 0xFFFF : 800 * 8 + l * 8 + l % 8); });            // one quarter active
  const uint32_t p6 = probe([](int l) { return (uint16_t)(900 * 8 + (l / 2) * 8 + (l / 2) % 8); });                  // all pairs merged (16 distinct)
  const uint32_t p7 = probe([](int l) { return (uint16_t)(l == 5 ? 1000 * 8 + 7 : 900 * 8 + (l / 2) * 8 + (l / 2) % 8); });  // ... but one lane breaks its pair
  const uint32_t probes[] = {p0, p1, p2, p3, p4, p5, p6, p7};
  const char *probe_name[] = {"odd lanes off (16 distinct)", "lanes 0-15 off", "half 0 paired, half 1 distinct", "quarter 0 paired, rest distinct",
                              "4 lanes of every quarter off", "one quarter active", "all pairs merged", "all pairs merged but lane 5"};

  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  uint16_t *d_tab;
  uint32_t *d_cyc, *sink;
  CK(cudaMalloc(&d_tab, tab.size() * 2));
  CK(cudaMemcpy(d_tab, tab.data(), tab.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&d_cyc, (size_t)n_rows * 16));
  CK(cudaMalloc(&sink, (size_t)sms * 1024 * 4));
  const uint32_t iters = 1024;
  CK(cudaFuncSetAttribute(k_replay, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 16));
  k_replay<<<sms, 1024, 8192 * 16>>>(d_tab, n_rows, iters, d_cyc, sink);
  CK(cudaDeviceSynchronize());
  std::vector<uint32_t> cyc((size_t)n_rows * 4);
  CK(cudaMemcpy(cyc.data(), d_cyc, cyc.size() * 4, cudaMemcpyDeviceToHost));
  const char *opn[4] = {"a", "b", "c", "dst"};
  for (size_t i = 0; i < 8; i++) {
    std::printf("probe %-34s", probe_name[i]);
    for (int op : {0, 3}) std::printf("  %s %.2f wavefronts", op ? "STS.128" : "LDS.128", cyc[probes[i] * 4 + op] / (double)(iters * 32));
    std::printf("\n");
  }
  const uint32_t real = n_rows - 8;
  for (int op = 0; op < 4; op++) {
    double meas = 0, pq = 0, pp = 0;
    std::map<int, int> hist;
    int shown = 0;
    for (uint32_t r = 0; r < real; r++) {
      const uint16_t *s = &tab[(r * 4 + op) * 32];
      const double m = cyc[r * 4 + op] / (double)(iters * 32);
      const int q = predict(s, true), p = predict(s, false);
      if (q == 0) continue;
      meas += m, pq += q, pp += p;
      const int d = (int)(m - q + 0.5);
      hist[d]++;
      if (d != 0 && shown < 6) {
        shown++;
        std::printf("  row %u op %s: measured %.2f, quarter model %d; slot %% 8 by lane:", r, opn[op], m, q);
        for (int l = 0; l < 32; l++) std::printf("%s%c", l % 8 ? "" : " ", s[l] == 0xFFFF ? '-' : '0' + (s[l] & 7));
        std::printf("\n");
      }
    }
    std::printf("replay %-3s (%s): measured %.0f wavefronts, quarter-warp model %.0f (x%.3f), with pair merging %.0f; measured - model histogram:", opn[op],
                op < 3 ? "LDS.128" : "STS.128", meas, pq, meas / pq, pp);
    for (auto &kv : hist) std::printf(" %+d:%d", kv.first, kv.second);
    std::printf("\n");
  }
  return 0;
}
