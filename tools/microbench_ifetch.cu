// How fast does straight-line code run on B200 when it is executed once per warp (no loop, no reuse inside a warp)?
// The specialised kernels of long programs (clg_spec_chunk: 12 k LOP3 per thread, 255 registers, 2 warps per scheduler)
// stop at 57 % of LOP3 issue with `no_instruction` as the top stall. This generates kernels of N independent-enough LOP3
// (16 accumulators) as PTX, assembles them with the driver, and runs them at 2 / 4 / 8 warps per scheduler, many waves
// of CTAs per SM, to separate the cache level the code comes from (N x 16 bytes) from the occupancy that hides the fetch.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mbi tools/microbench_ifetch.cu -lcuda && /tmp/mbi
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include <cuda.h>

#define CU(x)                                                                   \
  do {                                                                          \
    CUresult r_ = (x);                                                          \
    if (r_ != CUDA_SUCCESS) {                                                   \
      const char *s_ = nullptr;                                                 \
      cuGetErrorString(r_, &s_);                                                \
      std::fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, s_ ? s_ : "?"); \
      std::exit(1);                                                             \
    }                                                                           \
  } while (0)

static std::string make_ptx(int n) {
  std::string s = ".version 8.7\n.target sm_100a\n.address_size 64\n.visible .entry k(.param .u64 out, .param .u32 seed)\n{\n.reg .b32 r<20>;\n.reg .b64 a, b;\n";
  s += "mov.u32 r16, %tid.x;\nld.param.u32 r17, [seed];\n";
  for (int i = 0; i < 16; i++) s += "mad.lo.u32 r" + std::to_string(i) + ", r16, " + std::to_string(2654435761u + 97u * i) + ", r17;\n";
  for (int i = 0; i < n; i++) {
    const int d = i % 16;
    s += "lop3.b32 r" + std::to_string(d) + ", r" + std::to_string(d) + ", r" + std::to_string((d + 1) % 16) + ", r" + std::to_string((d + 5) % 16) + ", 0x96;\n";
  }
  s += "xor.b32 r0, r0, r1;\n";
  for (int i = 2; i < 16; i++) s += "xor.b32 r0, r0, r" + std::to_string(i) + ";\n";
  s += "setp.ne.u32 %p1, r0, 0x12345678;\n@%p1 bra DONE;\nld.param.u64 a, [out];\ncvta.to.global.u64 b, a;\nst.global.u32 [b], r0;\nDONE:\nret;\n}\n";
  s.insert(s.find(".reg .b64"), ".reg .pred %p<2>;\n");
  return s;
}

int main() {
  CU(cuInit(0));
  CUdevice dev;
  CU(cuDeviceGet(&dev, 0));
  CUcontext ctx;
  CU(cuDevicePrimaryCtxRetain(&ctx, dev));
  CU(cuCtxSetCurrent(ctx));
  int sms = 0, khz = 0;
  CU(cuDeviceGetAttribute(&sms, CU_DEVICE_ATTRIBUTE_MULTIPROCESSOR_COUNT, dev));
  CU(cuDeviceGetAttribute(&khz, CU_DEVICE_ATTRIBUTE_CLOCK_RATE, dev));
  CUdeviceptr out;
  CU(cuMemAlloc(&out, 4096));
  CUevent e0, e1;
  CU(cuEventCreate(&e0, CU_EVENT_DEFAULT));
  CU(cuEventCreate(&e1, CU_EVENT_DEFAULT));
  std::printf("SMs %d, %d kHz; LOP3 peak = 0.5 warp instructions per clock per scheduler\n", sms, khz);
  for (int n : {512, 2048, 4096, 6144, 8192, 12288, 16384, 32768}) {
    const std::string ptx = make_ptx(n);
    CUmodule mod;
    CU(cuModuleLoadData(&mod, ptx.c_str()));
    CUfunction fn;
    CU(cuModuleGetFunction(&fn, mod, "k"));
    CU(cuFuncSetAttribute(fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, 200 * 1024));
    for (int warps_per_sched : {1, 2, 4, 8}) {
      // one CTA per SM (the dynamic shared memory request keeps a second one out), warps_per_sched * 4 warps in it
      const unsigned threads = warps_per_sched * 4 * 32, smem = 120 * 1024, waves = 24, grid = sms * waves;
      unsigned seed = 1;
      void *args[] = {&out, &seed};
      for (int rep = 0; rep < 2; rep++) {
        CU(cuEventRecord(e0, nullptr));
        CU(cuLaunchKernel(fn, grid, 1, 1, threads, 1, 1, smem, nullptr, args, nullptr));
        CU(cuEventRecord(e1, nullptr));
        CU(cuEventSynchronize(e1));
      }
      float ms = 0;
      CU(cuEventElapsedTime(&ms, e0, e1));
      const double warp_instr = (double)n * waves * warps_per_sched;  // per scheduler
      const double clocks = ms * 1e-3 * khz * 1e3;
      std::printf("N = %6d LOP3 (%4d KB of SASS)  %d warps/scheduler  %.3f ms  %.3f warp instr/clk/scheduler = %.0f %% of LOP3 peak\n", n, n * 16 / 1024,
                  warps_per_sched, ms, warp_instr / clocks, 100.0 * warp_instr / clocks / 0.5);
    }
    CU(cuModuleUnload(mod));
  }
  return 0;
}
