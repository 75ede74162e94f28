// Host-side helpers shared by the .cu files of libovla_b200: error string, CUDA check, launch counter.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ovla {

// printf-style; stores the message for ovla_last_error() and returns -1
int set_error(const char* fmt, ...);
const char* last_error();
void count_launch(int n = 1);
long long launch_count();
void reset_launch_count();

// Live kernel timing for bench.py's roofline: when enabled, every launch site brackets its kernel with CUDA events
// recorded on the launch stream; prof_collect() synchronises and sums per category.
enum ProfCat : int { kCatGemm = 0, kCatGemv, kCatFlash, kCatDecodeAttn, kCatNorm, kCatPool, kCatOther, kNumCat };
void prof_enable(bool on);
bool prof_enabled();
void prof_begin(int cat, double flops, double bytes, cudaStream_t st);
void prof_end(cudaStream_t st);
// out arrays of kNumCat: launches, milliseconds, flops, bytes. Clears the records.
int prof_collect(long long* launches, double* ms, double* flops, double* bytes);

struct ProfScope {
  cudaStream_t st;
  bool on;
  ProfScope(int cat, double flops, double bytes, cudaStream_t s) : st(s), on(prof_enabled()) {
    if (on) prof_begin(cat, flops, bytes, st);
  }
  ~ProfScope() {
    if (on) prof_end(st);
  }
};

// Programmatic dependent launch (sm_90+): a kernel launched through launch_pdl may begin while its stream predecessor
// is still draining; such kernels call griddep_wait() (ptx.cuh) before touching anything the predecessor wrote.
// OVLA_PDL=0 disables the attribute (plain stream serialisation) for A/B measurements.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

#define CUDA_TRY(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      return ::ovla::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define OVLA_TRY(expr)       \
  do {                       \
    if ((expr) != 0) return -1; \
  } while (0)

}  // namespace ovla
