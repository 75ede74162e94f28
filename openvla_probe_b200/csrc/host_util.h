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
