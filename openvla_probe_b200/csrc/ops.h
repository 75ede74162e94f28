// Library-internal launch functions (one per kernel family). All return 0 / -1 (set_error).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "gemm.cuh"

namespace ovla {

int num_sms();

// gemm.cu
int gemm_launch(const void* A, long long lda, const void* W, long long ldw, int M, int N, int K, int mode, int kind,
                const GemmEpi& epi, int bn, int cg, int num_sms, cudaStream_t stream);

}  // namespace ovla
