// Direct 3-class probe (experiment_utils/train_3class_direct.py:147-212): nn.Linear(D, 3K), logits viewed as
// [B*K, 3], class-weighted CrossEntropyLoss over {N/A (-1), False (0), True (1)} -> class index {0, 1, 2}.
//   loss = sum_i w[t_i] * (-log softmax(z_i)[t_i]) / sum_i w[t_i]         (torch weighted-mean reduction)
//   dz_ic = w[t_i] * (softmax(z_i)_c - [c == t_i])                        (un-normalised; divide by sum_i w[t_i])
// Gradient rows are written transposed (dZT [3K padded, n]) for the dW GEMM, like the BCE kernels in probe.cu.
#include "host_util.h"
#include "ops.h"
#include "ptx.cuh"

namespace ovla {

__global__ void __launch_bounds__(256) ce3_grad_kernel(const float* __restrict__ Z, long long ldz,
                                                       const signed char* __restrict__ Y, long long ldy, int n, int K,
                                                       int rows_pad, float w0, float w1, float w2,
                                                       float* __restrict__ dZT, long long ldt,
                                                       float* __restrict__ stats) {
  __shared__ float tile[3][32][33];
  __shared__ float red[2][8];
  const int i0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  float acc_l = 0.f, acc_w = 0.f;
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int i = i0 + r, k = k0 + tx;
    float g[3] = {0.f, 0.f, 0.f};
    if (i < n && k < K) {
      const int y = Y[static_cast<long long>(i) * ldy + k];
      const int t = y + 1;  // -1 -> 0 (N/A), 0 -> 1 (False), 1 -> 2 (True)
      const float* z = Z + static_cast<long long>(i) * ldz + 3 * k;
      const float z0 = z[0], z1 = z[1], z2 = z[2];
      const float m = fmaxf(z0, fmaxf(z1, z2));
      const float e0 = expf(z0 - m), e1 = expf(z1 - m), e2 = expf(z2 - m);
      const float s = e0 + e1 + e2;
      const float w = t == 0 ? w0 : (t == 1 ? w1 : w2);
      const float zt = t == 0 ? z0 : (t == 1 ? z1 : z2);
      acc_l += w * (logf(s) + m - zt);
      acc_w += w;
      g[0] = w * (e0 / s - (t == 0));
      g[1] = w * (e1 / s - (t == 1));
      g[2] = w * (e2 / s - (t == 2));
    }
    tile[0][r][tx] = g[0];
    tile[1][r][tx] = g[1];
    tile[2][r][tx] = g[2];
  }
  __syncthreads();
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int k = k0 + r, i = i0 + tx;
    if (i < n) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int row = 3 * k + c;
        if (row < rows_pad) dZT[static_cast<long long>(row) * ldt + i] = (k < K) ? tile[c][tx][r] : 0.f;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    acc_l += __shfl_xor_sync(0xffffffffu, acc_l, o);
    acc_w += __shfl_xor_sync(0xffffffffu, acc_w, o);
  }
  if (tx == 0) { red[0][ty] = acc_l; red[1][ty] = acc_w; }
  __syncthreads();
  if (threadIdx.x < 2) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[threadIdx.x][w];
    if (s != 0.f) atomicAdd(stats + threadIdx.x, s);
  }
}

// Z fp32 [n, >= 3K] (pitch ldz), Y int8 [n, *] kept columns (pitch ldy); dZT [rows_pad, n] (pitch ldt);
// stats[0] += sum w*nll, stats[1] += sum w.
int probe_ce3_grad_launch(const float* Z, long long ldz, const signed char* Y, long long ldy, int n, int K,
                          int rows_pad, const float* class_w3_host, float* dZT, long long ldt, float* stats,
                          cudaStream_t st) {
  if (n <= 0) return 0;
  if (!class_w3_host) return set_error("probe: class weights required");
  if (rows_pad < 3 * K) return set_error("probe: rows_pad must cover 3*K logits");
  dim3 grid(((rows_pad + 2) / 3 + 31) / 32, (n + 31) / 32);
  ProfScope prof(kCatOther, 0.0, 25.0 * n * K, st);
  ce3_grad_kernel<<<grid, 256, 0, st>>>(Z, ldz, Y, ldy, n, K, rows_pad, class_w3_host[0], class_w3_host[1],
                                        class_w3_host[2], dZT, ldt, stats);
  CUDA_TRY(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace ovla
