// Direct 3-class probe (experiment_utils/train_3class_direct.py:147-212): nn.Linear(D, 3K), logits viewed as
// [B*K, 3], class-weighted CrossEntropyLoss over {N/A (-1), False (0), True (1)} -> class index {0, 1, 2}.
//   loss = sum_i w[t_i] * (-log softmax(z_i)[t_i]) / sum_i w[t_i]         (torch weighted-mean reduction)
//   dz_ic = w[t_i] * (softmax(z_i)_c - [c == t_i])                        (un-normalised; divide by sum_i w[t_i])
// Gradient rows are written transposed (dZT [3K padded, n]) for the dW GEMM, like the BCE kernels in probe.cu.
#include <algorithm>

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

// ------------------------------------------------------------------------------------------------ validation counters
// On-device confusion counts for the probe validation metrics (train_object_probes.py:190-206, train_spatial_probes.py,
// train_dual_head_final.py:196-232, train_3class_direct.py:196-207): the reference gathers logits and labels to the host
// and calls sklearn; accuracy and F1 only need these integer counts, which are exact and order-independent.
//   kind 0 (object)  mask y != -1, target y == 1, pred sigmoid(z) > thresh        counts[0..3]  = tp, fp, fn, tn
//   kind 1 (spatial) all elements, target y (0/1)                                   counts[0..3]
//   kind 2 (dual)    presence head z[:, k] vs (y != -1), all elements               counts[0..3]
//                    truth head z[:, Kpad + k] vs (y == 1) where y != -1            counts[4..7]
//   kind 3 (3-class) argmax(z[:, 3k..3k+2]) (first maximum) vs y + 1                counts[3*target + pred], 9 entries
// Y is the label matrix as stored ([n, *] int8), `keep` selects its K kept columns (nullptr: identity).
__global__ void probe_confusion_kernel(const float* __restrict__ Z, long long ldz, const signed char* __restrict__ Y,
                                       long long ldy, const int* __restrict__ keep, int n, int K, int Kpad, int kind,
                                       float thresh, unsigned long long* __restrict__ counts) {
  unsigned int c[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  const long long total = static_cast<long long>(n) * K;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / K;
    const int k = static_cast<int>(i - r * K);
    const int y = Y[r * ldy + (keep ? keep[k] : k)];
    if (kind == 3) {
      const float* z = Z + r * ldz + 3 * k;
      int pred = 0;
      float best = z[0];
      if (z[1] > best) { best = z[1]; pred = 1; }
      if (z[2] > best) { pred = 2; }
      c[3 * (y + 1) + pred]++;
      continue;
    }
    auto sig = [](float z) { return __fdiv_rn(1.f, __fadd_rn(1.f, expf(-z))); };
    if (kind == 2) {
      const int pt = (y != -1), pp = sig(Z[r * ldz + k]) > 0.5f;
      c[pt ? (pp ? 0 : 2) : (pp ? 1 : 3)]++;
      if (y != -1) {
        const int tt = (y == 1), tp = sig(Z[r * ldz + Kpad + k]) > 0.5f;
        c[4 + (tt ? (tp ? 0 : 2) : (tp ? 1 : 3))]++;
      }
      continue;
    }
    if (kind == 0 && y == -1) continue;
    const int t = (kind == 0) ? (y == 1) : (y != 0), pr = sig(Z[r * ldz + k]) > thresh;
    c[t ? (pr ? 0 : 2) : (pr ? 1 : 3)]++;
  }
#pragma unroll
  for (int j = 0; j < 9; ++j) {
    unsigned int v = c[j];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(counts + j, static_cast<unsigned long long>(v));
  }
}

int probe_confusion_launch(const float* Z, long long ldz, const signed char* Y, long long ldy, const int* keep, int n,
                           int K, int Kpad, int kind, float thresh, unsigned long long* counts, cudaStream_t st) {
  if (kind < 0 || kind > 3) return set_error("probe confusion: unknown kind %d", kind);
  CUDA_TRY(cudaMemsetAsync(counts, 0, 9 * sizeof(unsigned long long), st));
  const long long total = static_cast<long long>(n) * K;
  if (total <= 0) return 0;
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148LL * 8));
  probe_confusion_kernel<<<blocks, 256, 0, st>>>(Z, ldz, Y, ldy, keep, n, K, Kpad, kind, thresh, counts);
  CUDA_TRY(cudaGetLastError());
  count_launch();
  return 0;
}

// Per-label confusion counts for the reference's per-label evaluation (experiment_utils/eval_probes_per_label.py:59-96:
// for each kept label, mask y != -1, target y == 1, pred sigmoid(z) > 0.5, then precision / recall / F1 / MCC / balanced
// accuracy -- all functions of these four integers).  One block per label; counts[k*4 + {tp, fp, fn, tn}].
__global__ void __launch_bounds__(256) probe_confusion_per_label_kernel(
    const float* __restrict__ Z, long long ldz, const signed char* __restrict__ Y, long long ldy,
    const int* __restrict__ keep, int n, int K, float thresh, unsigned long long* __restrict__ counts) {
  const int k = blockIdx.x;
  if (k >= K) return;
  const int col = keep ? keep[k] : k;
  unsigned int c[4] = {0, 0, 0, 0};
  for (int r = threadIdx.x; r < n; r += blockDim.x) {
    const int y = Y[static_cast<long long>(r) * ldy + col];
    if (y == -1) continue;
    const float z = Z[static_cast<long long>(r) * ldz + k];
    const int t = (y == 1), pr = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-z))) > thresh;
    c[t ? (pr ? 0 : 2) : (pr ? 1 : 3)]++;
  }
  __shared__ unsigned int sm[4][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    unsigned int v = c[j];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) sm[j][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    unsigned long long s2 = 0;
    for (int w = 0; w < 8; ++w) s2 += sm[threadIdx.x][w];
    counts[static_cast<long long>(k) * 4 + threadIdx.x] = s2;
  }
}

int probe_confusion_per_label_launch(const float* Z, long long ldz, const signed char* Y, long long ldy, const int* keep,
                                     int n, int K, float thresh, unsigned long long* counts, cudaStream_t st) {
  if (K <= 0) return 0;
  probe_confusion_per_label_kernel<<<K, 256, 0, st>>>(Z, ldz, Y, ldy, keep, n, K, thresh, counts);
  CUDA_TRY(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace ovla
