// Probe-training kernels (experiment_utils/train_object_probes.py:177-206, train_spatial_probes.py:153-176,
// train_dual_head_final.py:147-232): everything around the two tcgen05 TF32 GEMMs of a step
//   logits = X_b W^T + b            (gemm, kind::tf32, fp32 out)
//   dW     = dZ^T X_b               (gemm on the transposed operands produced here)
// i.e. epoch permutation gather (row-major + transposed copies of the resident features), fused
// BCE-with-logits loss + gradient written transposed, bias-gradient row sums, and the AdamW update.
// All of it is HBM-bound fp32 streaming with coalesced 128-byte rows.
#include "host_util.h"
#include "ops.h"
#include "ptx.cuh"

#include <math.h>

namespace ovla {

// ------------------------------------------------------------------------------------------- epoch gather
// Xp[i, :] = X[perm[i], :]  and  XpT[:, i] = X[perm[i], :]   (X fp32 [N, D]; XpT [D, ldt])
// 32 x 32 tiles through shared memory so that both the row-major and the transposed store are coalesced.
__global__ void __launch_bounds__(256) gather_transpose_kernel(const float* __restrict__ X, long long ldx,
                                                               const long long* __restrict__ perm, int n, int D,
                                                               float* __restrict__ Xp, long long ldp,
                                                               float* __restrict__ XpT, long long ldt) {
  __shared__ float tile[32][33];
  const int i0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 8 row groups
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int i = i0 + r, c = c0 + tx;
    float v = 0.f;
    if (i < n && c < D) {
      v = X[perm[i] * ldx + c];
      Xp[static_cast<long long>(i) * ldp + c] = v;
    }
    tile[r][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r, i = i0 + tx;
    if (c < D && i < n) XpT[static_cast<long long>(c) * ldt + i] = tile[tx][r];
  }
}

int probe_gather_launch(const float* X, long long ldx, const long long* perm, int n, int D, float* Xp, long long ldp,
                        float* XpT, long long ldt, cudaStream_t st) {
  if (n <= 0) return 0;
  dim3 grid((D + 31) / 32, (n + 31) / 32);
  ProfScope prof(kCatOther, 0.0, 12.0 * n * D, st);
  gather_transpose_kernel<<<grid, 256, 0, st>>>(X, ldx, perm, n, D, Xp, ldp, XpT, ldt);
  CUDA_TRY(cudaGetLastError());
  count_launch();
  return 0;
}

// Yp[i, k] = Y[perm[i], keep[k]] for k < K, and -1 (ignored) in the padding columns [K, Kpad)
__global__ void gather_labels_kernel(const signed char* __restrict__ Y, long long ldy,
                                     const long long* __restrict__ perm, const int* __restrict__ keep, int n, int K,
                                     int Kpad, signed char* __restrict__ Yp) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(n) * Kpad) return;
  const int k = static_cast<int>(idx % Kpad);
  const long long i = idx / Kpad;
  Yp[idx] = (k < K) ? Y[perm[i] * ldy + keep[k]] : static_cast<signed char>(-1);
}

int probe_gather_labels_launch(const signed char* Y, long long ldy, const long long* perm, const int* keep, int n,
                               int K, int Kpad, signed char* Yp, cudaStream_t st) {
  const long long total = static_cast<long long>(n) * Kpad;
  if (total <= 0) return 0;
  gather_labels_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(Y, ldy, perm, keep, n, K, Kpad, Yp);
  CUDA_TRY(cudaGetLastError());
  count_launch();
  return 0;
}

// ------------------------------------------------------------------------------------------- BCE loss + gradient
// torch.nn.BCEWithLogitsLoss(pos_weight) element:  l = (1-t) z + lw * softplus(-z),  lw = 1 + (pw-1) t
//                                      gradient:  dl/dz = (1-t) - lw * (1 - sigmoid(z)) = sigmoid(z) lw - pw t
// head 0 (rows [0, Kpad) of dZT) -- kind:
//   0 "object" : target = (y==1), valid = (y!=-1), vector pos_weight, normaliser = sum(valid)      (train_object_probes.py:184-188)
//   1 "spatial": target = y (0/1), every real column valid, vector pos_weight, normaliser = B*K     (train_spatial_probes.py:153-163)
//   2 "presence": target = (y!=-1), every real column valid, scalar pos_weight, normaliser = B*K    (train_dual_head_final.py:158,180)
// head 1 (rows [Kpad, 2 Kpad), dual-head only) -- "truth": target = (y==1), valid = (y!=-1), no pos_weight,
//   normaliser = sum(valid)                                                                            (train_dual_head_final.py:183-186)
// Gradients are written UN-normalised (the normalisers are global sums that an allreduce completes); stats gets
// [loss_h0, count_h0, loss_h1, count_h1] accumulated with one atomic per CTA.
__global__ void __launch_bounds__(256) bce_grad_kernel(const float* __restrict__ Z, long long ldz,
                                                       const signed char* __restrict__ Y, int n, int K, int Kpad,
                                                       int kind0, int heads, const float* __restrict__ pos_weight,
                                                       float pos_weight_scalar, float* __restrict__ dZT, long long ldt,
                                                       float* __restrict__ stats) {
  __shared__ float tile[2][32][33];
  __shared__ float red[4][8];
  const int i0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int i = i0 + r, k = k0 + tx;
    float g0 = 0.f, g1 = 0.f;
    if (i < n && k < K) {
      const int y = Y[static_cast<long long>(i) * Kpad + k];
      {
        const float z = Z[static_cast<long long>(i) * ldz + k];
        float t, valid, pw;
        if (kind0 == 0) { t = (y == 1); valid = (y != -1); pw = pos_weight[k]; }
        else if (kind0 == 1) { t = static_cast<float>(y); valid = 1.f; pw = pos_weight[k]; }
        else { t = (y != -1); valid = 1.f; pw = pos_weight_scalar; }
        const float lw = 1.f + (pw - 1.f) * t;
        const float sp = log1pf(expf(-fabsf(z))) + fmaxf(-z, 0.f);  // softplus(-z), as torch computes it
        const float sig = 1.f / (1.f + expf(-z));
        acc[0] += valid * ((1.f - t) * z + lw * sp);
        acc[1] += valid;
        g0 = valid * (sig * lw - pw * t);
      }
      if (heads == 2) {
        const float z = Z[static_cast<long long>(i) * ldz + Kpad + k];
        const float t = (y == 1), valid = (y != -1);
        const float sp = log1pf(expf(-fabsf(z))) + fmaxf(-z, 0.f);
        const float sig = 1.f / (1.f + expf(-z));
        acc[2] += valid * ((1.f - t) * z + sp);
        acc[3] += valid;
        g1 = valid * (sig - t);
      }
    }
    tile[0][r][tx] = g0;
    tile[1][r][tx] = g1;
  }
  __syncthreads();
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int k = k0 + r, i = i0 + tx;
    if (k < Kpad && i < n) {  // padding label rows get zero gradients
      dZT[static_cast<long long>(k) * ldt + i] = tile[0][tx][r];
      if (heads == 2) dZT[static_cast<long long>(Kpad + k) * ldt + i] = tile[1][tx][r];
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
    if (tx == 0) red[j][ty] = acc[j];
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[threadIdx.x][w];
    if (s != 0.f) atomicAdd(stats + threadIdx.x, s);
  }
}

int probe_bce_grad_launch(const float* Z, long long ldz, const signed char* Y, int n, int K, int Kpad, int kind0,
                          int heads, const float* pos_weight, float pos_weight_scalar, float* dZT, long long ldt,
                          float* stats, cudaStream_t st) {
  if (n <= 0) return 0;
  if (heads != 1 && heads != 2) return set_error("probe: heads must be 1 or 2");
  if (kind0 < 0 || kind0 > 2) return set_error("probe: unknown loss kind %d", kind0);
  if (kind0 != 2 && !pos_weight) return set_error("probe: vector pos_weight required");
  dim3 grid((Kpad + 31) / 32, (n + 31) / 32);
  ProfScope prof(kCatOther, 0.0, (8.0 * heads + 1.0) * n * Kpad, st);
  bce_grad_kernel<<<grid, 256, 0, st>>>(Z, ldz, Y, n, K, Kpad, kind0, heads, pos_weight, pos_weight_scalar, dZT, ldt,
                                        stats);
  CUDA_TRY(cudaGetLastError());
  count_launch();
  return 0;
}

// ------------------------------------------------------------------------------------------- grouped, fused step
// The same loss / gradient for `groups` probes at once (one per captured layer: same labels, different logits), with
// the bias gradient (row sums of dZ^T) and the loss statistics produced by the same kernel, DETERMINISTICALLY:
// grid = (k-tiles of 32 labels, i-splits, groups); a CTA walks its share of the batch in [128 x 32] tiles, keeps
// per-label column sums and the four loss statistics in registers, and writes them as partials; the last CTA of a
// group to finish (ticket counter) adds the partials in a fixed order into db[groups][rows] and stats[groups][4].
// Nothing is accumulated with floating-point atomics, so a step is bit-reproducible.
//   Z    [groups][n][ldz]  (z_gs apart)      dZT [groups][heads*Kpad][ldt]  (dzt_gs apart)
//   db   = out_base + g*out_gs + db_off      stats = out_base + g*out_gs + stats_off   (the flat [dW | db | stats] buffer)
//   part [groups][isplits][heads*Kpad + 4*ktiles] floats, ticket [groups] ints (zero on entry, left zero)
// sp = softplus(-z) = log(1 + exp(-z)) and sig = sigmoid(z) from ONE exponential e = exp(-|z|) in (0, 1]:
//   sp = log(1 + e) + max(-z, 0),  sig = z >= 0 ? 1 / (1 + e) : e / (1 + e).
// ex2 / lg2 / rcp run on the SFU; 1 + e lies in (1, 2], where lg2.approx is accurate to ~1e-7 absolute, which is the
// error that matters for a loss that is a SUM of such terms (tests: loss 1e-3 relative, gradients at fp32 round-off).
__device__ __forceinline__ void softplus_neg_sigmoid(float z, float& sp, float& sig) {
  const float e = __expf(-fabsf(z));
  const float d = 1.f + e;
  const float r = __frcp_rn(d);
  sp = __logf(d) + fmaxf(-z, 0.f);
  sig = z >= 0.f ? r : e * r;
}

template <int KIND0, int HEADS>
__global__ void __launch_bounds__(256, 3) bce_grad_grouped_kernel(const float* __restrict__ Z, long long ldz, long long z_gs,
                                                                  const signed char* __restrict__ Y, int n, int K, int Kpad,
                                                                  const float* __restrict__ pos_weight,
                                                                  float pos_weight_scalar, float* __restrict__ dZT,
                                                                  long long ldt, long long dzt_gs, float* __restrict__ out_base,
                                                                  long long out_gs, long long db_off, long long stats_off,
                                                                  float* __restrict__ part, int* __restrict__ ticket) {
  constexpr int TI = 128;
  constexpr int kind0 = KIND0, heads = HEADS;
  __shared__ float tile[HEADS][TI][33];
  __shared__ float red[8][8][32];   // [value: db0, db1, stats 0..3 -> rows 2..5][warp][lane]
  __shared__ int s_last;
  const int ktiles = gridDim.x, isplits = gridDim.y, grp = blockIdx.z;
  const int k0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int rows = heads * Kpad;
  Z += grp * z_gs;
  dZT += grp * dzt_gs;
  const int n_it = (n + TI - 1) / TI;
  const int it0 = static_cast<int>(1LL * n_it * blockIdx.y / isplits), it1 = static_cast<int>(1LL * n_it * (blockIdx.y + 1) / isplits);
  const int k = k0 + tx;
  const bool k_ok = k < K;
  const float pw = (kind0 == 2) ? pos_weight_scalar : (k_ok ? pos_weight[k] : 1.f);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  float db0 = 0.f, db1 = 0.f;
  // store phase: thread -> (label row r0 + 2j, sample column sc) of the [32 x 128] transposed tile
  const int sr0 = threadIdx.x >> 7, sc = threadIdx.x & 127;
  const long long ldz8 = 8 * ldz, ldt2 = 2 * ldt;
  for (int it = it0; it < it1; ++it) {
    const int i0 = it * TI;
    const bool full = i0 + TI <= n;
    // all loads of the tile first (16 independent rows per thread), then the math
    int yv[16];
    float z0[16], z1[16];
    {
      const float* zp = Z + static_cast<long long>(i0 + ty) * ldz + k;
      const signed char* yp = Y + static_cast<long long>(i0 + ty) * Kpad + k;
      const int ystep = 8 * Kpad;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const bool ok = k_ok && (full || i0 + ty + 8 * j < n);
        yv[j] = ok ? static_cast<int>(*yp) : -2;
        z0[j] = ok ? __ldcs(zp) : 0.f;
        if (HEADS == 2) z1[j] = ok ? __ldcs(zp + Kpad) : 0.f;
        zp += ldz8;
        yp += ystep;
      }
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float g0 = 0.f, g1 = 0.f;
      const int y = yv[j];
      if (y != -2) {
        {
          const float z = z0[j];
          float t, valid;
          if (kind0 == 0) { t = (y == 1); valid = (y != -1); }
          else if (kind0 == 1) { t = static_cast<float>(y); valid = 1.f; }
          else { t = (y != -1); valid = 1.f; }
          const float lw = 1.f + (pw - 1.f) * t;
          float sp, sig;
          softplus_neg_sigmoid(z, sp, sig);
          acc[0] += valid * ((1.f - t) * z + lw * sp);
          acc[1] += valid;
          g0 = valid * (sig * lw - pw * t);
        }
        if (HEADS == 2) {
          const float z = z1[j];
          const float t = (y == 1), valid = (y != -1);
          float sp, sig;
          softplus_neg_sigmoid(z, sp, sig);
          acc[2] += valid * ((1.f - t) * z + sp);
          acc[3] += valid;
          g1 = valid * (sig - t);
        }
      }
      db0 += g0;
      db1 += g1;
      tile[0][ty + 8 * j][tx] = g0;
      if (HEADS == 2) tile[HEADS - 1][ty + 8 * j][tx] = g1;
    }
    __syncthreads();
    // transposed store: label row k0 + r, 128 consecutive samples (4 lanes-of-32 segments per row)
    if (i0 + sc < n) {
      float* dp = dZT + static_cast<long long>(k0 + sr0) * ldt + i0 + sc;
      const int r_end = min(32, Kpad - k0);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int r = sr0 + 2 * j;
        if (r < r_end) {
          __stcs(dp, tile[0][sc][r]);
          if (HEADS == 2) __stcs(dp + static_cast<long long>(Kpad) * ldt, tile[HEADS - 1][sc][r]);
        }
        dp += ldt2;
      }
    }
    __syncthreads();
  }
  // per-CTA partials, fixed order: db over the 8 row groups (warps), stats over lanes then warps
  red[0][ty][tx] = db0;
  red[1][ty][tx] = db1;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float v = acc[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (tx == 0) red[2 + j][ty][0] = v;
  }
  __syncthreads();
  float* my = part + (static_cast<long long>(grp) * isplits + blockIdx.y) * (rows + 4 * ktiles);
  if (threadIdx.x < 64) {
    const int h = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (h < heads && k0 + l < Kpad) {
      float sum = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) sum += red[h][w][l];
      my[h * Kpad + k0 + l] = sum;
    }
  } else if (threadIdx.x < 68) {
    const int j = threadIdx.x - 64;
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += red[2 + j][w][0];
    my[rows + 4 * blockIdx.x + j] = sum;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int total = ktiles * isplits;
    s_last = (atomicAdd(ticket + grp, 1) == total - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const float* gp = part + static_cast<long long>(grp) * isplits * (rows + 4 * ktiles);
  float* db = out_base + grp * out_gs + db_off;
  float* stats = out_base + grp * out_gs + stats_off;
  for (int r = threadIdx.x; r < rows; r += 256) {
    float sum = 0.f;
    for (int sidx = 0; sidx < isplits; ++sidx) sum += __ldcg(gp + static_cast<long long>(sidx) * (rows + 4 * ktiles) + r);
    db[r] = sum;
  }
  if (threadIdx.x < 4) {
    float sum = 0.f;
    for (int sidx = 0; sidx < isplits; ++sidx)
      for (int kt = 0; kt < ktiles; ++kt)
        sum += __ldcg(gp + static_cast<long long>(sidx) * (rows + 4 * ktiles) + rows + 4 * kt + threadIdx.x);
    stats[threadIdx.x] = sum;
  }
  if (threadIdx.x == 0) ticket[grp] = 0;
}

int probe_bce_grad_grouped_launch(const float* Z, long long ldz, long long z_gs, const signed char* Y, int n, int K,
                                  int Kpad, int kind0, int heads, const float* pos_weight, float pos_weight_scalar,
                                  float* dZT, long long ldt, long long dzt_gs, int groups, float* out_base,
                                  long long out_gs, long long db_off, long long stats_off, float* part, int isplits,
                                  int* ticket, cudaStream_t st) {
  if (n <= 0 || groups <= 0) return 0;
  if (heads != 1 && heads != 2) return set_error("probe: heads must be 1 or 2");
  if (kind0 < 0 || kind0 > 2) return set_error("probe: unknown loss kind %d", kind0);
  if (kind0 != 2 && !pos_weight) return set_error("probe: vector pos_weight required");
  if (isplits < 1 || isplits > 64) return set_error("probe: isplits %d out of range [1, 64]", isplits);
  if (!part || !ticket || !out_base) return set_error("probe: null workspace");
  dim3 grid((Kpad + 31) / 32, isplits, groups);
  ProfScope prof(kCatOther, 0.0, (8.0 * heads + 1.0) * n * Kpad * groups, st);
#define OVLA_BCE_GROUPED(KD, HD)                                                                                        \
  bce_grad_grouped_kernel<KD, HD><<<grid, 256, 0, st>>>(Z, ldz, z_gs, Y, n, K, Kpad, pos_weight, pos_weight_scalar, dZT, ldt, \
                                                        dzt_gs, out_base, out_gs, db_off, stats_off, part, ticket)
  if (heads == 2) {
    if (kind0 != 2) return set_error("probe: the two-head form is the dual probe (kind0 = 2)");
    OVLA_BCE_GROUPED(2, 2);
  } else if (kind0 == 0) OVLA_BCE_GROUPED(0, 1);
  else if (kind0 == 1) OVLA_BCE_GROUPED(1, 1);
  else OVLA_BCE_GROUPED(2, 1);
#undef OVLA_BCE_GROUPED
  CUDA_TRY(cudaGetLastError());
  count_launch();
  return 0;
}

// db[r] = sum_i dZT[r, i]  (one warp per row)
__global__ void rowsum_kernel(const float* __restrict__ A, long long lda, int rows, int cols, float* __restrict__ out) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  float s = 0.f;
  for (int c = lane; c < cols; c += 32) s += A[static_cast<long long>(r) * lda + c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[r] = s;
}

int probe_rowsum_launch(const float* A, long long lda, int rows, int cols, float* out, cudaStream_t st) {
  if (rows <= 0) return 0;
  rowsum_kernel<<<(rows + 7) / 8, 256, 0, st>>>(A, lda, rows, cols, out);
  CUDA_TRY(cudaGetLastError());
  count_launch();
  return 0;
}

// ------------------------------------------------------------------------------------------- AdamW
// torch.optim.AdamW (decoupled weight decay).  params = [W (rows x D) | b (rows)] in one flat buffer, grads likewise
// (un-normalised); row r of head h is divided by stats[2h+1], the global number of loss terms of that head -- read from
// device memory, so no host sync separates loss and update.
// Grouped form: `groups` probes (blockIdx.y) with parameters n_total apart and gradients / statistics g_gs / stats_gs
// apart.  One thread updates 4 consecutive parameters of W with 16-byte accesses (D % 4 == 0, so the 4 share a row);
// the bias tail (and everything, when a base pointer is not 16-byte aligned) goes through the scalar path.
__device__ __forceinline__ float adamw_one(float p, float grad, float& m, float& v, float lr, float beta1, float beta2,
                                           float eps, float wd, float bc1, float bc2_sqrt) {
  p *= (1.f - lr * wd);
  m = beta1 * m + (1.f - beta1) * grad;
  v = beta2 * v + (1.f - beta2) * grad * grad;
  return p - (lr / bc1) * m / (sqrtf(v) / bc2_sqrt + eps);
}

__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p_all, const float* __restrict__ g_all,
                                                    float* __restrict__ m_all, float* __restrict__ v_all, long long n_w, int D,
                                                    int rows_per_head, long long n_total, const float* __restrict__ stats_all,
                                                    float lr, float beta1, float beta2, float eps, float wd, float bc1,
                                                    float bc2_sqrt, long long g_gs, long long stats_gs, int vec) {
  const long long grp = blockIdx.y;
  float* p = p_all + grp * n_total;
  float* m = m_all + grp * n_total;
  float* v = v_all + grp * n_total;
  const float* g = g_all + grp * g_gs;
  const float* stats = stats_all + grp * stats_gs;
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long n_vec = vec ? n_w / 4 : 0;                     // float4 items of the W block
  if (t < n_vec) {
    const long long j = t * 4;
    const int head = static_cast<int>((j / D) / rows_per_head);
    const float denom = stats[2 * head + 1];
    float4 pp = *reinterpret_cast<const float4*>(p + j), gg = __ldcs(reinterpret_cast<const float4*>(g + j));
    float4 mm = *reinterpret_cast<const float4*>(m + j), vv = *reinterpret_cast<const float4*>(v + j);
    pp.x = adamw_one(pp.x, denom > 0.f ? gg.x / denom : 0.f, mm.x, vv.x, lr, beta1, beta2, eps, wd, bc1, bc2_sqrt);
    pp.y = adamw_one(pp.y, denom > 0.f ? gg.y / denom : 0.f, mm.y, vv.y, lr, beta1, beta2, eps, wd, bc1, bc2_sqrt);
    pp.z = adamw_one(pp.z, denom > 0.f ? gg.z / denom : 0.f, mm.z, vv.z, lr, beta1, beta2, eps, wd, bc1, bc2_sqrt);
    pp.w = adamw_one(pp.w, denom > 0.f ? gg.w / denom : 0.f, mm.w, vv.w, lr, beta1, beta2, eps, wd, bc1, bc2_sqrt);
    *reinterpret_cast<float4*>(p + j) = pp;
    *reinterpret_cast<float4*>(m + j) = mm;
    *reinterpret_cast<float4*>(v + j) = vv;
    return;
  }
  const long long i = n_vec * 4 + (t - n_vec);                  // scalar tail
  if (i >= n_total) return;
  const long long row = (i < n_w) ? i / D : i - n_w;
  const int head = static_cast<int>(row / rows_per_head);
  const float denom = stats[2 * head + 1];  // global count of the head's loss terms (after the allreduce)
  const float grad = denom > 0.f ? g[i] / denom : 0.f;
  float mi = m[i], vi = v[i];
  p[i] = adamw_one(p[i], grad, mi, vi, lr, beta1, beta2, eps, wd, bc1, bc2_sqrt);
  m[i] = mi;
  v[i] = vi;
}

int probe_adamw_launch(float* p, const float* g, float* m, float* v, long long n_w, int D, int rows_per_head,
                       long long n_total, const float* stats, float lr, float beta1, float beta2, float eps, float wd,
                       int step, cudaStream_t st, int groups, long long g_gs, long long stats_gs) {
  if (n_total <= 0 || groups <= 0) return 0;
  if (step < 1) return set_error("adamw: step must be >= 1");
  const float bc1 = static_cast<float>(1.0 - pow(static_cast<double>(beta1), step));
  const double bc2 = 1.0 - pow(static_cast<double>(beta2), step);
  ProfScope prof(kCatOther, 0.0, 28.0 * n_total * groups, st);
  // 16-byte path for the W block when every base and group stride keeps float4 alignment
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  const int vec = (D % 4 == 0 && n_w % 4 == 0 && al16(p) && al16(g) && al16(m) && al16(v) &&
                   (groups == 1 || (n_total % 4 == 0 && g_gs % 4 == 0))) ? 1 : 0;
  const long long items = (vec ? n_w / 4 : 0) + (n_total - (vec ? n_w : 0));
  if (groups > 65535) return set_error("adamw: too many groups (%d)", groups);
  dim3 grid(static_cast<unsigned>((items + 255) / 256), static_cast<unsigned>(groups));
  adamw_kernel<<<grid, 256, 0, st>>>(p, g, m, v, n_w, D, rows_per_head, n_total, stats, lr, beta1, beta2, eps, wd, bc1,
                                     static_cast<float>(sqrt(bc2)), g_gs, stats_gs, vec);
  CUDA_TRY(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace ovla
