// HBM-bound data-movement kernels of the predict_action path: patch im2col, ViT token assembly, tower concat,
// embedding gather + multimodal splice, RoPE + KV-cache write, hidden-state pooling (capture), argmax,
// de-tokenise / un-normalise.  All use 16-byte accesses along the contiguous dimension.
#include "host_util.h"
#include "ops.h"
#include "ptx.cuh"

namespace ovla {

// ------------------------------------------------------------------------------------------- im2col
// pixel_values [B, C_total, H, W] bf16 -> patches [B*np, Kpad], K index = (c, kh, kw) as Conv2d weight.flatten(1)
// (timm PatchEmbed: Conv2d(3, D, 14, 14) reached through modeling_prismatic.py:121). Columns >= 3*p*p are zero.
__global__ void im2col_kernel(const __nv_bfloat16* __restrict__ px, int c_total, int c0, int H, int Wd, int patch,
                              int Kpad, __nv_bfloat16* __restrict__ out, long long total) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int k = static_cast<int>(idx % Kpad);
  const long long row = idx / Kpad;
  const int gw = Wd / patch, gh = H / patch;
  const int np = gw * gh;
  const int b = static_cast<int>(row / np), p = static_cast<int>(row % np);
  __nv_bfloat16 v = __float2bfloat16(0.f);
  if (k < 3 * patch * patch) {
    const int c = k / (patch * patch), r = k % (patch * patch);
    const int kh = r / patch, kw = r % patch;
    const int y = (p / gw) * patch + kh, x = (p % gw) * patch + kw;
    v = px[((static_cast<long long>(b) * c_total + c0 + c) * H + y) * Wd + x];
  }
  out[idx] = v;
}

int im2col_launch(const void* px, int B, int c_total, int c0, int H, int W, int patch, int Kpad, void* out,
                  cudaStream_t st) {
  const long long total = static_cast<long long>(B) * (H / patch) * (W / patch) * Kpad;
  if (total <= 0) return 0;
  im2col_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(
      static_cast<const __nv_bfloat16*>(px), c_total, c0, H, W, patch, Kpad, static_cast<__nv_bfloat16*>(out), total);
  CUDA_TRY(cudaGetLastError());
  count_launch();
  return 0;
}

// ------------------------------------------------------------------------------------------- ViT token assembly
// tokens[b, n_prefix + p, :] = bf16(patch[b, p, :] + pos[p, :]);  tokens[b, 0] = cls; tokens[b, 1..] = reg
// (timm VisionTransformer._pos_embed with no_embed_class=True: pos-embed is added to the patches only.)
__global__ void assemble_tokens_kernel(const __nv_bfloat16* __restrict__ patch, const __nv_bfloat16* __restrict__ pos,
                                       const __nv_bfloat16* __restrict__ cls, const __nv_bfloat16* __restrict__ reg,
                                       int np, int n_prefix, int D, __nv_bfloat16* __restrict__ tokens,
                                       long long total_vec) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total_vec) return;
  const int dv = D / 8;
  const int c = static_cast<int>(idx % dv) * 8;
  const long long row = idx / dv;
  const int N = np + n_prefix;
  const int b = static_cast<int>(row / N), t = static_cast<int>(row % N);
  uint4 o;
  if (t == 0 && n_prefix) {
    o = *reinterpret_cast<const uint4*>(cls + c);
  } else if (t < n_prefix) {
    o = *reinterpret_cast<const uint4*>(reg + static_cast<long long>(t - 1) * D + c);
  } else {
    const int p = t - n_prefix;
    const uint4 a = *reinterpret_cast<const uint4*>(patch + (static_cast<long long>(b) * np + p) * D + c);
    const uint4 e = *reinterpret_cast<const uint4*>(pos + static_cast<long long>(p) * D + c);
    const uint32_t au[4] = {a.x, a.y, a.z, a.w}, eu[4] = {e.x, e.y, e.z, e.w};
    uint32_t r[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 x = unpack_bf16(au[i]), y = unpack_bf16(eu[i]);
      r[i] = pack_bf16(x.x + y.x, x.y + y.y);
    }
    o = make_uint4(r[0], r[1], r[2], r[3]);
  }
  *reinterpret_cast<uint4*>(tokens + row * D + c) = o;
}

int assemble_tokens_launch(const void* patch, const void* pos, const void* cls, const void* reg, int B, int np,
                           int n_prefix, int D, void* tokens, cudaStream_t st) {
  const long long total = static_cast<long long>(B) * (np + n_prefix) * (D / 8);
  if (total <= 0) return 0;
  assemble_tokens_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(
      static_cast<const __nv_bfloat16*>(patch), static_cast<const __nv_bfloat16*>(pos),
      static_cast<const __nv_bfloat16*>(cls), static_cast<const __nv_bfloat16*>(reg), np, n_prefix, D,
      static_cast<__nv_bfloat16*>(tokens), total);
  CUDA_TRY(cudaGetLastError());
  count_launch();
  return 0;
}

// ------------------------------------------------------------------------------------------- strided 2-D copy
// dst[b, r, dcol0 + c] = src[b, srow0 + r, c]  (prefix strip + channel concat, modeling_prismatic.py:123)
__global__ void copy_rows_kernel(const __nv_bfloat16* __restrict__ src, long long src_batch, long long src_ld,
                                 int srow0, __nv_bfloat16* __restrict__ dst, long long dst_batch, long long dst_ld,
                                 int dcol0, int rows, int cols, long long total_vec) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total_vec) return;
  const int cv = cols / 8;
  const int c = static_cast<int>(idx % cv) * 8;
  const long long rr = idx / cv;
  const int b = static_cast<int>(rr / rows), r = static_cast<int>(rr % rows);
  *reinterpret_cast<uint4*>(dst + b * dst_batch + r * dst_ld + dcol0 + c) =
      *reinterpret_cast<const uint4*>(src + b * src_batch + (srow0 + r) * src_ld + c);
}

int copy_rows_launch(const void* src, long long src_batch, long long src_ld, int srow0, void* dst, long long dst_batch,
                     long long dst_ld, int dcol0, int B, int rows, int cols, cudaStream_t st) {
  if (cols % 8 || dcol0 % 8) return set_error("copy_rows: cols/dcol0 must be multiples of 8");
  const long long total = static_cast<long long>(B) * rows * (cols / 8);
  if (total <= 0) return 0;
  copy_rows_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(
      static_cast<const __nv_bfloat16*>(src), src_batch, src_ld, srow0, static_cast<__nv_bfloat16*>(dst), dst_batch,
      dst_ld, dcol0, rows, cols, total);
  CUDA_TRY(cudaGetLastError());
  count_launch();
  return 0;
}

// ------------------------------------------------------------------------------------------- ragged prompts
// out[b, :] = x[b, last - (P - lens[b]), :]: the last real position of every right-padded row (lm_head input of the
// first generated token).  Also validates the lengths (err flag 2 for a length outside [1, P]).
__global__ void gather_last_rows_kernel(const __nv_bfloat16* __restrict__ x, long long batch_stride, long long ld, int last,
                                        const int* __restrict__ lens, int P, int D, __nv_bfloat16* __restrict__ out,
                                        int* __restrict__ err) {
  const int b = blockIdx.y;
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (c >= D) return;
  int len = lens[b];
  if (len < 1 || len > P) {
    if (c == 0) atomicExch(err, 2);
    len = min(max(len, 1), P);
  }
  const long long r = last - (P - len);
  *reinterpret_cast<uint4*>(out + static_cast<long long>(b) * D + c) =
      *reinterpret_cast<const uint4*>(x + b * batch_stride + r * ld + c);
}

int gather_last_rows_launch(const void* x, long long batch_stride, long long ld, int last, const int* lens, int P, int B,
                            int D, void* out, int* err_flag, cudaStream_t st) {
  if (B <= 0) return 0;
  if (D % 8) return set_error("gather_last_rows: D must be a multiple of 8");
  if (last - (P - 1) < 0) return set_error("gather_last_rows: shortest possible row starts before the buffer");
  dim3 grid((D / 8 + 127) / 128, B);
  gather_last_rows_kernel<<<grid, 128, 0, st>>>(static_cast<const __nv_bfloat16*>(x), batch_stride, ld, last, lens, P, D,
                                               static_cast<__nv_bfloat16*>(out), err_flag);
  CUDA_TRY(cudaGetLastError());
  count_launch();
  return 0;
}

// ------------------------------------------------------------------------------------------- embed + splice
// x[b, 0] = E[ids[b,0]];  x[b, 1..np] = proj[b, :];  x[b, np+j] = E[ids[b,j]], j >= 1
// (modeling_prismatic.py:380-385: [BOS | projected patches | text[1:]])
__global__ void embed_splice_kernel(const long long* __restrict__ ids, int P, const __nv_bfloat16* __restrict__ E,
                                    int vocab, const __nv_bfloat16* __restrict__ proj, int np, int D,
                                    __nv_bfloat16* __restrict__ x, long long total_vec, int* __restrict__ err) {
  griddep_launch_dependents();
  griddep_wait();  // the ids may come from the argmax of the previous decode step
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total_vec) return;
  const int dv = D / 8;
  const int c = static_cast<int>(idx % dv) * 8;
  const long long row = idx / dv;
  const int T = np + P;
  const int b = static_cast<int>(row / T), t = static_cast<int>(row % T);
  const __nv_bfloat16* src;
  if (t >= 1 && t <= np) {
    src = proj + (static_cast<long long>(b) * np + (t - 1)) * D;
  } else {
    const int j = (t == 0) ? 0 : t - np;
    long long id = ids[static_cast<long long>(b) * P + j];
    if (id < 0 || id >= vocab) {
      if (c == 0) atomicExch(err, 1);
      id = 0;
    }
    src = E + id * D;
  }
  *reinterpret_cast<uint4*>(x + row * D + c) = *reinterpret_cast<const uint4*>(src + c);
}

int embed_splice_launch(const void* ids, int B, int P, const void* E, int vocab, const void* proj, int np, int D,
                        void* x, int* err_flag, cudaStream_t st) {
  const long long total = static_cast<long long>(B) * (np + P) * (D / 8);
  if (total <= 0) return 0;
  CUDA_TRY(launch_pdl(embed_splice_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, st,
                      static_cast<const long long*>(ids), P, static_cast<const __nv_bfloat16*>(E), vocab,
                      static_cast<const __nv_bfloat16*>(proj), np, D, static_cast<__nv_bfloat16*>(x), total, err_flag));
  count_launch();
  return 0;
}

// ------------------------------------------------------------------------------------------- RoPE + KV write
// qkv [B*T, 3*D] (q | k | v, each [H, hd]) -> q rotated in place; k rotated and v copied into the layer's
// KV cache [B, H, Tmax, hd] at position pos0 + t.  cos/sin tables [Tmax, hd/2] are bf16 (HF casts the fp32
// cos/sin to the activation dtype), and  q*cos + rotate_half(q)*sin  rounds every product and the sum to bf16
// exactly like the reference's elementwise bf16 ops (transformers modeling_llama.py apply_rotary_pos_emb).
__global__ void rope_kv_kernel(__nv_bfloat16* __restrict__ qkv, int T, int H, int hd, int pos0,
                               const __nv_bfloat16* __restrict__ cos_t, const __nv_bfloat16* __restrict__ sin_t,
                               __nv_bfloat16* __restrict__ kc, __nv_bfloat16* __restrict__ vc, int Tmax,
                               long long total) {
  // one thread per (row, head, 8-element group of the first half)
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int half = hd / 2, gv = half / 8;
  const int g = static_cast<int>(idx % gv) * 8;
  long long r = idx / gv;
  const int h = static_cast<int>(r % H);
  r /= H;  // row = b*T + t
  const int t = static_cast<int>(r % T), b = static_cast<int>(r / T);
  const int pos = pos0 + t;
  const int D = H * hd;
  const uint4 cu = *reinterpret_cast<const uint4*>(cos_t + static_cast<long long>(pos) * half + g);
  const uint4 su = *reinterpret_cast<const uint4*>(sin_t + static_cast<long long>(pos) * half + g);
  const uint32_t cw[4] = {cu.x, cu.y, cu.z, cu.w}, sw[4] = {su.x, su.y, su.z, su.w};
  __nv_bfloat16* base = qkv + r * 3 * D + h * hd;
  const long long cache_off = ((static_cast<long long>(b) * H + h) * Tmax + pos) * hd;
#pragma unroll
  for (int which = 0; which < 2; ++which) {  // 0: q, 1: k
    __nv_bfloat16* p = base + which * D;
    const uint4 lo = *reinterpret_cast<const uint4*>(p + g);
    const uint4 hi = *reinterpret_cast<const uint4*>(p + half + g);
    const uint32_t lw[4] = {lo.x, lo.y, lo.z, lo.w}, hw[4] = {hi.x, hi.y, hi.z, hi.w};
    uint32_t ol[4], oh[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 x1 = unpack_bf16(lw[i]), x2 = unpack_bf16(hw[i]);
      const float2 c = unpack_bf16(cw[i]), s = unpack_bf16(sw[i]);
      // first half:  x1*cos + (-x2)*sin ; second half: x2*cos + x1*sin   (cos/sin identical for both halves)
      ol[i] = pack_bf16(bf16_round(x1.x * c.x) + bf16_round(-x2.x * s.x), bf16_round(x1.y * c.y) + bf16_round(-x2.y * s.y));
      oh[i] = pack_bf16(bf16_round(x2.x * c.x) + bf16_round(x1.x * s.x), bf16_round(x2.y * c.y) + bf16_round(x1.y * s.y));
    }
    const uint4 vlo = make_uint4(ol[0], ol[1], ol[2], ol[3]), vhi = make_uint4(oh[0], oh[1], oh[2], oh[3]);
    if (which == 0) {
      *reinterpret_cast<uint4*>(p + g) = vlo;
      *reinterpret_cast<uint4*>(p + half + g) = vhi;
    } else {
      *reinterpret_cast<uint4*>(kc + cache_off + g) = vlo;
      *reinterpret_cast<uint4*>(kc + cache_off + half + g) = vhi;
    }
  }
  const __nv_bfloat16* pv = base + 2 * D;
  *reinterpret_cast<uint4*>(vc + cache_off + g) = *reinterpret_cast<const uint4*>(pv + g);
  *reinterpret_cast<uint4*>(vc + cache_off + half + g) = *reinterpret_cast<const uint4*>(pv + half + g);
}

int rope_kv_launch(void* qkv, int B, int T, int H, int hd, int pos0, const void* cos_t, const void* sin_t, void* kc,
                   void* vc, int Tmax, cudaStream_t st) {
  if (hd % 16) return set_error("rope: head_dim %d must be a multiple of 16", hd);
  if (pos0 + T > Tmax) return set_error("rope: position %d exceeds KV capacity %d", pos0 + T, Tmax);
  const long long total = static_cast<long long>(B) * T * H * (hd / 16);
  if (total <= 0) return 0;
  ProfScope prof(kCatOther, 0.0, 12.0 * B * T * H * hd, st);
  rope_kv_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(
      static_cast<__nv_bfloat16*>(qkv), T, H, hd, pos0, static_cast<const __nv_bfloat16*>(cos_t),
      static_cast<const __nv_bfloat16*>(sin_t), static_cast<__nv_bfloat16*>(kc), static_cast<__nv_bfloat16*>(vc),
      Tmax, total);
  CUDA_TRY(cudaGetLastError());
  count_launch();
  return 0;
}

// ------------------------------------------------------------------------------------------- capture pooling
// experiments/robot/openvla_utils.py:126-131,193-199:  hs.float().mean(1)  or  hs[:, -1]  per layer.
// x [B, T(ld rows), D] bf16 -> out [B, D] fp32.  mode 0: mean over rows [0, n_rows); mode 1: row n_rows-1.
// CTA = 8 warps over 256 columns; warp w sums rows w, w+8, ... with 16-byte loads; smem tree over warps.
// Ragged (right-padded) prompts: `lens` (int32 [B], may be null) holds each row's true prompt length out of `P`; row b
// then pools over its own n_rows - (P - lens[b]) leading rows, in the same summation order as a uniform batch of that
// length.
__global__ void __launch_bounds__(256) pool_tokens_kernel(const __nv_bfloat16* __restrict__ x, long long batch_stride,
                                                          long long ld, int n_rows, int D, int mode,
                                                          float* __restrict__ out, long long out_batch_stride,
                                                          const int* __restrict__ lens, int P) {
  __shared__ float red[8][256];
  const int b = blockIdx.y;
  if (lens) n_rows = max(1, n_rows - (P - min(max(lens[b], 1), P)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col = blockIdx.x * 256 + lane * 8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const __nv_bfloat16* xb = x + b * batch_stride;
  if (col < D) {
    if (mode == 1) {
      if (warp == 0) {
        const uint4 v = *reinterpret_cast<const uint4*>(xb + static_cast<long long>(n_rows - 1) * ld + col);
        const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = unpack_bf16(u[i]);
          acc[2 * i] = f.x;
          acc[2 * i + 1] = f.y;
        }
      }
    } else {
#pragma unroll 4
      for (int r = warp; r < n_rows; r += 8) {
        const uint4 v = *reinterpret_cast<const uint4*>(xb + static_cast<long long>(r) * ld + col);
        const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = unpack_bf16(u[i]);
          acc[2 * i] += f.x;
          acc[2 * i + 1] += f.y;
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[warp][lane * 8 + i] = acc[i];
  __syncthreads();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c < D) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    out[b * out_batch_stride + c] = (mode == 0) ? s / static_cast<float>(n_rows) : s;
  }
}

int pool_tokens_launch(const void* x, long long batch_stride, long long ld, int B, int n_rows, int D, int mode,
                       float* out, long long out_batch_stride, cudaStream_t st, const int* lens, int P) {
  if (B <= 0) return 0;
  if (n_rows <= 0) return set_error("pool_tokens: empty token range");
  if (D % 8) return set_error("pool_tokens: D must be a multiple of 8");
  dim3 grid((D + 255) / 256, B);
  ProfScope prof(kCatPool, 0.0, (mode == 0 ? 2.0 * B * n_rows * D : 2.0 * B * D) + 4.0 * B * D, st);
  pool_tokens_kernel<<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), batch_stride, ld, n_rows, D, mode,
                                          out, out_batch_stride, lens, P);
  CUDA_TRY(cudaGetLastError());
  count_launch();
  return 0;
}

// ------------------------------------------------------------------------------------------- argmax
// torch.argmax semantics on fp32 rows: first index of the maximum; NaN counts as the maximum.
__device__ __forceinline__ bool better(float v, int i, float bv, int bi) {
  const bool vn = v != v, bn = bv != bv;
  if (vn != bn) return vn;
  if (vn) return i < bi;
  return v > bv || (v == bv && i < bi);
}

__global__ void __launch_bounds__(256) argmax_rows_kernel(const float* __restrict__ x, long long ld, int n,
                                                          long long* __restrict__ out) {
  __shared__ float sv[8];
  __shared__ int si[8];
  griddep_launch_dependents();
  griddep_wait();
  const float* row = x + blockIdx.x * ld;
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  const int nv = n / 4;
  for (int i = threadIdx.x; i < nv; i += 256) {
    const float4 v = *reinterpret_cast<const float4*>(row + 4 * i);
    const float a[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (better(a[j], 4 * i + j, bv, bi)) { bv = a[j]; bi = 4 * i + j; }
  }
  for (int i = nv * 4 + threadIdx.x; i < n; i += 256)
    if (better(row[i], i, bv, bi)) { bv = row[i]; bi = i; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = bv; si[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w)
      if (better(sv[w], si[w], bv, bi)) { bv = sv[w]; bi = si[w]; }
    out[blockIdx.x] = bi;
  }
}

int argmax_launch(const float* x, long long ld, int rows, int n, long long* out, cudaStream_t st) {
  if (rows <= 0) return 0;
  if (n <= 0) return set_error("argmax: empty rows");
  if (ld % 4 || (reinterpret_cast<uintptr_t>(x) & 15)) return set_error("argmax: rows must be 16-byte aligned");
  CUDA_TRY(launch_pdl(argmax_rows_kernel, dim3(rows), dim3(256), 0, st, x, ld, n, out));
  count_launch();
  return 0;
}

// ------------------------------------------------------------------------------------------- de-tokenise
// modeling_prismatic.py:521-534: disc = clip(vocab - id - 1, 0, n_centers-1); norm = centers[disc];
// action = mask ? 0.5*(norm+1)*(q99-q01)+q01 : norm -- float64, evaluated with explicit round-to-nearest
// mul/add (no FMA contraction) so the result is bit-identical to numpy.
__global__ void detok_unnorm_kernel(const long long* __restrict__ ids, int n, int action_dim, int vocab_size,
                                    const double* __restrict__ centers, int n_centers, const double* __restrict__ q01,
                                    const double* __restrict__ q99, const unsigned char* __restrict__ mask,
                                    double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int a = i % action_dim;
  long long disc = static_cast<long long>(vocab_size) - ids[i] - 1;
  disc = disc < 0 ? 0 : (disc > n_centers - 1 ? n_centers - 1 : disc);
  const double c = centers[disc];
  double r = c;
  if (mask[a]) {
    const double t0 = __dmul_rn(0.5, __dadd_rn(c, 1.0));
    const double t1 = __dmul_rn(t0, __dadd_rn(q99[a], -q01[a]));
    r = __dadd_rn(t1, q01[a]);
  }
  out[i] = r;
}

int detok_unnorm_launch(const long long* ids, int n, int action_dim, int vocab_size, const double* centers,
                        int n_centers, const double* q01, const double* q99, const unsigned char* mask, double* out,
                        cudaStream_t st) {
  if (n <= 0) return 0;
  detok_unnorm_kernel<<<(n + 127) / 128, 128, 0, st>>>(ids, n, action_dim, vocab_size, centers, n_centers, q01, q99,
                                                        mask, out);
  CUDA_TRY(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace ovla

namespace ovla {

// ------------------------------------------------------------------------------------------- frame pre-processing
// PrismaticImageProcessor.apply_transform for frames that already have the model resolution
// (processing_prismatic.py:128-145: to_tensor -> normalize per tower -> channel-stack) followed by the bf16 cast of
// get_vla_action (openvla_utils.py:186):  out[b, 3*tw + c, y, x] = bf16((float(u8[b, y, x, c]) / 255 - mean[tw][c]) / std[tw][c])
// with IEEE fp32 divisions, so the result is bit-identical to torchvision on the host.
__global__ void preprocess_frames_kernel(const unsigned char* __restrict__ frames, int S, int n_towers,
                                         const float* __restrict__ mean, const float* __restrict__ stdv,
                                         __nv_bfloat16* __restrict__ out, long long n_pix) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;  // (b, y, x)
  if (idx >= n_pix) return;
  const long long b = idx / (static_cast<long long>(S) * S);
  const long long yx = idx - b * S * S;
  const unsigned char* p = frames + idx * 3;
  const float v[3] = {__fdiv_rn(static_cast<float>(p[0]), 255.f), __fdiv_rn(static_cast<float>(p[1]), 255.f),
                      __fdiv_rn(static_cast<float>(p[2]), 255.f)};
  for (int tw = 0; tw < n_towers; ++tw) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float r = __fdiv_rn(__fsub_rn(v[c], mean[tw * 3 + c]), stdv[tw * 3 + c]);
      out[(b * 3 * n_towers + tw * 3 + c) * S * S + yx] = __float2bfloat16_rn(r);
    }
  }
}

int preprocess_frames_launch(const void* frames_u8, int B, int S, int n_towers, const float* mean, const float* stdv,
                             void* out, cudaStream_t st) {
  const long long n_pix = static_cast<long long>(B) * S * S;
  if (n_pix <= 0) return 0;
  if (n_towers < 1 || n_towers > 2) return set_error("preprocess: n_towers must be 1 or 2");
  ProfScope prof(kCatOther, 0.0, n_pix * (3.0 + 6.0 * n_towers), st);
  preprocess_frames_kernel<<<static_cast<unsigned>((n_pix + 255) / 256), 256, 0, st>>>(
      static_cast<const unsigned char*>(frames_u8), S, n_towers, mean, stdv, static_cast<__nv_bfloat16*>(out), n_pix);
  CUDA_TRY(cudaGetLastError());
  count_launch();
  return 0;
}

// ------------------------------------------------------------------------------------------------ center crop
// The reference's `center_crop=True` branch (experiments/robot/openvla_utils.py:155-175 + crop_and_resize :81-124):
// uint8 -> float32 / 255 -> tf.image.crop_and_resize(box = centred square of side sqrt(crop_scale), bilinear, to
// out x out) -> clip [0,1] -> uint8 by tf.image.convert_image_dtype(saturate=True) (scale 255.5, truncate).
// Arithmetic follows TensorFlow's CPU kernel (crop_and_resize_op.cc): in_y = y1 (H-1) + y * ((y2-y1) (H-1) / (out-1)),
// lerp across x first, then y, every step a separate fp32 operation (no FMA contraction).
__global__ void center_crop_kernel(const unsigned char* __restrict__ src, int H, int W, float y1, float x1,
                                   float hs, float ws, unsigned char* __restrict__ dst, int S, long long n_pix) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;  // (b, y, x)
  if (idx >= n_pix) return;
  const int x = static_cast<int>(idx % S), y = static_cast<int>((idx / S) % S);
  const long long b = idx / (static_cast<long long>(S) * S);
  const float in_y = __fadd_rn(__fmul_rn(y1, static_cast<float>(H - 1)), __fmul_rn(static_cast<float>(y), hs));
  const float in_x = __fadd_rn(__fmul_rn(x1, static_cast<float>(W - 1)), __fmul_rn(static_cast<float>(x), ws));
  unsigned char* o = dst + idx * 3;
  if (in_y < 0.f || in_y > static_cast<float>(H - 1) || in_x < 0.f || in_x > static_cast<float>(W - 1)) {
    o[0] = o[1] = o[2] = 0;   // extrapolation_value = 0
    return;
  }
  const int ty = static_cast<int>(floorf(in_y)), by = static_cast<int>(ceilf(in_y));
  const int lx = static_cast<int>(floorf(in_x)), rx = static_cast<int>(ceilf(in_x));
  const float yl = __fsub_rn(in_y, static_cast<float>(ty)), xl = __fsub_rn(in_x, static_cast<float>(lx));
  const unsigned char* img = src + b * H * W * 3;
  const float k = 1.0f / 255.0f;   // convert_image_dtype(uint8 -> float32): multiply by float32(1/255)
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float tl = __fmul_rn(static_cast<float>(img[(static_cast<long long>(ty) * W + lx) * 3 + c]), k);
    const float tr = __fmul_rn(static_cast<float>(img[(static_cast<long long>(ty) * W + rx) * 3 + c]), k);
    const float bl = __fmul_rn(static_cast<float>(img[(static_cast<long long>(by) * W + lx) * 3 + c]), k);
    const float br = __fmul_rn(static_cast<float>(img[(static_cast<long long>(by) * W + rx) * 3 + c]), k);
    const float top = __fadd_rn(tl, __fmul_rn(__fsub_rn(tr, tl), xl));
    const float bot = __fadd_rn(bl, __fmul_rn(__fsub_rn(br, bl), xl));
    float v = __fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), yl));
    v = fminf(fmaxf(v, 0.f), 1.f);
    const float sc = __fmul_rn(v, 255.5f);
    o[c] = static_cast<unsigned char>(fminf(fmaxf(sc, 0.f), 255.f));   // saturate, truncate toward zero
  }
}

int center_crop_launch(const void* src_u8, int B, int H, int W, float crop_scale, void* dst_u8, int S, cudaStream_t st) {
  const long long n_pix = static_cast<long long>(B) * S * S;
  if (n_pix <= 0) return 0;
  if (H < 2 || W < 2 || S < 2) return set_error("center crop: image and output sides must be >= 2");
  if (!(crop_scale > 0.f)) return set_error("center crop: crop_scale must be positive");
  // box as the reference builds it (float32): side = clip(sqrt(scale), 0, 1), offset = (1 - side) / 2
  const float side = fminf(fmaxf(sqrtf(crop_scale), 0.f), 1.f);
  const float off = (1.f - side) / 2.f;
  const float y2 = off + side;
  const float hs = (y2 - off) * static_cast<float>(H - 1) / static_cast<float>(S - 1);
  const float ws = (y2 - off) * static_cast<float>(W - 1) / static_cast<float>(S - 1);
  ProfScope prof(kCatOther, 0.0, n_pix * 3.0 * 5.0, st);
  center_crop_kernel<<<static_cast<unsigned>((n_pix + 255) / 256), 256, 0, st>>>(
      static_cast<const unsigned char*>(src_u8), H, W, off, off, hs, ws, static_cast<unsigned char*>(dst_u8), S, n_pix);
  CUDA_TRY(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace ovla
