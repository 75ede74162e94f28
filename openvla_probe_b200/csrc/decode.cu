// Fused cached-decode attention step (modeling_prismatic.py:325-341 -> LlamaAttention with past_key_values):
// one kernel per layer does  RoPE(q, k) at position `pos`  ->  append k, v to the KV cache  ->  softmax(q K^T / sqrt d) V
// over the pos+1 cached keys.  CTA per (head, batch row), 256 threads.  K and V are streamed exactly once with
// 16-byte loads, 4 independent loads in flight per lane, so the kernel is HBM-bound at large batch and
// latency-short (about ten dependent load rounds for ctx ~ 290) at batch 1.
// Rounding points follow the reference's bf16 ops: q/k RoPE products and sum rounded to bf16, probabilities rounded
// to bf16 before the PV product, output rounded to bf16.
#include "host_util.h"
#include "ops.h"
#include "ptx.cuh"

namespace ovla {

static constexpr int kDecThreads = 256;
static constexpr int kDecGroups = kDecThreads / 16;  // 16-lane groups, one 256-byte K/V row each
static constexpr int kDecR = 4;                       // rows per group per iteration = independent 16-byte loads per lane

template <int HD>
__global__ void __launch_bounds__(kDecThreads) decode_rope_attn_kernel(
    const __nv_bfloat16* __restrict__ qkv, long long qkv_ld, const __nv_bfloat16* __restrict__ cos_t,
    const __nv_bfloat16* __restrict__ sin_t, int pos, __nv_bfloat16* __restrict__ kc, __nv_bfloat16* __restrict__ vc,
    int Tmax, __nv_bfloat16* __restrict__ out, long long o_ld, float scale) {
  static_assert(HD == 128, "decode attention is specialised for head_dim 128");
  extern __shared__ float dyn[];            // [max(ctx, kDecGroups * HD)] scores, later the cross-group reduction
  __shared__ float sq[HD];
  __shared__ float red[kDecThreads / 32];
  const int h = blockIdx.x, b = blockIdx.y, H = gridDim.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int hl = tid & 15, grp = tid >> 4;
  const int ctx = pos + 1;
  const long long head_off = (static_cast<long long>(b) * H + h) * Tmax * HD;
  __nv_bfloat16* kb = kc + head_off;
  __nv_bfloat16* vb = vc + head_off;
  const __nv_bfloat16* row = qkv + b * qkv_ld + h * HD;
  const long long D = static_cast<long long>(H) * HD;
  // let a PDL-launched successor (the o_proj GEMV) start prefetching its weights while this kernel runs, then wait
  // for the QKV projection this kernel consumes
  griddep_launch_dependents();
  griddep_wait();

  // ---- RoPE on q and k (pairs i, i + HD/2), append k and v at `pos`
  if (tid < HD / 2) {
    const float c = __bfloat162float(cos_t[static_cast<long long>(pos) * (HD / 2) + tid]);
    const float s = __bfloat162float(sin_t[static_cast<long long>(pos) * (HD / 2) + tid]);
    const float q1 = __bfloat162float(row[tid]), q2 = __bfloat162float(row[tid + HD / 2]);
    const float k1 = __bfloat162float(row[D + tid]), k2 = __bfloat162float(row[D + tid + HD / 2]);
    sq[tid] = bf16_round(bf16_round(q1 * c) + bf16_round(-q2 * s));
    sq[tid + HD / 2] = bf16_round(bf16_round(q2 * c) + bf16_round(q1 * s));
    kb[static_cast<long long>(pos) * HD + tid] = __float2bfloat16_rn(bf16_round(k1 * c) + bf16_round(-k2 * s));
    kb[static_cast<long long>(pos) * HD + tid + HD / 2] = __float2bfloat16_rn(bf16_round(k2 * c) + bf16_round(k1 * s));
  } else if (tid < HD / 2 + HD / 8) {
    const int i = (tid - HD / 2) * 8;
    *reinterpret_cast<uint4*>(vb + static_cast<long long>(pos) * HD + i) =
        *reinterpret_cast<const uint4*>(row + 2 * D + i);
  }
  __syncthreads();  // q in smem; this CTA's own k/v stores are visible to its later loads
  float qv[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) qv[i] = sq[hl * 8 + i];

  // ---- scores: 16 lanes per key, kDecR keys per group per iteration
  // (the trip count is block-uniform: both 16-lane halves of a warp must reach the shuffles together)
  for (int base = 0; base < ctx; base += kDecGroups * kDecR) {
    const int j0 = base + grp * kDecR;
    uint4 u[kDecR];
#pragma unroll
    for (int r = 0; r < kDecR; ++r) {
      const int j = min(j0 + r, ctx - 1);
      u[r] = *reinterpret_cast<const uint4*>(kb + static_cast<long long>(j) * HD + hl * 8);
    }
#pragma unroll
    for (int r = 0; r < kDecR; ++r) {
      const uint32_t w[4] = {u[r].x, u[r].y, u[r].z, u[r].w};
      float d = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = unpack_bf16(w[i]);
        d += f.x * qv[2 * i] + f.y * qv[2 * i + 1];
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
      if (hl == 0 && j0 + r < ctx) dyn[j0 + r] = d * scale;
    }
  }
  __syncthreads();
  // ---- softmax over the scores
  float mx = -INFINITY;
  for (int j = tid; j < ctx; j += kDecThreads) mx = fmaxf(mx, dyn[j]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int w = 1; w < kDecThreads / 32; ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  float sum = 0.f;
  for (int j = tid; j < ctx; j += kDecThreads) {
    const float p = __expf(dyn[j] - mx);
    dyn[j] = p;
    sum += p;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < kDecThreads / 32; ++w) tot += red[w];
  const float inv = 1.f / tot;
  // ---- O = P V: group g takes kDecR consecutive rows of every block of kDecGroups*kDecR rows
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int base = 0; base < ctx; base += kDecGroups * kDecR) {
    const int j0 = base + grp * kDecR;
    uint4 u[kDecR];
    float p[kDecR];
#pragma unroll
    for (int r = 0; r < kDecR; ++r) {
      const int j = min(j0 + r, ctx - 1);
      u[r] = *reinterpret_cast<const uint4*>(vb + static_cast<long long>(j) * HD + hl * 8);
      p[r] = (j0 + r < ctx) ? bf16_round(dyn[j] * inv) : 0.f;
    }
#pragma unroll
    for (int r = 0; r < kDecR; ++r) {
      const uint32_t w[4] = {u[r].x, u[r].y, u[r].z, u[r].w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = unpack_bf16(w[i]);
        acc[2 * i] += p[r] * f.x;
        acc[2 * i + 1] += p[r] * f.y;
      }
    }
  }
  __syncthreads();  // scores are dead: reuse the buffer for the cross-group reduction
#pragma unroll
  for (int i = 0; i < 8; ++i) dyn[grp * HD + hl * 8 + i] = acc[i];
  __syncthreads();
  if (tid < HD) {
    float o = 0.f;
#pragma unroll
    for (int g = 0; g < kDecGroups; ++g) o += dyn[g * HD + tid];
    out[b * o_ld + h * HD + tid] = __float2bfloat16_rn(o);
  }
}

int decode_rope_attn_launch(const void* qkv, long long qkv_ld, const void* cos_t, const void* sin_t, int pos, void* kc,
                            void* vc, int B, int H, int head_dim, int Tmax, void* out, long long o_ld,
                            cudaStream_t st) {
  if (B <= 0) return 0;
  if (head_dim != 128) return set_error("decode attention: head_dim %d unsupported (128 only)", head_dim);
  if (pos < 0 || pos >= Tmax) return set_error("decode attention: position %d outside the KV capacity %d", pos, Tmax);
  const int ctx = pos + 1;
  const int n_f = ctx > kDecGroups * 128 ? ctx : kDecGroups * 128;
  const int smem = n_f * static_cast<int>(sizeof(float));
  if (smem > 48 * 1024) return set_error("decode attention: ctx=%d too long", ctx);
  dim3 grid(H, B);
  ProfScope prof(kCatDecodeAttn, 4.0 * B * H * ctx * head_dim, 4.0 * B * H * ctx * head_dim + 12.0 * B * H * head_dim, st);
  CUDA_TRY(launch_pdl(decode_rope_attn_kernel<128>, grid, dim3(kDecThreads), smem, st,
                      static_cast<const __nv_bfloat16*>(qkv), qkv_ld, static_cast<const __nv_bfloat16*>(cos_t),
                      static_cast<const __nv_bfloat16*>(sin_t), pos, static_cast<__nv_bfloat16*>(kc),
                      static_cast<__nv_bfloat16*>(vc), Tmax, static_cast<__nv_bfloat16*>(out), o_ld,
                      1.0f / sqrtf(static_cast<float>(head_dim))));
  count_launch();
  return 0;
}

}  // namespace ovla
