// Body of the fused cached-decode attention step for ONE (head, batch row), executed by a whole CTA of kDecThreads
// threads (shared by decode.cu's per-layer kernel and the persistent decode-step kernel in decode_mega.cu, which must
// return the same bits).  Rounding points follow the reference's bf16 ops: q/k RoPE products and sum rounded to bf16,
// probabilities rounded to bf16 before the PV product, output rounded to bf16.
// LDCG: read the qkv row with L2-only loads (a persistent kernel re-reads a buffer other SMs rewrote; L1 is not
// coherent across SMs).
#pragma once
#include "ptx.cuh"

namespace ovla {

static constexpr int kDecThreads = 256;
static constexpr int kDecGroups = kDecThreads / 16;  // 16-lane groups, one 256-byte K/V row each
static constexpr int kDecR = 4;                       // rows per group per iteration = independent 16-byte loads per lane
static constexpr int kDecFast = 5;                    // blocks of kDecGroups * kDecR rows whose loads are issued up front

template <bool LDCG>
__device__ __forceinline__ float ld_act(const __nv_bfloat16* p) {
  return __bfloat162float(LDCG ? __ldcg(p) : *p);
}

// the kDecThreads threads of the body synchronise with __syncthreads(), or -- inside a larger CTA whose other warps do
// something else (decode_mega.cu) -- with named barrier 1
template <bool NAMED_BAR>
__device__ __forceinline__ void dec_sync() {
  if (NAMED_BAR) asm volatile("bar.sync 1, %0;" ::"n"(kDecThreads) : "memory");
  else __syncthreads();
}

// dyn: max(ctx, kDecGroups * HD) floats, sq: HD floats, red: kDecThreads / 32 floats (all shared memory)
// FAST: short contexts issue all their K (then V) loads up front -- the latency-bound small-batch form; it holds
// kDecFast * kDecR 16-byte registers per lane, which costs occupancy the HBM-bound large-batch launch needs (bs = 256:
// 43 -> 65 ms per step with it), so the per-layer kernel instantiates both and picks by grid size.  Same bits either way.
template <int HD, bool LDCG, bool NAMED_BAR = false, bool FAST = true>
__device__ __forceinline__ void decode_rope_attn_body(const __nv_bfloat16* __restrict__ row, long long D,
                                                      const __nv_bfloat16* __restrict__ cos_t,
                                                      const __nv_bfloat16* __restrict__ sin_t, int pos,
                                                      __nv_bfloat16* __restrict__ kb, __nv_bfloat16* __restrict__ vb,
                                                      __nv_bfloat16* __restrict__ out_row, float scale, float* dyn, float* sq,
                                                      float* red) {
  static_assert(HD == 128, "decode attention is specialised for head_dim 128");
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int hl = tid & 15, grp = tid >> 4;
  const int ctx = pos + 1;
  // ---- RoPE on q and k (pairs i, i + HD/2), append k and v at `pos`
  if (tid < HD / 2) {
    const float c = __bfloat162float(cos_t[static_cast<long long>(pos) * (HD / 2) + tid]);
    const float s = __bfloat162float(sin_t[static_cast<long long>(pos) * (HD / 2) + tid]);
    const float q1 = ld_act<LDCG>(row + tid), q2 = ld_act<LDCG>(row + tid + HD / 2);
    const float k1 = ld_act<LDCG>(row + D + tid), k2 = ld_act<LDCG>(row + D + tid + HD / 2);
    sq[tid] = bf16_round(bf16_round(q1 * c) + bf16_round(-q2 * s));
    sq[tid + HD / 2] = bf16_round(bf16_round(q2 * c) + bf16_round(q1 * s));
    kb[static_cast<long long>(pos) * HD + tid] = __float2bfloat16_rn(bf16_round(k1 * c) + bf16_round(-k2 * s));
    kb[static_cast<long long>(pos) * HD + tid + HD / 2] = __float2bfloat16_rn(bf16_round(k2 * c) + bf16_round(k1 * s));
  } else if (tid < HD / 2 + HD / 8) {
    const int i = (tid - HD / 2) * 8;
    *reinterpret_cast<uint4*>(vb + static_cast<long long>(pos) * HD + i) =
        LDCG ? __ldcg(reinterpret_cast<const uint4*>(row + 2 * D + i)) : *reinterpret_cast<const uint4*>(row + 2 * D + i);
  }
  dec_sync<NAMED_BAR>();  // q in smem; this CTA's own k/v stores are visible to its later loads
  float qv[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) qv[i] = sq[hl * 8 + i];

  // Rows of K / V a 16-lane group touches: j = base + grp * kDecR + r for base = 0, 64, 128, ...  (this mapping fixes the
  // summation order of the PV product).  Short contexts (ctx <= kDecFast * 64 = 320: every OpenVLA decode step) issue
  // ALL their K loads up front and all their V loads before the softmax, so the phase is two memory round trips plus
  // the block reductions instead of ~10 dependent rounds; longer contexts loop.  Same arithmetic in both forms.
  constexpr int kStep = kDecGroups * kDecR;
  const bool fast = FAST && ctx <= kDecFast * kStep;
  auto score_block = [&](int base, const uint4* u) {
    const int j0 = base + grp * kDecR;
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
  };
  auto load_block = [&](const __nv_bfloat16* base_ptr, int base, uint4* u) {
    const int j0 = base + grp * kDecR;
#pragma unroll
    for (int r = 0; r < kDecR; ++r) {
      const int j = min(j0 + r, ctx - 1);
      u[r] = *reinterpret_cast<const uint4*>(base_ptr + static_cast<long long>(j) * HD + hl * 8);
    }
  };
  constexpr int kF = FAST ? kDecFast : 1;
  uint4 uf[kF][kDecR];
  // ---- scores: 16 lanes per key, kDecR keys per group per block
  // (trip counts are block-uniform: both 16-lane halves of a warp must reach the shuffles together)
  if (fast) {
#pragma unroll
    for (int it = 0; it < kF; ++it) load_block(kb, it * kStep, uf[it]);
#pragma unroll
    for (int it = 0; it < kF; ++it)
      if (it * kStep < ctx) score_block(it * kStep, uf[it]);
    // V rows on their way while the softmax runs
#pragma unroll
    for (int it = 0; it < kF; ++it) load_block(vb, it * kStep, uf[it]);
  } else {
    for (int base = 0; base < ctx; base += kStep) {
      uint4 u[kDecR];
      load_block(kb, base, u);
      score_block(base, u);
    }
  }
  dec_sync<NAMED_BAR>();
  // ---- softmax over the scores
  float mx = -INFINITY;
  for (int j = tid; j < ctx; j += kDecThreads) mx = fmaxf(mx, dyn[j]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) red[warp] = mx;
  dec_sync<NAMED_BAR>();
  mx = red[0];
#pragma unroll
  for (int w = 1; w < kDecThreads / 32; ++w) mx = fmaxf(mx, red[w]);
  dec_sync<NAMED_BAR>();
  float sum = 0.f;
  for (int j = tid; j < ctx; j += kDecThreads) {
    const float p = __expf(dyn[j] - mx);
    dyn[j] = p;
    sum += p;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) red[warp] = sum;
  dec_sync<NAMED_BAR>();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < kDecThreads / 32; ++w) tot += red[w];
  const float inv = 1.f / tot;
  // ---- O = P V: group g takes kDecR consecutive rows of every block of kDecGroups*kDecR rows
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  auto pv_block = [&](int base, const uint4* u) {
    const int j0 = base + grp * kDecR;
    float p[kDecR];
#pragma unroll
    for (int r = 0; r < kDecR; ++r) {
      const int j = min(j0 + r, ctx - 1);
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
  };
  if (fast) {
#pragma unroll
    for (int it = 0; it < kF; ++it)
      if (it * kStep < ctx) pv_block(it * kStep, uf[it]);
  } else {
    for (int base = 0; base < ctx; base += kStep) {
      uint4 u[kDecR];
      load_block(vb, base, u);
      pv_block(base, u);
    }
  }
  dec_sync<NAMED_BAR>();  // scores are dead: reuse the buffer for the cross-group reduction
#pragma unroll
  for (int i = 0; i < 8; ++i) dyn[grp * HD + hl * 8 + i] = acc[i];
  dec_sync<NAMED_BAR>();
  if (tid < HD) {
    float o = 0.f;
#pragma unroll
    for (int g = 0; g < kDecGroups; ++g) o += dyn[g * HD + tid];
    out_row[tid] = __float2bfloat16_rn(o);
  }
}

}  // namespace ovla
