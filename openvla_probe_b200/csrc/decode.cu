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
#include "decode_attn.cuh"

namespace ovla {

template <int HD, bool FAST>
__global__ void __launch_bounds__(kDecThreads) decode_rope_attn_kernel(
    const __nv_bfloat16* __restrict__ qkv, long long qkv_ld, const __nv_bfloat16* __restrict__ cos_t,
    const __nv_bfloat16* __restrict__ sin_t, int pos, __nv_bfloat16* __restrict__ kc, __nv_bfloat16* __restrict__ vc,
    int Tmax, __nv_bfloat16* __restrict__ out, long long o_ld, float scale, const int* __restrict__ lens, int P) {
  extern __shared__ float dyn[];            // [max(ctx, kDecGroups * HD)] scores, later the cross-group reduction
  __shared__ float sq[HD];
  __shared__ float red[kDecThreads / 32];
  const int h = blockIdx.x, b = blockIdx.y, H = gridDim.x;
  const long long head_off = (static_cast<long long>(b) * H + h) * Tmax * HD;
  // let a PDL-launched successor (the o_proj GEMV) start prefetching its weights while this kernel runs, then wait
  // for the QKV projection this kernel consumes
  griddep_launch_dependents();
  griddep_wait();
  // ragged (right-padded) prompts: row b sits (P - lens[b]) positions before the longest row
  if (lens) pos = max(0, pos - (P - min(max(lens[b], 1), P)));
  decode_rope_attn_body<HD, false, false, FAST>(qkv + b * qkv_ld + h * HD, static_cast<long long>(H) * HD, cos_t, sin_t, pos,
                                   kc + head_off, vc + head_off, out + b * o_ld + h * HD, scale, dyn, sq, red);
}

int decode_rope_attn_launch(const void* qkv, long long qkv_ld, const void* cos_t, const void* sin_t, int pos, void* kc,
                            void* vc, int B, int H, int head_dim, int Tmax, void* out, long long o_ld,
                            cudaStream_t st, const int* lens, int P) {
  if (B <= 0) return 0;
  if (head_dim != 128) return set_error("decode attention: head_dim %d unsupported (128 only)", head_dim);
  if (pos < 0 || pos >= Tmax) return set_error("decode attention: position %d outside the KV capacity %d", pos, Tmax);
  const int ctx = pos + 1;
  const int n_f = ctx > kDecGroups * 128 ? ctx : kDecGroups * 128;
  const int smem = n_f * static_cast<int>(sizeof(float));
  if (smem > 48 * 1024) return set_error("decode attention: ctx=%d too long", ctx);
  dim3 grid(H, B);
  ProfScope prof(kCatDecodeAttn, 4.0 * B * H * ctx * head_dim, 4.0 * B * H * ctx * head_dim + 12.0 * B * H * head_dim, st);
  // few CTAs (latency-bound): all K / V loads in flight at once; many CTAs (HBM-bound): the low-register looped form
  auto kern = static_cast<long long>(B) * H <= 4LL * num_sms() ? decode_rope_attn_kernel<128, true>
                                                               : decode_rope_attn_kernel<128, false>;
  CUDA_TRY(launch_pdl(kern, grid, dim3(kDecThreads), smem, st,
                      static_cast<const __nv_bfloat16*>(qkv), qkv_ld, static_cast<const __nv_bfloat16*>(cos_t),
                      static_cast<const __nv_bfloat16*>(sin_t), pos, static_cast<__nv_bfloat16*>(kc),
                      static_cast<__nv_bfloat16*>(vc), Tmax, static_cast<__nv_bfloat16*>(out), o_ld,
                      1.0f / sqrtf(static_cast<float>(head_dim)), lens, P));
  count_launch();
  return 0;
}

}  // namespace ovla
