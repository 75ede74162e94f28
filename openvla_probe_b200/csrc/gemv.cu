// Skinny GEMM for the batch<=8 cached decode steps (modeling_prismatic.py:325-341 -> LlamaForCausalLM decode):
//   out[m, n] = sum_k x[m, k] * W[n, k],  m < M <= 8.
// Pure weight streaming: every weight byte is read exactly once with 16-byte loads (L1 no-allocate), one warp per
// output column, 4 independent loads in flight per lane; the M activation rows (<= 176 KB total) stay L1/L2
// resident.  Same epilogues (and rounding points) as the tcgen05 GEMM.
#include "host_util.h"
#include "ops.h"
#include "ptx.cuh"

namespace ovla {

__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

template <int MB>
__device__ __forceinline__ void dot_row(const __nv_bfloat16* __restrict__ w, const __nv_bfloat16* __restrict__ x,
                                        long long ldx, int M, int K, int lane, float* acc) {
#pragma unroll
  for (int m = 0; m < MB; ++m) acc[m] = 0.f;
  const int kv = K / 8;
  int i = lane;
  for (; i + 96 < kv; i += 128) {
    uint4 wv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) wv[u] = ldg_stream(w + (i + 32 * u) * 8);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint32_t ww[4] = {wv[u].x, wv[u].y, wv[u].z, wv[u].w};
#pragma unroll
      for (int m = 0; m < MB; ++m) {
        const uint4 xv = *reinterpret_cast<const uint4*>(x + (m < M ? m : M - 1) * ldx + (i + 32 * u) * 8);
        const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 a = unpack_bf16(ww[j]), b = unpack_bf16(xw[j]);
          acc[m] = fmaf(a.x, b.x, acc[m]);
          acc[m] = fmaf(a.y, b.y, acc[m]);
        }
      }
    }
  }
  for (; i < kv; i += 32) {
    const uint4 wv = ldg_stream(w + i * 8);
    const uint32_t ww[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
    for (int m = 0; m < MB; ++m) {
      const uint4 xv = *reinterpret_cast<const uint4*>(x + (m < M ? m : M - 1) * ldx + i * 8);
      const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 a = unpack_bf16(ww[j]), b = unpack_bf16(xw[j]);
        acc[m] = fmaf(a.x, b.x, acc[m]);
        acc[m] = fmaf(a.y, b.y, acc[m]);
      }
    }
  }
#pragma unroll
  for (int m = 0; m < MB; ++m) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[m] += __shfl_xor_sync(0xffffffffu, acc[m], o);
  }
}

template <int MB, int MODE>
__global__ void __launch_bounds__(256) gemv_kernel(const __nv_bfloat16* __restrict__ x, long long ldx,
                                                   const __nv_bfloat16* __restrict__ W, long long ldw, int M, int N,
                                                   int K, GemmEpi epi) {
  const int lane = threadIdx.x & 31;
  const int warp_g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int n_warps = gridDim.x * (blockDim.x >> 5);
  const int n_out = (MODE == kModeSwiGLU) ? N / 2 : N;
  for (int n = warp_g; n < n_out; n += n_warps) {
    float acc[MB], acc2[MB];
    if (MODE == kModeSwiGLU) {
      const long long rg = static_cast<long long>(n / 32) * 64 + (n % 32);
      dot_row<MB>(W + rg * ldw, x, ldx, M, K, lane, acc);
      dot_row<MB>(W + (rg + 32) * ldw, x, ldx, M, K, lane, acc2);
    } else {
      dot_row<MB>(W + static_cast<long long>(n) * ldw, x, ldx, M, K, lane, acc);
    }
    if (lane == 0) {
#pragma unroll
      for (int m = 0; m < MB; ++m) {
        if (m >= M) break;
        if (MODE == kModeBf16) {
          float v = acc[m];
          if (epi.bias) v += __bfloat162float(epi.bias[n]);
          v = bf16_round(v);
          if (epi.gelu) v = bf16_round(gelu_erf(v));
          if (epi.scale) v = bf16_round(v * __bfloat162float(epi.scale[n]));
          if (epi.resid) v += __bfloat162float(epi.resid[m * epi.ldr + n]);
          reinterpret_cast<__nv_bfloat16*>(epi.out)[m * epi.ldo + n] = __float2bfloat16_rn(v);
        } else if (MODE == kModeSwiGLU) {
          const float g = bf16_round(acc[m]), u = bf16_round(acc2[m]);
          reinterpret_cast<__nv_bfloat16*>(epi.out)[m * epi.ldo + n] = __float2bfloat16_rn(bf16_round(silu(g)) * u);
        } else {
          float v = acc[m];
          if (epi.bias_f32) v += epi.bias_f32[n];
          if (epi.bias) v += __bfloat162float(epi.bias[n]);
          if (epi.round_bf16) v = bf16_round(v);
          reinterpret_cast<float*>(epi.out)[m * epi.ldo + n] = v;
        }
      }
    }
  }
}

template <int MODE>
static int gemv_dispatch(const void* x, long long ldx, const void* W, long long ldw, int M, int N, int K,
                         const GemmEpi& epi, cudaStream_t st) {
  const int n_out = (MODE == kModeSwiGLU) ? N / 2 : N;
  int blocks = (n_out + 7) / 8;
  const int max_blocks = num_sms() * 8;
  if (blocks > max_blocks) blocks = max_blocks;
  auto X = static_cast<const __nv_bfloat16*>(x);
  auto Wp = static_cast<const __nv_bfloat16*>(W);
  ProfScope prof(kCatGemv, 2.0 * M * N * K, 2.0 * N * K + 2.0 * M * (K + n_out), st);
  if (M <= 1) gemv_kernel<1, MODE><<<blocks, 256, 0, st>>>(X, ldx, Wp, ldw, M, N, K, epi);
  else if (M <= 2) gemv_kernel<2, MODE><<<blocks, 256, 0, st>>>(X, ldx, Wp, ldw, M, N, K, epi);
  else if (M <= 4) gemv_kernel<4, MODE><<<blocks, 256, 0, st>>>(X, ldx, Wp, ldw, M, N, K, epi);
  else gemv_kernel<8, MODE><<<blocks, 256, 0, st>>>(X, ldx, Wp, ldw, M, N, K, epi);
  CUDA_TRY(cudaGetLastError());
  count_launch();
  return 0;
}

int gemv_launch(const void* x, long long ldx, const void* W, long long ldw, int M, int N, int K, int mode,
                const GemmEpi& epi, cudaStream_t st) {
  if (M <= 0 || M > 8) return set_error("gemv: M=%d out of range [1,8]", M);
  if (K % 8 || ldx % 8 || ldw % 8) return set_error("gemv: K and pitches must be multiples of 8");
  if (mode == kModeSwiGLU && (N % 64)) return set_error("gemv: SwiGLU N must be a multiple of 64");
  if (mode == kModeBf16) return gemv_dispatch<kModeBf16>(x, ldx, W, ldw, M, N, K, epi, st);
  if (mode == kModeSwiGLU) return gemv_dispatch<kModeSwiGLU>(x, ldx, W, ldw, M, N, K, epi, st);
  if (mode == kModeF32) return gemv_dispatch<kModeF32>(x, ldx, W, ldw, M, N, K, epi, st);
  return set_error("gemv: unsupported mode %d", mode);
}

}  // namespace ovla
