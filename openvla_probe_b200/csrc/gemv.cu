// Skinny GEMM for the batch<=8 cached decode steps (modeling_prismatic.py:325-341 -> LlamaForCausalLM decode):
//   out[m, n] = sum_k x[m, k] * W[n, k],  m < M <= 8.
// Pure weight streaming: every weight byte is read exactly once with 16-byte L1-bypassing loads; one warp per output
// column; the loads are software-pipelined (the next 4 chunks are in flight while the current 4 are consumed), small
// register footprint so that ~32 warps per SM keep > 64 KB of loads in flight.  The M activation rows (<= 176 KB in
// total) are read through L1.  Launched with programmatic dependent launch: the first weight chunks are fetched
// before the dependency wait, i.e. while the preceding kernel is still finishing.  Same epilogues / rounding points
// as the tcgen05 GEMM.
#include <stdlib.h>

#include <algorithm>

#include "host_util.h"
#include "ops.h"
#include "ptx.cuh"

namespace ovla {

static constexpr int kGemvThreadsDefault = 128;

__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

template <int MB>
__device__ __forceinline__ void fma_chunk(const uint4& wv, const __nv_bfloat16* __restrict__ x, long long ldx, int M,
                                          int col, WsAcc* acc) {
#pragma unroll
  for (int m = 0; m < MB; ++m) {
    const uint4 xv = __ldg(reinterpret_cast<const uint4*>(x + (m < M ? m : M - 1) * ldx + col));
    wstream_fma8(wv, xv, acc[m]);
  }
}

template <int kCh>
__device__ __forceinline__ void load_stage(uint4* dst, const __nv_bfloat16* __restrict__ w, int i, int kv) {
#pragma unroll
  for (int u = 0; u < kCh; ++u)
    if (i + 32 * u < kv) dst[u] = ldg_stream(w + (i + 32 * u) * 8);
}

// dot products of one weight row with the M activation rows; `cur` already holds the row's first stage
template <int MB, int kCh>
__device__ __forceinline__ void dot_row(const __nv_bfloat16* __restrict__ w, uint4* cur,
                                        const __nv_bfloat16* __restrict__ x, long long ldx, int M, int K, int lane,
                                        float* acc) {
  WsAcc c[MB];
#pragma unroll
  for (int m = 0; m < MB; ++m) wstream_zero(c[m]);
  const int kv = K / 8;
  for (int i = lane; i < kv; i += 32 * kCh) {
    uint4 nxt[kCh];
    load_stage<kCh>(nxt, w, i + 32 * kCh, kv);  // next stage in flight while this one is consumed
#pragma unroll
    for (int u = 0; u < kCh; ++u)
      if (i + 32 * u < kv) fma_chunk<MB>(cur[u], x, ldx, M, (i + 32 * u) * 8, c);
#pragma unroll
    for (int u = 0; u < kCh; ++u) cur[u] = nxt[u];
  }
#pragma unroll
  for (int m = 0; m < MB; ++m) {
    acc[m] = wstream_combine(c[m]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[m] += __shfl_xor_sync(0xffffffffu, acc[m], o);
  }
}

// epilogue of output column n for the M activation rows (one lane): same rounding points as the tcgen05 GEMM epilogue
template <int MB, int MODE>
__device__ __forceinline__ void wstream_epilogue(const GemmEpi& epi, int M, int n, const float* acc, const float* acc2) {
#pragma unroll
  for (int m = 0; m < MB; ++m) {
    if (m >= M) break;
    if (MODE == kModeBf16) {
      float v = acc[m];
      if (epi.bias) v += __bfloat162float(epi.bias[n]);
      v = bf16_round(v);
      if (epi.gelu) {   // same GELU evaluation as the tensor-core epilogue, so results do not depend on the batch size
        float unused = 0.f;
        gelu_erf_x2(v, unused);
        v = bf16_round(v);
      }
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

template <int MB, int MODE, int kCh>
__global__ void __launch_bounds__(256) ovla_wstream_kernel(const __nv_bfloat16* __restrict__ x, long long ldx,
                                                            const __nv_bfloat16* __restrict__ W, long long ldw, int M,
                                                            int N, int K, GemmEpi epi) {
  const int lane = threadIdx.x & 31;
  const int warp_g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int n_warps = gridDim.x * (blockDim.x >> 5);
  const int n_out = (MODE == kModeSwiGLU) ? N / 2 : N;
  const int kv = K / 8;
  auto row_of = [](int n) -> long long {
    return (MODE == kModeSwiGLU) ? static_cast<long long>(n / 32) * 64 + (n % 32) : n;
  };
  // weights do not depend on the previous kernel: the first stage of this warp's first row is fetched before the
  // dependency wait and overlaps the predecessor's tail
  uint4 cur[kCh];
  if (warp_g < n_out) load_stage<kCh>(cur, W + row_of(warp_g) * ldw, lane, kv);
  griddep_launch_dependents();
  griddep_wait();  // activations / residual written by the predecessor are complete and visible from here on

  for (int n = warp_g; n < n_out; n += n_warps) {
    float acc[MB], acc2[MB];
    const long long r = row_of(n);
    if (n != warp_g) load_stage<kCh>(cur, W + r * ldw, lane, kv);
    dot_row<MB, kCh>(W + r * ldw, cur, x, ldx, M, K, lane, acc);
    if (MODE == kModeSwiGLU) {
      load_stage<kCh>(cur, W + (r + 32) * ldw, lane, kv);
      dot_row<MB, kCh>(W + (r + 32) * ldw, cur, x, ldx, M, K, lane, acc2);
    }
    if (lane == 0) wstream_epilogue<MB, MODE>(epi, M, n, acc, acc2);
  }
}

// ------------------------------------------------------------------------------------------ TMA-fed variant (M <= 4)
// The register-pipelined kernel above keeps one 2 KB stage per warp in flight (~48 KB per SM at its occupancy), which
// is about what Little's law asks for at the UNLOADED DRAM latency and not enough once HBM queues build up: ncu shows
// 36-55 % DRAM throughput for the Llama decode matrices.  Here the bytes in flight live in shared memory instead of
// registers: every warp owns a ring of S slots of 8 KB, lane 0 fills a slot with ONE cp.async.bulk (a whole 4096-element
// piece of a weight row) that completes on the slot's mbarrier, and the warp consumes slot u while S-1 more pieces
// (16-24 KB per warp, 128-192 KB per SM) are on their way.  The first S pieces are requested BEFORE the programmatic
// dependency wait, i.e. while the predecessor still runs.  x (M rows, <= 88 KB) is staged in shared memory once.
// Per-lane summation order is identical to the kernel above (ptx.cuh: wstream_fma8 / wstream_combine over the 16-byte
// chunks lane, lane+32, ... of the row in ascending order, then the xor-shuffle tree), so both return the same bits.
static bool use_pdl();
static constexpr int kTmaWarps = 8;
static constexpr int kTmaPiece = 4096;                         // elements per ring slot (8 KB)

template <int MB, int MODE>
__global__ void __launch_bounds__(kTmaWarps * 32, 1) ovla_wstream_tma_kernel(const __nv_bfloat16* __restrict__ x, long long ldx,
                                                                             const __nv_bfloat16* __restrict__ W, long long ldw,
                                                                             int M, int N, int K, GemmEpi epi, int S) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int warp_g = blockIdx.x * kTmaWarps + warp;
  const int n_warps = gridDim.x * kTmaWarps;
  const int n_out = (MODE == kModeSwiGLU) ? N / 2 : N;
  constexpr int kRowsPerJob = (MODE == kModeSwiGLU) ? 2 : 1;
  const int cpr = (K + kTmaPiece - 1) / kTmaPiece;               // pieces per weight row
  const int upj = kRowsPerJob * cpr;                             // pieces per output column
  const int n_jobs = warp_g < n_out ? (n_out - warp_g + n_warps - 1) / n_warps : 0;
  const int n_units = n_jobs * upj;
  const long long x_bytes = ((static_cast<long long>(MB) * K * 2 + 127) / 128) * 128;
  __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  uint8_t* ring = smem_raw + x_bytes + static_cast<long long>(warp) * S * (kTmaPiece * 2);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + x_bytes + static_cast<long long>(kTmaWarps) * S * (kTmaPiece * 2)) + warp * S;
  auto row_of = [](int n) -> long long {
    return (MODE == kModeSwiGLU) ? static_cast<long long>(n / 32) * 64 + (n % 32) : n;
  };
  // piece u of this warp: job u / upj, weight row (gate, then up for SwiGLU), piece of the row
  auto issue = [&](int u) {
    const int job = u / upj, within = u - job * upj;
    const int rsel = within / cpr, piece = within - rsel * cpr;
    const long long r = row_of(warp_g + job * n_warps) + 32 * rsel;
    const int k0 = piece * kTmaPiece;
    const uint32_t bytes = static_cast<uint32_t>(min(kTmaPiece, K - k0)) * 2u;
    const int slot = u % S;
    mbar_expect_tx(&bars[slot], bytes);
    bulk_load_1d(ring + slot * (kTmaPiece * 2), W + r * ldw + k0, bytes, &bars[slot], kL2EvictFirst);
  };
  if (lane == 0) {
    for (int sidx = 0; sidx < S; ++sidx) mbar_init(&bars[sidx], 1);
    fence_barrier_init();
    fence_proxy_async();
    // weights do not depend on the predecessor: the first S pieces are on their way before the dependency wait
    for (int u = 0; u < S && u < n_units; ++u) issue(u);
  }
  griddep_launch_dependents();
  griddep_wait();  // activations / residual written by the predecessor are complete and visible from here on
  {
    const int kv = K / 8;
    for (int i = threadIdx.x; i < MB * kv; i += kTmaWarps * 32) {
      const int m = i / kv, c = i - m * kv;
      reinterpret_cast<uint4*>(xs)[i] = __ldg(reinterpret_cast<const uint4*>(x + (m < M ? m : M - 1) * ldx + c * 8));
    }
  }
  __syncthreads();

  WsAcc acc[MB], acc2[MB];
  for (int u = 0; u < n_units; ++u) {
    const int job = u / upj, within = u - job * upj;
    const int rsel = within / cpr, piece = within - rsel * cpr;
    const int slot = u % S;
    if (within == 0) {
#pragma unroll
      for (int m = 0; m < MB; ++m) {
        wstream_zero(acc[m]);
        wstream_zero(acc2[m]);
      }
    }
    mbar_wait(&bars[slot], (u / S) & 1);
    const int k0 = piece * kTmaPiece;
    const int n16 = min(kTmaPiece, K - k0) / 8;
    const uint4* wp = reinterpret_cast<const uint4*>(ring + slot * (kTmaPiece * 2));
    const __nv_bfloat16* xk = xs + k0;
    WsAcc* a = (MODE == kModeSwiGLU && rsel == 1) ? acc2 : acc;
#pragma unroll 4
    for (int i = lane; i < n16; i += 32) {
      const uint4 wv = wp[i];
#pragma unroll
      for (int m = 0; m < MB; ++m) wstream_fma8(wv, *reinterpret_cast<const uint4*>(xk + m * K + i * 8), a[m]);
    }
    // the slot has been read by every lane (generic proxy) before the bulk copy (async proxy) refills it
    __syncwarp();
    if (lane == 0 && u + S < n_units) {
      fence_proxy_async();
      issue(u + S);
    }
    if (within == upj - 1) {
      float r[MB], r2[MB];
#pragma unroll
      for (int m = 0; m < MB; ++m) {
        r[m] = wstream_combine(acc[m]);
        r2[m] = (MODE == kModeSwiGLU) ? wstream_combine(acc2[m]) : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          r[m] += __shfl_xor_sync(0xffffffffu, r[m], o);
          if (MODE == kModeSwiGLU) r2[m] += __shfl_xor_sync(0xffffffffu, r2[m], o);
        }
      }
      if (lane == 0) wstream_epilogue<MB, MODE>(epi, M, warp_g + job * n_warps, r, r2);
    }
  }
}

template <int MB, int MODE>
static int gemv_tma_launch_t(const __nv_bfloat16* X, long long ldx, const __nv_bfloat16* Wp, long long ldw, int M, int N,
                             int K, const GemmEpi& epi, cudaStream_t st) {
  const int n_out = (MODE == kModeSwiGLU) ? N / 2 : N;
  const long long x_bytes = ((static_cast<long long>(MB) * K * 2 + 127) / 128) * 128;
  const long long budget = 227LL * 1024 - x_bytes - kTmaWarps * 8 * 8 - 2048;   // 2 KB: static shared memory
  int S = static_cast<int>(budget / (static_cast<long long>(kTmaWarps) * kTmaPiece * 2));
  if (S > 4) S = 4;
  if (S < 2) return -2;                                          // does not fit: caller falls back to the register kernel
  const size_t smem = static_cast<size_t>(x_bytes + static_cast<long long>(kTmaWarps) * S * kTmaPiece * 2 + kTmaWarps * S * 8);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncAttributes fa;
    CUDA_TRY(cudaFuncGetAttributes(&fa, ovla_wstream_tma_kernel<MB, MODE>));
    CUDA_TRY(cudaFuncSetAttribute(ovla_wstream_tma_kernel<MB, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  227 * 1024 - static_cast<int>(fa.sharedSizeBytes)));
    attr_set = true;
  }
  int blocks = (n_out + kTmaWarps - 1) / kTmaWarps;
  if (blocks > num_sms()) blocks = num_sms();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(blocks);
  cfg.blockDim = dim3(kTmaWarps * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr_pdl[1];
  attr_pdl[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr_pdl[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr_pdl;
  cfg.numAttrs = (use_pdl() && pdl_enabled()) ? 1 : 0;
  CUDA_TRY(cudaLaunchKernelEx(&cfg, ovla_wstream_tma_kernel<MB, MODE>, X, ldx, Wp, ldw, M, N, K, epi, S));
  count_launch();
  return 0;
}

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return (e && e[0]) ? atoi(e) : dflt;
}
// tuning knobs (defaults chosen from tools/gemv_microbench.py on B200; see profiles/)
static int gemv_threads() { static int v = env_int("OVLA_GEMV_THREADS", kGemvThreadsDefault); return v; }
static int gemv_ch() { static int v = env_int("OVLA_GEMV_CH", 4); return v; }
static int gemv_bps() { static int v = env_int("OVLA_GEMV_BPS", 16); return v; }

static bool use_pdl() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("OVLA_GEMV_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

template <int MB, int MODE, int kCh>
static int gemv_launch_t(const __nv_bfloat16* X, long long ldx, const __nv_bfloat16* Wp, long long ldw, int M, int N,
                         int K, const GemmEpi& epi, int blocks, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(blocks);
  cfg.blockDim = dim3(gemv_threads());
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr_pdl[1];
  attr_pdl[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr_pdl[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr_pdl;
  cfg.numAttrs = (use_pdl() && pdl_enabled()) ? 1 : 0;
  CUDA_TRY(cudaLaunchKernelEx(&cfg, ovla_wstream_kernel<MB, MODE, kCh>, X, ldx, Wp, ldw, M, N, K, epi));
  count_launch();
  return 0;
}

template <int MODE>
static int gemv_dispatch(const void* x, long long ldx, const void* W, long long ldw, int M, int N, int K,
                         const GemmEpi& epi, cudaStream_t st) {
  const int n_out = (MODE == kModeSwiGLU) ? N / 2 : N;
  const int wpb = gemv_threads() / 32;
  int blocks = (n_out + wpb - 1) / wpb;
  const int max_blocks = num_sms() * gemv_bps();
  if (blocks > max_blocks) blocks = max_blocks;
  auto X = static_cast<const __nv_bfloat16*>(x);
  auto Wp = static_cast<const __nv_bfloat16*>(W);
  ProfScope prof(kCatGemv, 2.0 * M * N * K, 2.0 * N * K + 2.0 * M * (K + n_out), st);
  static const int use_tma = env_int("OVLA_GEMV_TMA", 1);
  if (use_tma && M <= 4) {   // shared-memory ring fed by cp.async.bulk; -2 = x + ring do not fit, use the register kernel
    int rc;
    if (M <= 1) rc = gemv_tma_launch_t<1, MODE>(X, ldx, Wp, ldw, M, N, K, epi, st);
    else if (M <= 2) rc = gemv_tma_launch_t<2, MODE>(X, ldx, Wp, ldw, M, N, K, epi, st);
    else rc = gemv_tma_launch_t<4, MODE>(X, ldx, Wp, ldw, M, N, K, epi, st);
    if (rc != -2) return rc;
  }
  const int ch = gemv_ch();
#define OVLA_GEMV_CASE(MBV)                                                                          \
  {                                                                                                  \
    if (ch == 2) return gemv_launch_t<MBV, MODE, 2>(X, ldx, Wp, ldw, M, N, K, epi, blocks, st);      \
    if (ch == 8) return gemv_launch_t<MBV, MODE, 8>(X, ldx, Wp, ldw, M, N, K, epi, blocks, st);      \
    return gemv_launch_t<MBV, MODE, 4>(X, ldx, Wp, ldw, M, N, K, epi, blocks, st);                   \
  }
  if (M <= 1) OVLA_GEMV_CASE(1)
  if (M <= 2) OVLA_GEMV_CASE(2)
  if (M <= 4) OVLA_GEMV_CASE(4)
  OVLA_GEMV_CASE(8)
#undef OVLA_GEMV_CASE
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
