// Split-K for the weight-streaming GEMM shapes (M <= 512: bs=1 prefill, batched decode steps).
// With few output tiles a plain tiling leaves most SMs idle and each active CTA limited by its own shared-memory
// fill rate (~100 GB/s per SM); slicing K gives every SM a (tile, K-slice) pair.  Each slice writes its raw fp32
// partial tile (kModePartial) into a workspace [S, M, N]; this kernel sums the S partials IN SLICE ORDER
// (deterministic) and applies the same epilogue / rounding points as the fused GEMM epilogue.
#include "host_util.h"
#include "ops.h"
#include "ptx.cuh"

#include <map>
#include <mutex>

namespace ovla {

// Workspaces are OWNED BY THE CALLER of gemm_launch (SplitKWs): an OvlaEngine allocates one per stream it issues GEMMs
// on at ovla_create time (engine.cu), so two engines -- or two host threads -- on one device never share partial tiles
// and nothing is allocated under stream capture.  The stand-alone ovla_gemm entry has no engine: it uses one lazily
// allocated buffer per (device, stream), created outside capture only (a capturing stream without a buffer simply does
// not split).
static constexpr long long kStandaloneWsFloats = 48LL << 20;  // 192 MB

namespace {
struct WsKey {
  int dev;
  cudaStream_t st;
  bool operator<(const WsKey& o) const { return dev != o.dev ? dev < o.dev : st < o.st; }
};
std::mutex g_ws_mu;
std::map<WsKey, float*> g_ws;
}  // namespace

SplitKWs splitk_stream_workspace(cudaStream_t st) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return {nullptr, 0};
  std::lock_guard<std::mutex> lk(g_ws_mu);
  auto it = g_ws.find(WsKey{dev, st});
  if (it != g_ws.end()) return {it->second, kStandaloneWsFloats};
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) {
    cudaGetLastError();
    return {nullptr, 0};
  }
  float* p = nullptr;
  if (cudaMalloc(&p, kStandaloneWsFloats * sizeof(float)) != cudaSuccess) {
    cudaGetLastError();
    return {nullptr, 0};
  }
  g_ws[WsKey{dev, st}] = p;
  return {p, kStandaloneWsFloats};
}

// one thread = one row x 8 consecutive OUTPUT columns
template <int MODE>
__global__ void __launch_bounds__(256) splitk_epilogue_kernel(const float* __restrict__ ws, long long slice_stride,
                                                              long long ldw, int S, int M, int N, GemmEpi epi) {
  griddep_launch_dependents();
  griddep_wait();
  const int n_out = (MODE == kModeSwiGLU) ? N / 2 : N;
  const int cv = n_out / 8;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(M) * cv) return;
  const int row = static_cast<int>(idx / cv), j = static_cast<int>(idx % cv) * 8;
  auto sum8 = [&](int col, float* a) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = 0.f;
    const float* p = ws + static_cast<long long>(row) * ldw + col;
    for (int s = 0; s < S; ++s) {
      const float4 u = *reinterpret_cast<const float4*>(p + s * slice_stride);
      const float4 v = *reinterpret_cast<const float4*>(p + s * slice_stride + 4);
      a[0] += u.x; a[1] += u.y; a[2] += u.z; a[3] += u.w;
      a[4] += v.x; a[5] += v.y; a[6] += v.z; a[7] += v.w;
    }
  };
  if (MODE == kModeBf16) {
    float x[8];
    sum8(j, x);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (epi.bias) x[i] += __bfloat162float(epi.bias[j + i]);
      x[i] = bf16_round(x[i]);
    }
    if (epi.gelu) {   // the same pairwise GELU as the fused GEMM epilogue: split and unsplit results are bit-identical
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        gelu_erf_x2(x[i], x[i + 1]);
        x[i] = bf16_round(x[i]);
        x[i + 1] = bf16_round(x[i + 1]);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (epi.scale) x[i] = bf16_round(x[i] * __bfloat162float(epi.scale[j + i]));
      if (epi.resid) x[i] += __bfloat162float(epi.resid[static_cast<long long>(row) * epi.ldr + j + i]);
    }
    uint4 o;
    o.x = pack_bf16(x[0], x[1]); o.y = pack_bf16(x[2], x[3]); o.z = pack_bf16(x[4], x[5]); o.w = pack_bf16(x[6], x[7]);
    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(epi.out) + static_cast<long long>(row) * epi.ldo + j) = o;
  } else if (MODE == kModeSwiGLU) {
    // W rows are interleaved [32 gate | 32 up]: output column j lives in block j / 32
    const int gcol = (j / 32) * 64 + (j % 32);
    float g[8], u[8], y[8];
    sum8(gcol, g);
    sum8(gcol + 32, u);
#pragma unroll
    for (int i = 0; i < 8; ++i) y[i] = bf16_round(silu(bf16_round(g[i]))) * bf16_round(u[i]);
    uint4 o;
    o.x = pack_bf16(y[0], y[1]); o.y = pack_bf16(y[2], y[3]); o.z = pack_bf16(y[4], y[5]); o.w = pack_bf16(y[6], y[7]);
    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(epi.out) + static_cast<long long>(row) * epi.ldo + j) = o;
  } else {  // kModeF32
    float x[8];
    sum8(j, x);
    float* out = reinterpret_cast<float*>(epi.out) + static_cast<long long>(row) * epi.ldo + j;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (epi.bias_f32) x[i] += epi.bias_f32[j + i];
      if (epi.bias) x[i] += __bfloat162float(epi.bias[j + i]);
      if (epi.round_bf16) x[i] = bf16_round(x[i]);
    }
    *reinterpret_cast<float4*>(out) = make_float4(x[0], x[1], x[2], x[3]);
    *reinterpret_cast<float4*>(out + 4) = make_float4(x[4], x[5], x[6], x[7]);
  }
}

int splitk_epilogue_launch(int mode, const float* ws, long long slice_stride, long long ldw, int S, int M, int N,
                           const GemmEpi& epi, cudaStream_t st) {
  const int n_out = (mode == kModeSwiGLU) ? N / 2 : N;
  if (n_out % 8) return set_error("split-K epilogue: output width %d must be a multiple of 8", n_out);
  const long long total = static_cast<long long>(M) * (n_out / 8);
  const dim3 grid(static_cast<unsigned>((total + 255) / 256)), block(256);
  ProfScope prof(kCatOther, 0.0, 4.0 * S * M * N + 2.0 * M * n_out, st);
  if (mode == kModeBf16) CUDA_TRY(launch_pdl(splitk_epilogue_kernel<kModeBf16>, grid, block, 0, st, ws, slice_stride, ldw, S, M, N, epi));
  else if (mode == kModeSwiGLU) CUDA_TRY(launch_pdl(splitk_epilogue_kernel<kModeSwiGLU>, grid, block, 0, st, ws, slice_stride, ldw, S, M, N, epi));
  else if (mode == kModeF32) CUDA_TRY(launch_pdl(splitk_epilogue_kernel<kModeF32>, grid, block, 0, st, ws, slice_stride, ldw, S, M, N, epi));
  else return set_error("split-K epilogue: unsupported mode %d", mode);
  count_launch();
  return 0;
}

}  // namespace ovla
