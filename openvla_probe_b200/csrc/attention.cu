// Attention kernels.
//  * flash_attn_kernel: tiled online-softmax attention for the 256/261-token bidirectional ViT blocks (head_dim 64
//    and 72) and the causal Llama prefill (head_dim 128).  One CTA = BR (64 or 128) query rows of one (batch, head),
//    one warp per 16 rows, so a 128-row CTA shares each K/V tile between 8 warps; K/V tiles
//    of 64 keys are double-buffered in shared memory with cp.async; S = QK^T and O += PV run on the warp-level
//    tensor-core path (mma.sync m16n8k16 bf16, fp32 accumulate).  Attention is < 1 % of the path's flops
//    (21.8 of 4157 GFLOP per action), so this kernel is sized for correctness and low traffic; the
//    tcgen05 budget goes to the GEMMs that carry 99 % of the work.
//  (the cached-decode attention lives in decode.cu, fused with RoPE and the KV append)
// Softmax is computed in fp32 and probabilities are rounded to bf16 before the PV product, as flash-attn does.
#include <stdlib.h>

#include "host_util.h"
#include "ops.h"
#include "ptx.cuh"

namespace ovla {

__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid) {
  const uint32_t d = smem_u32(dst);
  const int sz = valid ? 16 : 0;  // src-size 0 => zero-fill the 16 bytes
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma_bf16_16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct AttnStrides {
  long long q_b, q_t, q_h;  // element strides of Q[b][t][h][:]
  long long k_b, k_t, k_h;
  long long v_b, v_t, v_h;
  long long o_b, o_t, o_h;
};

// HD: real head dim; HDP: head dim padded to a multiple of 16 (zero-filled in smem)
// LITE: single-buffered K/V tiles + Q fragments re-read from shared memory: 52 KB and <= 128 registers per 64-row
// CTA at head_dim 128, so four CTAs (16 warps) share an SM and hide each other's softmax / load phases.
template <int HD, int HDP, bool CAUSAL, int BR, bool LITE>
__global__ void __launch_bounds__(BR * 2, LITE ? 4 : 1) flash_attn_kernel(const __nv_bfloat16* __restrict__ Q,
                                                         const __nv_bfloat16* __restrict__ K,
                                                         const __nv_bfloat16* __restrict__ V,
                                                         __nv_bfloat16* __restrict__ O, AttnStrides st, int Tq, int Tk,
                                                         float scale_log2) {
  constexpr int LD = HDP + 8;           // padded smem row (elements): conflict-free ldmatrix
  constexpr int CH = HDP / 8;           // 16-byte chunks per row
  constexpr int KS = HDP / 16;          // k-steps of QK^T
  constexpr int NT = HDP / 8;           // n-tiles of the output
  extern __shared__ __align__(16) uint8_t smem_attn[];
  constexpr int NTHR = BR * 2;
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_attn);  // [BR][LD]
  __nv_bfloat16* sK = sQ + BR * LD;                                  // [2][64][LD]
  constexpr int NBUF = LITE ? 1 : 2;
  __nv_bfloat16* sV = sK + NBUF * 64 * LD;                           // [NBUF][64][LD]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * BR, h = blockIdx.y, b = blockIdx.z;
  const __nv_bfloat16* Qb = Q + b * st.q_b + h * st.q_h;
  const __nv_bfloat16* Kb = K + b * st.k_b + h * st.k_h;
  const __nv_bfloat16* Vb = V + b * st.v_b + h * st.v_h;

  auto load_tile = [&](__nv_bfloat16* dst, const __nv_bfloat16* src, long long tstride, int t0, int tmax, int rows) {
    for (int i = tid; i < rows * CH; i += NTHR) {
      const int r = i / CH, c = (i % CH) * 8;
      const bool ok = (t0 + r < tmax) && (c < HD);
      const __nv_bfloat16* g = src + static_cast<long long>(ok ? t0 + r : 0) * tstride + (ok ? c : 0);
      cp_async16(dst + r * LD + c, g, ok);
    }
  };

  int n_kt = (Tk + 63) / 64;
  if (CAUSAL) {
    const int last_q = min(q0 + BR - 1, Tq - 1);
    n_kt = min(n_kt, last_q / 64 + 1);
  }
  load_tile(sQ, Qb, st.q_t, q0, Tq, BR);
  load_tile(sK, Kb, st.k_t, 0, Tk, 64);
  load_tile(sV, Vb, st.v_t, 0, Tk, 64);
  cp_async_commit();

  uint32_t qf[LITE ? 1 : KS][4];  // LITE re-reads the Q fragments from shared memory instead of pinning 32 registers
  float o_acc[NT][4];
#pragma unroll
  for (int n = 0; n < NT; ++n) o_acc[n][0] = o_acc[n][1] = o_acc[n][2] = o_acc[n][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  const int g = lane >> 2, tq = lane & 3;
  const int qrow0 = q0 + warp * 16 + g;  // this thread's two query rows: qrow0, qrow0 + 8

  for (int kt = 0; kt < n_kt; ++kt) {
    const int buf = LITE ? 0 : (kt & 1);
    if (LITE) {
      if (kt > 0) {  // the barrier that ended the previous iteration freed the single buffer
        load_tile(sK, Kb, st.k_t, kt * 64, Tk, 64);
        load_tile(sV, Vb, st.v_t, kt * 64, Tk, 64);
        cp_async_commit();
      }
      cp_async_wait<0>();
    } else if (kt + 1 < n_kt) {
      load_tile(sK + (buf ^ 1) * 64 * LD, Kb, st.k_t, (kt + 1) * 64, Tk, 64);
      load_tile(sV + (buf ^ 1) * 64 * LD, Vb, st.v_t, (kt + 1) * 64, Tk, 64);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (!LITE && kt == 0) {
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const int r = warp * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
        const int c = ks * 16 + 8 * (lane >> 4);
        ldsm_x4(qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], sQ + r * LD + c);
      }
    }
    const __nv_bfloat16* k_s = sK + buf * 64 * LD;
    const __nv_bfloat16* v_s = sV + buf * 64 * LD;

    // a warp whose 16 rows all lie above this key tile has nothing to do here (causal); it only keeps the barriers
    if (!CAUSAL || kt * 64 <= q0 + warp * 16 + 15) {
      // ---- S = Q K^T for 64 keys: 8 n-tiles of 8 keys
      float s[8][4];
  #pragma unroll
      for (int n = 0; n < 8; ++n) s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
  #pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        if (LITE) {
          const int r = warp * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
          const int c = ks * 16 + 8 * (lane >> 4);
          ldsm_x4(qf[0][0], qf[0][1], qf[0][2], qf[0][3], sQ + r * LD + c);
        }
        const uint32_t* qa = qf[LITE ? 0 : ks];
  #pragma unroll
        for (int np = 0; np < 4; ++np) {  // two key n-tiles per ldmatrix.x4
          uint32_t b0, b1, b2, b3;
          const int r = np * 16 + (lane & 7) + 8 * (lane >> 4);
          const int c = ks * 16 + 8 * ((lane >> 3) & 1);
          ldsm_x4(b0, b1, b2, b3, k_s + r * LD + c);
          mma_bf16_16816(s[2 * np], qa, b0, b1);
          mma_bf16_16816(s[2 * np + 1], qa, b2, b3);
        }
      }
      // ---- mask + online softmax (rows qrow0 and qrow0+8; this thread holds cols n*8 + 2*tq + {0,1})
      const int key0 = kt * 64;
      float mx[2] = {-INFINITY, -INFINITY};
  #pragma unroll
      for (int n = 0; n < 8; ++n) {
  #pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int key = key0 + n * 8 + 2 * tq + (j & 1);
          const int qr = qrow0 + 8 * (j >> 1);
          const bool ok = key < Tk && (!CAUSAL || key <= qr);
          s[n][j] = ok ? s[n][j] * scale_log2 : -INFINITY;
          mx[j >> 1] = fmaxf(mx[j >> 1], s[n][j]);
        }
      }
      float corr[2], m_new[2];
  #pragma unroll
      for (int r = 0; r < 2; ++r) {
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
        m_new[r] = fmaxf(m_run[r], mx[r]);
        const float m_use = (m_new[r] == -INFINITY) ? 0.f : m_new[r];
        corr[r] = exp2f(m_run[r] - m_use);  // m_run = -inf -> 0
        m_run[r] = m_new[r];
        m_new[r] = m_use;
      }
      float rs[2] = {0.f, 0.f};
      uint32_t pf[4][4];  // P as A fragments: 4 k-steps of 16 keys
  #pragma unroll
      for (int n = 0; n < 8; ++n) {
        const float p0 = exp2f(s[n][0] - m_new[0]), p1 = exp2f(s[n][1] - m_new[0]);
        const float p2 = exp2f(s[n][2] - m_new[1]), p3 = exp2f(s[n][3] - m_new[1]);
        const uint32_t lo = pack_bf16(p0, p1), hi = pack_bf16(p2, p3);
        const float2 l2 = unpack_bf16(lo), h2 = unpack_bf16(hi);  // row sums use the rounded probabilities
        rs[0] += l2.x + l2.y;
        rs[1] += h2.x + h2.y;
        pf[n >> 1][(n & 1) * 2 + 0] = lo;
        pf[n >> 1][(n & 1) * 2 + 1] = hi;
      }
  #pragma unroll
      for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * corr[r] + rs[r];
  #pragma unroll
      for (int n = 0; n < NT; ++n) {
        o_acc[n][0] *= corr[0];
        o_acc[n][1] *= corr[0];
        o_acc[n][2] *= corr[1];
        o_acc[n][3] *= corr[1];
      }
      // ---- O += P V
  #pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
  #pragma unroll
        for (int np = 0; np < NT / 2; ++np) {
          uint32_t b0, b1, b2, b3;
          const int r = kk * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
          const int c = np * 16 + 8 * (lane >> 4);
          ldsm_x4_t(b0, b1, b2, b3, v_s + r * LD + c);
          mma_bf16_16816(o_acc[2 * np], pf[kk], b0, b1);
          mma_bf16_16816(o_acc[2 * np + 1], pf[kk], b2, b3);
        }
      }
    }
    __syncthreads();  // everyone is done with `buf` before the next iteration's prefetch overwrites it
  }

  // ---- normalise and write (stage the warp's 16 x HD tile through its own rows of sQ for 16-byte stores)
  float inv[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    inv[r] = l_run[r] > 0.f ? 1.f / l_run[r] : 0.f;
  }
  __nv_bfloat16* sO = sQ + warp * 16 * LD;
#pragma unroll
  for (int n = 0; n < NT; ++n) {
    *reinterpret_cast<uint32_t*>(sO + g * LD + n * 8 + 2 * tq) = pack_bf16(o_acc[n][0] * inv[0], o_acc[n][1] * inv[0]);
    *reinterpret_cast<uint32_t*>(sO + (g + 8) * LD + n * 8 + 2 * tq) =
        pack_bf16(o_acc[n][2] * inv[1], o_acc[n][3] * inv[1]);
  }
  __syncwarp();
  __nv_bfloat16* Ob = O + b * st.o_b + h * st.o_h;
  constexpr int OCH = HD / 8;
  for (int i = lane; i < 16 * OCH; i += 32) {
    const int r = i / OCH, c = (i % OCH) * 8;
    const int t = q0 + warp * 16 + r;
    if (t < Tq) *reinterpret_cast<uint4*>(Ob + t * st.o_t + c) = *reinterpret_cast<const uint4*>(sO + r * LD + c);
  }
}

template <int HD, int HDP, bool CAUSAL, int BR, bool LITE>
static int flash_launch_t(const void* Q, const void* K, const void* V, void* O, const AttnStrides& s, int B, int H,
                          int Tq, int Tk, cudaStream_t st) {
  constexpr int LD = HDP + 8;
  constexpr int smem = (BR + (LITE ? 2 : 4) * 64) * LD * 2;
  auto kern = flash_attn_kernel<HD, HDP, CAUSAL, BR, LITE>;
  static bool attr = false;
  if (!attr) {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr = true;
  }
  const float scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(HD));
  dim3 grid((Tq + BR - 1) / BR, H, B);
  ProfScope prof(kCatFlash, (CAUSAL ? 2.0 : 4.0) * B * H * Tq * Tk * HD, 2.0 * B * H * HD * (2.0 * Tq + 2.0 * Tk), st);
  kern<<<grid, BR * 2, smem, st>>>(static_cast<const __nv_bfloat16*>(Q), static_cast<const __nv_bfloat16*>(K),
                               static_cast<const __nv_bfloat16*>(V), static_cast<__nv_bfloat16*>(O), s, Tq, Tk,
                               scale_log2);
  CUDA_TRY(cudaGetLastError());
  count_launch();
  return 0;
}

// Q/K/V/O element strides (batch, token, head); head vectors contiguous and 16-byte aligned.
int flash_attn_launch(const void* Q, const void* K, const void* V, void* O, const long long* strides12, int B, int H,
                      int Tq, int Tk, int head_dim, int causal, cudaStream_t st) {
  if (B <= 0 || H <= 0 || Tq <= 0) return 0;
  if (Tk <= 0) return set_error("attention: empty key range");
  AttnStrides s;
  s.q_b = strides12[0]; s.q_t = strides12[1]; s.q_h = strides12[2];
  s.k_b = strides12[3]; s.k_t = strides12[4]; s.k_h = strides12[5];
  s.v_b = strides12[6]; s.v_t = strides12[7]; s.v_h = strides12[8];
  s.o_b = strides12[9]; s.o_t = strides12[10]; s.o_h = strides12[11];
  for (int i = 0; i < 12; ++i)
    if (strides12[i] % 8) return set_error("attention: strides must be multiples of 8 elements (16 bytes)");
  // 64-row CTAs (2 per SM) beat 128-row CTAs (1 per SM at head_dim 128: 170 registers x 256 threads) by ~20 % on
  // B200 (profiles/r01 A/B); OVLA_ATTN_BR=128 selects the larger CTA for measurements
  static int br = -1;
  if (br < 0) { const char* ev = getenv("OVLA_ATTN_BR"); br = (ev && atoi(ev) == 128) ? 128 : 64; }
  static int lite = -1;  // OVLA_ATTN_LITE=0 selects the double-buffered 2-CTA/SM variant
  if (lite < 0) { const char* ev = getenv("OVLA_ATTN_LITE"); lite = (ev && ev[0] == '0') ? 0 : 1; }
  const bool big = br == 128 && Tq > 64;
#define OVLA_ATTN(HDv, HDPv, Cv)                                                                         \
  return big ? flash_launch_t<HDv, HDPv, Cv, 128, false>(Q, K, V, O, s, B, H, Tq, Tk, st)                \
       : lite ? flash_launch_t<HDv, HDPv, Cv, 64, true>(Q, K, V, O, s, B, H, Tq, Tk, st)                 \
              : flash_launch_t<HDv, HDPv, Cv, 64, false>(Q, K, V, O, s, B, H, Tq, Tk, st)
  if (head_dim == 64 && !causal) OVLA_ATTN(64, 64, false);
  if (head_dim == 72 && !causal) OVLA_ATTN(72, 80, false);
  if (head_dim == 128 && causal) OVLA_ATTN(128, 128, true);
  if (head_dim == 128 && !causal) OVLA_ATTN(128, 128, false);
  if (head_dim == 64 && causal) OVLA_ATTN(64, 64, true);
#undef OVLA_ATTN
  return set_error("attention: unsupported head_dim=%d causal=%d", head_dim, causal);
}

}  // namespace ovla
