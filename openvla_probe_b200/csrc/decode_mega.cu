// One cached decode step of the Llama stack as ONE persistent kernel (batch <= 4 rows): the reference's
// modeling_prismatic.py:325-341 -> LlamaForCausalLM single-token forward with past_key_values.
//
// Why: at batch 1 a decode step streams every weight matrix exactly once (13.4 GB) and is nothing else, but as ~230
// dependent kernels each boundary costs ~5 us of idle HBM (tail of one grid, launch + ramp-up of the next, the small
// norm / attention kernels in between): 3.4-3.6 ms per step against 2.05 ms of pure streaming at the measured HBM peak.
// Here one CTA per SM stays resident for the whole step and the weight stream never stops:
//   * every consumer warp owns a ring of S shared-memory slots of 8 KB; a slot is filled by ONE cp.async.bulk (a
//     4096-element piece of a weight row, L2 evict-first) that completes on the slot's "full" mbarrier.  The pieces a warp
//     will need are a fixed sequence over (layer, linear, output column, piece) known up front, so a ninth PRODUCER warp
//     (lane w serves ring w: waits for the slot's "empty" mbarrier, issues the copy, advances its iterator) refills the
//     rings ACROSS phase boundaries: while a grid barrier or the attention phase holds the consumers, 128-192 KB per SM
//     (19-28 MB over the chip, 3-4 us of HBM time) of the next linear's weights are in flight or landed.  Keeping the
//     address arithmetic and the copy issue off the consumers matters: measured on the timeline (tools/decode_trace.py),
//     a consumer that also refills its ring spends 0.75 us consuming + 0.6 us issuing per 8 KB piece, i.e. exactly the
//     HBM pace (1.5 us per piece per warp) with no slack to catch up after a barrier;
//   * RMSNorm is not a phase: every CTA recomputes the (identical) row statistics of the 8 KB residual row while it
//     stages the activation vector into shared memory;
//   * phases are separated by a grid barrier (one atomic per CTA, acquire/release at gpu scope); activations written by
//     other SMs are read with L2-only loads (L1 is not coherent across SMs).
// Per layer: [norm1 + QKV] | RoPE + KV append + attention | [o_proj + residual] | [norm2 + gate/up + SwiGLU] |
// [down + residual]; then [final norm + lm_head].  Summation orders and rounding points are those of the multi-kernel
// path (norm.cu rmsnorm_split_kernel, gemv.cu, decode_attn.cuh), so both paths return the same bits -- tested.
#include <algorithm>
#include <type_traits>

#include "decode_attn.cuh"
#include "gemm.cuh"
#include "host_util.h"
#include "ops.h"
#include "ptx.cuh"

#ifndef OVLA_MEGA_STAGE_BATCH
#define OVLA_MEGA_STAGE_BATCH 6      // A/B knobs (tools/decode_trace.py on variant builds)
#endif
#ifndef OVLA_MEGA_LDS_GROUP
#define OVLA_MEGA_LDS_GROUP 8
#endif
#ifndef OVLA_MEGA_PIECE_TRACE
#define OVLA_MEGA_PIECE_TRACE 0      // 1: per-piece (wait, consume) stamps of warp 0 in layer 1 for tools/decode_trace.py
#endif
#ifndef OVLA_MEGA_NORM_PREFETCH
#define OVLA_MEGA_NORM_PREFETCH 0      // measured: no gain (qkv.stage +0.4 us), off
#endif

namespace ovla {

namespace {

constexpr int kWarps = 8;                  // == kDecThreads / 32 (the attention body needs 256 threads)
constexpr int kPiece = 4096;               // elements per ring slot (8 KB)
constexpr int kPieceBytes = kPiece * 2;

enum : int { kLinQkv = 0, kLinO = 1, kLinGateUp = 2, kLinDown = 3, kLinPerLayer = 4 };

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add_u32(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// all CTAs of the (co-resident, cooperative) grid; `target` counts arrivals since the counter was zeroed
__device__ __forceinline__ void cons_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }   // the 8 consumer warps

__device__ __forceinline__ void grid_barrier(unsigned* ctr, unsigned& target) {
  cons_sync();
  target += gridDim.x;
  if (threadIdx.x == 0) {
    red_release_add_u32(ctr, 1u);
    const long long t0 = clock64();
    while (ld_acquire_u32(ctr) < target) {
      if (clock64() - t0 > 4000000000LL) __trap();      // a lost CTA is a bug: trap instead of hanging the GPU
    }
  }
  cons_sync();
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// the weight pieces of one warp, in consumption order
struct PieceIter {
  int lin;        // global linear index: layer * 4 + {qkv, o, gate_up, down}; n_layers * 4 = lm_head; beyond = done
  int job;        // index into this warp's output columns of the linear
  int within;     // piece within the job (rows_per_job * pieces_per_row)
  // cached parameters of `lin`
  const __nv_bfloat16* W;
  int K, cpr, upj, n_jobs, swiglu;
};

}  // namespace

template <int MB>
__global__ void __launch_bounds__(kWarps * 32 + 32, 1) decode_step_kernel(const DecodeStepArgs a) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ float s_part[MB][4];
  __shared__ float s_sq[128];
  __shared__ float s_red[kWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // (Measured and dropped, round 2: warp-major numbering of the global warps, which spreads the warps that own one output
  // column more than the others -- N = 4096 on 1184 warps: 544 warps with 4 columns, 640 with 3 -- evenly over the CTAs
  // instead of giving them all to CTAs 0-67.  The step did not move (2849 vs 2842 us): a phase lasts as long as ONE warp
  // needs for its 4 columns at its latency-bound pace, whatever its neighbours on the SM do; evening that out needs a
  // split along K, i.e. another summation order.)
  const int warp_g = blockIdx.x * kWarps + warp;
  const int n_warps = gridDim.x * kWarps;
  const int D = a.D, I = a.I, S = a.S;
  const int n_lin = a.n_layers * kLinPerLayer + 1;
  // shared memory: [x: MB * max(D, I) bf16] [attention scratch: kDecGroups * 128 floats] [ring] [mbarriers]
  const long long x_bytes = ((static_cast<long long>(MB) * (I > D ? I : D) * 2 + 127) / 128) * 128;
  __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  float* dyn = reinterpret_cast<float*>(smem_raw + x_bytes);
  const long long attn_bytes = static_cast<long long>(a.attn_floats) * 4;
  uint8_t* ring_base = smem_raw + x_bytes + attn_bytes;
  uint64_t* bar_base = reinterpret_cast<uint64_t*>(ring_base + static_cast<long long>(kWarps) * S * kPieceBytes);
  // full[w][s] at bar_base[(w * S + s) * 2], empty[w][s] right after it

  auto lin_params = [&](PieceIter& it, int wg) {
    // skip linears in which this warp owns no output column
    while (it.lin < n_lin) {
      int N;
      if (it.lin == n_lin - 1) { it.W = a.lm_head; N = a.vocab; it.K = D; it.swiglu = 0; }
      else {
        const DecodeLayerPtrs& l = a.layers[it.lin / kLinPerLayer];
        const int which = it.lin % kLinPerLayer;
        if (which == kLinQkv) { it.W = l.qkv; N = 3 * D; it.K = D; it.swiglu = 0; }
        else if (which == kLinO) { it.W = l.o; N = D; it.K = D; it.swiglu = 0; }
        else if (which == kLinGateUp) { it.W = l.gate_up; N = I; it.K = D; it.swiglu = 1; }   // N = output columns
        else { it.W = l.down; N = D; it.K = I; it.swiglu = 0; }
      }
      it.cpr = (it.K + kPiece - 1) / kPiece;
      it.upj = (it.swiglu ? 2 : 1) * it.cpr;
      it.n_jobs = wg < N ? (N - wg + n_warps - 1) / n_warps : 0;
      if (it.n_jobs > 0) return;
      ++it.lin;
    }
  };
  auto advance = [&](PieceIter& it, int wg) {
    if (++it.within == it.upj) {
      it.within = 0;
      if (++it.job == it.n_jobs) {
        it.job = 0;
        ++it.lin;
        lin_params(it, wg);
      }
    }
  };
  if (threadIdx.x == 0) {
    for (int i = 0; i < kWarps * S * 2; ++i) mbar_init(&bar_base[i], 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  __syncthreads();   // the only CTA-wide barrier: afterwards consumers use named barrier 1, the producer none

  if (warp == kWarps) {
    // ---- producer warp: lane w streams the piece sequence of consumer warp w into its ring
    if (lane < kWarps) {
      const int wg = blockIdx.x * kWarps + lane;
      uint8_t* ring = ring_base + static_cast<long long>(lane) * S * kPieceBytes;
      uint64_t* bars = bar_base + lane * S * 2;
      PieceIter it = {};
      lin_params(it, wg);
      unsigned u = 0;
      while (it.lin < n_lin) {
        const int slot = static_cast<int>(u % static_cast<unsigned>(S));
        const unsigned fill = u / static_cast<unsigned>(S);
        if (fill > 0) {
          mbar_wait(&bars[slot * 2 + 1], (fill - 1) & 1u);     // the consumer has released the previous fill
          fence_proxy_async();                                  // its generic-proxy reads precede this async-proxy write
        }
        const int n = wg + it.job * n_warps;
        const int rsel = it.within / it.cpr, piece = it.within - rsel * it.cpr;
        const long long r = it.swiglu ? static_cast<long long>(n / 32) * 64 + (n % 32) + 32 * rsel : n;
        const int k0 = piece * kPiece;
        const uint32_t bytes = static_cast<uint32_t>(min(kPiece, it.K - k0)) * 2u;
        mbar_expect_tx(&bars[slot * 2], bytes);
        bulk_load_1d(ring + slot * kPieceBytes, it.W + r * static_cast<long long>(it.K) + k0, bytes, &bars[slot * 2], kL2EvictFirst);
        ++u;
        advance(it, wg);
      }
    }
    return;
  }
  uint8_t* ring = ring_base + static_cast<long long>(warp) * S * kPieceBytes;
  uint64_t* bars = bar_base + warp * S * 2;
  unsigned n_consumed = 0;
  unsigned bar_target = 0;

  // ---- activation staging: xs[m, :K] = src[m, :K], optionally RMS-normalised with `gamma` (same arithmetic and
  // summation order as norm.cu's rmsnorm_split_kernel: four quarter-row partial sums combined in a fixed order)
  auto stage = [&](const __nv_bfloat16* src, long long ld, int K, const __nv_bfloat16* gamma) {
    const int kv = K / 8;
    if (gamma && MB > 1) {
      // several rows: two passes (statistics, then normalise on re-read) keep the register footprint small
      if (warp < 4) {
        const int q = K / 4;
#pragma unroll 1
        for (int m = 0; m < MB; ++m) {
          const __nv_bfloat16* xr = src + static_cast<long long>(m < a.M ? m : a.M - 1) * ld + warp * q;
          float sq = 0.f;
#pragma unroll
          for (int c = 0; c < 5; ++c) {
            const int col = (c * 32 + lane) * 8;
            if (col < q) {
              const uint4 v = __ldcg(reinterpret_cast<const uint4*>(xr + col));
              const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float2 f = unpack_bf16(u[i]);
                sq += f.x * f.x + f.y * f.y;
              }
            }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
          if (lane == 0) s_part[m][warp] = sq;
        }
      }
      cons_sync();
      for (int i = threadIdx.x; i < MB * kv; i += kWarps * 32) {
        const int m = i / kv, c = i - m * kv;
        const uint4 v = __ldcg(reinterpret_cast<const uint4*>(src + static_cast<long long>(m < a.M ? m : a.M - 1) * ld + c * 8));
        const float rstd = rsqrtf((s_part[m][0] + s_part[m][1] + s_part[m][2] + s_part[m][3]) / K + a.eps);
        const uint4 wv = __ldg(reinterpret_cast<const uint4*>(gamma + c * 8));
        const uint32_t u[4] = {v.x, v.y, v.z, v.w};
        const uint32_t wu[4] = {wv.x, wv.y, wv.z, wv.w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = unpack_bf16(u[j]), wf = unpack_bf16(wu[j]);
          o[j] = pack_bf16(wf.x * bf16_round(f.x * rstd), wf.y * bf16_round(f.y * rstd));
        }
        reinterpret_cast<uint4*>(xs)[i] = make_uint4(o[0], o[1], o[2], o[3]);
      }
    } else if (gamma) {
      // one memory round trip: warps 0-3 hold their quarter rows in registers across the block reduction
      const int q = K / 4;
      uint4 v[MB][5], gv[5];
      if (warp < 4) {
#pragma unroll
        for (int c = 0; c < 5; ++c) {
          const int col = (c * 32 + lane) * 8;
          // out-of-range chunks load column 0 and are zeroed: unconditional loads keep v / gv in registers (predicated
          // ones turned them into local-memory arrays), and zeros add nothing to the sum of squares
          gv[c] = __ldg(reinterpret_cast<const uint4*>(gamma + warp * q + (col < q ? col : 0)));
        }
#pragma unroll
        for (int m = 0; m < MB; ++m) {
          const __nv_bfloat16* xr = src + static_cast<long long>(m < a.M ? m : a.M - 1) * ld + warp * q;
#pragma unroll
          for (int c = 0; c < 5; ++c) {
            const int col = (c * 32 + lane) * 8;
            v[m][c] = __ldcg(reinterpret_cast<const uint4*>(xr + (col < q ? col : 0)));
            if (col >= q) v[m][c] = make_uint4(0u, 0u, 0u, 0u);
          }
        }
#pragma unroll
        for (int m = 0; m < MB; ++m) {
          float sq = 0.f;
#pragma unroll
          for (int c = 0; c < 5; ++c) {
            const uint32_t u[4] = {v[m][c].x, v[m][c].y, v[m][c].z, v[m][c].w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 f = unpack_bf16(u[i]);
              sq += f.x * f.x + f.y * f.y;
            }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
          if (lane == 0) s_part[m][warp] = sq;
        }
      }
      cons_sync();
      if (warp < 4) {
#pragma unroll
        for (int m = 0; m < MB; ++m) {
          const float rstd = rsqrtf((s_part[m][0] + s_part[m][1] + s_part[m][2] + s_part[m][3]) / K + a.eps);
#pragma unroll
          for (int c = 0; c < 5; ++c) {
            const int col = (c * 32 + lane) * 8;
            if (col < q) {
              const uint32_t u[4] = {v[m][c].x, v[m][c].y, v[m][c].z, v[m][c].w};
              const uint32_t wu[4] = {gv[c].x, gv[c].y, gv[c].z, gv[c].w};
              uint32_t o[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 f = unpack_bf16(u[j]), wf = unpack_bf16(wu[j]);
                o[j] = pack_bf16(wf.x * bf16_round(f.x * rstd), wf.y * bf16_round(f.y * rstd));
              }
              *reinterpret_cast<uint4*>(xs + m * K + warp * q + col) = make_uint4(o[0], o[1], o[2], o[3]);
            }
          }
        }
      }
    } else {
      // plain copy: the L2 round trips of a thread's loads overlap (issued six at a time), they do not queue up behind
      // each other's shared-memory stores (22 KB of SwiGLU output per row: 2.1 -> 0.6 us on the timeline)
      constexpr int kT = kWarps * 32;
      const int n = MB * kv, rounds = (n + kT - 1) / kT;
      auto copy_rounds = [&](auto nb_tag, int r0) {   // NB rounds of kT 16-byte loads, all issued before the first store
        constexpr int NB = decltype(nb_tag)::value;
        uint4 v[NB];
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          const int i = min((r0 + j) * kT + static_cast<int>(threadIdx.x), n - 1);   // clamped: every load is valid
          const int m = i / kv, c = i - m * kv;
          v[j] = __ldcg(reinterpret_cast<const uint4*>(src + static_cast<long long>(m < a.M ? m : a.M - 1) * ld + c * 8));
        }
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          const int i = (r0 + j) * kT + static_cast<int>(threadIdx.x);
          if (i < n) reinterpret_cast<uint4*>(xs)[i] = v[j];
        }
      };
      int r0 = 0;
      for (; r0 + OVLA_MEGA_STAGE_BATCH <= rounds; r0 += OVLA_MEGA_STAGE_BATCH)
        copy_rounds(std::integral_constant<int, OVLA_MEGA_STAGE_BATCH>{}, r0);
      for (; r0 + 2 <= rounds; r0 += 2) copy_rounds(std::integral_constant<int, 2>{}, r0);
      for (; r0 < rounds; ++r0) copy_rounds(std::integral_constant<int, 1>{}, r0);
    }
    cons_sync();
  };

  // piece-level timeline (debug): lane 0 of warp 0 of the first / last CTA, one layer, [wait begin, wait end, consumed]
  unsigned long long* ptr = nullptr;
  int ptr_n = 0;
  bool ptr_on = false;
  if (a.trace && threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1))
    ptr = a.trace + (blockIdx.x == 0 ? 0 : a.trace_stride) + 512;

  // ---- one linear: this warp's output columns n = warp_g, warp_g + n_warps, ...; epilogue by lane 0
  // mode 0: bf16 out (+ optional in-place residual), 1: SwiGLU, 2: fp32 out of bf16-rounded values (lm_head)
  auto linear = [&](int N, int K, int mode, void* out, long long ldo, const __nv_bfloat16* resid, long long ldr) {
    const int cpr = (K + kPiece - 1) / kPiece;
    const int rows_per_job = mode == 1 ? 2 : 1;
    const int n_jobs = warp_g < N ? (N - warp_g + n_warps - 1) / n_warps : 0;
    for (int job = 0; job < n_jobs; ++job) {
      float res[MB];                                   // residual values of this column, fetched ahead of the dot product
#pragma unroll
      for (int m = 0; m < MB; ++m)
        res[m] = (resid && lane == 0 && m < a.M) ? __bfloat162float(__ldcg(resid + m * ldr + warp_g + job * n_warps)) : 0.f;
      WsAcc c1[MB], c2[MB];
#pragma unroll
      for (int m = 0; m < MB; ++m) {
        wstream_zero(c1[m]);
        wstream_zero(c2[m]);
      }
      for (int rsel = 0; rsel < rows_per_job; ++rsel) {
        WsAcc* ac = rsel ? c2 : c1;
        for (int piece = 0; piece < cpr; ++piece) {
          const int slot = static_cast<int>(n_consumed % static_cast<unsigned>(S));
#if OVLA_MEGA_PIECE_TRACE
          if (ptr_on && ptr_n + 3 <= 500) ptr[ptr_n++] = globaltimer_ns();
#endif
          mbar_wait(&bars[slot * 2], (n_consumed / static_cast<unsigned>(S)) & 1u);
#if OVLA_MEGA_PIECE_TRACE
          if (ptr_on && ptr_n + 2 <= 500) ptr[ptr_n++] = globaltimer_ns();
#endif
          const int k0 = piece * kPiece;
          const int n16 = min(kPiece, K - k0) / 8;
          const uint4* wp = reinterpret_cast<const uint4*>(ring + slot * kPieceBytes);
          const __nv_bfloat16* xk = xs + k0;
          if (!(a.dbg & 1)) {
            // kGroup steps of a lane are loaded together and then multiplied (same FMA sequence as a plain loop): with
            // two consumer warps per scheduler the shared-memory latency of a load-use-load-use chain is exposed.
            // Measured per layer (tools/decode_trace.py, same box): 91.5 us ungrouped, 88.5 us with groups of 8
            // (4 and 2 were slower).  What bounds a consumer after that is the SM's one 128 B/clk shared-memory pipe:
            // per 8 KB piece it carries the bulk-copy write, the weight read and the activation read (1536 clk per
            // round of 8 warps = 0.8 us, against 1.45 us of HBM time per round) -- keeping the activations in
            // registers would remove a third of it, but 64 more registers do not fit under the 168 a 9-warp CTA gets.
            constexpr int kGroup = MB == 1 ? OVLA_MEGA_LDS_GROUP : (MB == 2 ? 2 : 1);
            const int n_it = n16 >> 5;                      // full steps per lane
            int it = 0;
            for (; it + kGroup <= n_it; it += kGroup) {
              uint4 wv[kGroup], xv[MB][kGroup];
#pragma unroll
              for (int j = 0; j < kGroup; ++j) {
                const int i = lane + 32 * (it + j);
                wv[j] = wp[i];
#pragma unroll
                for (int m = 0; m < MB; ++m) xv[m][j] = *reinterpret_cast<const uint4*>(xk + m * K + i * 8);
              }
#pragma unroll
              for (int j = 0; j < kGroup; ++j) {
#pragma unroll
                for (int m = 0; m < MB; ++m) wstream_fma8(wv[j], xv[m][j], ac[m]);
              }
            }
            for (; it * 32 < n16; ++it) {                   // leftover steps, the last one possibly ragged
              const int i = lane + 32 * it;
              if (i < n16) {
                const uint4 wv = wp[i];
#pragma unroll
                for (int m = 0; m < MB; ++m) wstream_fma8(wv, *reinterpret_cast<const uint4*>(xk + m * K + i * 8), ac[m]);
              }
            }
          }
          ++n_consumed;
          // every lane has read the slot: hand it back to the producer
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars[slot * 2 + 1]);
#if OVLA_MEGA_PIECE_TRACE
          if (ptr_on && ptr_n + 1 <= 500) ptr[ptr_n++] = globaltimer_ns();
#endif
        }
      }
      float acc[MB], acc2[MB];
#pragma unroll
      for (int m = 0; m < MB; ++m) {
        acc[m] = wstream_combine(c1[m]);
        acc2[m] = wstream_combine(c2[m]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          acc[m] += __shfl_xor_sync(0xffffffffu, acc[m], o);
          acc2[m] += __shfl_xor_sync(0xffffffffu, acc2[m], o);
        }
      }
      if (lane == 0) {
        const int n = warp_g + job * n_warps;
#pragma unroll
        for (int m = 0; m < MB; ++m) {
          if (m >= a.M) break;
          if (mode == 0) {
            float v = bf16_round(acc[m]);
            if (resid) v += res[m];
            static_cast<__nv_bfloat16*>(out)[m * ldo + n] = __float2bfloat16_rn(v);
          } else if (mode == 1) {
            const float g = bf16_round(acc[m]), u = bf16_round(acc2[m]);
            static_cast<__nv_bfloat16*>(out)[m * ldo + n] = __float2bfloat16_rn(bf16_round(silu(g)) * u);
          } else {
            static_cast<float*>(out)[m * ldo + n] = bf16_round(acc[m]);
          }
        }
      }
    }
  };

  const int H = a.H;
  const float scale = a.attn_scale;
  // ragged (right-padded) prompts: row b sits (P - lens[b]) positions before the longest row
  auto row_pos = [&](int b) { return a.lens ? max(0, a.pos - (a.P - min(max(a.lens[b], 1), a.P))) : a.pos; };
  // optional timeline (OVLA_MEGA_TRACE): CTA 0 and the last CTA stamp %globaltimer at every phase edge
  unsigned long long* tr = nullptr;
  int tr_n = 0;
  if (a.trace && threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1))
    tr = a.trace + (blockIdx.x == 0 ? 0 : a.trace_stride);
  auto stamp = [&]() {
    if (tr && tr_n < 512) tr[tr_n++] = globaltimer_ns();
  };
  stamp();
  for (int layer = 0; layer < a.n_layers; ++layer) {
    const DecodeLayerPtrs& l = a.layers[layer];
    ptr_on = ptr != nullptr && layer == 1;
    // The cached K / V rows of this layer do not depend on this step: the CTAs that will run attention pull their head's
    // history into L2 now, so that the attention phase reads L2 instead of queueing behind the weight stream in HBM.
    // the norm weights this layer's gate/up stage and the next layer's QKV stage (or the final norm) will read: one
    // 8 KB row each, cold in L2 otherwise -- a full HBM round trip inside a phase where the weight stream is paused
    if (OVLA_MEGA_NORM_PREFETCH && blockIdx.x == 0 && threadIdx.x == 2 && !(a.dbg & 4)) {
      l2_prefetch_bulk(l.ln2, static_cast<uint32_t>(D) * 2u);
      l2_prefetch_bulk(layer + 1 < a.n_layers ? a.layers[layer + 1].ln1 : a.final_norm, static_cast<uint32_t>(D) * 2u);
    }
    if (threadIdx.x < 2 && !(a.dbg & 4)) {
      for (int p = blockIdx.x; p < a.M * H; p += gridDim.x) {
        const int b = p / H, h = p - b * H;
        const long long head_off = (static_cast<long long>(b) * H + h) * a.Tmax * 128;
        const __nv_bfloat16* base = a.kv + layer * a.kv_layer_elems + (threadIdx.x ? a.kv_layer_elems / 2 : 0) + head_off;
        const int pos_b = row_pos(b);
        if (pos_b > 0) l2_prefetch_bulk(base, static_cast<uint32_t>(pos_b) * 128u * 2u);
      }
    }
    // [norm1 + QKV]
    stage(a.x, D, D, l.ln1);
    stamp();
    linear(3 * D, D, 0, a.qkv, 3LL * D, nullptr, 0);
    stamp();
    grid_barrier(a.barrier, bar_target);
    stamp();
    // RoPE + KV append + attention: one (batch row, head) per CTA at a time
    for (int p = blockIdx.x; p < ((a.dbg & 2) ? 0 : a.M * H); p += gridDim.x) {
      const int b = p / H, h = p - b * H;
      const long long head_off = (static_cast<long long>(b) * H + h) * a.Tmax * 128;
      decode_rope_attn_body<128, true, true>(a.qkv + b * 3LL * D + h * 128, D, a.rope_cos, a.rope_sin, row_pos(b),
                                       a.kv + layer * a.kv_layer_elems + head_off,
                                       a.kv + layer * a.kv_layer_elems + a.kv_layer_elems / 2 + head_off,
                                       a.attn + b * static_cast<long long>(D) + h * 128, scale, dyn, s_sq, s_red);
      cons_sync();
    }
    stamp();
    grid_barrier(a.barrier, bar_target);
    stamp();
    // [o_proj + residual] (in place: column n of x is read and written by its one owner lane)
    stage(a.attn, D, D, nullptr);
    stamp();
    linear(D, D, 0, a.x, D, a.x, D);
    stamp();
    grid_barrier(a.barrier, bar_target);
    stamp();
    // [norm2 + gate/up + SwiGLU]
    stage(a.x, D, D, l.ln2);
    stamp();
    linear(I, D, 1, a.act, I, nullptr, 0);
    stamp();
    grid_barrier(a.barrier, bar_target);
    stamp();
    // [down + residual]
    stage(a.act, I, I, nullptr);
    stamp();
    linear(D, I, 0, a.x, D, a.x, D);
    stamp();
    grid_barrier(a.barrier, bar_target);
    stamp();
  }
  // [final norm + lm_head]: fp32 storage of bf16-rounded logits, as HF's `.float()`
  stage(a.x, D, D, a.final_norm);
  stamp();
  linear(a.vocab, D, 2, a.logits, a.vocab, nullptr, 0);
  stamp();
}

template <int MB>
static int launch_t(const DecodeStepArgs& a, int blocks, size_t smem, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncAttributes fa;
    CUDA_TRY(cudaFuncGetAttributes(&fa, decode_step_kernel<MB>));
    CUDA_TRY(cudaFuncSetAttribute(decode_step_kernel<MB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  227 * 1024 - static_cast<int>(fa.sharedSizeBytes)));
    attr_set = true;
  }
  {
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decode_step_kernel<MB>, kWarps * 32 + 32, smem));
    if (per_sm < 1) {
      cudaFuncAttributes fa;
      CUDA_TRY(cudaFuncGetAttributes(&fa, decode_step_kernel<MB>));
      return set_error("decode step: persistent kernel does not fit an SM (regs %d x %d threads, smem %zu + %zu static, max "
                       "threads %d)", fa.numRegs, kWarps * 32 + 32, smem, fa.sharedSizeBytes, fa.maxThreadsPerBlock);
    }
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(blocks);
  cfg.blockDim = dim3(kWarps * 32 + 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;      // all CTAs co-resident, or the launch fails (never a deadlock)
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CUDA_TRY(cudaLaunchKernelEx(&cfg, decode_step_kernel<MB>, a));
  count_launch();
  return 0;
}

// returns -2 when the shape is outside what the persistent kernel covers (the caller uses the per-layer kernels)
int decode_step_launch(DecodeStepArgs a, cudaStream_t st) {
  if (a.M < 1 || a.M > 4 || a.head_dim != 128 || a.D % 32 || a.I % 64 || a.D / 4 > 5 * 256 || a.D % 8 || a.vocab < 1)
    return -2;
  if (a.pos < 0 || a.pos >= a.Tmax) return set_error("decode step: position %d outside the KV capacity %d", a.pos, a.Tmax);
  const int MB = a.M <= 1 ? 1 : (a.M <= 2 ? 2 : 4);
  const long long x_bytes = ((static_cast<long long>(MB) * std::max(a.D, a.I) * 2 + 127) / 128) * 128;
  const int ctx = a.pos + 1;
  a.attn_floats = std::max(ctx, kDecGroups * 128);
  a.attn_floats = (a.attn_floats + 31) / 32 * 32;
  const long long fixed = x_bytes + 4LL * a.attn_floats + kWarps * 8 * 8 * 2 + 2048;   // + static shared memory
  int S = static_cast<int>((227LL * 1024 - fixed) / (static_cast<long long>(kWarps) * kPieceBytes));
  if (S > 4) S = 4;
  if (S < 2) return -2;
  a.S = S;
  a.attn_scale = 1.0f / sqrtf(static_cast<float>(a.head_dim));
  const size_t smem = static_cast<size_t>(x_bytes + 4LL * a.attn_floats + static_cast<long long>(kWarps) * S * kPieceBytes + kWarps * S * 16);
  const int blocks = num_sms();
  CUDA_TRY(cudaMemsetAsync(a.barrier, 0, sizeof(unsigned), st));
  const double w_bytes = 2.0 * (a.n_layers * (4.0 * a.D * a.D + 3.0 * a.D * a.I) + 1.0 * a.vocab * a.D);
  ProfScope prof(kCatGemv, a.M * w_bytes, w_bytes, st);
  if (MB == 1) return launch_t<1>(a, blocks, smem, st);
  if (MB == 2) return launch_t<2>(a, blocks, smem, st);
  return launch_t<4>(a, blocks, smem, st);
}

}  // namespace ovla
