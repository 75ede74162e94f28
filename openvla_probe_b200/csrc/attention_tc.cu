// tcgen05 / TMEM flash attention: the causal Llama prefill (head_dim 128, K/V in the cache layout) and the
// non-causal ViT towers (q, k, v packed in one [B*T, 3D] buffer): head_dim 64 (DINOv2) and 72 (SigLIP).
//
// A work item = 128 query rows of one (batch, head); keys are consumed in chunks of 64 with an online softmax.
// CTAs are persistent (two per SM) and walk a static, cost-balanced list of items.
//   warp 4      TMA producer: Q tile per item, K chunks and V chunks into two separate rings (128B-swizzled)
//   warp 5      MMA issuer:   S_c[128 x 64]  = Q . K_c^T   (SS, both K-major)            -> TMEM cols [64*(c&1), +64)
//                             O  [128 x HD] += P_c . V_c   (TS: P_c read from TMEM; V as an MN-major B operand, i.e.
//                                                           exactly as it lies in memory) -> TMEM cols [128, 128+HD)
//               S_{c+1} is issued before the PV product of chunk c, so the tensor core computes the next scores
//               while the softmax warps work on the current ones.
//   warps 0-3   softmax: thread r owns query row r (= TMEM lane r): tcgen05.ld the 64 scores, mask, p = exp2(..),
//               P_c -> packed bf16 written back with tcgen05.st over the first 32 columns of S_c.
// Neither S nor P ever touches shared memory, so a K stage is free as soon as its QK^T product retires and the rings
// run two to three chunks ahead of the tensor core (the L2 / HBM latency of a chunk is ~2 chunk times).
// O accumulates in TMEM across chunks.  The running max is only refreshed (and O, l rescaled with a TMEM
// load-multiply-store) when a row's max grows by more than 2^8 relative to the reference it is using, so the rescale
// is rare; probabilities are bounded by 256 in that frame, which bf16 and the fp32 accumulators hold exactly as well
// as values <= 1.  Probabilities are rounded to bf16 before the PV product, as flash-attn does.
// Tiles are aligned to the END of the sequence (the ragged tile is the first one, which under the causal mask has the
// fewest keys), and the last chunk of a tile is shortened to a multiple of 16 keys.
// head_dim 128: Q 32 KB + 3 K stages + 2 V stages of 16 KB = 112 KB and 256 TMEM columns: two CTAs per SM.
// head_dim 72: every operand is a 128B-swizzled 64-column slab plus a 32B-swizzled 16-column slab (3-D tensor maps
// zero-fill columns 72..79): QK^T takes a fifth K = 16 step from the tail slabs, PV a second N = 16 product.
// A wait on a multi-phase mbarrier that some threads skip would alias phases: the end-of-item wait has its own
// barrier (bar_done), and the optional rescale wait is argued safe where it is issued.
#include <stdio.h>
#include <stdlib.h>

#include "host_util.h"
#include "ops.h"
#include "ptx.cuh"

namespace ovla {

int make_tmap_2d(CUtensorMap* m, const void* ptr, int elem_bytes, long long rows, long long cols, long long ld,
                 int box_rows);  // gemm.cu
int make_tmap_3d_slots(CUtensorMap* m, const void* ptr, long long rows, int n_slots, int slot_cols, long long ld,
                       int box_rows, int box_cols);  // gemm.cu

static constexpr int kTcThreads = 192;       // warps 0..3 softmax / epilogue, warp 4 TMA, warp 5 MMA + TMEM alloc
static constexpr int kTcBC = 64;             // keys per chunk
static constexpr int kSlabQ = 128 * 128;     // bytes of a [128 rows x 128 B] swizzled slab (Q, P)
static constexpr int kSlabKV = kTcBC * 128;  // bytes of a [64 rows x 128 B] slab (K, V chunk)
static constexpr float kRescaleThreshold = 8.f;  // log2 units

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// MN-major B operand (rows = K index, each 128-byte row = 64 contiguous N elements), SWIZZLE_128B: the N extent of an
// MMA spans 64-element slabs LBO bytes apart; groups of 8 K rows are SBO = 1024 bytes apart.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// 32-byte-swizzled slab (rows of 16 bf16): the 16-column tail of a head_dim that is not a multiple of 64 (72 -> 64 + 8,
// zero-padded to 16 by the TMA box).  As a K-major operand it is one K = 16 step, as an MN-major one it is N = 16;
// 8-row groups are 256 bytes apart either way.
__device__ __forceinline__ uint64_t umma_desc_sw32(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(256 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(6) << 61;
  return d;
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}

struct AttnTcParams {
  __nv_bfloat16* out;
  long long ldo;
  int T, H, n_batch, rot;
  int q_col_per_h;                        // query columns of head h start at h * q_col_per_h
  int kv_rows_per_b, kv_rows_per_h;       // first key row of (b, h) in the K / V tensor maps
  int k_col0, v_col0, kv_col_per_h;       // first column of head h: col0 + h * kv_col_per_h
  float scale_log2;
};

template <int HD>
struct AttnTcCfg {
  static constexpr int kNS = HD / 64;                       // 64-column (128-byte swizzled) slabs per operand
  static constexpr bool kExt = (HD % 64) != 0;              // + one 16-column (32-byte swizzled) slab: head_dim 72
  static_assert(!kExt || HD % 64 <= 16, "tail slab holds at most 16 columns");
  static constexpr int kHDP = kNS * 64 + (kExt ? 16 : 0);   // head_dim as the tensor core sees it
  static constexpr int kNK = HD == 128 ? 3 : 4;             // K ring depth
  static constexpr int kNV = HD == 128 ? 2 : 4;             // V ring depth
  static constexpr int kQBytes = kNS * kSlabQ + (kExt ? 128 * 32 : 0);
  static constexpr int kChunkBytes = kNS * kSlabKV + (kExt ? kTcBC * 32 : 0);   // one K (or V) chunk
  static constexpr int kKOff = kQBytes;
  static constexpr int kVOff = kKOff + kNK * kChunkBytes;
  static constexpr int kBarOff = kVOff + kNV * kChunkBytes;
  static constexpr int kSmemBytes = kBarOff + 256;          // + barriers; the base must be 1024-byte aligned (checked)
  static constexpr int kTmemCols = 256;                     // S/P ping-pong [0,128) | O [128, 128+kHDP)
  static_assert(kQBytes % 1024 == 0 && kChunkBytes % 1024 == 0, "slabs stay 1024-byte aligned");
  static_assert(2 * (kSmemBytes + 1024) <= 233472, "two CTAs per SM");
};

// D[tmem] (+)= A[tmem] * B[smem desc]: A = packed bf16 pairs, row r in lane r, 8 columns per K = 16 step
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int HD, bool CAUSAL>
__global__ void __launch_bounds__(kTcThreads, 2)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
               const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_qx,
               const __grid_constant__ CUtensorMap tmap_kvx, const AttnTcParams p) {
  using Cfg = AttnTcCfg<HD>;
  constexpr int NS = Cfg::kNS;
  constexpr int NK = Cfg::kNK, NV = Cfg::kNV;
  constexpr bool EXT = Cfg::kExt;      // head_dim 72: operands are [64-column slab | 16-column slab], loaded through
  constexpr int HDP = Cfg::kHDP;       // 3-D tensor maps (column in head, head slot, row) so that the tail zero-fills
  extern __shared__ __align__(1024) uint8_t smem[];
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023)) {
    printf("attn_tc_kernel: dynamic shared memory base is not 1024-byte aligned\n");
    __trap();
  }
  uint8_t* sQ = smem;
  auto sK = [&](int s) { return smem + Cfg::kKOff + s * Cfg::kChunkBytes; };
  auto sV = [&](int s) { return smem + Cfg::kVOff + s * Cfg::kChunkBytes; };
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kBarOff);
  uint64_t* bar_q = bars + 0;       // Q tile landed
  uint64_t* bar_qfree = bars + 1;   // last QK^T product of the item retired: Q may be overwritten
  uint64_t* bar_done = bars + 2;    // last PV product of the item retired
  uint64_t* bar_ofree = bars + 3;   // the item's O has been read out of TMEM (128 arrivals)
  uint64_t* bar_s = bars + 4;       // [2] scores ready in TMEM
  uint64_t* bar_p = bars + 6;       // [2] probabilities written back to TMEM (128 arrivals)
  uint64_t* bar_kfull = bars + 8;   // [NK] K chunk landed
  uint64_t* bar_kfree = bars + 12;  // [NK] its QK^T product retired
  uint64_t* bar_vfull = bars + 16;  // [NV] V chunk landed
  uint64_t* bar_vfree = bars + 20;  // [NV] its PV product retired (O holds chunks up to and including it)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 24);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // Persistent CTA: work item w = (batch*H + head) * n_tiles + reversed tile index (longest tile of a head first, the
  // tiles of one head adjacent so that they run at about the same time on different SMs and share K/V through L2).
  // Round k hands CTA x the item k*G + (x + k*rot) % G: every round covers a contiguous block of items, and the host
  // picks `rot` so that a CTA's list walks through all tile lengths (causal tiles cost 1..2*n_tiles-1 chunks): the
  // static lists then carry equal work.
  const uint32_t n_tiles = (p.T + 127) / 128;
  const uint32_t total = n_tiles * p.H * p.n_batch;
  const uint32_t G = gridDim.x, cta = blockIdx.x;
  struct Item { int q0, h, b, k_limit, n_chunks, n_last; };
  auto item_of = [&](uint32_t w) {
    Item it;
    const uint32_t bh = w / n_tiles;
    const int tile = static_cast<int>(n_tiles - 1 - (w - bh * n_tiles));
    it.b = static_cast<int>(bh / static_cast<uint32_t>(p.H));
    it.h = static_cast<int>(bh - static_cast<uint32_t>(it.b) * p.H);
    it.q0 = p.T - 128 * (static_cast<int>(n_tiles) - tile);      // negative for tile 0 of a ragged T: rows dropped
    it.k_limit = CAUSAL ? it.q0 + 128 : p.T;                     // keys [0, k_limit); q0 + 128 <= T by construction
    it.n_chunks = (it.k_limit + kTcBC - 1) / kTcBC;
    it.n_last = ((it.k_limit - (it.n_chunks - 1) * kTcBC) + 15) & ~15;  // keys issued for the last chunk (16..64)
    return it;
  };
  auto work = [&](uint32_t k) { return k * G + (cta + k * p.rot) % G; };

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    mbar_init(bar_q, 1);
    mbar_init(bar_qfree, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_s + s, 1);
      mbar_init(bar_p + s, 128);
    }
    for (int s = 0; s < NK; ++s) {
      mbar_init(bar_kfull + s, 1);
      mbar_init(bar_kfree + s, 1);
    }
    for (int s = 0; s < NV; ++s) {
      mbar_init(bar_vfull + s, 1);
      mbar_init(bar_vfree + s, 1);
    }
    mbar_init(bar_done, 1);
    mbar_init(bar_ofree, 128);
    fence_barrier_init();
  }
  if (warp == 5) {
    tmem_alloc<1>(tmem_ptr_smem, Cfg::kTmemCols);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr_smem;

  // `cur` numbers the key chunks this CTA processes across all its items; ring stage = cur % depth,
  // phase = (cur / depth) & 1.
  if (warp == 4) {
    // ------------------------------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t cur0 = 0;
      for (int k = 0;; ++k) {
        const uint32_t w = work(k);
        if (w >= total) break;
        const Item it = item_of(w);
        if (k > 0) mbar_wait(bar_qfree, (k - 1) & 1);   // every QK^T product of the previous item has read Q
        mbar_expect_tx(bar_q, Cfg::kQBytes);
        if constexpr (EXT) {
          tma_load_3d(&tmap_q, bar_q, sQ, 0, it.h, it.b * p.T + it.q0);
          tma_load_3d(&tmap_qx, bar_q, sQ + NS * kSlabQ, 64, it.h, it.b * p.T + it.q0);
        } else {
#pragma unroll
          for (int sl = 0; sl < NS; ++sl)
            tma_load_2d(&tmap_q, bar_q, sQ + sl * kSlabQ, it.h * p.q_col_per_h + sl * 64, it.b * p.T + it.q0);
        }
        const int kv_row = it.b * p.kv_rows_per_b + it.h * p.kv_rows_per_h;
        const int k_col = p.k_col0 + it.h * p.kv_col_per_h, v_col = p.v_col0 + it.h * p.kv_col_per_h;
        auto load_k = [&](int c) {
          const uint32_t cur = cur0 + c, s = cur % NK;
          if (cur >= NK) mbar_wait(bar_kfree + s, (cur / NK - 1) & 1);
          mbar_expect_tx(bar_kfull + s, Cfg::kChunkBytes);
          if constexpr (EXT) {   // packed qkv: head slots [0,H) q, [H,2H) k, [2H,3H) v
            tma_load_3d(&tmap_k, bar_kfull + s, sK(s), 0, p.H + it.h, kv_row + c * kTcBC);
            tma_load_3d(&tmap_kvx, bar_kfull + s, sK(s) + NS * kSlabKV, 64, p.H + it.h, kv_row + c * kTcBC);
          } else {
#pragma unroll
            for (int sl = 0; sl < NS; ++sl)
              tma_load_2d(&tmap_k, bar_kfull + s, sK(s) + sl * kSlabKV, k_col + sl * 64, kv_row + c * kTcBC);
          }
        };
        auto load_v = [&](int c) {
          const uint32_t cur = cur0 + c, s = cur % NV;
          if (cur >= NV) mbar_wait(bar_vfree + s, (cur / NV - 1) & 1);
          mbar_expect_tx(bar_vfull + s, Cfg::kChunkBytes);
          if constexpr (EXT) {
            tma_load_3d(&tmap_v, bar_vfull + s, sV(s), 0, 2 * p.H + it.h, kv_row + c * kTcBC);
            tma_load_3d(&tmap_kvx, bar_vfull + s, sV(s) + NS * kSlabKV, 64, 2 * p.H + it.h, kv_row + c * kTcBC);
          } else {
#pragma unroll
            for (int sl = 0; sl < NS; ++sl)
              tma_load_2d(&tmap_v, bar_vfull + s, sV(s) + sl * kSlabKV, v_col + sl * 64, kv_row + c * kTcBC);
          }
        };
        // K runs one chunk ahead of V: the order of the waits then matches the order in which the products retire
        load_k(0);
        for (int c = 0; c < it.n_chunks; ++c) {
          if (c + 1 < it.n_chunks) load_k(c + 1);
          load_v(c);
        }
        cur0 += it.n_chunks;
      }
    }
  } else if (warp == 5) {
    // ------------------------------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc(1, 128, 0);                    // N filled in per chunk
      constexpr uint32_t idesc_o = umma_idesc(1, 128, NS * 64) | (1u << 16);   // B (= V) is MN-major
      constexpr uint32_t idesc_ox = umma_idesc(1, 128, 16) | (1u << 16);       // the 16-column tail of V
      uint32_t cur0 = 0;
      for (int k = 0;; ++k) {
        const uint32_t w = work(k);
        if (w >= total) break;
        const Item it = item_of(w);
        mbar_wait(bar_q, k & 1);
        for (int c = 0; c <= it.n_chunks; ++c) {
          if (c < it.n_chunks) {
            const uint32_t cur = cur0 + c;
            const int s = cur & 1;
            const int nc = (c == it.n_chunks - 1) ? it.n_last : kTcBC;
            mbar_wait(bar_kfull + cur % NK, (cur / NK) & 1);
            tc_fence_after();
            const uint32_t idesc = idesc_s | (static_cast<uint32_t>(nc >> 3) << 17);
#pragma unroll
            for (int ks = 0; ks < NS * 4; ++ks) {
              const uint64_t a = umma_desc_sw128(smem_u32(sQ + (ks >> 2) * kSlabQ)) + 2 * (ks & 3);
              const uint64_t bd = umma_desc_sw128(smem_u32(sK(cur % NK) + (ks >> 2) * kSlabKV)) + 2 * (ks & 3);
              umma_bf16<1>(tmem + s * kTcBC, a, bd, idesc, ks != 0);
            }
            if constexpr (EXT)   // columns 64..79 of the head (72..79 are zeros)
              umma_bf16<1>(tmem + s * kTcBC, umma_desc_sw32(smem_u32(sQ + NS * kSlabQ)),
                           umma_desc_sw32(smem_u32(sK(cur % NK) + NS * kSlabKV)), idesc, 1u);
            umma_commit(bar_kfree + cur % NK);
            umma_commit(bar_s + s);
            if (c == it.n_chunks - 1) umma_commit(bar_qfree);
          }
          if (c > 0) {
            const int cp = c - 1;
            const uint32_t cur = cur0 + cp;
            const int s = cur & 1;
            const int nc = (cp == it.n_chunks - 1) ? it.n_last : kTcBC;
            mbar_wait(bar_vfull + cur % NV, (cur / NV) & 1);
            mbar_wait(bar_p + s, (cur >> 1) & 1);
            if (cp == 0 && k > 0) mbar_wait(bar_ofree, (k - 1) & 1);   // the previous item's O has been read out
            tc_fence_after();
            for (int j = 0; j < nc / 16; ++j) {
              const uint64_t bd = umma_desc_mn_sw128(smem_u32(sV(cur % NV) + j * 16 * 128), kSlabKV);
              umma_bf16_ts(tmem + 128, tmem + s * kTcBC + 8 * j, bd, idesc_o, (cp > 0 || j > 0) ? 1u : 0u);
              if constexpr (EXT)
                umma_bf16_ts(tmem + 128 + NS * 64, tmem + s * kTcBC + 8 * j,
                             umma_desc_sw32(smem_u32(sV(cur % NV) + NS * kSlabKV + j * 16 * 32)), idesc_ox,
                             (cp > 0 || j > 0) ? 1u : 0u);
            }
            umma_commit(bar_vfree + cur % NV);
            if (cp == it.n_chunks - 1) umma_commit(bar_done);
          }
        }
        cur0 += it.n_chunks;
      }
    }
  } else {
    // ------------------------------------------------------------------------------------------ softmax: thread = row
    const int r = tid;
    const uint32_t lane_base = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    uint32_t cur = 0;
    for (int k = 0;; ++k) {
      const uint32_t w = work(k);
      if (w >= total) break;
      const Item it = item_of(w);
      const int t_row = it.q0 + r;
      float m_ref = 0.f, l_run = 0.f;
      for (int c = 0; c < it.n_chunks; ++c, ++cur) {
        const int s = cur & 1;
        const int key0 = c * kTcBC;
        int n_valid = it.k_limit - key0;
        if (CAUSAL) n_valid = max(0, min(n_valid, t_row - key0 + 1));
        const bool all_valid = __all_sync(0xffffffffu, n_valid >= kTcBC);
        const bool none_valid = __all_sync(0xffffffffu, n_valid <= 0 || t_row < 0);
        mbar_wait(bar_s + s, (cur >> 1) & 1);
        tc_fence_after();
        if (none_valid) {
          // this warp's 32 rows see none of the chunk's keys (above the causal diagonal) or are all dropped rows of a
          // ragged first tile: P = 0, nothing to accumulate
          uint32_t z[kTcBC / 2];
#pragma unroll
          for (int i = 0; i < kTcBC / 2; ++i) z[i] = 0u;
          tmem_st32(lane_base + s * kTcBC, z);
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(bar_p + s);
          continue;
        }
        uint32_t v[kTcBC];
        tmem_ld32(lane_base + s * kTcBC, v);
        tmem_ld32(lane_base + s * kTcBC + 32, v + 32);
        tmem_ld_wait();
        if (!all_valid) {   // warp-uniform: only warps that straddle the diagonal / the end of the sequence pay this
#pragma unroll
          for (int i = 0; i < kTcBC; ++i) v[i] = (i < n_valid) ? v[i] : 0xff800000u;   // -inf -> p = 0
        }
        float mx = __uint_as_float(v[0]);
#pragma unroll
        for (int i = 1; i < kTcBC; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
        const float mxs = mx * p.scale_log2;
        float factor = 1.f;
        bool need = false;
        if (c == 0) {
          m_ref = (mx == -INFINITY) ? 0.f : mxs;
        } else if (mxs > m_ref + kRescaleThreshold) {
          factor = ex2_approx(m_ref - mxs);
          m_ref = mxs;
          l_run *= factor;
          need = true;
        }
        if (c > 0 && __any_sync(0xffffffffu, need)) {
          // O holds chunks [0, c) once the PV product of chunk c-1 retires; that of chunk c waits for our arrive below.
          // A parity wait is safe although most chunks skip it: bar_s of this chunk was observed, so the PV product
          // two phases back on this barrier (chunk cur-3) retired long ago -- the barrier is in the phase of chunk
          // cur-1 or just past it.
          mbar_wait(bar_vfree + (cur - 1) % NV, ((cur - 1) / NV) & 1);
          tc_fence_after();
#pragma unroll 1
          for (int cc = 0; cc < HDP / 32; ++cc) {
            uint32_t o[32];
            tmem_ld32(lane_base + 128 + cc * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
            tmem_st32(lane_base + 128 + cc * 32, o);
          }
          if constexpr (HDP % 32 != 0) {
            uint32_t o[16];
            tmem_ld16(lane_base + 128 + (HDP / 32) * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
            tmem_st16(lane_base + 128 + (HDP / 32) * 32, o);
          }
          tmem_st_wait();
        }
        const uint64_t scale2 = pack_f32x2(p.scale_log2, p.scale_log2), negm2 = pack_f32x2(-m_ref, -m_ref);
        uint64_t l2 = pack_f32x2(0.f, 0.f);
        uint32_t pk[kTcBC / 2];
#pragma unroll
        for (int i = 0; i < kTcBC / 2; ++i) {
          float x0, x1;
          unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), scale2, negm2), x0, x1);
          const float p0 = ex2_approx(x0), p1 = ex2_approx(x1);
          l2 = add_f32x2(l2, pack_f32x2(p0, p1));
          pk[i] = pack_bf16(p0, p1);
        }
        float l_lo, l_hi;
        unpack_f32x2(l2, l_lo, l_hi);
        l_run += l_lo + l_hi;
        tmem_st32(lane_base + s * kTcBC, pk);   // P_c over the first half of S_c (already in registers)
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(bar_p + s);
      }
      mbar_wait(bar_done, k & 1);   // waited every item, so the parity cannot alias
      tc_fence_after();
      const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
      __nv_bfloat16* dst = p.out + (static_cast<long long>(it.b) * p.T + t_row) * p.ldo + it.h * HD;
#pragma unroll 1
      for (int cc = 0; cc < HDP / 32; ++cc) {
        uint32_t o[32];
        tmem_ld32(lane_base + 128 + cc * 32, o);
        tmem_ld_wait();
        if (HDP % 32 == 0 && cc == HDP / 32 - 1) {   // O is in registers: the next item's first PV product may overwrite it
          tc_fence_before();
          mbar_arrive(bar_ofree);
        }
        if (t_row >= 0) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 o4;
            o4.x = pack_bf16(__uint_as_float(o[g * 8 + 0]) * inv, __uint_as_float(o[g * 8 + 1]) * inv);
            o4.y = pack_bf16(__uint_as_float(o[g * 8 + 2]) * inv, __uint_as_float(o[g * 8 + 3]) * inv);
            o4.z = pack_bf16(__uint_as_float(o[g * 8 + 4]) * inv, __uint_as_float(o[g * 8 + 5]) * inv);
            o4.w = pack_bf16(__uint_as_float(o[g * 8 + 6]) * inv, __uint_as_float(o[g * 8 + 7]) * inv);
            *reinterpret_cast<uint4*>(dst + cc * 32 + g * 8) = o4;
          }
        }
      }
      if constexpr (HDP % 32 != 0) {   // tail: 16 accumulator columns, of which HD - 64*NS (= 8) are real
        uint32_t o[16];
        tmem_ld16(lane_base + 128 + (HDP / 32) * 32, o);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(bar_ofree);
        if (t_row >= 0) {
          static_assert(HDP % 32 == 0 || HD - (HDP / 32) * 32 == 8, "tail store writes exactly 8 columns");
          uint4 o4;
          o4.x = pack_bf16(__uint_as_float(o[0]) * inv, __uint_as_float(o[1]) * inv);
          o4.y = pack_bf16(__uint_as_float(o[2]) * inv, __uint_as_float(o[3]) * inv);
          o4.z = pack_bf16(__uint_as_float(o[4]) * inv, __uint_as_float(o[5]) * inv);
          o4.w = pack_bf16(__uint_as_float(o[6]) * inv, __uint_as_float(o[7]) * inv);
          *reinterpret_cast<uint4*>(dst + (HDP / 32) * 32) = o4;
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc<1>(tmem, Cfg::kTmemCols);
  }
}

template <int HD, bool CAUSAL>
static int attn_tc_launch_one(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& tqx,
                              const CUtensorMap& tkvx, const AttnTcParams& p, int B, cudaStream_t st) {
  using Cfg = AttnTcCfg<HD>;
  auto kern = attn_tc_kernel<HD, CAUSAL>;
  static bool attr = false;
  if (!attr) {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    attr = true;
  }
  AttnTcParams pp = p;
  pp.n_batch = B;
  const long long total = 1LL * ((p.T + 127) / 128) * p.H * B;
  int grid_x = 2 * num_sms();   // two resident CTAs per SM
  if (const char* gx = getenv("OVLA_ATTN_TC_GRID")) grid_x = atoi(gx) > 0 ? atoi(gx) : grid_x;
  const dim3 grid(static_cast<unsigned>(total < grid_x ? total : grid_x));
  const int n_tiles = (p.T + 127) / 128;
  auto gcd = [](int a, int b) { while (b) { const int t = a % b; a = b; b = t; } return a; };
  pp.rot = 0;   // kind of CTA x's k-th item ~ (x + k * (G + rot)) mod n_tiles: make the step coprime with n_tiles
  while (n_tiles > 1 && gcd((static_cast<int>(grid.x) + pp.rot) % n_tiles, n_tiles) != 1) ++pp.rot;
  const double pairs = CAUSAL ? 0.5 * p.T * (p.T + 1.0) : 1.0 * p.T * p.T;
  ProfScope prof(kCatFlash, 4.0 * B * p.H * pairs * HD, 2.0 * B * p.H * HD * (4.0 * p.T), st);
  kern<<<grid, kTcThreads, Cfg::kSmemBytes, st>>>(tq, tk, tv, tqx, tkvx, pp);
  CUDA_TRY(cudaGetLastError());
  count_launch();
  return 0;
}

// Causal prefill, head_dim 128.  q: rotated queries inside the fused qkv buffer [B*T, ld_q] (head h at columns
// h*128); caches [B, H, Tmax, 128] (rows >= T must be finite); out [B*T, ldo].
int attn_tc_prefill_launch(const void* q, long long ld_q, const void* kc, const void* vc, void* out, long long ldo, int B,
                           int H, int T, int Tmax, cudaStream_t st) {
  if (B <= 0 || T <= 0) return 0;
  CUtensorMap tq, tk, tv;
  if (make_tmap_2d(&tq, q, 2, 1LL * B * T, 1LL * H * 128, ld_q, 128)) return -1;
  if (make_tmap_2d(&tk, kc, 2, 1LL * B * H * Tmax, 128, 128, kTcBC)) return -1;
  if (make_tmap_2d(&tv, vc, 2, 1LL * B * H * Tmax, 128, 128, kTcBC)) return -1;
  AttnTcParams p = {};
  p.out = static_cast<__nv_bfloat16*>(out);
  p.ldo = ldo;
  p.T = T;
  p.H = H;
  p.q_col_per_h = 128;
  p.kv_rows_per_b = H * Tmax;
  p.kv_rows_per_h = Tmax;
  p.scale_log2 = 1.4426950408889634f / sqrtf(128.f);
  return attn_tc_launch_one<128, true>(tq, tk, tv, tq, tk, p, B, st);
}

// q, k, v packed as [B*T, 3*H*hd] (q | k | v, head h at columns h*hd of each third); head_dim 64, 72 (non-causal) or 128.
int attn_tc_qkv_launch(const void* qkv, long long ld, void* out, long long ldo, int B, int H, int T, int hd, int causal,
                       cudaStream_t st) {
  if (B <= 0 || T <= 0) return 0;
  if (hd != 64 && hd != 128 && !(hd == 72 && !causal))
    return set_error("tcgen05 attention: head_dim %d (causal %d) not supported (64, 128, or 72 non-causal)", hd, causal);
  const long long D = 1LL * H * hd;
  AttnTcParams p = {};
  p.out = static_cast<__nv_bfloat16*>(out);
  p.ldo = ldo;
  p.T = T;
  p.H = H;
  p.q_col_per_h = hd;
  p.kv_rows_per_b = T;
  p.kv_rows_per_h = 0;
  p.k_col0 = static_cast<int>(D);
  p.v_col0 = static_cast<int>(2 * D);
  p.kv_col_per_h = hd;
  p.scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(hd));
  if (hd == 72) {
    if ((ldo * 2) & 15) return set_error("tcgen05 attention: output pitch must be a multiple of 8 elements");
    CUtensorMap tq, tkv, tqx, tkvx;
    if (make_tmap_3d_slots(&tq, qkv, 1LL * B * T, 3 * H, hd, ld, 128, 64)) return -1;
    if (make_tmap_3d_slots(&tkv, qkv, 1LL * B * T, 3 * H, hd, ld, kTcBC, 64)) return -1;
    if (make_tmap_3d_slots(&tqx, qkv, 1LL * B * T, 3 * H, hd, ld, 128, 16)) return -1;
    if (make_tmap_3d_slots(&tkvx, qkv, 1LL * B * T, 3 * H, hd, ld, kTcBC, 16)) return -1;
    return attn_tc_launch_one<72, false>(tq, tkv, tkv, tqx, tkvx, p, B, st);
  }
  CUtensorMap tq, tkv;
  if (make_tmap_2d(&tq, qkv, 2, 1LL * B * T, 3 * D, ld, 128)) return -1;
  if (make_tmap_2d(&tkv, qkv, 2, 1LL * B * T, 3 * D, ld, kTcBC)) return -1;
  if (hd == 64) return causal ? attn_tc_launch_one<64, true>(tq, tkv, tkv, tq, tkv, p, B, st)
                              : attn_tc_launch_one<64, false>(tq, tkv, tkv, tq, tkv, p, B, st);
  return causal ? attn_tc_launch_one<128, true>(tq, tkv, tkv, tq, tkv, p, B, st)
                : attn_tc_launch_one<128, false>(tq, tkv, tkv, tq, tkv, p, B, st);
}

}  // namespace ovla
