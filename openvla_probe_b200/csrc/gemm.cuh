// Persistent warp-specialised tcgen05 GEMM for sm_100a:  C[M,N] = A[M,K] . W[N,K]^T  (+ fused epilogue)
//
//   A  : activations, row-major (K contiguous)            -> TMA tile [BM x 128 B], SWIZZLE_128B
//   W  : nn.Linear weight layout [N, K] (K contiguous)    -> TMA tile [BN x 128 B], SWIZZLE_128B
//   acc: fp32 in TMEM, two accumulator buffers so the epilogue of tile i overlaps the MMAs of tile i+1
//
// Roles (256 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one lane), warp 2 = TMEM allocator,
// warps 4..11 = epilogue (tcgen05.ld -> registers -> fused math -> 16-byte global stores): two warps per TMEM lane
// quarter, each taking every other 32-column chunk, so that math-heavy epilogues (erf-GELU, RoPE) stay hidden under
// short-K mainloops (ViT K = 1024 / 1152).
// CG = 1: one CTA per 128 x BN tile.  CG = 2: a CTA pair (cluster of 2) computes a 256 x BN tile with
// tcgen05.mma.cta_group::2; each CTA stages its own 128 rows of A and half of the W tile, which halves the
// shared-memory / L2 traffic per flop.
//
// Fused epilogues reproduce the rounding points of the reference's bf16 HF/timm forward (every torch op
// output is rounded to bf16): linear(+bias) -> round -> [GELU -> round] -> [LayerScale -> round] ->
// [residual add -> round];  SwiGLU: silu(round(g)) -> round -> * round(u) -> round.
#pragma once
#include "ptx.cuh"
#ifdef OVLA_DBG_EPI_TIMELINE
#include <cstdio>
#endif

namespace ovla {

enum GemmMode : int {
  kModeBf16 = 0,    // bf16 out; optional bias / GELU / LayerScale / residual
  kModeSwiGLU = 1,  // W rows interleaved [32 gate | 32 up]; out[M, N/2] = silu(g) * u
  kModeF32 = 2,     // fp32 out; optional fp32-or-bf16 bias; optional rounding of the value to bf16
  kModeQkvRope = 3, // fused Llama QKV projection: RoPE on q/k, q written back, rotated k and v written to the KV cache
  kModePartial = 4, // split-K: raw fp32 partial sums of this tile's K slice -> out + slice * ldr (reduced by splitk_epilogue)
};
enum GemmKind : int { kKindBf16 = 0, kKindTf32 = 1 };

struct GemmEpi {
  void* out;
  long long ldo;  // elements
  const __nv_bfloat16* bias;
  const __nv_bfloat16* scale;
  const __nv_bfloat16* resid;
  long long ldr;
  const float* bias_f32;
  int gelu;
  int round_bf16;
  // kModeQkvRope only (head_dim 128): rows are (b, t) with t = row % T at position pos0 + t
  const __nv_bfloat16* rope_cos;  // [Tmax, 64] bf16
  const __nv_bfloat16* rope_sin;
  __nv_bfloat16* k_cache;         // [B, H, Tmax, 128]
  __nv_bfloat16* v_cache;
  int T, pos0, Tmax, H;
  int tma_epi;  // kModeBf16: output (and residual) tiles move through shared memory with TMA (tensor maps passed beside)
  // grouped GEMM (GemmShape::groups > 1, kModeF32 only): element strides between the groups' outputs / fp32 biases
  long long out_gs, bias_gs;
  // RMSNorm fused across two GEMMs (LlamaRMSNorm between o_proj / down_proj and the next q,k,v / gate,up projection):
  //   producer (kModeBf16): ss_out[row * ss_ld + slot] = sum of squares of the bf16 values this GEMM stores in the
  //     128-column group g = col / 128, over the 32-column chunks of parity p = (col / 32) & 1; slot = 2 g + p
  //     (a fixed partition, independent of the tile shape, written once by one thread: deterministic);
  //   consumer (kModeQkvRope / kModeSwiGLU): every accumulator of `row` is multiplied by
  //     rsqrt(sum_{s < ss_parts} ss_in[row * ss_ld + s] * ss_inv_k + ss_eps) before its first bf16 rounding; the
  //     norm weight is folded into W beforehand (fold_norm_weight_launch).
  float* ss_out;
  const float* ss_in;
  int ss_ld, ss_parts;
  float ss_inv_k, ss_eps;
};

// consumer side of the fused RMSNorm: the row's 1/rms from the producer's partial sums (fixed summation order).  Called
// BEFORE the wait for the accumulator, so that the loads (issued together, 16 at a time) ride under the tile's main loop.
__device__ __forceinline__ float fused_norm_rstd(const GemmEpi& epi, int row, bool row_ok) {
  if (!epi.ss_in) return 1.f;
  float ss = 0.f;
  if (row_ok) {
    const float4* p = reinterpret_cast<const float4*>(epi.ss_in + static_cast<long long>(row) * epi.ss_ld);
    const int n4 = epi.ss_parts / 4;
    for (int j0 = 0; j0 < n4; j0 += 16) {
      float4 v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = (j0 + j < n4) ? p[j0 + j] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        ss += v[j].x;
        ss += v[j].y;
        ss += v[j].z;
        ss += v[j].w;
      }
    }
  }
  return rsqrtf(ss * epi.ss_inv_k + epi.ss_eps);
}

static constexpr int kRasterSerpentine = 1 << 30;  // or-ed into GemmShape::group_n

struct GemmShape {
  int M, N, K;  // N = rows of W (pre-epilogue output columns); K in elements
  int group_m;  // rasterisation group size (row-tiles)
  int split_k;  // K slices (each tile of a slice accumulates kb_per_split K blocks); 1 = no split
  int kb_per_split;
  unsigned long long l2_a, l2_b;  // L2 eviction-priority hints of the A / W tile loads (kL2Evict*)
  // > 1: `groups` independent problems of the same shape in one launch (the probes of all captured layers): A and W
  // are 3-D tensor maps (k, row, group) whose row boxes clip at the group's own M / N, tiles are numbered group-major
  int groups;
  // rasterisation super-group (column-tiles): the row groups sweep `group_n` column-tiles at a time, so that a W slab
  // of group_n * BN rows stays L2-resident while the activations stream; 0 = all column-tiles (one super-group)
  int group_n;
  // wave alignment (long-K problems): the TMA producers of all CTAs meet every `sync_seg` K blocks (bounded skew), so
  // that CTAs that share operand tiles fetch them from L2 at the same time and one DRAM read serves them all.
  // sync = two zero-initialised words (arrivals, exits) that the last CTA to leave zeroes again; nullptr = off
  unsigned int* sync;
  int sync_seg;
};

#ifndef OVLA_EPI_SLOTS
#define OVLA_EPI_SLOTS 1
#endif
#ifndef OVLA_RES_FENCE
#define OVLA_RES_FENCE 1
#endif
static constexpr int kGemmThreads = 384;
static constexpr int kEpiWarps = 8;
static constexpr int kBM = 128;           // rows per CTA
static constexpr int kStageABytes = kBM * 128;

template <int BN, int CG>
struct GemmCfg {
  static constexpr int kBRows = BN / CG;                 // W rows staged per CTA
  static constexpr int kStageBBytes = kBRows * 128;
  static constexpr int kStageBytes = kStageABytes + kStageBBytes;
  // staging slots per epilogue warp and direction: 2 where the operand ring keeps >= 5 stages, else 1
  static constexpr int kEpiSlots = (OVLA_EPI_SLOTS >= 2 && kStageBytes <= 32 * 1024) ? 2 : 1;
  static constexpr int kRing = (232448 - 1024 - 512 - 2 * kEpiSlots * 8 * 2048);   // bytes left for the operand ring
  static constexpr int kStages = kRing / kStageBytes > 8 ? 8 : kRing / kStageBytes;
  static constexpr int kTmemCols = 2 * BN;               // 256 or 512 (power of two)
  static constexpr int kEpiStageBytes = 32 * 64;         // one [32 rows x 32 bf16] chunk, 64B-swizzled, per epilogue warp
  static constexpr int kEpiSmemBytes = 2 * kEpiSlots * kEpiWarps * kEpiStageBytes;  // output + residual staging
  static constexpr int kSmemBytes = kStages * kStageBytes + kEpiSmemBytes + 1024 /*align slack*/ + 512 /*barriers*/;
  static_assert(kSmemBytes <= 232448, "shared memory budget");
};

// Tile rasterisation: groups of G row-tiles sweep the column-tiles of a super-group of NC column-tiles (NC = 0: all of
// them), so that a wave of concurrent CTAs touches a near-square block of the output and both operands are re-used
// from L2.  Which operand stays resident across waves is the launcher's choice: a wide G keeps the activation slab
// [G * tile rows, K] while W streams; a narrow G with NC column-tiles keeps the W slab [NC * BN, K] while the
// activations stream (once per super-group).
__host__ __device__ __forceinline__ void gemm_tile_coords(int t, int num_m, int num_n, int G, int NC, int& mb, int& nb) {
  const bool serp = (NC >> 30) & 1;
  NC &= ~(1 << 30);
  int first_n = 0, ncs = num_n;
  if (NC > 0 && NC < num_n) {
    const int sg = t / (num_m * NC);
    first_n = sg * NC;
    ncs = NC < num_n - first_n ? NC : num_n - first_n;
    t -= sg * num_m * NC;
  }
  const int per_group = G * ncs;
  const int g = t / per_group;
  const int first_m = g * G;
  const int gsz = G < num_m - first_m ? G : num_m - first_m;
  const int in_g = t - g * per_group;
  mb = first_m + in_g % gsz;
  const int c = in_g / gsz;
  // serpentine (bit 30 of NC, kRasterSerpentine): odd row groups sweep the columns backwards, so the W tiles the
  // previous group used last -- still in L2 -- are the first ones this group needs
  nb = first_n + ((serp && (g & 1)) ? ncs - 1 - c : c);
}

template <int BN, int CG, int MODE, int KIND>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_res,
                    const GemmShape shape, const GemmEpi epi) {
  using Cfg = GemmCfg<BN, CG>;
  constexpr int kStages = Cfg::kStages;
  constexpr int kKElems = (KIND == kKindBf16) ? 64 : 32;  // elements per 128-byte K block
  constexpr int kTileM = kBM * CG;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi_stage = smem + kStages * Cfg::kStageBytes;   // [out | resid][slot][kEpiWarps][2 KB], 1024-byte aligned
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_stage + Cfg::kEpiSmemBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* res_bar = tmem_empty + 2;                        // [slot][kEpiWarps] residual chunk landed
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(res_bar + Cfg::kEpiSlots * kEpiWarps);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0;
  const bool leader = (cta_rank == 0);

  const int num_m = (shape.M + kTileM - 1) / kTileM;
  const int num_n = (shape.N + BN - 1) / BN;
  const int tiles_mn = num_m * num_n;
  const int n_groups = shape.groups > 1 ? shape.groups : 1;
  const int num_tiles = tiles_mn * shape.split_k * n_groups;
  const int num_k = (shape.K + kKElems - 1) / kKElems;
  const int worker = blockIdx.x / CG;
  const int num_workers = gridDim.x / CG;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], kEpiWarps * CG);  // one arrive per epilogue warp of every CTA in the group
    }
    for (int w = 0; w < Cfg::kEpiSlots * kEpiWarps; ++w) mbar_init(&res_bar[w], 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc<CG>(tmem_ptr_smem, Cfg::kTmemCols);
    tmem_relinquish<CG>();
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      // wave alignment: segment g of the launch (tile iteration x segments per tile) starts once every CTA has issued
      // the loads of segment g - 1; a CTA that runs out of tiles hands in its remaining arrivals at once.
      // (Measured and dropped, profiles/r03e_producer_ab.md: sending the arrival 4 K blocks early to hide the counter
      // round trip, and taking this thread's ~10 integer divisions per tile off the tile boundary -- into the middle
      // of the tile, or out of the loop with the lane-parallel scheme of the epilogue warps -- all LOST 4-6 % on the
      // Llama shapes: the divisions run in the shadow of the barrier round trip, and anything that lets the CTAs of a
      // wave start a tile less exactly together costs L2 hits.)
      unsigned int* const sync = shape.sync;
      const int segs_per_tile = sync ? (num_k + shape.sync_seg - 1) / shape.sync_seg : 0;
      unsigned int seg_done = 0;
      bool gave_up = false;
      for (int t = worker; t < num_tiles; t += num_workers) {
        int mb, nb;
        const int outer = t / tiles_mn;
        const int slice = outer % shape.split_k, gidx = outer / shape.split_k;
        gemm_tile_coords(t - outer * tiles_mn, num_m, num_n, shape.group_m, shape.group_n, mb, nb);
        const int row_a = mb * kTileM + static_cast<int>(cta_rank) * kBM;
        const int row_b = nb * BN + static_cast<int>(cta_rank) * Cfg::kBRows;
        const int kb0 = slice * shape.kb_per_split, kb1 = min(num_k, kb0 + shape.kb_per_split);
        int next_sync = kb0;
        for (int kb = kb0; kb < kb1; ++kb) {
          if (sync && kb == next_sync) {
            if (kb != kb0 || seg_done) {
              if (kb != kb0) { red_add_relaxed_gpu(sync, 1u); ++seg_done; }
              const unsigned int want = seg_done * gridDim.x;
              if (!gave_up && ld_relaxed_gpu(sync) < want) {
                // the alignment is a timing aid only: if part of the grid is not resident (SMs held by another
                // kernel), stop waiting after ~4 M cycles and never wait again -- no CTA can block another for good
                const long long t0 = clock64();
                while (ld_relaxed_gpu(sync) < want) {
                  if (clock64() - t0 > (4LL << 20)) { gave_up = true; break; }
                }
              }
            }
            next_sync += shape.sync_seg;
          }
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + kStageABytes;
          if (shape.groups > 1) {
            if constexpr (CG == 1) {
              mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
              tma_load_3d(&tmap_a, &full_bar[stage], sa, kb * kKElems, row_a, gidx);
              tma_load_3d(&tmap_b, &full_bar[stage], sb, kb * kKElems, row_b, gidx);
            } else {
              if (leader) mbar_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
              tma_load_3d_pair(&tmap_a, &full_bar[stage], sa, kb * kKElems, row_a, gidx);
              tma_load_3d_pair(&tmap_b, &full_bar[stage], sb, kb * kKElems, row_b, gidx);
            }
          } else if constexpr (CG == 1) {
            mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
            tma_load_2d_hint(&tmap_a, &full_bar[stage], sa, kb * kKElems, row_a, shape.l2_a);
            tma_load_2d_hint(&tmap_b, &full_bar[stage], sb, kb * kKElems, row_b, shape.l2_b);
          } else {
            if (leader) mbar_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
            tma_load_2d_pair_hint(&tmap_a, &full_bar[stage], sa, kb * kKElems, row_a, shape.l2_a);
            tma_load_2d_pair_hint(&tmap_b, &full_bar[stage], sb, kb * kKElems, row_b, shape.l2_b);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (sync) { red_add_relaxed_gpu(sync, 1u); ++seg_done; }   // last segment of this tile issued
      }
      if (sync) {
        const unsigned int total = static_cast<unsigned int>((num_tiles + num_workers - 1) / num_workers) * segs_per_tile;
        if (seg_done < total) red_add_relaxed_gpu(sync, total - seg_done);
        if (atom_add_relaxed_gpu(sync + 1, 1u) == gridDim.x - 1) {   // everyone is past its last wait: clean up
          st_relaxed_gpu(sync, 0u);
          st_relaxed_gpu(sync + 1, 0u);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA, one lane)
    if (leader && lane == 0) {
      constexpr uint32_t idesc = umma_idesc(KIND == kKindBf16 ? 1 : 2, kTileM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = worker; t < num_tiles; t += num_workers) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        const int slice = shape.split_k == 1 ? 0 : (t / tiles_mn) % shape.split_k;   // no divisions on the MMA issue path
        const int kb0 = slice * shape.kb_per_split, kb1 = min(num_k, kb0 + shape.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint64_t adesc = umma_desc_sw128(sa);
          const uint64_t bdesc = umma_desc_sw128(sa + kStageABytes);
#pragma unroll
          for (int k = 0; k < 4; ++k) {  // 4 MMAs of 32 bytes of K each per 128-byte block
            if constexpr (KIND == kKindBf16)
              umma_bf16<CG>(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, ((kb - kb0) | k) != 0);
            else
              umma_tf32<CG>(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, ((kb - kb0) | k) != 0);
          }
          if constexpr (CG == 1) umma_commit(&empty_bar[stage]); else umma_commit_pair(&empty_bar[stage], 3);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if constexpr (CG == 1) umma_commit(&tmem_full[acc]); else umma_commit_pair(&tmem_full[acc], 3);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue
    const int q = warp & 3;             // TMEM lane quarter this warp may access
    const int part = (warp - 4) >> 2;   // which half of the column chunks this warp handles
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t res_phase = 0;   // bit s: parity of this warp's residual-slot barrier s
    uint32_t n_out = 0, n_res = 0;   // chunks this warp has stored / residual chunks it has requested and consumed
    uint32_t n_req = 0;
#ifdef OVLA_DBG_EPI_TIMELINE   // debug variant only: where does an epilogue warp spend its cycles? (tools/gemm_epilogue_timeline.py)
    long long tl[7] = {0, 0, 0, 0, 0, 0, 0};
    long long tl_t = clock64();
    int tl_tiles = 0;
#define OVLA_TL(i) { const long long now_ = clock64(); tl[i] += now_ - tl_t; tl_t = now_; }
#else
#define OVLA_TL(i)
#endif
    // Tile coordinates cost ~10 integer divisions (no hardware divider: ~2000 cycles), and with short K the epilogue
    // warps ARE the critical path (profiles/r03a_epilogue_timeline.md): every 32 iterations lane l works out the
    // coordinates of this worker's iteration i + l -- all lanes at once -- and each iteration fetches its own by shuffle.
    int l_mb = 0, l_nb = 0, l_slice = 0, l_gidx = 0, iter = 0;
    for (int t = worker; t < num_tiles; t += num_workers, ++iter) {
      if ((iter & 31) == 0) {
        const long long tl = static_cast<long long>(t) + static_cast<long long>(lane) * num_workers;
        if (tl < num_tiles) {
          const int outer = static_cast<int>(tl) / tiles_mn;
          l_slice = outer % shape.split_k;
          l_gidx = outer / shape.split_k;
          gemm_tile_coords(static_cast<int>(tl) - outer * tiles_mn, num_m, num_n, shape.group_m, shape.group_n, l_mb, l_nb);
        }
      }
      const int mb = __shfl_sync(0xffffffffu, l_mb, iter & 31), nb = __shfl_sync(0xffffffffu, l_nb, iter & 31);
      const int slice = __shfl_sync(0xffffffffu, l_slice, iter & 31), gidx = __shfl_sync(0xffffffffu, l_gidx, iter & 31);
      const int row = mb * kTileM + static_cast<int>(cta_rank) * kBM + q * 32 + lane;
      const bool row_ok = row < shape.M;
      float rstd = 1.f;
      if constexpr (MODE == kModeSwiGLU || MODE == kModeQkvRope) rstd = fused_norm_rstd(epi, row, row_ok);
      OVLA_TL(6)
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      OVLA_TL(0)
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
      const int col0 = nb * BN;

      if (MODE == kModeBf16 && epi.tma_epi) {
        // Output (and residual) chunks of [32 rows x 32 columns] move through a 64B-swizzled shared-memory slot per
        // warp with TMA: a thread owns one ROW of the accumulator, so direct global accesses put every lane on its
        // own 128-byte line (32 L1 wavefronts per instruction, on the same data pipe that feeds the tensor core);
        // TMA moves whole lines and clips the ragged edges.  In-place residual (out == resid) is safe: a chunk's
        // residual has landed before the same chunk is stored, and chunks do not overlap.
        constexpr int SL = Cfg::kEpiSlots;
        const int ew = warp - 4;
        auto out_slot = [&](uint32_t i) { return epi_stage + ((i % SL) * kEpiWarps + ew) * Cfg::kEpiStageBytes; };
        auto res_slot = [&](uint32_t i) { return epi_stage + ((SL + i % SL) * kEpiWarps + ew) * Cfg::kEpiStageBytes; };
        const int row0 = mb * kTileM + static_cast<int>(cta_rank) * kBM + q * 32;
        const bool has_res = epi.resid != nullptr;
        const uint32_t sw = static_cast<uint32_t>((lane >> 1) & 3);       // 64B swizzle: 16-byte piece ^= (row >> 1) & 3
        // residual chunks are requested SL chunks ahead (slot = request number % SL; a slot is refilled only after
        // every lane has read it)
        auto fetch_res = [&](int c) {
          if (has_res && c < BN / 32 && col0 + c * 32 < shape.N) {
            if (lane == 0) {
              uint64_t* bar = &res_bar[(n_req % SL) * kEpiWarps + ew];
#ifdef OVLA_GEMM_RELEASE_ARRIVE
              mbar_expect_tx(bar, Cfg::kEpiStageBytes);
#else
              mbar_expect_tx_relaxed(bar, Cfg::kEpiStageBytes);
#endif
              tma_load_2d(&tmap_res, bar, res_slot(n_req), col0 + c * 32, row0);
            }
            ++n_req;
          }
        };
#ifdef OVLA_DBG_EPI_SKIP      // timing experiment only (no output): is a short-K GEMM bound by its epilogue or by its operand feed?
        if (false)
#endif
#pragma unroll
        for (int i = 0; i < SL; ++i) fetch_res(part + 2 * i);
        float ss0 = 0.f, ss1 = 0.f;   // fused RMSNorm producer: this row's sums of squares, first / second 128 columns
#pragma unroll 1
        for (int c = part; c < BN / 32; c += 2) {
          const int col = col0 + c * 32;
          if (col >= shape.N) break;
#ifdef OVLA_DBG_EPI_SKIP
          break;
#endif
          uint32_t v[32];
          tmem_ld32(taddr + c * 32, v);
          // the chunk's bias and LayerScale (four 8-column pieces each) are requested before the wait for
          // the accumulator: inside the 8-column loop each of these loads sat behind a branch, with its full latency
          // exposed (4 x ~250 cycles per chunk; the epilogue warps of a short-K GEMM are the critical path)
          uint4 bq[4] = {}, sq[4] = {};
          if (epi.bias) {
#pragma unroll
            for (int g = 0; g < 4; ++g)
              if (col + g * 8 < shape.N) bq[g] = *reinterpret_cast<const uint4*>(epi.bias + col + g * 8);
          }
          if (epi.scale) {
#pragma unroll
            for (int g = 0; g < 4; ++g)
              if (col + g * 8 < shape.N) sq[g] = *reinterpret_cast<const uint4*>(epi.scale + col + g * 8);
          }
          tmem_ld_wait();
          OVLA_TL(1)
          float cs = 0.f;
          uint4 rr[4] = {};
          if (has_res) {
            const uint32_t sl = n_res % SL;
            mbar_wait(&res_bar[sl * kEpiWarps + ew], (res_phase >> sl) & 1);
            res_phase ^= 1u << sl;
            const uint8_t* sr = res_slot(n_res);
#pragma unroll
            for (int g = 0; g < 4; ++g) rr[g] = *reinterpret_cast<const uint4*>(sr + lane * 64 + ((g ^ sw) << 4));
            ++n_res;
#if OVLA_RES_FENCE
            fence_proxy_async();   // order this lane's generic-proxy reads before the async-proxy (TMA) refill
#endif
            __syncwarp();          // every lane has read its row: the slot may be refilled
            fetch_res(c + 2 * SL);
          }
          OVLA_TL(2)
          // Rounding points of the reference's bf16 ops: linear (+ bias) -> [GELU] -> [LayerScale] -> [residual].  After
          // the first rounding the values travel as packed bf16 pairs: LayerScale and the residual add are single
          // bf16x2 instructions (bit-identical to fp32 arithmetic + rounding, see mul_bf16x2 / add_bf16x2) -- the
          // epilogue of a short-K GEMM is bound by its instruction count (profiles/r02q_gemm_epilogue_cost.md).
          // The four phases run over the whole 32-column chunk one after the other (one run-time branch per phase, not
          // per 8 columns): the erf-GELU of 16 independent pairs then schedules as one straight-line block
          // (profiles/r03a_epilogue_timeline.md: fc1 spent 8500 of its 13000 epilogue cycles per tile here).
          uint32_t pk[16];
          if (epi.bias) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const uint32_t bw[4] = {bq[g].x, bq[g].y, bq[g].z, bq[g].w};   // zeros beyond N (those columns are clipped)
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float2 f = unpack_bf16(bw[i]);
                v[g * 8 + 2 * i] = __float_as_uint(__uint_as_float(v[g * 8 + 2 * i]) + f.x);
                v[g * 8 + 2 * i + 1] = __float_as_uint(__uint_as_float(v[g * 8 + 2 * i + 1]) + f.y);
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[j] = pack_bf16(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
          if (epi.gelu) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float2 f = unpack_bf16(pk[j]);
              gelu_erf_x2(f.x, f.y);
              pk[j] = pack_bf16(f.x, f.y);
            }
          }
          if (epi.scale) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              pk[4 * g] = mul_bf16x2(pk[4 * g], sq[g].x);
              pk[4 * g + 1] = mul_bf16x2(pk[4 * g + 1], sq[g].y);
              pk[4 * g + 2] = mul_bf16x2(pk[4 * g + 2], sq[g].z);
              pk[4 * g + 3] = mul_bf16x2(pk[4 * g + 3], sq[g].w);
            }
          }
          if (has_res) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              pk[4 * g] = add_bf16x2(pk[4 * g], rr[g].x);
              pk[4 * g + 1] = add_bf16x2(pk[4 * g + 1], rr[g].y);
              pk[4 * g + 2] = add_bf16x2(pk[4 * g + 2], rr[g].z);
              pk[4 * g + 3] = add_bf16x2(pk[4 * g + 3], rr[g].w);
            }
          }
          if (epi.ss_out) {   // squares of the values as stored (bf16), in column order
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float2 r = unpack_bf16(pk[j]);
              cs = fmaf(r.x, r.x, cs);
              cs = fmaf(r.y, r.y, cs);
            }
          }
          uint4 o4[4];
#pragma unroll
          for (int g = 0; g < 4; ++g) o4[g] = make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
          if (c & 4) ss1 += cs; else ss0 += cs;
          uint8_t* so = out_slot(n_out);
#ifdef OVLA_DBG_EPI_TIMELINE
          asm volatile("" :: "r"(o4[0].x), "r"(o4[1].y), "r"(o4[2].z), "r"(o4[3].w) : "memory");   // the math is done here
#endif
          OVLA_TL(3)
          if (lane == 0) tma_store_wait_read<SL - 1>();   // the store that last used this slot has finished reading it
          __syncwarp();
          OVLA_TL(4)
#pragma unroll
          for (int g = 0; g < 4; ++g) *reinterpret_cast<uint4*>(so + lane * 64 + ((g ^ sw) << 4)) = o4[g];
#ifndef OVLA_DBG_NO_STORE_FENCE   // timing experiments only
          fence_proxy_async();
#endif
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmap_out, so, col, row0);
            tma_store_commit();
          }
          ++n_out;
          OVLA_TL(5)
        }
        if (epi.ss_out && row_ok && col0 < shape.N) {
          float* sp = epi.ss_out + static_cast<long long>(row) * epi.ss_ld + (col0 >> 7) * 2 + part;
          sp[0] = ss0;
          if (BN > 128 && col0 + 128 < shape.N) sp[2] = ss1;
        }
      } else if constexpr (MODE == kModeBf16) {
        __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(epi.out) + static_cast<long long>(row) * epi.ldo;
        const __nv_bfloat16* res = (epi.resid && row_ok) ? epi.resid + static_cast<long long>(row) * epi.ldr : nullptr;
        // The residual row segment of a chunk (64 B per thread, rows a full pitch apart) is fetched one chunk ahead:
        // issued back to back with the accumulator wait these loads cost ~1 us each and, with short K (the ViT
        // proj / fc2 GEMMs), made the epilogue longer than the next tile's main loop.
        uint4 rr[4] = {};
        auto fetch_resid = [&](int c, uint4* dst) {
          const int col = col0 + c * 32;
#pragma unroll
          for (int g = 0; g < 4; ++g)
            if (res && c < BN / 32 && col + g * 8 < shape.N) dst[g] = *reinterpret_cast<const uint4*>(res + col + g * 8);
        };
        fetch_resid(part, rr);
        float ss0 = 0.f, ss1 = 0.f;
#pragma unroll 1
        for (int c = part; c < BN / 32; c += 2) {
          uint32_t v[32];
          tmem_ld32(taddr + c * 32, v);
          uint4 rn[4] = {};
          fetch_resid(c + 2, rn);
          tmem_ld_wait();
          const int col = col0 + c * 32;
          float cs = 0.f;
          if (col < shape.N) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {  // 8 columns -> one 16-byte store
              const int cg = col + g * 8;
              if (cg >= shape.N) break;
              float x[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) x[i] = __uint_as_float(v[g * 8 + i]);
              if (epi.bias) {
                const uint4 bb = *reinterpret_cast<const uint4*>(epi.bias + cg);
                const uint32_t bw[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float2 f = unpack_bf16(bw[i]);
                  x[2 * i] += f.x;
                  x[2 * i + 1] += f.y;
                }
              }
#pragma unroll
              for (int i = 0; i < 8; ++i) x[i] = bf16_round(x[i]);
              if (epi.gelu) {
#pragma unroll
                for (int i = 0; i < 8; i += 2) {
                  gelu_erf_x2(x[i], x[i + 1]);
                  x[i] = bf16_round(x[i]);
                  x[i + 1] = bf16_round(x[i + 1]);
                }
              }
              if (epi.scale) {
                const uint4 ss = *reinterpret_cast<const uint4*>(epi.scale + cg);
                const uint32_t sw[4] = {ss.x, ss.y, ss.z, ss.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float2 f = unpack_bf16(sw[i]);
                  x[2 * i] = bf16_round(x[2 * i] * f.x);
                  x[2 * i + 1] = bf16_round(x[2 * i + 1] * f.y);
                }
              }
              if (row_ok) {
                if (res) {
                  const uint32_t rw[4] = {rr[g].x, rr[g].y, rr[g].z, rr[g].w};
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    const float2 f = unpack_bf16(rw[i]);
                    x[2 * i] += f.x;
                    x[2 * i + 1] += f.y;
                  }
                }
                uint4 o;
                o.x = pack_bf16(x[0], x[1]);
                o.y = pack_bf16(x[2], x[3]);
                o.z = pack_bf16(x[4], x[5]);
                o.w = pack_bf16(x[6], x[7]);
                *reinterpret_cast<uint4*>(out + cg) = o;
                if (epi.ss_out) {
                  const uint32_t ow[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    const float2 r = unpack_bf16(ow[i]);
                    cs = fmaf(r.x, r.x, cs);
                    cs = fmaf(r.y, r.y, cs);
                  }
                }
              }
            }
          }
          if (c & 4) ss1 += cs; else ss0 += cs;
#pragma unroll
          for (int g = 0; g < 4; ++g) rr[g] = rn[g];
        }
        if (epi.ss_out && row_ok && col0 < shape.N) {
          float* sp = epi.ss_out + static_cast<long long>(row) * epi.ss_ld + (col0 >> 7) * 2 + part;
          sp[0] = ss0;
          if (BN > 128 && col0 + 128 < shape.N) sp[2] = ss1;
        }
      } else if (MODE == kModeSwiGLU && epi.tma_epi) {
        // same shared-memory + TMA store as the bf16 mode: a 64-column accumulator chunk ([32 gate | 32 up]) gives one
        // [32 x 32] output chunk
        constexpr int SL = Cfg::kEpiSlots;
        const int ew = warp - 4;
        const int row0 = mb * kTileM + static_cast<int>(cta_rank) * kBM + q * 32;
        const uint32_t sw = static_cast<uint32_t>((lane >> 1) & 3);
        const int n_cols = shape.N / 2;
#pragma unroll 1
        for (int c = part; c < BN / 64; c += 2) {
          const int col = (col0 + c * 64) / 2;
          if (col >= n_cols) break;
          uint32_t g[32], u[32];
          tmem_ld32(taddr + c * 64, g);
          tmem_ld32(taddr + c * 64 + 32, u);
          tmem_ld_wait();
          uint4 o4[4];
#pragma unroll
          for (int grp = 0; grp < 4; ++grp) {
            float y[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float gv = bf16_round(__uint_as_float(g[grp * 8 + i]) * rstd);
              const float uv = bf16_round(__uint_as_float(u[grp * 8 + i]) * rstd);
              y[i] = bf16_round(silu(gv)) * uv;
            }
            o4[grp].x = pack_bf16(y[0], y[1]);
            o4[grp].y = pack_bf16(y[2], y[3]);
            o4[grp].z = pack_bf16(y[4], y[5]);
            o4[grp].w = pack_bf16(y[6], y[7]);
          }
          uint8_t* so = epi_stage + ((n_out % SL) * kEpiWarps + ew) * Cfg::kEpiStageBytes;
          if (lane == 0) tma_store_wait_read<SL - 1>();
          __syncwarp();
#pragma unroll
          for (int grp = 0; grp < 4; ++grp) *reinterpret_cast<uint4*>(so + lane * 64 + ((grp ^ sw) << 4)) = o4[grp];
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmap_out, so, col, row0);
            tma_store_commit();
          }
          ++n_out;
        }
      } else if constexpr (MODE == kModeSwiGLU) {
        __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(epi.out) + static_cast<long long>(row) * epi.ldo;
        const int n_out = shape.N / 2;
#pragma unroll 1
        for (int c = part; c < BN / 64; c += 2) {
          uint32_t g[32], u[32];
          tmem_ld32(taddr + c * 64, g);
          tmem_ld32(taddr + c * 64 + 32, u);
          tmem_ld_wait();
          const int col = (col0 + c * 64) / 2;
          if (col >= n_out) continue;
#pragma unroll
          for (int grp = 0; grp < 4; ++grp) {
            const int cg = col + grp * 8;
            if (cg >= n_out) break;
            float y[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float gv = bf16_round(__uint_as_float(g[grp * 8 + i]) * rstd);
              const float uv = bf16_round(__uint_as_float(u[grp * 8 + i]) * rstd);
              y[i] = bf16_round(silu(gv)) * uv;
            }
            if (row_ok) {
              uint4 o;
              o.x = pack_bf16(y[0], y[1]);
              o.y = pack_bf16(y[2], y[3]);
              o.z = pack_bf16(y[4], y[5]);
              o.w = pack_bf16(y[6], y[7]);
              *reinterpret_cast<uint4*>(out + cg) = o;
            }
          }
        }
      } else if constexpr (MODE == kModeQkvRope) {
        // W = [q | k | v] rows, each [H, 128]; a 128-column group of the tile is one head of q, k or v.  rotate_half
        // pairs column j with j + 64, i.e. 32-column chunk c with chunk c + 2 of the same head.  Rounding points as
        // the reference's bf16 ops: linear output, each RoPE product, and the sum (modeling_llama apply_rotary_pos_emb).
        const int bidx = row / epi.T, t = row - bidx * epi.T, pos = epi.pos0 + t;
        const int Dm = epi.H * 128;
        __nv_bfloat16* qrow = reinterpret_cast<__nv_bfloat16*>(epi.out) + static_cast<long long>(row) * epi.ldo;
#pragma unroll 1
        for (int idx = part; idx < BN / 64; idx += 2) {  // (head, half) pairs of this tile
          const int hh = idx >> 1, half = idx & 1;
          const int colh = col0 + hh * 128;
          const int which = colh / Dm, h = (colh - which * Dm) >> 7;
          const long long coff = ((static_cast<long long>(bidx) * epi.H + h) * epi.Tmax + pos) * 128;
          {
            uint32_t lo[32], hi[32];
            tmem_ld32(taddr + hh * 128 + half * 32, lo);
            tmem_ld32(taddr + hh * 128 + 64 + half * 32, hi);
            tmem_ld_wait();
            if (!row_ok || colh >= shape.N) continue;
            __nv_bfloat16* d1;
            if (which == 0) d1 = qrow + h * 128 + half * 32;
            else if (which == 1) d1 = epi.k_cache + coff + half * 32;
            else d1 = epi.v_cache + coff + half * 32;
            const __nv_bfloat16* cs = epi.rope_cos + static_cast<long long>(pos) * 64 + half * 32;
            const __nv_bfloat16* sn = epi.rope_sin + static_cast<long long>(pos) * 64 + half * 32;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              // the projection output rounded to bf16 (x 1/rms of the fused input norm), packed in column pairs
              uint32_t a[4], b[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                a[i] = pack_bf16(__uint_as_float(lo[g * 8 + 2 * i]) * rstd, __uint_as_float(lo[g * 8 + 2 * i + 1]) * rstd);
                b[i] = pack_bf16(__uint_as_float(hi[g * 8 + 2 * i]) * rstd, __uint_as_float(hi[g * 8 + 2 * i + 1]) * rstd);
              }
              uint4 w1, w2;
              if (which < 2) {
                // x cos + rotate_half(x) sin on packed bf16 pairs: every product and the sum rounded to bf16 as the
                // reference's bf16 ops do (one rounding each, bit-identical to fp32 arithmetic + rounding)
                const uint4 cu = *reinterpret_cast<const uint4*>(cs + g * 8);
                const uint4 su = *reinterpret_cast<const uint4*>(sn + g * 8);
                const uint32_t cw[4] = {cu.x, cu.y, cu.z, cu.w}, sw[4] = {su.x, su.y, su.z, su.w};
                uint32_t r1[4], r2[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  r1[i] = sub_bf16x2(mul_bf16x2(a[i], cw[i]), mul_bf16x2(b[i], sw[i]));
                  r2[i] = add_bf16x2(mul_bf16x2(b[i], cw[i]), mul_bf16x2(a[i], sw[i]));
                }
                w1 = make_uint4(r1[0], r1[1], r1[2], r1[3]);
                w2 = make_uint4(r2[0], r2[1], r2[2], r2[3]);
              } else {
                w1 = make_uint4(a[0], a[1], a[2], a[3]);
                w2 = make_uint4(b[0], b[1], b[2], b[3]);
              }
              *reinterpret_cast<uint4*>(d1 + g * 8) = w1;
              *reinterpret_cast<uint4*>(d1 + 64 + g * 8) = w2;
            }
          }
        }
      } else if constexpr (MODE == kModePartial) {
        float* out = reinterpret_cast<float*>(epi.out) + static_cast<long long>(slice) * epi.ldr +
                     static_cast<long long>(row) * epi.ldo;
#pragma unroll 1
        for (int c = part; c < BN / 32; c += 2) {
          uint32_t v[32];
          tmem_ld32(taddr + c * 32, v);
          tmem_ld_wait();
          const int col = col0 + c * 32;
          if (col >= shape.N || !row_ok) continue;
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const int cg = col + g * 4;
            if (cg >= shape.N) break;
            *reinterpret_cast<float4*>(out + cg) =
                make_float4(__uint_as_float(v[g * 4]), __uint_as_float(v[g * 4 + 1]), __uint_as_float(v[g * 4 + 2]),
                            __uint_as_float(v[g * 4 + 3]));
          }
        }
      } else {  // kModeF32
        float* out = reinterpret_cast<float*>(epi.out) + gidx * epi.out_gs + static_cast<long long>(row) * epi.ldo;
        const float* bias_f32 = epi.bias_f32 ? epi.bias_f32 + gidx * epi.bias_gs : nullptr;
#pragma unroll 1
        for (int c = part; c < BN / 32; c += 2) {
          uint32_t v[32];
          tmem_ld32(taddr + c * 32, v);
          tmem_ld_wait();
          const int col = col0 + c * 32;
          if (col >= shape.N) continue;
#pragma unroll
          for (int g = 0; g < 8; ++g) {  // 4 columns -> one 16-byte store
            const int cg = col + g * 4;
            if (cg >= shape.N) break;
            float x[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              x[i] = __uint_as_float(v[g * 4 + i]);
              if (bias_f32) x[i] += bias_f32[cg + i];
              if (epi.bias) x[i] += __bfloat162float(epi.bias[cg + i]);
              if (epi.round_bf16) x[i] = bf16_round(x[i]);
            }
            if (row_ok) *reinterpret_cast<float4*>(out + cg) = make_float4(x[0], x[1], x[2], x[3]);
          }
        }
      }

      // release this accumulator buffer to the MMA issuer (of the leader CTA)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
#ifdef OVLA_GEMM_RELEASE_ARRIVE   // A/B: the round-2a form (generic release in front of the arrive: MEMBAR + ERRBAR per tile)
        if constexpr (CG == 1) mbar_arrive(&tmem_empty[acc]); else mbar_arrive_cluster(&tmem_empty[acc], 0);
#else
        if constexpr (CG == 1) mbar_arrive_relaxed(&tmem_empty[acc]); else mbar_arrive_cluster_relaxed(&tmem_empty[acc], 0);
#endif
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
#ifdef OVLA_DBG_EPI_TIMELINE
      ++tl_tiles;
#endif
    }
#ifdef OVLA_DBG_EPI_TIMELINE
    if (MODE == kModeBf16 && lane == 0 && (warp == 4 || warp == 11) && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1) && tl_tiles > 1)
      printf("EPI_TL cta %d warp %d tiles %d cycles/tile: acc_wait %lld tmem_ld %lld resid %lld math %lld slot_wait %lld store %lld release %lld\n",
             blockIdx.x, warp, tl_tiles, tl[0] / tl_tiles, tl[1] / tl_tiles, tl[2] / tl_tiles, tl[3] / tl_tiles, tl[4] / tl_tiles,
             tl[5] / tl_tiles, tl[6] / tl_tiles);
#endif
  }

  // ------------------------------------------------------------ teardown
  if (warp >= 4 && lane == 0) tma_store_wait<0>();   // bulk stores read shared memory: finish before the CTA exits
  __syncwarp();
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) tmem_dealloc<CG>(tmem_base, Cfg::kTmemCols);
}

}  // namespace ovla
