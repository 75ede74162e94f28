// Host side of the tcgen05 GEMM: TMA tensor-map encoding and launch dispatch.
#include <stdlib.h>

#include "gemm.cuh"
#include <algorithm>

#include "host_util.h"
#include "ops.h"

#include <map>
#include <mutex>
#include <utility>

namespace ovla {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// 2-D row-major [rows, cols] tensor with `ld` elements between rows; box = [box_rows, box_bytes] with the swizzle whose
// span equals the box width (128 or 64 bytes)
static int make_tmap_2d_box(CUtensorMap* m, const void* ptr, int elem_bytes, long long rows, long long cols, long long ld,
                            int box_rows, int box_bytes) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return set_error("cuTensorMapEncodeTiled entry point not found");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((ld * elem_bytes) & 15))
    return set_error("TMA operand must be 16-byte aligned with a 16-byte multiple row pitch");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * elem_bytes};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_bytes / elem_bytes), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = enc(m, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   box_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error("cuTensorMapEncodeTiled failed (%d)", static_cast<int>(r));
  return 0;
}
int make_tmap_2d(CUtensorMap* m, const void* ptr, int elem_bytes, long long rows, long long cols, long long ld,
                 int box_rows) {
  return make_tmap_2d_box(m, ptr, elem_bytes, rows, cols, ld, box_rows, 128);
}

// bf16 [rows, n_slots * slot_cols] (pitch `ld` elements) viewed as 3-D (column in slot, slot, row); box =
// [box_rows, 1 slot, box_cols] with the swizzle whose span equals the box width (128 or 32 bytes).  Columns of a box
// that lie beyond slot_cols read as zeros: this is how a head of 72 columns is padded to 64 + 16.
int make_tmap_3d_slots(CUtensorMap* m, const void* ptr, long long rows, int n_slots, int slot_cols, long long ld,
                       int box_rows, int box_cols) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return set_error("cuTensorMapEncodeTiled entry point not found");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((ld * 2) & 15) || ((slot_cols * 2) & 15))
    return set_error("TMA operand must be 16-byte aligned with 16-byte multiple pitches");
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(slot_cols), static_cast<cuuint64_t>(n_slots), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(slot_cols) * 2, static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(box_cols), 1, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[3] = {1, 1, 1};
  const int box_bytes = box_cols * 2;
  if (box_bytes != 128 && box_bytes != 32) return set_error("make_tmap_3d_slots: box of %d bytes", box_bytes);
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, box_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error("cuTensorMapEncodeTiled (3-D) failed (%d)", static_cast<int>(r));
  return 0;
}

template <int BN, int CG, int MODE, int KIND>
static int launch_one(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& tr,
                      const GemmShape& s, const GemmEpi& e, int num_sms, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, CG>;
  auto kern = gemm_tcgen05_kernel<BN, CG, MODE, KIND>;
  static bool attr_set = false;
  if (!attr_set) {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  const int tile_m = kBM * CG;
  const int tiles = ((s.M + tile_m - 1) / tile_m) * ((s.N + BN - 1) / BN) * s.split_k * (s.groups > 1 ? s.groups : 1);
  int workers = num_sms / CG;
  if (workers > tiles) workers = tiles;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(workers * CG);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  ProfScope prof(kCatGemm, 2.0 * s.M * s.N * s.K * (s.groups > 1 ? s.groups : 1), 0.0, stream);
  CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, ta, tb, to, tr, s, e));
  count_launch();
  return 0;
}

template <int MODE, int KIND>
static int dispatch_tile(int bn, int cg, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to,
                         const CUtensorMap& tr, const GemmShape& s, const GemmEpi& e, int num_sms, cudaStream_t stream) {
  if (bn == 256 && cg == 2) return launch_one<256, 2, MODE, KIND>(ta, tb, to, tr, s, e, num_sms, stream);
  if (bn == 256 && cg == 1) return launch_one<256, 1, MODE, KIND>(ta, tb, to, tr, s, e, num_sms, stream);
  if (bn == 128 && cg == 2) return launch_one<128, 2, MODE, KIND>(ta, tb, to, tr, s, e, num_sms, stream);
  if (bn == 128 && cg == 1) return launch_one<128, 1, MODE, KIND>(ta, tb, to, tr, s, e, num_sms, stream);
  if (bn == 64 && cg == 1) return launch_one<64, 1, MODE, KIND>(ta, tb, to, tr, s, e, num_sms, stream);
  return set_error("gemm: unsupported tile config bn=%d cg=%d", bn, cg);
}

// Shapes that may be split along K (the caller then has to pass a workspace): the small-M weight-streaming bf16
// GEMMs (bs=1 prefill, batched decode) and the single-probe TF32 GEMMs (64 tiles on 148 SMs without a split).
bool splitk_eligible(int M, int N, int kind) {
  return kind == kKindBf16 ? M <= 512 : 1LL * M * N <= (4LL << 20);
}

// 3-D view (k, row, group) of `groups` equally shaped row-major [rows, cols] matrices `group_stride` elements apart;
// box = [1 group, box_rows, 128 bytes]: rows beyond `rows` read as zeros instead of running into the next group
static int make_tmap_3d_groups(CUtensorMap* m, const void* ptr, int elem_bytes, long long rows, long long cols,
                               long long ld, long long group_stride, int groups, int box_rows) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return set_error("cuTensorMapEncodeTiled entry point not found");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((ld * elem_bytes) & 15) || ((group_stride * elem_bytes) & 15))
    return set_error("grouped TMA operand must be 16-byte aligned with 16-byte multiple row / group pitches");
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(groups)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld) * elem_bytes, static_cast<cuuint64_t>(group_stride) * elem_bytes};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(128 / elem_bytes), static_cast<cuuint32_t>(box_rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = enc(m, dt, 3, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error("cuTensorMapEncodeTiled (grouped) failed (%d)", static_cast<int>(r));
  return 0;
}

// `groups` independent fp32-output GEMMs of one shape in a single persistent launch (probe training over all captured
// layers at once: 33 x 64 tiles fill the 148 SMs that a single probe's 64 tiles cannot):
//   out_g[M, N] = A_g[M, K] . W_g[N, K]^T (+ bias_g),  g < groups,  operands `*_gs` elements apart.
int gemm_grouped_launch(const void* A, long long lda, long long a_gs, const void* W, long long ldw, long long w_gs,
                        int groups, int M, int N, int K, int kind, float* out, long long ldo, long long out_gs,
                        const float* bias_f32, long long bias_gs, int bn, int cg, int num_sms, cudaStream_t stream) {
  if (groups <= 0 || M <= 0 || N <= 0 || K <= 0) return set_error("grouped gemm: empty shape G=%d M=%d N=%d K=%d", groups, M, N, K);
  if (N % 4) return set_error("grouped gemm: N=%d must be a multiple of 4", N);
  if ((reinterpret_cast<uintptr_t>(out) & 15) || ((ldo * 4) & 15) || ((out_gs * 4) & 15))
    return set_error("grouped gemm: output must be 16-byte aligned with 16-byte multiple pitches");
  if (groups == 1) {   // plain 2-D path
    GemmEpi e1 = {};
    e1.out = out;
    e1.ldo = ldo;
    e1.bias_f32 = bias_f32;
    return gemm_launch(A, lda, W, ldw, M, N, K, kModeF32, kind, e1, bn, cg, num_sms, stream);
  }
  const int eb = kind == kKindBf16 ? 2 : 4;
  if (bn <= 0) {
    const long long m128 = (M + 127) / 128, m256 = (M + 255) / 256;
    if (groups * m256 * ((N + 255LL) / 256) * 2 >= num_sms) { bn = 256; cg = 2; }
    else if (groups * m128 * ((N + 127LL) / 128) >= num_sms) { bn = 128; cg = 1; }
    else { bn = 64; cg = 1; }
  }
  CUtensorMap ta, tb;
  if (make_tmap_3d_groups(&ta, A, eb, M, K, lda, a_gs, groups, kBM)) return -1;
  if (make_tmap_3d_groups(&tb, W, eb, N, K, ldw, w_gs, groups, bn / cg)) return -1;
  const int num_k = (K + (128 / eb) - 1) / (128 / eb);
  GemmShape s{M, N, K, 16, 1, num_k, kL2EvictNormal, kL2EvictNormal, groups, 0, nullptr, 0};
  GemmEpi e = {};
  e.out = out;
  e.ldo = ldo;
  e.bias_f32 = bias_f32;
  e.out_gs = out_gs;
  e.bias_gs = bias_gs;
  if (kind == kKindBf16) return dispatch_tile<kModeF32, kKindBf16>(bn, cg, ta, tb, ta, ta, s, e, num_sms, stream);
  return dispatch_tile<kModeF32, kKindTf32>(bn, cg, ta, tb, ta, ta, s, e, num_sms, stream);
}

// Public (library-internal) entry. A: [M,K] lda; W: [N,K] ldw. kind: 0 bf16, 1 tf32(fp32 storage).
// Rasterisation of the persistent tile loop: (row-tiles per group, column-tiles per super-group, L2 hints of A / W).
// Overrides, in order: gemm_raster_override (ovla_debug_gemm_raster, used by the sweep tool), the environment
// (OVLA_GEMM_GROUP, OVLA_GEMM_GROUP_N, OVLA_GEMM_L2 = two letters n/f/l for the A and W loads), the heuristic.
static int g_raster_override[6] = {-1, -1, -1, -1, -1, -1};
void gemm_raster_override(int group_m, int group_n, int l2_a, int l2_b, int sync_seg, int serpentine) {
  g_raster_override[4] = sync_seg;
  g_raster_override[5] = serpentine;
  g_raster_override[0] = group_m;
  g_raster_override[1] = group_n;
  g_raster_override[2] = l2_a;
  g_raster_override[3] = l2_b;
}
// the kernel's tile walk, callable on the host (CPU test: every (row-tile, column-tile) exactly once for any knobs)
void gemm_tile_coords_host(int t, int num_m, int num_n, int group_m, int group_n, int serpentine, int* mb, int* nb) {
  gemm_tile_coords(t, num_m, num_n, group_m, group_n | (serpentine ? kRasterSerpentine : 0), *mb, *nb);
}
static unsigned long long l2_code(int c) { return c == 2 ? kL2EvictLast : c == 1 ? kL2EvictFirst : kL2EvictNormal; }
static void gemm_raster_choice(int M, int N, int K, int eb, int bn, int cg, int num_sms, int& group, int& group_n,
                               unsigned long long& l2_a, unsigned long long& l2_b, int& sync_seg) {
  static int env_group = -1, env_group_n = -1, env_l2[2] = {-1, -1}, env_sync = -1;
  if (env_group < 0) {
    const char* ev = getenv("OVLA_GEMM_GROUP");
    env_group = (ev && atoi(ev) > 0) ? atoi(ev) : 0;
    ev = getenv("OVLA_GEMM_GROUP_N");
    env_group_n = (ev && atoi(ev) > 0) ? atoi(ev) : 0;
    ev = getenv("OVLA_GEMM_WAVESYNC");
    env_sync = ev ? atoi(ev) : -1;
    ev = getenv("OVLA_GEMM_L2");
    for (int i = 0; i < 2; ++i) {
      const char c = (ev && ev[0] && ev[1]) ? ev[i] : 'n';
      env_l2[i] = c == 'l' ? 2 : c == 'f' ? 1 : 0;
    }
  }
  // Heuristic (profiles/r02t_gemm_raster.md).  A persistent grid drifts out of phase: CTAs that share an operand tile
  // read it up to a whole tile duration apart, the L2 (whose useful capacity under this streaming load is a few tens of
  // MB) has dropped it by then and DRAM serves every reader.  Multi-wave problems therefore align their producers
  // (sync_seg K blocks apart: once per tile up to K = 6144, ~60 K blocks beyond), which cut the DRAM reads of the four
  // Llama prefill GEMMs by 2-2.6x and their time by 4-10 %.  With aligned waves: N <= 4096 keeps W resident under
  // narrow row groups (and, for long K, half of the columns at a time); wider N keeps a 16-row-tile activation slab.
  const int num_k = (K + (128 / eb) - 1) / (128 / eb);
  const long long tiles = ((M + 128LL * cg - 1) / (128 * cg)) * ((N + bn - 1) / bn);
  const bool multi_wave = M >= 8192 && tiles >= 2LL * (num_sms / cg);
  group = (N <= 4096 && K <= 4096) ? 8 : 16;
  group_n = 0;
  sync_seg = 0;
  bool serp = false;
  if (multi_wave && K >= 2048) {
    const int parts = (num_k + 95) / 96 > 1 ? (num_k + 32) / 64 : 1;
    sync_seg = (num_k + parts - 1) / parts;
    if (N <= 4096) {
      group = 2;
      if (K > 4096) group_n = 8;
    } else {
      serp = true;   // the W tiles a row group used last are the first ones the next group needs (-5 % DRAM reads)
    }
  }
  int la = 0, lb = 0;
  if (env_group > 0) group = env_group;
  if (env_group_n > 0) group_n = env_group_n;
  if (env_l2[0] > 0) la = env_l2[0];
  if (env_l2[1] > 0) lb = env_l2[1];
  if (g_raster_override[0] > 0) group = g_raster_override[0];
  if (g_raster_override[1] >= 0) group_n = g_raster_override[1];
  if (g_raster_override[2] >= 0) la = g_raster_override[2];
  if (g_raster_override[3] >= 0) lb = g_raster_override[3];
  l2_a = l2_code(la);
  l2_b = l2_code(lb);
  if (env_sync >= 0) sync_seg = env_sync;
  if (g_raster_override[4] >= 0) sync_seg = g_raster_override[4];
  if (g_raster_override[5] >= 0) serp = g_raster_override[5] != 0;
  if (serp) group_n |= kRasterSerpentine;
}

// two zero-initialised words per (device, stream) for the wave alignment of gemm_tcgen05_kernel: launches on one stream
// run one after the other and each leaves the words zeroed; allocated outside stream capture only
static unsigned int* gemm_sync_words(cudaStream_t st) {
  static std::mutex mu;
  static std::map<std::pair<int, cudaStream_t>, unsigned int*> words;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> lk(mu);
  auto it = words.find({dev, st});
  if (it != words.end()) return it->second;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) {
    cudaGetLastError();
    return nullptr;
  }
  unsigned int* p = nullptr;
  if (cudaMalloc(&p, 256) != cudaSuccess || cudaMemset(p, 0, 256) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  words[{dev, st}] = p;
  return p;
}

int gemm_launch(const void* A, long long lda, const void* W, long long ldw, int M, int N, int K, int mode, int kind,
                const GemmEpi& epi, int bn, int cg, int num_sms, cudaStream_t stream, SplitKWs ws) {
  if (M <= 0 || N <= 0 || K <= 0) return set_error("gemm: empty shape M=%d N=%d K=%d", M, N, K);
  const int eb = kind == kKindBf16 ? 2 : 4;
  if (mode == kModeBf16 && (N % 8)) return set_error("gemm: N=%d must be a multiple of 8", N);
  if (mode == kModeSwiGLU && (N % 64)) return set_error("gemm: SwiGLU N=%d must be a multiple of 64", N);
  if (mode == kModeF32 && (N % 4)) return set_error("gemm: N=%d must be a multiple of 4", N);
  if (mode == kModeQkvRope) {
    if (N != 3 * epi.H * 128) return set_error("gemm: fused QKV+RoPE needs N = 3 * H * 128 (head_dim 128)");
    if (!epi.rope_cos || !epi.rope_sin || !epi.k_cache || !epi.v_cache || epi.T <= 0 || (M % epi.T))
      return set_error("gemm: fused QKV+RoPE needs cos/sin tables, KV caches and M = B * T");
    if (epi.pos0 + epi.T > epi.Tmax) return set_error("gemm: position %d exceeds KV capacity %d", epi.pos0 + epi.T, epi.Tmax);
  }
  const bool fused_norm = epi.ss_out || epi.ss_in;
  if (epi.ss_out) {
    if (mode != kModeBf16 || kind != kKindBf16 || (N % 128)) return set_error("gemm: sum-of-squares output needs a bf16 GEMM with N %% 128 == 0");
    if (epi.ss_ld < N / 64 || (epi.ss_ld % 4)) return set_error("gemm: sum-of-squares pitch %d below %d slots (or not a multiple of 4)", epi.ss_ld, N / 64);
  }
  if (epi.ss_in) {
    if (mode != kModeQkvRope && mode != kModeSwiGLU) return set_error("gemm: a row-norm input needs the QKV+RoPE or SwiGLU epilogue");
    if (epi.ss_parts <= 0 || (epi.ss_parts % 4) || epi.ss_parts > epi.ss_ld || (epi.ss_ld % 4) || (reinterpret_cast<uintptr_t>(epi.ss_in) & 15))
      return set_error("gemm: row-norm input needs 16-byte aligned rows of a multiple of 4 partial sums");
  }
  const bool auto_tile = bn <= 0;
  if (bn <= 0) {
    // Tile heuristic.  Large problems: CTA-pair 256x256 tiles (tensor-bound).  Small M (bs=1 prefill, M = 261..288,
    // and the batched decode steps, M = B): the GEMM is a weight stream whose speed is set by shared-memory fill
    // traffic (every CTA re-loads the activation tile), so wide single-CTA tiles win
    // (tools/gemm_smallm_bench.py, profiles/r01_gemm_smallm.jsonl).
    const long long t256 = ((M + 255LL) / 256) * ((N + 255LL) / 256);
    const int m128 = (M + 127) / 128, m256 = (M + 255) / 256;
    if (M <= 512) {
      if (N >= 8192) { bn = 256; cg = (m256 * 2 == m128) ? 2 : 1; }   // a CTA pair only if it adds no empty 128-row tile
      else { bn = (m128 >= 3) ? 128 : 64; cg = 1; }
    }
    else if (t256 * 2 >= num_sms) { bn = 256; cg = 2; }
    else if (1LL * m128 * ((N + 127LL) / 128) >= num_sms) { bn = 128; cg = 1; }
    else { bn = 64; cg = 1; }
  }
  if (mode == kModeSwiGLU && bn < 64) return set_error("gemm: SwiGLU needs bn >= 64");
  if (mode == kModeQkvRope && bn < 128) { bn = 128; cg = 1; }  // a tile must hold whole 128-wide heads
  if (epi.ss_out && bn < 128) { bn = 128; cg = 1; }            // a thread must own whole (128-column group, parity) slots
  // rasterisation (tools/gemm_raster_sweep.py, profiles/r02s_gemm_raster.md): see gemm_raster_choice
  static int g_split = -2;
  if (g_split == -2) { const char* ev = getenv("OVLA_SPLITK"); g_split = ev ? atoi(ev) : -1; }  // -1 auto, 0/1 off, n forced
  int group, group_n, sync_seg;
  unsigned long long l2_a, l2_b;
  gemm_raster_choice(M, N, K, eb, bn, cg, num_sms, group, group_n, l2_a, l2_b, sync_seg);

  // split-K for the small-M weight-streaming shapes (see splitk.cu)
  const int num_k = (K + (128 / eb) - 1) / (128 / eb);
  int split = 1;
  if (auto_tile && !fused_norm && g_split != 0 && g_split != 1 && splitk_eligible(M, N, kind) &&
      (kind == kKindBf16 ? (mode == kModeBf16 || mode == kModeSwiGLU || mode == kModeF32) : mode == kModeF32) && (mode == kModeSwiGLU ? (N / 2) % 8 == 0 : N % 8 == 0)) {
    if (kind == kKindBf16 && N < 8192) { bn = 128; cg = 1; }
    const long long tiles_mn = ((M + 128LL * cg - 1) / (128 * cg)) * ((N + bn - 1) / bn);
    // pick the slice count that minimises (waves of CTAs) x (K blocks per slice) + the reduce kernel (fixed cost +
    // its fp32 workspace traffic at ~2 MB per unit), in units of one K block (~0.35 us); split only long-K problems
    // and only for a >= 15 % modelled gain (calibrated on tools/gemm_smallm_bench.py)
    int want = 1;
    if (g_split > 1) {
      want = g_split;
    } else if (num_k >= 32) {
      const long long slots = num_sms / cg;
      long long best = ((tiles_mn + slots - 1) / slots) * num_k;
      const long long base = best;
      for (int S = 2; S <= 8 && num_k / S >= 8; ++S) {
        const long long cost = ((tiles_mn * S + slots - 1) / slots) * ((num_k + S - 1) / S) + 8 +
                               (8LL * S * M * N) / (2 << 20);
        if (cost < best) { best = cost; want = S; }
      }
      if (best * 100 > base * 85) want = 1;
    }
    want = std::min(want, 8);
    while (want >= 2 && 1LL * want * M * N > ws.floats) --want;
    if (want >= 2 && ws.ptr) {
      const int kps = (num_k + want - 1) / want;
      split = (num_k + kps - 1) / kps;
    }
  }
  CUtensorMap ta, tb;
  if (make_tmap_2d(&ta, A, eb, M, K, lda, kBM)) return -1;
  if (make_tmap_2d(&tb, W, eb, N, K, ldw, bn / cg)) return -1;
  if (split >= 2) {
    const int kps = (num_k + split - 1) / split;
    GemmShape s{M, N, K, group, split, kps, l2_a, l2_b, 1, group_n, nullptr, 0};
    GemmEpi pe = {};
    pe.out = ws.ptr;
    pe.ldo = N;
    pe.ldr = 1LL * M * N;  // slice stride
    if (kind == kKindBf16) OVLA_TRY((dispatch_tile<kModePartial, kKindBf16>(bn, cg, ta, tb, ta, ta, s, pe, num_sms, stream)));
    else OVLA_TRY((dispatch_tile<kModePartial, kKindTf32>(bn, cg, ta, tb, ta, ta, s, pe, num_sms, stream)));
    return splitk_epilogue_launch(mode, ws.ptr, 1LL * M * N, N, split, M, N, epi, stream);
  }
  GemmShape s{M, N, K, group, 1, num_k, l2_a, l2_b, 1, group_n, nullptr, 0};
  if (sync_seg > 0 && (s.sync = gemm_sync_words(stream)) != nullptr) s.sync_seg = sync_seg;
  if (kind == kKindBf16 && (mode == kModeBf16 || mode == kModeSwiGLU)) {
    // epilogue through shared memory + TMA (OVLA_GEMM_TMA_EPI=0 keeps the direct per-row stores); needs 16-byte
    // aligned bases and pitches, otherwise the direct path runs
    static int g_tma_epi = -1;
    if (g_tma_epi < 0) { const char* ev = getenv("OVLA_GEMM_TMA_EPI"); g_tma_epi = (ev && ev[0] == '0') ? 0 : 1; }
    auto aligned = [](const void* ptr, long long ld) { return !(reinterpret_cast<uintptr_t>(ptr) & 15) && !((ld * 2) & 15); };
    GemmEpi e2 = epi;
    CUtensorMap to = ta, tr = ta;
    if (g_tma_epi && aligned(epi.out, epi.ldo) && (!epi.resid || aligned(epi.resid, epi.ldr))) {
      if (make_tmap_2d_box(&to, epi.out, 2, M, mode == kModeSwiGLU ? N / 2 : N, epi.ldo, 32, 64)) return -1;
      if (epi.resid && make_tmap_2d_box(&tr, epi.resid, 2, M, N, epi.ldr, 32, 64)) return -1;
      e2.tma_epi = 1;
    }
    if (mode == kModeSwiGLU) return dispatch_tile<kModeSwiGLU, kKindBf16>(bn, cg, ta, tb, to, tr, s, e2, num_sms, stream);
    return dispatch_tile<kModeBf16, kKindBf16>(bn, cg, ta, tb, to, tr, s, e2, num_sms, stream);
  }
  if (kind == kKindBf16) {
    if (mode == kModeF32) return dispatch_tile<kModeF32, kKindBf16>(bn, cg, ta, tb, ta, ta, s, epi, num_sms, stream);
    if (mode == kModeQkvRope) return dispatch_tile<kModeQkvRope, kKindBf16>(bn, cg, ta, tb, ta, ta, s, epi, num_sms, stream);
  } else {
    if (mode == kModeF32) return dispatch_tile<kModeF32, kKindTf32>(bn, cg, ta, tb, ta, ta, s, epi, num_sms, stream);
  }
  return set_error("gemm: unsupported mode=%d kind=%d", mode, kind);
}

}  // namespace ovla
