// extern "C" surface of libovla_b200 (see include/ovla_b200.h).
#include "../../include/ovla_b200.h"

#include "gemm.cuh"
#include "host_util.h"
#include "ops.h"

using namespace ovla;

static int g_num_sms = 0;
int ovla::num_sms() {
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

static GemmEpi to_epi(void* out, long long ldo, const OvlaGemmEpilogue* epi) {
  GemmEpi e = {};
  e.out = out;
  e.ldo = ldo;
  if (epi) {
    e.bias = static_cast<const __nv_bfloat16*>(epi->bias_bf16);
    e.scale = static_cast<const __nv_bfloat16*>(epi->scale_bf16);
    e.resid = static_cast<const __nv_bfloat16*>(epi->resid_bf16);
    e.ldr = epi->ld_resid;
    e.bias_f32 = epi->bias_f32;
    e.gelu = epi->gelu;
    e.round_bf16 = epi->round_bf16;
  }
  return e;
}


extern "C" {

int ovla_abi_version(void) { return OVLA_ABI_VERSION; }
const char* ovla_last_error(void) { return last_error(); }
long long ovla_launch_count(void) { return launch_count(); }
void ovla_reset_launch_count(void) { reset_launch_count(); }
int ovla_debug_gemm_tile_order(int num_m, int num_n, int group_m, int group_n, int serpentine, int* mb_out, int* nb_out) {
  if (num_m <= 0 || num_n <= 0 || group_m <= 0 || group_n < 0 || !mb_out || !nb_out) return set_error("tile order: bad arguments");
  for (int t = 0; t < num_m * num_n; ++t) gemm_tile_coords_host(t, num_m, num_n, group_m, group_n, serpentine, mb_out + t, nb_out + t);
  return 0;
}
void ovla_debug_gemm_raster(int group_m, int group_n, int l2_a, int l2_b, int sync_seg, int serpentine) {
  gemm_raster_override(group_m, group_n, l2_a, l2_b, sync_seg, serpentine);
}
void ovla_profile_enable(int on) { prof_enable(on != 0); }
int ovla_profile_collect(long long* launches, double* ms, double* flops, double* bytes) {
  static_assert(static_cast<int>(OVLA_NUM_CAT) == static_cast<int>(kNumCat), "category enums out of sync");
  if (!launches || !ms || !flops || !bytes) return set_error("ovla_profile_collect: null argument");
  return prof_collect(launches, ms, flops, bytes);
}

int ovla_gemm(const void* a_dev, long long lda, const void* w_dev, long long ldw, int M, int N, int K, int mode,
              int kind, void* out_dev, long long ldo, const OvlaGemmEpilogue* epi, int tile_n, int cta_group,
              void* stream) {
  GemmEpi e = {};
  e.out = out_dev;
  e.ldo = ldo;
  if (epi) {
    e.bias = static_cast<const __nv_bfloat16*>(epi->bias_bf16);
    e.scale = static_cast<const __nv_bfloat16*>(epi->scale_bf16);
    e.resid = static_cast<const __nv_bfloat16*>(epi->resid_bf16);
    e.ldr = epi->ld_resid;
    e.bias_f32 = epi->bias_f32;
    e.gelu = epi->gelu;
    e.round_bf16 = epi->round_bf16;
    e.ss_out = epi->row_sumsq_out;
    e.ss_in = epi->row_sumsq_in;
    e.ss_ld = epi->row_sumsq_ld;
    e.ss_parts = epi->row_sumsq_parts;
    e.ss_inv_k = 1.0f / static_cast<float>(K);
    e.ss_eps = epi->norm_eps;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // do not create a workspace for shapes that never split (gemm.cu)
  const SplitKWs ws = (tile_n <= 0 && splitk_eligible(M, N, kind)) ? splitk_stream_workspace(st) : SplitKWs{nullptr, 0};
  return gemm_launch(a_dev, lda, w_dev, ldw, M, N, K, mode, kind, e, tile_n, cta_group, num_sms(), st, ws);
}


int ovla_gemm_grouped(const void* a_dev, long long lda, long long a_gs, const void* w_dev, long long ldw, long long w_gs,
                      int groups, int M, int N, int K, int kind, float* out_dev, long long ldo, long long out_gs,
                      const float* bias_f32_dev, long long bias_gs, int tile_n, int cta_group, int sm_limit,
                      void* stream) {
  if (!a_dev || !w_dev || !out_dev) return set_error("ovla_gemm_grouped: null buffer");
  int sms = num_sms();
  if (sm_limit > 0 && sm_limit < sms) sms = sm_limit < 2 ? 2 : sm_limit;
  return gemm_grouped_launch(a_dev, lda, a_gs, w_dev, ldw, w_gs, groups, M, N, K, kind, out_dev, ldo, out_gs,
                             bias_f32_dev, bias_gs, tile_n, cta_group, sms, static_cast<cudaStream_t>(stream));
}

int ovla_qkv_rope_gemm(const void* a_dev, long long lda, const void* w_dev, long long ldw, int M, int H, int K, int T,
                       int pos0, const void* cos_dev, const void* sin_dev, void* qkv_out_dev, long long ldo,
                       void* k_cache_dev, void* v_cache_dev, int Tmax, int tile_n, int cta_group, void* stream) {
  GemmEpi e = {};
  e.out = qkv_out_dev;
  e.ldo = ldo;
  e.rope_cos = static_cast<const __nv_bfloat16*>(cos_dev);
  e.rope_sin = static_cast<const __nv_bfloat16*>(sin_dev);
  e.k_cache = static_cast<__nv_bfloat16*>(k_cache_dev);
  e.v_cache = static_cast<__nv_bfloat16*>(v_cache_dev);
  e.T = T;
  e.pos0 = pos0;
  e.Tmax = Tmax;
  e.H = H;
  return gemm_launch(a_dev, lda, w_dev, ldw, M, 3 * H * 128, K, kModeQkvRope, kKindBf16, e, tile_n, cta_group, num_sms(),
                     static_cast<cudaStream_t>(stream));
}

int ovla_qkv_rope_gemm_rownorm(const void* a_dev, long long lda, const void* w_dev, long long ldw, int M, int H, int K,
                               int T, int pos0, const void* cos_dev, const void* sin_dev, void* qkv_out_dev,
                               long long ldo, void* k_cache_dev, void* v_cache_dev, int Tmax, const float* ss_in, int ss_ld,
                               int ss_parts, float norm_eps, int tile_n, int cta_group, void* stream) {
  if (!ss_in) return set_error("ovla_qkv_rope_gemm_rownorm: null row sums");
  GemmEpi e = {};
  e.out = qkv_out_dev;
  e.ldo = ldo;
  e.rope_cos = static_cast<const __nv_bfloat16*>(cos_dev);
  e.rope_sin = static_cast<const __nv_bfloat16*>(sin_dev);
  e.k_cache = static_cast<__nv_bfloat16*>(k_cache_dev);
  e.v_cache = static_cast<__nv_bfloat16*>(v_cache_dev);
  e.T = T;
  e.pos0 = pos0;
  e.Tmax = Tmax;
  e.H = H;
  e.ss_in = ss_in;
  e.ss_ld = ss_ld;
  e.ss_parts = ss_parts;
  e.ss_inv_k = 1.0f / static_cast<float>(K);
  e.ss_eps = norm_eps;
  return gemm_launch(a_dev, lda, w_dev, ldw, M, 3 * H * 128, K, kModeQkvRope, kKindBf16, e, tile_n, cta_group, num_sms(),
                     static_cast<cudaStream_t>(stream));
}
int ovla_row_sumsq(const void* x, long long ldx, int rows, int D, float* ss, int ss_ld, void* stream) {
  if (!x || !ss) return set_error("ovla_row_sumsq: null buffer");
  return row_sumsq_launch(x, ldx, rows, D, ss, ss_ld, static_cast<cudaStream_t>(stream));
}
int ovla_fold_norm_weight(const void* w, const void* gamma, void* out, long long N, int K, void* stream) {
  if (!w || !gamma || !out) return set_error("ovla_fold_norm_weight: null buffer");
  return fold_norm_weight_launch(w, gamma, out, N, K, static_cast<cudaStream_t>(stream));
}

int ovla_gemv(const void* x_dev, long long ldx, const void* w_dev, long long ldw, int M, int N, int K, int mode,
              void* out_dev, long long ldo, const OvlaGemmEpilogue* epi, void* stream) {
  return gemv_launch(x_dev, ldx, w_dev, ldw, M, N, K, mode, to_epi(out_dev, ldo, epi),
                     static_cast<cudaStream_t>(stream));
}
int ovla_layernorm(const void* x, long long ldx, const void* w, const void* b, float eps, void* out, long long ldo,
                   int rows, int D, void* stream) {
  return layernorm_launch(x, ldx, w, b, eps, out, ldo, rows, D, static_cast<cudaStream_t>(stream));
}
int ovla_rmsnorm(const void* x, long long ldx, const void* w, float eps, void* out, long long ldo, int rows, int D,
                 void* stream) {
  return rmsnorm_launch(x, ldx, w, eps, out, ldo, rows, D, static_cast<cudaStream_t>(stream));
}
int ovla_flash_attention(const void* q, const void* k, const void* v, void* o, const long long* strides12, int B, int H,
                         int Tq, int Tk, int head_dim, int causal, void* stream) {
  if (!strides12) return set_error("ovla_flash_attention: null strides");
  return flash_attn_launch(q, k, v, o, strides12, B, H, Tq, Tk, head_dim, causal, static_cast<cudaStream_t>(stream));
}
int ovla_prefill_attention_tc(const void* q, long long ld_q, const void* kc, const void* vc, void* out, long long ldo,
                              int B, int H, int T, int Tmax, void* stream) {
  return attn_tc_prefill_launch(q, ld_q, kc, vc, out, ldo, B, H, T, Tmax, static_cast<cudaStream_t>(stream));
}
int ovla_attention_tc_qkv(const void* qkv, long long ld, void* out, long long ldo, int B, int H, int T, int head_dim,
                          int causal, void* stream) {
  return attn_tc_qkv_launch(qkv, ld, out, ldo, B, H, T, head_dim, causal, static_cast<cudaStream_t>(stream));
}
int ovla_resize_frames(const void* src_u8, int B, int H, int W, int strategy, int fill_r, int fill_g, int fill_b,
                       void* dst_u8, int out_size, void* stream) {
  if (B > 0 && (!src_u8 || !dst_u8)) return set_error("ovla_resize_frames: null buffer");
  return resize_frames_launch(src_u8, B, H, W, strategy, fill_r, fill_g, fill_b, dst_u8, out_size,
                              static_cast<cudaStream_t>(stream));
}
int ovla_center_crop_frames(const void* src_u8, int B, int H, int W, float crop_scale, void* dst_u8, int out_size,
                            void* stream) {
  if (!src_u8 || !dst_u8) return set_error("ovla_center_crop_frames: null buffer");
  return center_crop_launch(src_u8, B, H, W, crop_scale, dst_u8, out_size, static_cast<cudaStream_t>(stream));
}
int ovla_probe_confusion(const float* z, long long ldz, const signed char* y, long long ldy, const int* keep, int n, int K,
                         int Kpad, int kind, float thresh, unsigned long long* counts9, void* stream) {
  if (!z || !y || !counts9) return set_error("ovla_probe_confusion: null buffer");
  return probe_confusion_launch(z, ldz, y, ldy, keep, n, K, Kpad, kind, thresh, counts9, static_cast<cudaStream_t>(stream));
}
int ovla_probe_confusion_per_label(const float* z, long long ldz, const signed char* y, long long ldy, const int* keep,
                                   int n, int K, float thresh, unsigned long long* counts_k4, void* stream) {
  if (!z || !y || !counts_k4) return set_error("ovla_probe_confusion_per_label: null buffer");
  return probe_confusion_per_label_launch(z, ldz, y, ldy, keep, n, K, thresh, counts_k4, static_cast<cudaStream_t>(stream));
}
int ovla_decode_rope_attention(const void* qkv, long long qkv_ld, const void* cos_dev, const void* sin_dev, int pos,
                               void* kc, void* vc, int B, int H, int head_dim, int Tmax, void* out, long long o_ld,
                               void* stream) {
  return decode_rope_attn_launch(qkv, qkv_ld, cos_dev, sin_dev, pos, kc, vc, B, H, head_dim, Tmax, out, o_ld,
                                 static_cast<cudaStream_t>(stream));
}
int ovla_decode_rope_attention_ragged(const void* qkv, long long qkv_ld, const void* cos_dev, const void* sin_dev, int pos,
                                      const int* lens, int P, void* kc, void* vc, int B, int H, int head_dim, int Tmax,
                                      void* out, long long o_ld, void* stream) {
  if (!lens || P < 1) return set_error("ovla_decode_rope_attention_ragged: lens / P missing");
  return decode_rope_attn_launch(qkv, qkv_ld, cos_dev, sin_dev, pos, kc, vc, B, H, head_dim, Tmax, out, o_ld,
                                 static_cast<cudaStream_t>(stream), lens, P);
}
int ovla_rope_kv(void* qkv, int B, int T, int H, int head_dim, int pos0, const void* cos_dev, const void* sin_dev,
                 void* kc, void* vc, int Tmax, void* stream) {
  return rope_kv_launch(qkv, B, T, H, head_dim, pos0, cos_dev, sin_dev, kc, vc, Tmax, static_cast<cudaStream_t>(stream));
}
int ovla_pool_tokens(const void* x, long long batch_stride, long long ld, int B, int n_rows, int D, int mode,
                     float* out, long long out_batch_stride, void* stream) {
  return pool_tokens_launch(x, batch_stride, ld, B, n_rows, D, mode, out, out_batch_stride,
                            static_cast<cudaStream_t>(stream));
}
int ovla_pool_tokens_ragged(const void* x, long long batch_stride, long long ld, int B, int n_rows, int D, int mode,
                            const int* lens, int P, float* out, long long out_batch_stride, void* stream) {
  if (!lens || P < 1) return set_error("ovla_pool_tokens_ragged: lens / P missing");
  if (n_rows - (P - 1) < 1) return set_error("ovla_pool_tokens_ragged: a row of length 1 would pool nothing");
  return pool_tokens_launch(x, batch_stride, ld, B, n_rows, D, mode, out, out_batch_stride,
                            static_cast<cudaStream_t>(stream), lens, P);
}
int ovla_argmax(const float* logits, long long ld, int rows, int n, long long* out, void* stream) {
  return argmax_launch(logits, ld, rows, n, out, static_cast<cudaStream_t>(stream));
}
int ovla_detokenize(const long long* ids, int n, int action_dim, int vocab_size, const double* centers, int n_centers,
                    const double* q01, const double* q99, const unsigned char* mask, double* out, void* stream) {
  return detok_unnorm_launch(ids, n, action_dim, vocab_size, centers, n_centers, q01, q99, mask, out,
                             static_cast<cudaStream_t>(stream));
}


int ovla_preprocess_frames(const void* frames_u8_dev, int B, int S, int n_towers, const float* mean_dev,
                           const float* std_dev, void* pixel_values_out_dev, void* stream) {
  return preprocess_frames_launch(frames_u8_dev, B, S, n_towers, mean_dev, std_dev, pixel_values_out_dev,
                                  static_cast<cudaStream_t>(stream));
}

int ovla_probe_gather(const float* x, long long ldx, const long long* perm, int n, int D, float* xp, long long ldp,
                      float* xpt, long long ldt, void* stream) {
  return probe_gather_launch(x, ldx, perm, n, D, xp, ldp, xpt, ldt, static_cast<cudaStream_t>(stream));
}
int ovla_probe_gather_labels(const signed char* y, long long ldy, const long long* perm, const int* keep, int n, int K,
                             int Kpad, signed char* yp, void* stream) {
  return probe_gather_labels_launch(y, ldy, perm, keep, n, K, Kpad, yp, static_cast<cudaStream_t>(stream));
}
int ovla_probe_bce_grad(const float* z, long long ldz, const signed char* y, int n, int K, int Kpad, int kind0,
                        int heads, const float* pos_weight, float pos_weight_scalar, float* dzt, long long ldt,
                        float* stats, void* stream) {
  return probe_bce_grad_launch(z, ldz, y, n, K, Kpad, kind0, heads, pos_weight, pos_weight_scalar, dzt, ldt, stats,
                               static_cast<cudaStream_t>(stream));
}
int ovla_probe_ce3_grad(const float* z, long long ldz, const signed char* y, long long ldy, int n, int K, int rows_pad,
                        const float* class_w3_host, float* dzt, long long ldt, float* stats, void* stream) {
  return probe_ce3_grad_launch(z, ldz, y, ldy, n, K, rows_pad, class_w3_host, dzt, ldt, stats,
                               static_cast<cudaStream_t>(stream));
}
int ovla_probe_rowsum(const float* a, long long lda, int rows, int cols, float* out, void* stream) {
  return probe_rowsum_launch(a, lda, rows, cols, out, static_cast<cudaStream_t>(stream));
}
int ovla_probe_adamw(float* p, const float* g, float* m, float* v, long long n_w, int D, int rows_per_head,
                     long long n_total, const float* stats, float lr, float beta1, float beta2, float eps, float wd,
                     int step, void* stream) {
  return probe_adamw_launch(p, g, m, v, n_w, D, rows_per_head, n_total, stats, lr, beta1, beta2, eps, wd, step,
                            static_cast<cudaStream_t>(stream));
}

int ovla_probe_bce_grad_grouped(const float* z, long long ldz, long long z_gs, const signed char* y, int n, int K,
                                int Kpad, int kind0, int heads, const float* pos_weight, float pos_weight_scalar,
                                float* dzt, long long ldt, long long dzt_gs, int groups, float* out_base, long long out_gs,
                                long long db_off, long long stats_off, float* part, int isplits, int* ticket,
                                void* stream) {
  return probe_bce_grad_grouped_launch(z, ldz, z_gs, y, n, K, Kpad, kind0, heads, pos_weight, pos_weight_scalar, dzt, ldt,
                                       dzt_gs, groups, out_base, out_gs, db_off, stats_off, part, isplits, ticket,
                                       static_cast<cudaStream_t>(stream));
}
int ovla_probe_adamw_grouped(float* p, const float* g, float* m, float* v, int groups, long long n_w, int D,
                             int rows_per_head, long long n_total, long long g_gs, const float* stats, long long stats_gs,
                             float lr, float beta1, float beta2, float eps, float wd, int step, void* stream) {
  return probe_adamw_launch(p, g, m, v, n_w, D, rows_per_head, n_total, stats, lr, beta1, beta2, eps, wd, step,
                            static_cast<cudaStream_t>(stream), groups, g_gs, stats_gs);
}

}  // extern "C"
