// Library-internal launch functions (one per kernel family). All return 0 / -1 (set_error).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "gemm.cuh"

namespace ovla {

int num_sms();

// gemm.cu -- tcgen05 GEMM
// `ws`: caller-owned fp32 scratch for split-K partial tiles (one per stream that may run GEMMs concurrently); an empty
// workspace disables split-K
struct SplitKWs {
  float* ptr;
  long long floats;
};
int gemm_launch(const void* A, long long lda, const void* W, long long ldw, int M, int N, int K, int mode, int kind,
                const GemmEpi& epi, int bn, int cg, int num_sms, cudaStream_t stream, SplitKWs ws = SplitKWs{nullptr, 0});
bool splitk_eligible(int M, int N, int kind);
void gemm_raster_override(int group_m, int group_n, int l2_a, int l2_b, int sync_seg, int serpentine);
void gemm_tile_coords_host(int t, int num_m, int num_n, int group_m, int group_n, int serpentine, int* mb, int* nb);
int gemm_grouped_launch(const void* A, long long lda, long long a_gs, const void* W, long long ldw, long long w_gs,
                        int groups, int M, int N, int K, int kind, float* out, long long ldo, long long out_gs,
                        const float* bias_f32, long long bias_gs, int bn, int cg, int num_sms, cudaStream_t stream);
// splitk.cu -- reduce/epilogue of the split-K partial tiles; workspace of the engine-less ovla_gemm entry
SplitKWs splitk_stream_workspace(cudaStream_t st);
int splitk_epilogue_launch(int mode, const float* ws, long long slice_stride, long long ldw, int S, int M, int N,
                           const GemmEpi& epi, cudaStream_t st);
// decode_mega.cu -- one cached decode step (batch <= 4) as one persistent kernel
struct DecodeLayerPtrs {
  const __nv_bfloat16 *ln1, *qkv, *o, *ln2, *gate_up, *down;   // packed [N, K] weights (gate_up rows interleaved 32/32)
};
struct DecodeStepArgs {
  const DecodeLayerPtrs* layers;        // device array [n_layers]
  const __nv_bfloat16 *final_norm, *lm_head, *rope_cos, *rope_sin;
  __nv_bfloat16 *x, *qkv, *attn, *act;  // [M, D] residual (in/out), [M, 3D], [M, D], [M, I] scratch
  __nv_bfloat16* kv;                    // KV cache base; layer stride kv_layer_elems, V half at + kv_layer_elems / 2
  long long kv_layer_elems;
  float* logits;                        // [M, vocab] fp32
  unsigned* barrier;                    // one counter, zeroed by the launcher
  int n_layers, M, D, I, H, head_dim, vocab, Tmax, pos;
  const int* lens;                      // ragged prompts: int32 [M] true prompt lengths out of P (row m decodes at
  int P;                                // pos - (P - lens[m])); null = every row at `pos`
  float eps;
  unsigned long long* trace;            // optional timeline buffer [2][trace_stride] (debug), else null
  int trace_stride;
  int dbg;                              // debug bits (OVLA_MEGA_DBG): 1 = skip the dot products, 2 = skip attention (wrong results)
  // filled by the launcher
  int S, attn_floats;
  float attn_scale;
};
int decode_step_launch(DecodeStepArgs a, cudaStream_t st);
// gemv.cu -- M <= 8 weight streaming
int gemv_launch(const void* x, long long ldx, const void* W, long long ldw, int M, int N, int K, int mode,
                const GemmEpi& epi, cudaStream_t st);
// decode.cu -- fused RoPE + KV append + single-query attention
int decode_rope_attn_launch(const void* qkv, long long qkv_ld, const void* cos_t, const void* sin_t, int pos, void* kc,
                            void* vc, int B, int H, int head_dim, int Tmax, void* out, long long o_ld,
                            cudaStream_t st, const int* lens = nullptr, int P = 0);

// norm.cu
int layernorm_launch(const void* x, long long ldx, const void* w, const void* b, float eps, void* out, long long ldo,
                     int rows, int D, cudaStream_t st);
int rmsnorm_launch(const void* x, long long ldx, const void* w, float eps, void* out, long long ldo, int rows, int D,
                   cudaStream_t st);

// fused RMSNorm support (GemmEpi::ss_out / ss_in): per-row partial sums of squares in the GEMM epilogue's slot layout
// ([rows, ss_ld] floats, D / 64 slots per row), and the norm weight folded into the consuming projection
int row_sumsq_launch(const void* x, long long ldx, int rows, int D, float* ss, int ss_ld, cudaStream_t st);
int fold_norm_weight_launch(const void* w, const void* gamma, void* out, long long N, int K, cudaStream_t st);

// attention.cu
int flash_attn_launch(const void* Q, const void* K, const void* V, void* O, const long long* strides12, int B, int H,
                      int Tq, int Tk, int head_dim, int causal, cudaStream_t st);

// image.cu -- letterbox / PIL-bicubic resize / center crop of uint8 frames (PrismaticImageProcessor.apply_transform)
int resize_frames_launch(const void* frames_u8, int B, int H, int W, int strategy, int fill_r, int fill_g, int fill_b,
                         void* out_u8, int S, cudaStream_t st);
int center_crop_launch(const void* src_u8, int B, int H, int W, float crop_scale, void* dst_u8, int S, cudaStream_t st);

// attention_tc.cu: causal prefill attention on tcgen05 (head_dim 128)
int attn_tc_prefill_launch(const void* q, long long ld_q, const void* kc, const void* vc, void* out, long long ldo, int B,
                           int H, int T, int Tmax, cudaStream_t st);
int attn_tc_qkv_launch(const void* qkv, long long ld, void* out, long long ldo, int B, int H, int T, int hd, int causal,
                       cudaStream_t st);

// elementwise.cu
int im2col_launch(const void* px, int B, int c_total, int c0, int H, int W, int patch, int Kpad, void* out,
                  cudaStream_t st);
int assemble_tokens_launch(const void* patch, const void* pos, const void* cls, const void* reg, int B, int np,
                           int n_prefix, int D, void* tokens, cudaStream_t st);
int copy_rows_launch(const void* src, long long src_batch, long long src_ld, int srow0, void* dst, long long dst_batch,
                     long long dst_ld, int dcol0, int B, int rows, int cols, cudaStream_t st);
int embed_splice_launch(const void* ids, int B, int P, const void* E, int vocab, const void* proj, int np, int D,
                        void* x, int* err_flag, cudaStream_t st);
int rope_kv_launch(void* qkv, int B, int T, int H, int hd, int pos0, const void* cos_t, const void* sin_t, void* kc,
                   void* vc, int Tmax, cudaStream_t st);
int pool_tokens_launch(const void* x, long long batch_stride, long long ld, int B, int n_rows, int D, int mode,
                       float* out, long long out_batch_stride, cudaStream_t st, const int* lens = nullptr, int P = 0);
// out[b, :] = x[b, last - (P - lens[b]), :]  (the last real position of each right-padded row)
int gather_last_rows_launch(const void* x, long long batch_stride, long long ld, int last, const int* lens, int P, int B,
                            int D, void* out, int* err_flag, cudaStream_t st);
int argmax_launch(const float* x, long long ld, int rows, int n, long long* out, cudaStream_t st);
int detok_unnorm_launch(const long long* ids, int n, int action_dim, int vocab_size, const double* centers,
                        int n_centers, const double* q01, const double* q99, const unsigned char* mask, double* out,
                        cudaStream_t st);

int preprocess_frames_launch(const void* frames_u8, int B, int S, int n_towers, const float* mean, const float* stdv,
                             void* out, cudaStream_t st);

// probe.cu
int probe_gather_launch(const float* X, long long ldx, const long long* perm, int n, int D, float* Xp, long long ldp,
                        float* XpT, long long ldt, cudaStream_t st);
int probe_gather_labels_launch(const signed char* Y, long long ldy, const long long* perm, const int* keep, int n,
                               int K, int Kpad, signed char* Yp, cudaStream_t st);
int probe_bce_grad_launch(const float* Z, long long ldz, const signed char* Y, int n, int K, int Kpad, int kind0,
                          int heads, const float* pos_weight, float pos_weight_scalar, float* dZT, long long ldt,
                          float* stats, cudaStream_t st);
int probe_rowsum_launch(const float* A, long long lda, int rows, int cols, float* out, cudaStream_t st);
int probe_adamw_launch(float* p, const float* g, float* m, float* v, long long n_w, int D, int rows_per_head,
                       long long n_total, const float* stats, float lr, float beta1, float beta2, float eps, float wd,
                       int step, cudaStream_t st, int groups = 1, long long g_gs = 0, long long stats_gs = 0);
int probe_bce_grad_grouped_launch(const float* Z, long long ldz, long long z_gs, const signed char* Y, int n, int K,
                                  int Kpad, int kind0, int heads, const float* pos_weight, float pos_weight_scalar,
                                  float* dZT, long long ldt, long long dzt_gs, int groups, float* out_base,
                                  long long out_gs, long long db_off, long long stats_off, float* part, int isplits,
                                  int* ticket, cudaStream_t st);

// probe3.cu
int probe_ce3_grad_launch(const float* Z, long long ldz, const signed char* Y, long long ldy, int n, int K,
                          int rows_pad, const float* class_w3_host, float* dZT, long long ldt, float* stats,
                          cudaStream_t st);
int probe_confusion_launch(const float* Z, long long ldz, const signed char* Y, long long ldy, const int* keep, int n,
                           int K, int Kpad, int kind, float thresh, unsigned long long* counts, cudaStream_t st);
int probe_confusion_per_label_launch(const float* Z, long long ldz, const signed char* Y, long long ldy, const int* keep,
                                     int n, int K, float thresh, unsigned long long* counts, cudaStream_t st);

}  // namespace ovla
