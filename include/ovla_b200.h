/*
 * libovla_b200 -- C ABI of the B200-native OpenVLA predict_action + hidden-state-capture + probe path.
 *
 * The reference (helenlu66/openvla-probe) is 100 % Python and has no FFI / plugin layer; its boundary for
 * this path is a Python method surface.  Each entry point below names the reference interface it stands
 * behind (paths relative to the reference tree).  All pointers named *_dev are CUDA device pointers on the
 * engine's device, *_host are host pointers; `stream` is a cudaStream_t passed as void* (NULL = default
 * stream).  Every function returns 0 on success and -1 on failure; ovla_last_error() then holds the
 * message for the calling thread.  Handles are not thread-safe; the caller owns every buffer it passes.
 * No torch types cross this boundary.
 */
#ifndef OVLA_B200_H_
#define OVLA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OVLA_ABI_VERSION 2

/* ------------------------------------------------------------------ library */
int ovla_abi_version(void);
const char* ovla_last_error(void);
/* number of CUDA kernels this library has launched since the last reset (bench.py "gpu_launches") */
long long ovla_launch_count(void);
void ovla_reset_launch_count(void);

/* Live kernel timing (bench.py roofline): when enabled, each kernel launch is bracketed by CUDA events recorded on
 * its launch stream.  ovla_profile_collect synchronises those events and sums, per category, the number of launches,
 * the device milliseconds, and the algorithmic flops / bytes the launches were asked to do; it then clears the log. */
enum { OVLA_CAT_GEMM = 0, OVLA_CAT_GEMV, OVLA_CAT_FLASH_ATTN, OVLA_CAT_DECODE_ATTN, OVLA_CAT_NORM, OVLA_CAT_POOL,
       OVLA_CAT_OTHER, OVLA_NUM_CAT };
void ovla_profile_enable(int on);
int ovla_profile_collect(long long* launches, double* ms, double* flops, double* bytes);
/* Measurement knob (tools/gemm_raster_sweep.py): override the tile rasterisation of every following ovla_gemm launch
 * of this process: row-tiles per group, column-tiles per super-group (0 = all), L2 eviction hint of the A / W tile
 * loads (0 normal, 1 evict-first, 2 evict-last), K blocks between two alignment points of the CTAs' TMA producers
 * (0 = off), serpentine column order of alternate row groups (0 / 1); a negative value returns that knob to the
 * launcher's heuristic.                                                                                          */
void ovla_debug_gemm_raster(int group_m, int group_n, int l2_a, int l2_b, int sync_seg, int serpentine);
/* HOST-side evaluation of the GEMM kernel's tile walk (no GPU needed): tile t = 0 .. num_m * num_n - 1 of a problem of
 * num_m x num_n output tiles is (mb_out[t], nb_out[t]) for the given rasterisation knobs.                        */
int ovla_debug_gemm_tile_order(int num_m, int num_n, int group_m, int group_n, int serpentine, int* mb_out, int* nb_out);

/* ------------------------------------------------------------------ operators (device pointers)
 * Building blocks of PrismaticForConditionalGeneration.forward (prismatic/extern/hf/modeling_prismatic.py:291-447),
 * exposed one by one so that parity tests can check each kernel against the CPU oracle.            */

/* epilogue of a tcgen05 GEMM  out[M,N] = A[M,K] . W[N,K]^T */
typedef struct OvlaGemmEpilogue {
  const void* bias_bf16;  /* [N] or NULL                         (nn.Linear bias)                      */
  const void* scale_bf16; /* [N] or NULL: LayerScale.scale_factor (modeling_prismatic.py:52-59)        */
  const void* resid_bf16; /* [M,N] or NULL: residual added last   (timm Block: x = x + ls(f(norm(x)))) */
  long long ld_resid;     /* row pitch of resid in elements                                           */
  const float* bias_f32;  /* [N] or NULL (fp32-output modes only)                                      */
  int gelu;               /* 1: exact-erf GELU after the bias (nn.GELU, modeling_prismatic.py:139-144) */
  int round_bf16;         /* fp32 output: round the value to bf16 first (HF logits.float())            */
  /* LlamaRMSNorm (transformers modeling_llama.py, reached from modeling_prismatic.py:404-415) fused ACROSS two GEMMs:
   * the GEMM that writes the residual stream (o_proj / down_proj, OVLA_GEMM_BF16 with resid) also writes per-row
   * partial sums of squares of the bf16 values it stores: row_sumsq_out[row * row_sumsq_ld + 2 g + p] for the
   * 128-column group g and 32-column-chunk parity p (N / 64 floats per row, N % 128 == 0);  the GEMM that consumes the
   * normalised rows (OVLA_GEMM_SWIGLU here, ovla_qkv_rope_gemm_rownorm) is given A = the UN-normalised rows, W = the
   * weight with the norm weight folded in (ovla_fold_norm_weight) and row_sumsq_in: it scales row r of its fp32
   * accumulators by rsqrt(sum_{s < row_sumsq_parts} in[r * ld + s] / K + norm_eps) before their first bf16 rounding.
   * Rounding-point change vs the un-fused path (ovla_rmsnorm, then a plain GEMM): bf16(x * rstd) and the product with
   * the norm weight are no longer rounded per element; instead the folded weight is rounded once at bind time.  */
  float* row_sumsq_out;
  const float* row_sumsq_in;
  int row_sumsq_ld, row_sumsq_parts;
  float norm_eps;
} OvlaGemmEpilogue;

enum { OVLA_GEMM_BF16 = 0, OVLA_GEMM_SWIGLU = 1, OVLA_GEMM_F32OUT = 2 };
enum { OVLA_KIND_BF16 = 0, OVLA_KIND_TF32 = 1 };

/* mode: OVLA_GEMM_*; kind: OVLA_KIND_* (TF32 only with F32OUT); tile_n/cta_group: 0,0 = heuristic.
 * OVLA_GEMM_SWIGLU expects W rows interleaved [32 gate rows | 32 up rows] (see ovla_interleave_gate_up)
 * and writes out[M, N/2] = silu(gate) * up (LlamaMLP, transformers modeling_llama.py).              */
int ovla_gemm(const void* a_dev, long long lda, const void* w_dev, long long ldw, int M, int N, int K, int mode,
              int kind, void* out_dev, long long ldo, const OvlaGemmEpilogue* epi, int tile_n, int cta_group,
              void* stream);
/* `groups` independent fp32-output GEMMs of one shape in ONE persistent launch (the probes of all captured layers
 * train concurrently: train_object_probes.py:208-232 loops over the layers one at a time):
 *   out_g[M,N] = a_g[M,K] . w_g[N,K]^T (+ bias_f32_g),  operands of consecutive groups `*_gs` ELEMENTS apart.
 * kind: OVLA_KIND_*; every base / pitch / group stride must be a multiple of 16 bytes.  sm_limit > 0 caps the number
 * of SMs the persistent grid occupies (leave room for a collective running beside it), 0 = all SMs.            */
int ovla_gemm_grouped(const void* a_dev, long long lda, long long a_gs, const void* w_dev, long long ldw, long long w_gs,
                      int groups, int M, int N, int K, int kind, float* out_dev, long long ldo, long long out_gs,
                      const float* bias_f32_dev, long long bias_gs, int tile_n, int cta_group, int sm_limit,
                      void* stream);

/* LayerNorm(eps, affine) over rows of a bf16 [rows, D] matrix (timm Block.norm1/norm2). */
int ovla_layernorm(const void* x_dev, long long ldx, const void* w_dev, const void* b_dev, float eps, void* out_dev,
                   long long ldo, int rows, int D, void* stream);
/* LlamaRMSNorm.forward (transformers modeling_llama.py): w * bf16(x * rsqrt(mean(x^2) + eps)). */
int ovla_rmsnorm(const void* x_dev, long long ldx, const void* w_dev, float eps, void* out_dev, long long ldo,
                 int rows, int D, void* stream);
/* softmax(QK^T/sqrt(d))V. strides12 = element strides (batch, token, head) of Q, K, V, O in that order.
 * head_dim 64 / 72 (timm Attention via SDPA) or 128 (Llama, causal=1). */
int ovla_flash_attention(const void* q_dev, const void* k_dev, const void* v_dev, void* o_dev,
                         const long long* strides12, int B, int H, int Tq, int Tk, int head_dim, int causal,
                         void* stream);
/* Causal self-attention of a prefill on the tcgen05 tensor cores (head_dim 128 only): rotated queries in the fused
 * qkv buffer [B*T, ld_q] (head h at columns h*128), keys / values in the cache [B, H, Tmax, 128] whose rows >= T must
 * hold finite values; out [B*T, ldo].  Same math as ovla_flash_attention(causal=1) (LlamaAttention via SDPA,
 * transformers modeling_llama.py; called from prismatic/extern/hf/modeling_prismatic.py:330-341). */
int ovla_prefill_attention_tc(const void* q_dev, long long ld_q, const void* k_cache_dev, const void* v_cache_dev,
                              void* out_dev, long long ldo, int B, int H, int T, int Tmax, void* stream);
/* The same tensor-core kernel for q | k | v packed in one [B*T, 3*H*hd] buffer (timm Attention.qkv output,
 * modeling_prismatic.py:85-87 towers): head_dim 64 or 128, causal 0/1. */
int ovla_attention_tc_qkv(const void* qkv_dev, long long ld, void* out_dev, long long ldo, int B, int H, int T,
                          int head_dim, int causal, void* stream);
/* in-place RoPE on q of a fused [B*T, 3*H*hd] qkv buffer + rotated-k / v write into the KV cache at pos0+t */
int ovla_rope_kv(void* qkv_dev, int B, int T, int H, int head_dim, int pos0, const void* cos_dev, const void* sin_dev,
                 void* k_cache_dev, void* v_cache_dev, int Tmax, void* stream);
/* pool_tokens (experiments/robot/openvla_utils.py:126-137) for a whole batch: x bf16 [B, *, D] -> out fp32 [B, D].
 * mode 0 = mean over rows [0, n_rows), 1 = "final" (row n_rows-1). */
int ovla_pool_tokens(const void* x_dev, long long batch_stride, long long ld, int B, int n_rows, int D, int mode,
                     float* out_dev, long long out_batch_stride, void* stream);
/* the same for right-padded rows: row b pools over n_rows - (P - lens_dev[b]) rows ("pool over each
 * sample's true length") */
int ovla_pool_tokens_ragged(const void* x_dev, long long batch_stride, long long ld, int B, int n_rows, int D, int mode,
                            const int* lens_dev, int P, float* out_dev, long long out_batch_stride, void* stream);
/* torch.argmax(logits, -1) per row: first index of the maximum, NaN is maximal (greedy step of generate). */
int ovla_argmax(const float* logits_dev, long long ld, int rows, int n, long long* out_dev, void* stream);
/* de-tokenise + un-normalise (modeling_prismatic.py:521-534), float64, bit-identical to numpy. */
int ovla_detokenize(const long long* ids_dev, int n, int action_dim, int vocab_size, const double* centers_dev,
                    int n_centers, const double* q01_dev, const double* q99_dev, const unsigned char* mask_dev,
                    double* out_dev, void* stream);
/* Fused Llama QKV projection for the prefill (LlamaAttention.forward: q/k/v_proj + apply_rotary_pos_emb + cache update):
 * [M = B*T, K] x [3*H*128, K]^T; RoPE(q) is written to qkv_out[:, 0:H*128] (pitch ldo), RoPE(k) and v go straight into
 * the KV cache [B, H, Tmax, 128] at position pos0 + (row % T).  Rounding points as the reference's bf16 ops. */
int ovla_qkv_rope_gemm(const void* a_dev, long long lda, const void* w_dev, long long ldw, int M, int H, int K, int T,
                       int pos0, const void* cos_dev, const void* sin_dev, void* qkv_out_dev, long long ldo,
                       void* k_cache_dev, void* v_cache_dev, int Tmax, int tile_n, int cta_group, void* stream);
/* ovla_qkv_rope_gemm with the preceding input_layernorm fused (see OvlaGemmEpilogue.row_sumsq_in): a_dev holds the
 * un-normalised residual rows, w_dev the q|k|v weight with the norm weight folded in. */
int ovla_qkv_rope_gemm_rownorm(const void* a_dev, long long lda, const void* w_dev, long long ldw, int M, int H, int K,
                               int T, int pos0, const void* cos_dev, const void* sin_dev, void* qkv_out_dev,
                               long long ldo, void* k_cache_dev, void* v_cache_dev, int Tmax, const float* row_sumsq_in_dev,
                               int row_sumsq_ld, int row_sumsq_parts, float norm_eps, int tile_n, int cta_group,
                               void* stream);
/* partial sums of squares of bf16 rows [rows, D] (D % 128 == 0) in the slot layout of OvlaGemmEpilogue.row_sumsq_out
 * (the first layer's input comes from the embedding splice, not from a GEMM) */
int ovla_row_sumsq(const void* x_dev, long long ldx, int rows, int D, float* row_sumsq_out_dev, int row_sumsq_ld,
                   void* stream);
/* out[n, k] = bf16(w[n, k] * gamma[k]) for a bf16 [N, K] weight (K % 8 == 0): the RMSNorm weight folded into the
 * projection that follows it */
int ovla_fold_norm_weight(const void* w_dev, const void* gamma_dev, void* out_dev, long long N, int K, void* stream);
/* PrismaticImageProcessor.apply_transform (processing_prismatic.py:128-145) for frames already at model resolution +
 * the bf16 cast of get_vla_action (openvla_utils.py:186): uint8 HWC [B,S,S,3] -> bf16 [B, 3*n_towers, S, S];
 * mean/std: fp32 [n_towers*3] on the device.  Bit-identical to torchvision's to_tensor + normalize on the host. */
int ovla_preprocess_frames(const void* frames_u8_dev, int B, int S, int n_towers, const float* mean_dev,
                           const float* std_dev, void* pixel_values_out_dev, void* stream);
/* PrismaticImageProcessor.apply_transform up to the uint8 frame (processing_prismatic.py:128-135, set-up :70-123):
 * [letterbox_pad_transform :23-29] -> TVF.resize(bicubic, antialias; on a PIL image = Pillow's ImagingResample) ->
 * TVF.center_crop, uint8 HWC [B,H,W,3] -> uint8 HWC [B,out_size,out_size,3], bit-identical to Pillow / torchvision.
 * strategy (`image_resize_strategy`): OVLA_RESIZE_NAIVE stretch to out_size x out_size (openvla-7b), OVLA_RESIZE_CROP
 * short side -> out_size then center crop, OVLA_RESIZE_LETTERBOX pad to square with fill_{r,g,b} (the reference uses
 * int(255 * mean) of the last tower) then as OVLA_RESIZE_CROP.  Follow with ovla_preprocess_frames.              */
enum { OVLA_RESIZE_NAIVE = 0, OVLA_RESIZE_CROP = 1, OVLA_RESIZE_LETTERBOX = 2 };
int ovla_resize_frames(const void* frames_u8_dev, int B, int H, int W, int strategy, int fill_r, int fill_g, int fill_b,
                       void* out_u8_dev, int out_size, void* stream);
/* get_vla_action(center_crop=True) (experiments/robot/openvla_utils.py:155-175, crop_and_resize :81-124): centred
 * square crop of area crop_scale (side sqrt(crop_scale)), bilinear tf.image.crop_and_resize back to out_size x out_size,
 * uint8 HWC [B,H,W,3] -> uint8 HWC [B,out_size,out_size,3] (float32 [0,1] in between, uint8 by scale 255.5 + truncate
 * as tf.image.convert_image_dtype(saturate=True)). */
int ovla_center_crop_frames(const void* frames_u8_dev, int B, int H, int W, float crop_scale, void* out_u8_dev,
                            int out_size, void* stream);
/* small-batch (M <= 8) weight-streaming GEMM with the same epilogues as ovla_gemm */
int ovla_gemv(const void* x_dev, long long ldx, const void* w_dev, long long ldw, int M, int N, int K, int mode,
              void* out_dev, long long ldo, const OvlaGemmEpilogue* epi, void* stream);
/* one cached decode step of LlamaAttention for B single-token rows: RoPE(q,k) at `pos`, append k/v to the cache
 * [B, H, Tmax, 128], attention over the pos+1 keys; qkv is the raw fused projection [B, 3*H*128] */
int ovla_decode_rope_attention(const void* qkv_dev, long long qkv_ld, const void* cos_dev, const void* sin_dev, int pos,
                               void* k_cache_dev, void* v_cache_dev, int B, int H, int head_dim, int Tmax,
                               void* out_dev, long long o_ld, void* stream);
/* the same step for a ragged batch: row b works at position pos - (P - lens_dev[b]) (int32 [B] lengths out of P) */
int ovla_decode_rope_attention_ragged(const void* qkv_dev, long long qkv_ld, const void* cos_dev, const void* sin_dev,
                                      int pos, const int* lens_dev, int P, void* k_cache_dev, void* v_cache_dev, int B,
                                      int H, int head_dim, int Tmax, void* out_dev, long long o_ld, void* stream);

/* ------------------------------------------------------------------ engine
 * Stands behind OpenVLAForActionPrediction (modeling_prismatic.py:491-562) + get_vla_action's capture
 * (experiments/robot/openvla_utils.py:186-203).  One fused pass: vision towers -> projector -> splice ->
 * Llama prefill (KV cache + per-layer pooled hidden states) -> greedy cached decode.                       */
typedef struct OvlaTower {
  int dim, depth, heads, mlp, n_prefix, layerscale;
} OvlaTower;

typedef struct OvlaDims {
  int image_size, patch;
  int n_towers;        /* 2 = fused DINOv2+SigLIP (prism-dinosiglip-224px), 1 = single backbone */
  OvlaTower towers[2];
  int llm_dim, llm_inter, llm_layers, llm_heads, vocab;
  float rms_eps;
  int max_batch;       /* workspace is sized for this many observations per call */
  int max_seq;         /* n_patches + prompt tokens + generated tokens, upper bound (KV capacity) */
} OvlaDims;

typedef struct OvlaEngine OvlaEngine;

int ovla_create(const OvlaDims* dims, int device, OvlaEngine** out);
void ovla_destroy(OvlaEngine* e);
/* Copy one bf16 tensor of the HF state_dict (names fixed by vla-scripts/extern/convert_openvla_weights_to_hf.py:73-115,
 * e.g. "projector.fc1.weight", "language_model.model.layers.3.mlp.up_proj.weight") from device memory into the
 * engine's packed layout (q/k/v stacked, gate/up interleaved, patch-embed K padded).  Extra names:
 * "rope.cos" / "rope.sin" = bf16 [max_seq, head_dim/2] tables (LlamaRotaryEmbedding, cast to bf16).       */
int ovla_bind_weight(OvlaEngine* e, const char* name, const void* src_dev, const long long* shape, int ndim);
/* Execution options of one engine (A/B measurements and tests; results do not depend on them beyond what the tests
 * state): "decode_mega" (1: a cached decode step at batch <= 4 is one persistent kernel; 0: per-layer kernels),
 * "attn_tc", "fuse_rope", "two_streams", "graph_max_batch" (largest batch replayed from a CUDA graph; 0 = eager).
 * Drops the engine's captured CUDA graphs.                                                                     */
int ovla_set_option(OvlaEngine* e, const char* name, int value);
/* Debug: %globaltimer stamps (ns) of the last persistent decode step, [2][1024] = first and last CTA, one stamp per
 * phase edge (engine created with OVLA_MEGA_TRACE=1 in the environment); tools/decode_trace.py prints the timeline. */
int ovla_debug_decode_trace(OvlaEngine* e, unsigned long long* out_host, int n);
/* verifies that every tensor the path needs has been bound */
int ovla_finalize(OvlaEngine* e);
long long ovla_workspace_bytes(const OvlaEngine* e);
long long ovla_weight_bytes(const OvlaEngine* e);
/* number of passes ovla_run has replayed from a captured CUDA graph (small batches: first call eager, second captures) */
long long ovla_graph_replays(const OvlaEngine* e);

typedef struct OvlaRunArgs {
  const long long* input_ids_dev; /* int64 [B, P], first id = BOS; the host has already appended 29871 */
  const void* pixel_values_dev;   /* bf16 [B, 3*n_towers, S, S]                                         */
  int B, P;
  int pool_len;        /* capture pools hidden-state rows [0, pool_len); 0 = no capture               */
  int pool_mode;       /* 0 mean, 1 final                                                              */
  int n_new_tokens;    /* greedy tokens to generate (action_dim); 0 = prefill only                     */
  float* pooled_out_dev;          /* fp32 [llm_layers+1, B, llm_dim] or NULL                            */
  long long* tokens_out_dev;      /* int64 [B, n_new_tokens] or NULL                                    */
  float* step_logits_out_dev;     /* fp32 [n_new_tokens, B, vocab] or NULL (last-position logits)       */
  void* hidden_out_dev;           /* bf16 [llm_layers+1, B, T, llm_dim] or NULL (forward(output_hidden_states)) */
  void* projector_out_dev;        /* bf16 [B, n_patches, llm_dim] or NULL                               */
  void* patches_out_dev;          /* bf16 [B, n_patches, vision_dim] or NULL (vision backbone output)   */
  /* Ragged batch (the reference is batch-1 only, modeling_prismatic.py:326; its padded-batch splice is :388-390, SURVEY
   * Appendix B): int32 [B] true prompt lengths in [1, P] of RIGHT-padded rows of input_ids (each row: its ids, then
   * 29871, then any valid pad id), or NULL = every row has P ids.  Row b then pools over pool_len - (P - len_b) rows,
   * takes its first token from position n_patches + len_b - 1 and decodes at its own positions, i.e. gives what a
   * batch-1 call on its un-padded prompt gives (causal attention never looks at the pads). The lengths are read on
   * the device at run time (a captured pass may be replayed with new lengths in the same buffer).                 */
  const int* prompt_lens_dev;
} OvlaRunArgs;

/* device-resident inputs/outputs */
int ovla_run(OvlaEngine* e, const OvlaRunArgs* args, void* stream);
/* Same, with HOST buffers (pinned or pageable): copies inputs H2D, runs, copies pooled/tokens back D2H and
 * synchronises the stream.  This is the call the reference-facing predict_action makes.
 * prompt_lens_host: int32 [B] or NULL, see OvlaRunArgs.prompt_lens_dev.                                */
int ovla_run_host(OvlaEngine* e, const long long* input_ids_host, const void* pixel_values_host, int B, int P,
                  int pool_len, int pool_mode, int n_new_tokens, float* pooled_out_host, long long* tokens_out_host,
                  const int* prompt_lens_host, void* stream);

/* ------------------------------------------------------------------ probe training
 * Kernels around the two TF32 tcgen05 GEMMs (ovla_gemm, OVLA_KIND_TF32) of one probe step:
 *   experiment_utils/train_object_probes.py:177-206 (masked, per-label pos_weight), train_spatial_probes.py:153-176
 *   (un-masked mean), train_dual_head_final.py:147-232 (presence + truth heads).  Features X are fp32 [N, D]
 *   resident in HBM; labels int8 in {-1, 0, 1}.                                                              */
/* epoch shuffle: xp[i,:] = x[perm[i],:] (row-major, pitch ldp) and xpt[:,i] = x[perm[i],:] (transposed, pitch ldt) */
int ovla_probe_gather(const float* x_dev, long long ldx, const long long* perm_dev, int n, int D, float* xp_dev,
                      long long ldp, float* xpt_dev, long long ldt, void* stream);
/* yp[i,k] = y[perm[i], keep[k]] for k < K; padding columns [K, Kpad) are -1 */
int ovla_probe_gather_labels(const signed char* y_dev, long long ldy, const long long* perm_dev, const int* keep_dev,
                             int n, int K, int Kpad, signed char* yp_dev, void* stream);
/* BCEWithLogits loss + UN-normalised gradient, written transposed: dzt [heads*Kpad, n] (pitch ldt).
 * kind0: 0 object (masked, vector pos_weight), 1 spatial (all valid), 2 presence (scalar pos_weight);
 * heads = 2 adds the truth head (masked, no pos_weight) reading z[:, Kpad:2Kpad].
 * stats[4] += {loss_h0, count_h0, loss_h1, count_h1} (caller zeroes it).                                    */
int ovla_probe_bce_grad(const float* z_dev, long long ldz, const signed char* yp_dev, int n, int K, int Kpad,
                        int kind0, int heads, const float* pos_weight_dev, float pos_weight_scalar, float* dzt_dev,
                        long long ldt, float* stats_dev, void* stream);
/* The same loss / gradient for `groups` probes that share the labels (one probe per captured layer), fused with the
 * bias gradient and the loss statistics, all deterministic (no floating-point atomics): z [groups][n][ldz] (z_gs
 * apart), dzt [groups][heads*Kpad][ldt] (dzt_gs apart); for group g the bias gradient goes to
 * out_base + g*out_gs + db_off [heads*Kpad] and {loss_h0, count_h0, loss_h1, count_h1} to out_base + g*out_gs +
 * stats_off (overwritten, not accumulated) -- i.e. straight into the flat [dW | db | stats] gradient buffer that is
 * all-reduced.  part_dev: groups * isplits * (heads*Kpad + 4*ceil(Kpad/32)) floats of scratch; ticket_dev: `groups`
 * ints, zero on entry (left zero); isplits (1..64) CTAs share the batch of one (group, label tile).             */
int ovla_probe_bce_grad_grouped(const float* z_dev, long long ldz, long long z_gs, const signed char* yp_dev, int n,
                                int K, int Kpad, int kind0, int heads, const float* pos_weight_dev,
                                float pos_weight_scalar, float* dzt_dev, long long ldt, long long dzt_gs, int groups,
                                float* out_base_dev, long long out_gs, long long db_off, long long stats_off,
                                float* part_dev, int isplits, int* ticket_dev, void* stream);
/* Direct 3-class probe (train_3class_direct.py:147-212): z fp32 [n, 3K] viewed as [n*K, 3], class-weighted CE over
 * {-1 -> 0, 0 -> 1, 1 -> 2}; UN-normalised gradient transposed into dzt [rows_pad >= 3K, n]; stats[0] += sum w*nll,
 * stats[1] += sum w (the weighted-mean normaliser).  class_w3_host: 3 floats on the HOST.                      */
int ovla_probe_ce3_grad(const float* z_dev, long long ldz, const signed char* yp_dev, long long ldy, int n, int K,
                        int rows_pad, const float* class_w3_host, float* dzt_dev, long long ldt, float* stats_dev,
                        void* stream);
/* Validation confusion counts on the device (the reference gathers to the host and calls sklearn:
 * train_object_probes.py:190-206, train_dual_head_final.py:196-232, train_3class_direct.py:196-207).  kind 0 object
 * (mask y != -1), 1 spatial, 2 dual head (presence counts[0..3], truth counts[4..7]), 3 three-class (counts[3*t+p]).
 * Binary layout: tp, fp, fn, tn.  y: int8 [n, *] with `keep_dev` (int32 [K], may be NULL) selecting its columns.   */
int ovla_probe_confusion(const float* z_dev, long long ldz, const signed char* y_dev, long long ldy, const int* keep_dev,
                         int n, int K, int Kpad, int kind, float thresh, unsigned long long* counts9_dev, void* stream);
/* Per-label counts for eval_probes_per_label.py:59-96 (mask y != -1, target y == 1, pred sigmoid(z) > thresh):
 * counts_k4_dev[k*4 + {tp, fp, fn, tn}] for each kept label k (uint64 [K, 4]).                                     */
int ovla_probe_confusion_per_label(const float* z_dev, long long ldz, const signed char* y_dev, long long ldy,
                                   const int* keep_dev, int n, int K, float thresh, unsigned long long* counts_k4_dev,
                                   void* stream);
/* out[r] = sum_c a[r, c]  (bias gradient from dzt) */
int ovla_probe_rowsum(const float* a_dev, long long lda, int rows, int cols, float* out_dev, void* stream);
/* torch.optim.AdamW step on the flat [W (rows x D) | b (rows)] buffer; gradient rows of head h are divided by
 * stats[2h+1] (the global count of that head's loss terms) on the device.                                  */
int ovla_probe_adamw(float* p_dev, const float* g_dev, float* m_dev, float* v_dev, long long n_w, int D,
                     int rows_per_head, long long n_total, const float* stats_dev, float lr, float beta1, float beta2,
                     float eps, float wd, int step, void* stream);

/* AdamW for `groups` probes in one launch: parameters / moments [groups][n_total] contiguous, gradients g_gs apart and
 * statistics stats_gs apart (both inside the flat all-reduced buffer).                                          */
int ovla_probe_adamw_grouped(float* p_dev, const float* g_dev, float* m_dev, float* v_dev, int groups, long long n_w,
                             int D, int rows_per_head, long long n_total, long long g_gs, const float* stats_dev,
                             long long stats_gs, float lr, float beta1, float beta2, float eps, float wd, int step,
                             void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OVLA_B200_H_ */
