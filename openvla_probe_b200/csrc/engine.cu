// OvlaEngine: owns the packed bf16 weights and the HBM workspace of one GPU and runs the fused
// predict_action + hidden-state-capture pass (see include/ovla_b200.h).
//
// HBM layout (all bf16 unless noted; B = observations per call, N_t = prefix + patches of tower t):
//   weights   : per tower  patch_w [D, Kpad], pos [np, D], cls/reg, per block {ln1, qkv_w [3D, D], proj_w, ls1,
//               ln2, fc1_w [mlp, D], fc2_w [D, mlp], ls2};  projector fc1..3;  LLM embed [V, D], per layer
//               {ln1, qkv_w [3D, D] (q|k|v stacked), o_w, ln2, gate_up_w [2I, D] (32-row interleave), down_w [D, I]},
//               final norm, lm_head [V, D];  RoPE cos/sin [max_seq, hd/2].
//   workspace : ViT   im2col [B*np, Kpad] | patch [B*np, D] | x [B*N, D] | h [B*N, D] | qkv [B*N, 3D] |
//                     attn [B*N, D] | mlp [B*N, mlp]       (sized for the wider tower, reused by both)
//               proj  patches [B*np, Dv] | p1 [B*np, 4Dv] | p2 [B*np, D_llm] | p3 [B*np, D_llm]
//               LLM   x [B*T, D] | h [B*T, D] | qkv [B*T, 3D] | attn [B*T, D] | act [B*T, I]
//               KV    [layers][2][B, H, max_seq, hd]
//               out   pooled fp32 [layers+1, B, D] | logits fp32 [B, V] | tokens int64 [n_new, B]
#include <math.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/ovla_b200.h"
#include "host_util.h"
#include "ops.h"

using namespace ovla;
typedef __nv_bfloat16 bf16;

namespace {

struct Slot {
  bf16* ptr = nullptr;
  long long rows = 0, cols = 0;  // logical 2-D shape of the packed destination
  bool bound = false;
  bool required = true;
};

struct BlockW {
  Slot ln1_w, ln1_b, qkv_w, qkv_b, proj_w, proj_b, ls1, ln2_w, ln2_b, fc1_w, fc1_b, fc2_w, fc2_b, ls2;
};
struct TowerW {
  Slot patch_w, patch_b, pos, cls, reg;
  std::vector<BlockW> blocks;
};
struct LayerW {
  Slot ln1, qkv_w, o_w, ln2, gate_up_w, down_w;
  // copies of qkv_w / gate_up_w with the RMSNorm weight of the norm in front folded in (made by ovla_finalize): the
  // large-batch prefill reads these and applies 1/rms as a row scale in the GEMM epilogue (no stand-alone norm kernel)
  bf16 *qkv_wn = nullptr, *gate_up_wn = nullptr;
  bool q_bound = false, k_bound = false, v_bound = false, gate_bound = false, up_bound = false;
};

inline long long round_up(long long x, long long m) { return (x + m - 1) / m * m; }

struct GraphKey {
  OvlaRunArgs a;  // every pointer and size the captured pass depends on
  bool operator<(const GraphKey& o) const { return memcmp(&a, &o.a, sizeof(a)) < 0; }
};
struct GraphEntry {
  cudaGraphExec_t exec = nullptr;
  long long launches = 0;
};

}  // namespace

struct OvlaEngine {
  OvlaDims d;
  int device = 0;
  int np = 0, kpad = 0, vision_dim = 0, head_dim = 0;
  int n_run[2] = {0, 0};  // blocks actually executed per tower (depth - 1: the last block's output is never used)

  // weights
  std::vector<void*> allocs;
  long long weight_bytes = 0, workspace_bytes = 0;
  TowerW tower[2];
  Slot pj_w[3], pj_b[3];
  int n_proj = 0;
  Slot embed, final_norm, lm_head, rope_cos, rope_sin;
  std::vector<LayerW> layers;

  // workspace
  struct VitBufs { bf16 *im2col = nullptr, *patch = nullptr, *x = nullptr, *h = nullptr, *qkv = nullptr, *attn = nullptr, *mlp = nullptr; };
  VitBufs vb[2];            // [1] is a small second set: at B <= kTwoStreamBatch the two towers run on two streams
  static constexpr int kTwoStreamBatch = 8;
  cudaStream_t side_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bf16 *p_cat = nullptr, *p_1 = nullptr, *p_2 = nullptr, *p_3 = nullptr;
  bf16 *l_x = nullptr, *l_h = nullptr, *l_qkv = nullptr, *l_attn = nullptr, *l_act = nullptr;
  bf16* kv = nullptr;
  float *pooled = nullptr, *logits = nullptr;
  long long* tokens = nullptr;
  int* err_flag = nullptr;
  // split-K partial-tile scratch, one per stream this engine issues GEMMs on (never shared with another engine)
  SplitKWs ws_main = {nullptr, 0}, ws_side = {nullptr, 0};
  // persistent decode-step kernel (decode_mega.cu): per-layer weight pointers + the grid-barrier counter on the device
  DecodeLayerPtrs* mega_layers = nullptr;
  unsigned* mega_barrier = nullptr;
  unsigned long long* mega_trace = nullptr;   // OVLA_MEGA_TRACE=1: phase timeline of the last persistent decode step
  bool decode_mega = true;   // OVLA_DECODE_MEGA=0: per-layer kernels
  // CUDA-graph cache for launch-bound small batches (see ovla_run)
  int graph_max_batch = 16;
  bool two_streams = true;  // OVLA_TWO_STREAMS=0: towers back to back on one stream
  bool attn_tc = true;    // OVLA_ATTN_TC=0 falls back to the mma.sync flash kernel for head_dim 64 / 128 (A/B runs)
  bool fuse_rope = true;  // OVLA_FUSE_ROPE=0 keeps the stand-alone RoPE kernel (A/B measurements)
  // OVLA_FUSE_NORM=0 keeps the stand-alone RMSNorm kernels in the large-batch prefill.  Fused: o_proj / down_proj write
  // per-row partial sums of squares of the residual stream from their epilogues, the following QKV / gate-up GEMM
  // (norm weight folded into its weights) scales its accumulators by 1/rms.
  bool fuse_norm = true;
  int fuse_norm_min_rows = 513;   // below: split-K / weight-streaming shapes keep the un-fused kernels
                                  // (OVLA_FUSE_NORM_MIN_ROWS at engine creation; tests lower it to reach the path)
  float* l_ss = nullptr;                         // [max_batch * max_seq, llm_dim / 64] partial sums of squares
  int ss_ld = 0;
  // ovla_run_host, eager (non-graph) passes: each layer's pooled row block goes to the host on a copy stream as soon
  // as its pooling kernel is done, under the rest of the pass
  float* pooled_host_async = nullptr;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_pool = nullptr, ev_copied = nullptr;
  cudaStream_t own_stream = nullptr;
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;
  std::map<GraphKey, GraphEntry> graphs;
  long long graph_replays = 0;  // cudaGraphLaunch calls made by ovla_run (tests assert that a replay happened)
  // staging for ovla_run_host
  long long* in_ids = nullptr;
  int* in_lens = nullptr;
  bf16* in_px = nullptr;
  long long* out_tokens = nullptr;
  int max_P = 0;

  template <typename T>
  int alloc(T** p, long long n_elems, bool is_weight) {
    void* q = nullptr;
    const long long bytes = round_up(n_elems * static_cast<long long>(sizeof(T)), 256);
    CUDA_TRY(cudaMalloc(&q, bytes));
    allocs.push_back(q);
    (is_weight ? weight_bytes : workspace_bytes) += bytes;
    *p = static_cast<T*>(q);
    return 0;
  }
  int alloc_slot(Slot& s, long long rows, long long cols, bool required = true) {
    s.rows = rows;
    s.cols = cols;
    s.required = required;
    return alloc(&s.ptr, rows * cols, true);
  }
  long long kv_layer_elems() const {
    return 2LL * d.max_batch * d.llm_heads * d.max_seq * head_dim;
  }
  bf16* k_cache(int layer) const { return kv + layer * kv_layer_elems(); }
  bf16* v_cache(int layer) const { return k_cache(layer) + kv_layer_elems() / 2; }
};

static int check_dims(const OvlaDims& d) {
  if (d.n_towers < 1 || d.n_towers > 2) return set_error("n_towers must be 1 or 2");
  if (d.image_size % d.patch) return set_error("image_size must be a multiple of patch");
  for (int t = 0; t < d.n_towers; ++t) {
    const OvlaTower& w = d.towers[t];
    if (w.dim % w.heads) return set_error("tower %d: dim %% heads != 0", t);
    const int hd = w.dim / w.heads;
    if (hd != 64 && hd != 72 && hd != 128) return set_error("tower %d: head_dim %d unsupported (64/72/128)", t, hd);
    if (w.dim % 8 || w.mlp % 8) return set_error("tower %d: dim/mlp must be multiples of 8", t);
    if (w.depth < 2) return set_error("tower %d: depth must be >= 2", t);
  }
  if (d.llm_dim % d.llm_heads || d.llm_dim / d.llm_heads != 128) return set_error("LLM head_dim must be 128");
  if (d.llm_inter % 32) return set_error("llm_inter must be a multiple of 32");
  if (d.vocab % 8) return set_error("vocab must be a multiple of 8");
  if (d.max_batch < 1 || d.max_seq < 2) return set_error("max_batch / max_seq too small");
  return 0;
}

extern "C" int ovla_create(const OvlaDims* dims, int device, OvlaEngine** out) {
  if (!dims || !out) return set_error("ovla_create: null argument");
  OVLA_TRY(check_dims(*dims));
  CUDA_TRY(cudaSetDevice(device));
  OvlaEngine* e = new OvlaEngine();
  e->d = *dims;
  if (const char* tc = getenv("OVLA_ATTN_TC")) e->attn_tc = tc[0] != '0';
  if (const char* fr = getenv("OVLA_FUSE_ROPE")) e->fuse_rope = fr[0] != '0';
  if (const char* ts = getenv("OVLA_TWO_STREAMS")) e->two_streams = ts[0] != '0';
  if (const char* fn = getenv("OVLA_FUSE_NORM")) e->fuse_norm = fn[0] != '0';
  if (const char* fr = getenv("OVLA_FUSE_NORM_MIN_ROWS")) e->fuse_norm_min_rows = std::max(1, atoi(fr));
  if (const char* dm = getenv("OVLA_DECODE_MEGA")) e->decode_mega = dm[0] != '0';
  if (const char* gm = getenv("OVLA_GRAPH_MAX_BATCH")) e->graph_max_batch = std::max(0, atoi(gm));   // A/B knob
  if (const char* gr = getenv("OVLA_GRAPHS")) { if (gr[0] == '0') e->graph_max_batch = 0; }  // eager launches (for ncu)
  e->device = device;
  const OvlaDims& d = e->d;
  const int g = d.image_size / d.patch;
  e->np = g * g;
  e->kpad = static_cast<int>(round_up(3 * d.patch * d.patch, 8));
  e->head_dim = d.llm_dim / d.llm_heads;
  e->vision_dim = 0;
  int rc = 0;
  auto A = [&](Slot& s, long long r, long long c, bool req = true) { if (!rc) rc = e->alloc_slot(s, r, c, req); };
  for (int t = 0; t < d.n_towers && !rc; ++t) {
    const OvlaTower& w = d.towers[t];
    e->vision_dim += w.dim;
    e->n_run[t] = w.depth - 1;
    TowerW& tw = e->tower[t];
    A(tw.patch_w, w.dim, e->kpad);
    A(tw.patch_b, 1, w.dim);
    A(tw.pos, e->np, w.dim);
    if (w.n_prefix) {
      A(tw.cls, 1, w.dim);
      if (w.n_prefix > 1) A(tw.reg, w.n_prefix - 1, w.dim);
    }
    tw.blocks.resize(e->n_run[t]);
    for (BlockW& b : tw.blocks) {
      A(b.ln1_w, 1, w.dim); A(b.ln1_b, 1, w.dim);
      A(b.qkv_w, 3LL * w.dim, w.dim); A(b.qkv_b, 1, 3LL * w.dim);
      A(b.proj_w, w.dim, w.dim); A(b.proj_b, 1, w.dim);
      A(b.ln2_w, 1, w.dim); A(b.ln2_b, 1, w.dim);
      A(b.fc1_w, w.mlp, w.dim); A(b.fc1_b, 1, w.mlp);
      A(b.fc2_w, w.dim, w.mlp); A(b.fc2_b, 1, w.dim);
      if (w.layerscale) { A(b.ls1, 1, w.dim); A(b.ls2, 1, w.dim); }
    }
    if (!rc) rc = cudaMemset(tw.patch_w.ptr, 0, sizeof(bf16) * w.dim * e->kpad) == cudaSuccess ? 0 : set_error("memset");
  }
  const int Dv = e->vision_dim, Dl = d.llm_dim;
  if (d.n_towers == 2) {  // fused-gelu-mlp (modeling_prismatic.py:139-144)
    e->n_proj = 3;
    A(e->pj_w[0], 4LL * Dv, Dv); A(e->pj_b[0], 1, 4LL * Dv);
    A(e->pj_w[1], Dl, 4LL * Dv); A(e->pj_b[1], 1, Dl);
    A(e->pj_w[2], Dl, Dl);       A(e->pj_b[2], 1, Dl);
  } else {                // gelu-mlp (modeling_prismatic.py:135-137)
    e->n_proj = 2;
    A(e->pj_w[0], Dl, Dv); A(e->pj_b[0], 1, Dl);
    A(e->pj_w[1], Dl, Dl); A(e->pj_b[1], 1, Dl);
  }
  A(e->embed, d.vocab, Dl);
  e->layers.resize(d.llm_layers);
  // the fused-norm prefill needs whole 128-column slots per row and only pays above the split-K regime
  const bool want_fused_norm = e->fuse_norm && (Dl % 128 == 0) && (1LL * d.max_batch * d.max_seq >= e->fuse_norm_min_rows);
  e->fuse_norm = want_fused_norm;
  for (LayerW& l : e->layers) {
    A(l.ln1, 1, Dl);
    A(l.qkv_w, 3LL * Dl, Dl);
    A(l.o_w, Dl, Dl);
    A(l.ln2, 1, Dl);
    A(l.gate_up_w, 2LL * d.llm_inter, Dl);
    A(l.down_w, Dl, d.llm_inter);
    if (want_fused_norm && !rc) rc = e->alloc(&l.qkv_wn, 3LL * Dl * Dl, true);
    if (want_fused_norm && !rc) rc = e->alloc(&l.gate_up_wn, 2LL * d.llm_inter * Dl, true);
  }
  A(e->final_norm, 1, Dl);
  A(e->lm_head, d.vocab, Dl);
  A(e->rope_cos, d.max_seq, e->head_dim / 2);
  A(e->rope_sin, d.max_seq, e->head_dim / 2);

  // ---- workspace
  const long long B = d.max_batch;
  long long maxN = 0, maxD = 0, maxMlp = 0;
  for (int t = 0; t < d.n_towers; ++t) {
    maxN = std::max<long long>(maxN, e->np + d.towers[t].n_prefix);
    maxD = std::max<long long>(maxD, d.towers[t].dim);
    maxMlp = std::max<long long>(maxMlp, d.towers[t].mlp);
  }
  auto W = [&](auto** p, long long n) { if (!rc) rc = e->alloc(p, n, false); };
  for (int set = 0; set < (d.n_towers == 2 ? 2 : 1); ++set) {
    const long long Bv = set == 0 ? B : std::min<long long>(B, OvlaEngine::kTwoStreamBatch);
    OvlaEngine::VitBufs& v = e->vb[set];
    W(&v.im2col, Bv * e->np * e->kpad);
    W(&v.patch, Bv * e->np * maxD);
    W(&v.x, Bv * maxN * maxD);
    W(&v.h, Bv * maxN * maxD);
    W(&v.qkv, Bv * maxN * 3 * maxD);
    W(&v.attn, Bv * maxN * maxD);
    W(&v.mlp, Bv * maxN * maxMlp);
  }
  W(&e->p_cat, B * e->np * Dv);
  W(&e->p_1, B * e->np * std::max<long long>(4LL * Dv, Dl));
  W(&e->p_2, B * e->np * Dl);
  W(&e->p_3, B * e->np * Dl);
  const long long T = d.max_seq;
  W(&e->l_x, B * T * Dl);
  W(&e->l_h, B * T * Dl);
  W(&e->l_qkv, B * T * 3 * Dl);
  W(&e->l_attn, B * T * Dl);
  W(&e->l_act, B * T * d.llm_inter);
  if (e->fuse_norm) {
    e->ss_ld = Dl / 64;
    W(&e->l_ss, B * T * e->ss_ld);
  }
  W(&e->kv, e->kv_layer_elems() * d.llm_layers);
  W(&e->pooled, (d.llm_layers + 1LL) * B * Dl);
  W(&e->logits, B * d.vocab);
  W(&e->tokens, 64LL * B);
  W(&e->err_flag, 4);
  W(&e->mega_layers, std::max(1, d.llm_layers));
  W(&e->mega_barrier, 4);
  if (getenv("OVLA_MEGA_TRACE")) W(&e->mega_trace, 2 * 1024);
  {
    // split-K runs only for M <= 512 rows (gemm.cu): S <= 8 slices of [M, N] fp32; sized for the widest such GEMM
    // (the vocabulary / gate_up for the LLM stream, the tower MLP for the side stream), capped at 192 MB
    const long long n_main = std::max<long long>(2LL * d.llm_inter, d.vocab);
    e->ws_main.floats = std::min<long long>(48LL << 20, 8LL * 512 * n_main);
    W(&e->ws_main.ptr, e->ws_main.floats);
    if (d.n_towers == 2) {
      e->ws_side.floats = std::min<long long>(48LL << 20, 8LL * 512 * std::max<long long>(maxMlp, 3 * maxD));
      W(&e->ws_side.ptr, e->ws_side.floats);
    }
  }
  e->max_P = static_cast<int>(T - e->np);
  W(&e->in_ids, B * std::max(1, e->max_P));
  W(&e->in_lens, B);
  W(&e->in_px, B * 3LL * d.n_towers * d.image_size * d.image_size);
  W(&e->out_tokens, 64LL * B);
  if (!rc) rc = cudaMemset(e->err_flag, 0, 16) == cudaSuccess ? 0 : set_error("memset");
  // cache rows >= T are read (and multiplied by zero probabilities) by the tensor-core prefill attention: keep them finite
  if (!rc) rc = cudaMemset(e->kv, 0, sizeof(bf16) * e->kv_layer_elems() * d.llm_layers) == cudaSuccess ? 0 : set_error("memset");
  if (rc) {
    std::string msg = last_error();
    ovla_destroy(e);
    set_error("ovla_create: %s", msg.c_str());
    return -1;
  }
  *out = e;
  return 0;
}

extern "C" void ovla_destroy(OvlaEngine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  for (auto& kv : e->graphs)
    if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
  if (e->ev_pool) cudaEventDestroy(e->ev_pool);
  if (e->ev_copied) cudaEventDestroy(e->ev_copied);
  if (e->own_stream) cudaStreamDestroy(e->own_stream);
  if (e->side_stream) cudaStreamDestroy(e->side_stream);
  if (e->ev_fork) cudaEventDestroy(e->ev_fork);
  if (e->ev_join) cudaEventDestroy(e->ev_join);
  if (e->ev_in) cudaEventDestroy(e->ev_in);
  if (e->ev_out) cudaEventDestroy(e->ev_out);
  for (void* p : e->allocs) cudaFree(p);
  delete e;
}

extern "C" long long ovla_workspace_bytes(const OvlaEngine* e) { return e ? e->workspace_bytes : 0; }
extern "C" long long ovla_weight_bytes(const OvlaEngine* e) { return e ? e->weight_bytes : 0; }
extern "C" long long ovla_graph_replays(const OvlaEngine* e) { return e ? e->graph_replays : 0; }

// ----------------------------------------------------------------------------------------------- weight binding
static long long numel(const long long* shape, int ndim) {
  long long n = 1;
  for (int i = 0; i < ndim; ++i) n *= shape[i];
  return n;
}

static int copy_plain(Slot& s, const void* src, const long long* shape, int ndim, const char* name) {
  if (numel(shape, ndim) != s.rows * s.cols)
    return set_error("bind %s: expected %lld elements, got %lld", name, s.rows * s.cols, numel(shape, ndim));
  CUDA_TRY(cudaMemcpy(s.ptr, src, sizeof(bf16) * s.rows * s.cols, cudaMemcpyDeviceToDevice));
  s.bound = true;
  return 0;
}

static bool starts_with(const std::string& s, const char* p) { return s.rfind(p, 0) == 0; }

static int bind_tower(OvlaEngine* e, int t, const std::string& rest, const void* src, const long long* shape, int ndim,
                      const char* full) {
  TowerW& tw = e->tower[t];
  const OvlaTower& w = e->d.towers[t];
  if (rest == "patch_embed.proj.weight") {
    const long long k = 3LL * e->d.patch * e->d.patch;
    if (numel(shape, ndim) != w.dim * k) return set_error("bind %s: bad shape", full);
    CUDA_TRY(cudaMemcpy2D(tw.patch_w.ptr, sizeof(bf16) * e->kpad, src, sizeof(bf16) * k, sizeof(bf16) * k, w.dim,
                          cudaMemcpyDeviceToDevice));
    tw.patch_w.bound = true;
    return 0;
  }
  if (rest == "patch_embed.proj.bias") return copy_plain(tw.patch_b, src, shape, ndim, full);
  if (rest == "pos_embed") return copy_plain(tw.pos, src, shape, ndim, full);
  if (rest == "cls_token" && w.n_prefix) return copy_plain(tw.cls, src, shape, ndim, full);
  if (rest == "reg_token" && w.n_prefix > 1) return copy_plain(tw.reg, src, shape, ndim, full);
  if (starts_with(rest, "blocks.")) {
    const size_t dot = rest.find('.', 7);
    if (dot == std::string::npos) return set_error("bind %s: malformed block name", full);
    const int i = atoi(rest.substr(7, dot - 7).c_str());
    if (i < 0 || i >= w.depth) return set_error("bind %s: block index out of range", full);
    if (i >= e->n_run[t]) return 0;  // last block: executed by timm, output discarded (modeling_prismatic.py:85-87)
    BlockW& b = tw.blocks[i];
    const std::string leaf = rest.substr(dot + 1);
    struct { const char* n; Slot* s; } tab[] = {
        {"norm1.weight", &b.ln1_w}, {"norm1.bias", &b.ln1_b}, {"attn.qkv.weight", &b.qkv_w},
        {"attn.qkv.bias", &b.qkv_b}, {"attn.proj.weight", &b.proj_w}, {"attn.proj.bias", &b.proj_b},
        {"norm2.weight", &b.ln2_w}, {"norm2.bias", &b.ln2_b}, {"mlp.fc1.weight", &b.fc1_w},
        {"mlp.fc1.bias", &b.fc1_b}, {"mlp.fc2.weight", &b.fc2_w}, {"mlp.fc2.bias", &b.fc2_b},
        {"ls1.scale_factor", &b.ls1}, {"ls2.scale_factor", &b.ls2}};
    for (auto& x : tab)
      if (leaf == x.n) {
        if (!x.s->ptr) return set_error("bind %s: tower has no LayerScale", full);
        return copy_plain(*x.s, src, shape, ndim, full);
      }
  }
  // attn_pool.*, norm.*, fc_norm.* exist in checkpoints but are never executed on this path
  if (starts_with(rest, "attn_pool.") || starts_with(rest, "norm.") || starts_with(rest, "fc_norm.")) return 0;
  return set_error("bind %s: unknown vision tensor", full);
}

extern "C" int ovla_bind_weight(OvlaEngine* e, const char* cname, const void* src, const long long* shape, int ndim) {
  if (!e || !cname || !src || !shape) return set_error("ovla_bind_weight: null argument");
  CUDA_TRY(cudaSetDevice(e->device));
  const std::string name(cname);
  const OvlaDims& d = e->d;
  if (starts_with(name, "vision_backbone.featurizer."))
    return bind_tower(e, 0, name.substr(27), src, shape, ndim, cname);
  if (starts_with(name, "vision_backbone.fused_featurizer.")) {
    if (d.n_towers < 2) return set_error("bind %s: engine has a single tower", cname);
    return bind_tower(e, 1, name.substr(33), src, shape, ndim, cname);
  }
  if (starts_with(name, "projector.fc")) {
    const int i = name[12] - '1';
    if (i < 0 || i >= e->n_proj) return set_error("bind %s: projector has %d layers", cname, e->n_proj);
    if (name.substr(13) == ".weight") return copy_plain(e->pj_w[i], src, shape, ndim, cname);
    if (name.substr(13) == ".bias") return copy_plain(e->pj_b[i], src, shape, ndim, cname);
    return set_error("bind %s: unknown projector tensor", cname);
  }
  if (name == "rope.cos") return copy_plain(e->rope_cos, src, shape, ndim, cname);
  if (name == "rope.sin") return copy_plain(e->rope_sin, src, shape, ndim, cname);
  if (name == "language_model.model.embed_tokens.weight") return copy_plain(e->embed, src, shape, ndim, cname);
  if (name == "language_model.model.norm.weight") return copy_plain(e->final_norm, src, shape, ndim, cname);
  if (name == "language_model.lm_head.weight") return copy_plain(e->lm_head, src, shape, ndim, cname);
  const char* lp = "language_model.model.layers.";
  if (starts_with(name, lp)) {
    const size_t p0 = strlen(lp), dot = name.find('.', p0);
    if (dot == std::string::npos) return set_error("bind %s: malformed layer name", cname);
    const int i = atoi(name.substr(p0, dot - p0).c_str());
    if (i < 0 || i >= d.llm_layers) return set_error("bind %s: layer index out of range", cname);
    LayerW& l = e->layers[i];
    const std::string leaf = name.substr(dot + 1);
    const long long Dl = d.llm_dim, I = d.llm_inter;
    if (leaf == "input_layernorm.weight") return copy_plain(l.ln1, src, shape, ndim, cname);
    if (leaf == "post_attention_layernorm.weight") return copy_plain(l.ln2, src, shape, ndim, cname);
    if (leaf == "self_attn.o_proj.weight") return copy_plain(l.o_w, src, shape, ndim, cname);
    if (leaf == "mlp.down_proj.weight") return copy_plain(l.down_w, src, shape, ndim, cname);
    for (int j = 0; j < 3; ++j) {
      const char* nm[3] = {"self_attn.q_proj.weight", "self_attn.k_proj.weight", "self_attn.v_proj.weight"};
      if (leaf == nm[j]) {
        if (numel(shape, ndim) != Dl * Dl) return set_error("bind %s: bad shape", cname);
        CUDA_TRY(cudaMemcpy(l.qkv_w.ptr + j * Dl * Dl, src, sizeof(bf16) * Dl * Dl, cudaMemcpyDeviceToDevice));
        (j == 0 ? l.q_bound : j == 1 ? l.k_bound : l.v_bound) = true;
        l.qkv_w.bound = l.q_bound && l.k_bound && l.v_bound;
        return 0;
      }
    }
    for (int j = 0; j < 2; ++j) {
      const char* nm[2] = {"mlp.gate_proj.weight", "mlp.up_proj.weight"};
      if (leaf == nm[j]) {  // rows [32b, 32b+32) of gate -> rows [64b, 64b+32); of up -> [64b+32, 64b+64)
        if (numel(shape, ndim) != I * Dl) return set_error("bind %s: bad shape", cname);
        const size_t blk = sizeof(bf16) * 32 * Dl;
        CUDA_TRY(cudaMemcpy2D(l.gate_up_w.ptr + j * 32 * Dl, 2 * blk, src, blk, blk, I / 32, cudaMemcpyDeviceToDevice));
        (j == 0 ? l.gate_bound : l.up_bound) = true;
        l.gate_up_w.bound = l.gate_bound && l.up_bound;
        return 0;
      }
    }
    if (leaf == "self_attn.rotary_emb.inv_freq") return 0;
  }
  return set_error("bind %s: unknown tensor name", cname);
}

extern "C" int ovla_debug_decode_trace(OvlaEngine* e, unsigned long long* out_host, int n) {
  if (!e || !out_host) return set_error("ovla_debug_decode_trace: null argument");
  if (!e->mega_trace) return set_error("ovla_debug_decode_trace: create the engine with OVLA_MEGA_TRACE=1");
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpy(out_host, e->mega_trace, sizeof(unsigned long long) * std::min(n, 2048), cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int ovla_set_option(OvlaEngine* e, const char* name, int value) {
  if (!e || !name) return set_error("ovla_set_option: null argument");
  const std::string n(name);
  if (n == "decode_mega") e->decode_mega = value != 0;
  else if (n == "attn_tc") e->attn_tc = value != 0;
  else if (n == "fuse_rope") e->fuse_rope = value != 0;
  else if (n == "two_streams") e->two_streams = value != 0;
  else if (n == "fuse_norm") {
    if (value && !e->l_ss) return set_error("ovla_set_option: engine was created without the fused-norm weights (OVLA_FUSE_NORM=0 or a small engine)");
    e->fuse_norm = value != 0;
  }
  else if (n == "graph_max_batch") e->graph_max_batch = value;
  else return set_error("ovla_set_option: unknown option '%s'", name);
  // captured passes embed the old choice
  for (auto& kv : e->graphs)
    if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  e->graphs.clear();
  return 0;
}

extern "C" int ovla_finalize(OvlaEngine* e) {
  if (!e) return set_error("ovla_finalize: null engine");
  std::string missing;
  int n_missing = 0;
  auto chk = [&](const Slot& s, const std::string& nm) {
    if (s.ptr && s.required && !s.bound) {
      if (n_missing < 6) missing += (missing.empty() ? "" : ", ") + nm;
      ++n_missing;
    }
  };
  for (int t = 0; t < e->d.n_towers; ++t) {
    const std::string p = t == 0 ? "featurizer." : "fused_featurizer.";
    TowerW& tw = e->tower[t];
    chk(tw.patch_w, p + "patch_embed.proj.weight"); chk(tw.patch_b, p + "patch_embed.proj.bias");
    chk(tw.pos, p + "pos_embed"); chk(tw.cls, p + "cls_token"); chk(tw.reg, p + "reg_token");
    for (size_t i = 0; i < tw.blocks.size(); ++i) {
      BlockW& b = tw.blocks[i];
      const std::string q = p + "blocks." + std::to_string(i) + ".";
      chk(b.ln1_w, q + "norm1.weight"); chk(b.ln1_b, q + "norm1.bias"); chk(b.qkv_w, q + "attn.qkv.weight");
      chk(b.qkv_b, q + "attn.qkv.bias"); chk(b.proj_w, q + "attn.proj.weight"); chk(b.proj_b, q + "attn.proj.bias");
      chk(b.ln2_w, q + "norm2.weight"); chk(b.ln2_b, q + "norm2.bias"); chk(b.fc1_w, q + "mlp.fc1.weight");
      chk(b.fc1_b, q + "mlp.fc1.bias"); chk(b.fc2_w, q + "mlp.fc2.weight"); chk(b.fc2_b, q + "mlp.fc2.bias");
      chk(b.ls1, q + "ls1.scale_factor"); chk(b.ls2, q + "ls2.scale_factor");
    }
  }
  for (int i = 0; i < e->n_proj; ++i) {
    chk(e->pj_w[i], "projector.fc" + std::to_string(i + 1) + ".weight");
    chk(e->pj_b[i], "projector.fc" + std::to_string(i + 1) + ".bias");
  }
  chk(e->embed, "embed_tokens.weight"); chk(e->final_norm, "model.norm.weight"); chk(e->lm_head, "lm_head.weight");
  chk(e->rope_cos, "rope.cos"); chk(e->rope_sin, "rope.sin");
  for (size_t i = 0; i < e->layers.size(); ++i) {
    LayerW& l = e->layers[i];
    const std::string q = "layers." + std::to_string(i) + ".";
    chk(l.ln1, q + "input_layernorm.weight"); chk(l.qkv_w, q + "self_attn.{q,k,v}_proj.weight");
    chk(l.o_w, q + "self_attn.o_proj.weight"); chk(l.ln2, q + "post_attention_layernorm.weight");
    chk(l.gate_up_w, q + "mlp.{gate,up}_proj.weight"); chk(l.down_w, q + "mlp.down_proj.weight");
  }
  if (n_missing) return set_error("ovla_finalize: %d tensors not bound (%s%s)", n_missing, missing.c_str(),
                                  n_missing > 6 ? ", ..." : "");
  if (e->fuse_norm) {  // fold the norm weights into the copies the fused-norm prefill multiplies with
    for (LayerW& l : e->layers) {
      OVLA_TRY(fold_norm_weight_launch(l.qkv_w.ptr, l.ln1.ptr, l.qkv_wn, l.qkv_w.rows, static_cast<int>(l.qkv_w.cols), nullptr));
      OVLA_TRY(fold_norm_weight_launch(l.gate_up_w.ptr, l.ln2.ptr, l.gate_up_wn, l.gate_up_w.rows,
                                       static_cast<int>(l.gate_up_w.cols), nullptr));
    }
  }
  {  // weight-pointer table of the persistent decode-step kernel
    std::vector<DecodeLayerPtrs> tab(e->layers.size());
    for (size_t i = 0; i < e->layers.size(); ++i) {
      LayerW& l = e->layers[i];
      tab[i] = DecodeLayerPtrs{l.ln1.ptr, l.qkv_w.ptr, l.o_w.ptr, l.ln2.ptr, l.gate_up_w.ptr, l.down_w.ptr};
    }
    if (e->mega_layers && !tab.empty())
      CUDA_TRY(cudaMemcpy(e->mega_layers, tab.data(), sizeof(DecodeLayerPtrs) * tab.size(), cudaMemcpyHostToDevice));
  }
  CUDA_TRY(cudaDeviceSynchronize());
  return 0;
}

// ----------------------------------------------------------------------------------------------- forward
namespace {

// out = A . W^T with epilogue; picks the weight-streaming kernel for M <= 8
int linear(OvlaEngine* e, const bf16* A, long long lda, const Slot& W, int M, int mode, void* out, long long ldo, const bf16* bias,
           const bf16* scale, const bf16* resid, long long ldr, int gelu, int round_bf16, cudaStream_t st) {
  GemmEpi epi = {};
  epi.out = out;
  epi.ldo = ldo;
  epi.bias = bias;
  epi.scale = scale;
  epi.resid = resid;
  epi.ldr = ldr;
  epi.gelu = gelu;
  epi.round_bf16 = round_bf16;
  const int N = static_cast<int>(W.rows), K = static_cast<int>(W.cols);
  if (M <= 8) return gemv_launch(A, lda, W.ptr, K, M, N, K, mode, epi, st);
  // split-K partial tiles go to the workspace of the stream the GEMM runs on (the two towers overlap in time)
  const SplitKWs ws = (e->side_stream && st == e->side_stream) ? e->ws_side : e->ws_main;
  return gemm_launch(A, lda, W.ptr, K, M, N, K, mode, kKindBf16, epi, 0, 0, num_sms(), st, ws);
}

int run_tower(OvlaEngine* e, int t, const bf16* px, int B, const OvlaEngine::VitBufs& v, cudaStream_t st) {
  const OvlaDims& d = e->d;
  const OvlaTower& w = d.towers[t];
  TowerW& tw = e->tower[t];
  const int np = e->np, N = np + w.n_prefix, D = w.dim, hd = D / w.heads;
  const int rows = B * N;
  // patch embed: im2col -> GEMM(+bias) -> (+pos_embed, prefix tokens)
  OVLA_TRY(im2col_launch(px, B, 3 * d.n_towers, 3 * t, d.image_size, d.image_size, d.patch, e->kpad, v.im2col, st));
  OVLA_TRY(linear(e, v.im2col, e->kpad, tw.patch_w, B * np, kModeBf16, v.patch, D, tw.patch_b.ptr, nullptr, nullptr,
                  0, 0, 0, st));
  OVLA_TRY(assemble_tokens_launch(v.patch, tw.pos.ptr, tw.cls.ptr, tw.reg.ptr, B, np, w.n_prefix, D, v.x, st));
  const long long s12[12] = {3LL * D * N, 3LL * D, hd, 3LL * D * N, 3LL * D, hd, 3LL * D * N, 3LL * D, hd,
                             1LL * D * N, D, hd};
  for (int i = 0; i < e->n_run[t]; ++i) {
    BlockW& b = tw.blocks[i];
    OVLA_TRY(layernorm_launch(v.x, D, b.ln1_w.ptr, b.ln1_b.ptr, 1e-6f, v.h, D, rows, D, st));
    OVLA_TRY(linear(e, v.h, D, b.qkv_w, rows, kModeBf16, v.qkv, 3LL * D, b.qkv_b.ptr, nullptr, nullptr, 0, 0, 0, st));
    if (e->attn_tc && (hd == 64 || hd == 72))
      OVLA_TRY(attn_tc_qkv_launch(v.qkv, 3LL * D, v.attn, D, B, w.heads, N, hd, 0, st));
    else
      OVLA_TRY(flash_attn_launch(v.qkv, v.qkv + D, v.qkv + 2 * D, v.attn, s12, B, w.heads, N, N, hd, 0, st));
    // x = x + ls1(proj(attn))   (in place: each epilogue thread reads its residual before writing)
    OVLA_TRY(linear(e, v.attn, D, b.proj_w, rows, kModeBf16, v.x, D, b.proj_b.ptr, b.ls1.ptr, v.x, D, 0, 0, st));
    OVLA_TRY(layernorm_launch(v.x, D, b.ln2_w.ptr, b.ln2_b.ptr, 1e-6f, v.h, D, rows, D, st));
    OVLA_TRY(linear(e, v.h, D, b.fc1_w, rows, kModeBf16, v.mlp, w.mlp, b.fc1_b.ptr, nullptr, nullptr, 0, 1, 0, st));
    OVLA_TRY(linear(e, v.mlp, w.mlp, b.fc2_w, rows, kModeBf16, v.x, D, b.fc2_b.ptr, b.ls2.ptr, v.x, D, 0, 0, st));
  }
  // strip prefix tokens, concat on the feature dim (modeling_prismatic.py:123)
  int col0 = 0;
  for (int u = 0; u < t; ++u) col0 += d.towers[u].dim;
  OVLA_TRY(copy_rows_launch(v.x, 1LL * N * D, D, w.n_prefix, e->p_cat, 1LL * np * e->vision_dim, e->vision_dim, col0,
                            B, np, D, st));
  return 0;
}

// Llama prefill over all B*T rows: RMSNorm -> QKV GEMM -> RoPE + KV write -> causal flash attention -> o_proj(+res)
// -> RMSNorm -> gate/up GEMM with SwiGLU epilogue -> down(+res); the capture kernel pools the residual stream
// (hidden_states[i], i < L) before each layer and the post-final-norm states (hidden_states[L]) at the end.
// D2H of one layer's pooled block [B, D] behind its pooling kernel (host runs only; see pooled_host_async)
static int pooled_to_host(OvlaEngine* e, int layer, int B, cudaStream_t st) {
  if (!e->pooled_host_async) return 0;
  const long long D = e->d.llm_dim, off = 1LL * layer * B * D;
  CUDA_TRY(cudaEventRecord(e->ev_pool, st));
  CUDA_TRY(cudaStreamWaitEvent(e->copy_stream, e->ev_pool, 0));
  CUDA_TRY(cudaMemcpyAsync(e->pooled_host_async + off, e->pooled + off, sizeof(float) * B * D, cudaMemcpyDeviceToHost,
                           e->copy_stream));
  return 0;
}

int run_prefill(OvlaEngine* e, int B, int T, const OvlaRunArgs* a, cudaStream_t st) {
  const OvlaDims& d = e->d;
  const int D = d.llm_dim, H = d.llm_heads, hd = e->head_dim, rows = B * T;
  const int Tmax = d.max_seq;
  const long long s12[12] = {3LL * D * T, 3LL * D, hd,                                   // q in the fused buffer
                             1LL * H * Tmax * hd, hd, 1LL * Tmax * hd,                   // k cache [B,H,Tmax,hd]
                             1LL * H * Tmax * hd, hd, 1LL * Tmax * hd, 1LL * D * T, D, hd};
  // RMSNorm fused into the GEMMs on both sides of it (large batches): see OvlaEngine::fuse_norm
  const bool fnorm = e->fuse_norm && e->fuse_rope && e->l_ss && rows >= e->fuse_norm_min_rows && d.llm_layers > 0;
  auto norm_in = [&](GemmEpi& epi) {
    epi.ss_in = e->l_ss; epi.ss_ld = e->ss_ld; epi.ss_parts = D / 64; epi.ss_inv_k = 1.0f / static_cast<float>(D);
    epi.ss_eps = d.rms_eps;
  };
  if (fnorm) OVLA_TRY(row_sumsq_launch(e->l_x, D, rows, D, e->l_ss, e->ss_ld, st));   // layer 0's input comes from no GEMM
  for (int i = 0; i < d.llm_layers; ++i) {
    LayerW& l = e->layers[i];
    if (a->pool_len > 0)
      OVLA_TRY(pool_tokens_launch(e->l_x, 1LL * T * D, D, B, a->pool_len, D, a->pool_mode,
                                  e->pooled + 1LL * i * B * D, D, st, a->prompt_lens_dev, a->P));
    if (a->pool_len > 0) OVLA_TRY(pooled_to_host(e, i, B, st));
    if (a->hidden_out_dev)
      CUDA_TRY(cudaMemcpyAsync(static_cast<bf16*>(a->hidden_out_dev) + 1LL * i * rows * D, e->l_x,
                               sizeof(bf16) * rows * D, cudaMemcpyDeviceToDevice, st));
    if (!fnorm) OVLA_TRY(rmsnorm_launch(e->l_x, D, l.ln1.ptr, d.rms_eps, e->l_h, D, rows, D, st));
    if (e->fuse_rope) {  // QKV projection with RoPE + KV-cache write fused into the GEMM epilogue
      GemmEpi epi = {};
      epi.out = e->l_qkv;
      epi.ldo = 3LL * D;
      epi.rope_cos = e->rope_cos.ptr;
      epi.rope_sin = e->rope_sin.ptr;
      epi.k_cache = e->k_cache(i);
      epi.v_cache = e->v_cache(i);
      epi.T = T;
      epi.pos0 = 0;
      epi.Tmax = Tmax;
      epi.H = H;
      if (fnorm) norm_in(epi);
      OVLA_TRY(gemm_launch(fnorm ? e->l_x : e->l_h, D, fnorm ? l.qkv_wn : l.qkv_w.ptr, D, rows, 3 * D, D, kModeQkvRope,
                           kKindBf16, epi, 0, 0, num_sms(), st));
    } else {
      OVLA_TRY(linear(e, e->l_h, D, l.qkv_w, rows, kModeBf16, e->l_qkv, 3LL * D, nullptr, nullptr, nullptr, 0, 0, 0, st));
      OVLA_TRY(rope_kv_launch(e->l_qkv, B, T, H, hd, 0, e->rope_cos.ptr, e->rope_sin.ptr, e->k_cache(i), e->v_cache(i),
                              Tmax, st));
    }
    if (e->attn_tc && hd == 128)
      OVLA_TRY(attn_tc_prefill_launch(e->l_qkv, 3LL * D, e->k_cache(i), e->v_cache(i), e->l_attn, D, B, H, T, Tmax, st));
    else
      OVLA_TRY(flash_attn_launch(e->l_qkv, e->k_cache(i), e->v_cache(i), e->l_attn, s12, B, H, T, T, hd, 1, st));
    if (fnorm) {
      // x += o_proj(attn), the epilogue also leaving the row sums of squares of the new x; gate/up reads x itself
      GemmEpi eo = {};
      eo.out = e->l_x; eo.ldo = D; eo.resid = e->l_x; eo.ldr = D;
      eo.ss_out = e->l_ss; eo.ss_ld = e->ss_ld;
      OVLA_TRY(gemm_launch(e->l_attn, D, l.o_w.ptr, D, rows, D, D, kModeBf16, kKindBf16, eo, 0, 0, num_sms(), st));
      GemmEpi eg = {};
      eg.out = e->l_act; eg.ldo = d.llm_inter;
      norm_in(eg);
      OVLA_TRY(gemm_launch(e->l_x, D, l.gate_up_wn, D, rows, 2 * d.llm_inter, D, kModeSwiGLU, kKindBf16, eg, 0, 0, num_sms(), st));
      GemmEpi ed = eo;   // x += down(act), again with the row sums for the next layer's input norm
      if (i + 1 == d.llm_layers) ed.ss_out = nullptr;   // the final norm is a stand-alone kernel (its output is pooled)
      OVLA_TRY(gemm_launch(e->l_act, d.llm_inter, l.down_w.ptr, d.llm_inter, rows, D, d.llm_inter, kModeBf16, kKindBf16, ed, 0,
                           0, num_sms(), st));
      continue;
    }
    OVLA_TRY(linear(e, e->l_attn, D, l.o_w, rows, kModeBf16, e->l_x, D, nullptr, nullptr, e->l_x, D, 0, 0, st));
    OVLA_TRY(rmsnorm_launch(e->l_x, D, l.ln2.ptr, d.rms_eps, e->l_h, D, rows, D, st));
    OVLA_TRY(linear(e, e->l_h, D, l.gate_up_w, rows, kModeSwiGLU, e->l_act, d.llm_inter, nullptr, nullptr, nullptr, 0, 0,
                    0, st));
    OVLA_TRY(linear(e, e->l_act, d.llm_inter, l.down_w, rows, kModeBf16, e->l_x, D, nullptr, nullptr, e->l_x, D, 0, 0, st));
  }
  // final norm (hidden_states[L] is post-norm)
  OVLA_TRY(rmsnorm_launch(e->l_x, D, e->final_norm.ptr, d.rms_eps, e->l_h, D, rows, D, st));
  if (a->pool_len > 0)
    OVLA_TRY(pool_tokens_launch(e->l_h, 1LL * T * D, D, B, a->pool_len, D, a->pool_mode,
                                e->pooled + 1LL * d.llm_layers * B * D, D, st, a->prompt_lens_dev, a->P));
  if (a->pool_len > 0) OVLA_TRY(pooled_to_host(e, d.llm_layers, B, st));
  if (a->hidden_out_dev)
    CUDA_TRY(cudaMemcpyAsync(static_cast<bf16*>(a->hidden_out_dev) + 1LL * d.llm_layers * rows * D, e->l_h,
                             sizeof(bf16) * rows * D, cudaMemcpyDeviceToDevice, st));
  return 0;
}

// One cached decode step for B single-token rows at position `pos`: 7 kernels per layer (RMSNorm, QKV, fused
// RoPE + KV append + attention, o_proj(+res), RMSNorm, gate/up SwiGLU, down(+res)); for B <= 8 the linears are
// weight-streaming GEMVs chained with programmatic dependent launch.  Ends with final norm + lm_head (fp32 storage of
// bf16-rounded logits, as HF's `.float()`).
int run_decode_step(OvlaEngine* e, int B, int pos, float* logits, const int* lens, int P, cudaStream_t st) {
  const OvlaDims& d = e->d;
  const int D = d.llm_dim, H = d.llm_heads, hd = e->head_dim;
  const int Tmax = d.max_seq;
  if (e->decode_mega && B <= 4 && e->mega_layers && !prof_enabled()) {
    // batch <= 4: the whole step as one persistent kernel with a continuous weight stream (decode_mega.cu)
    DecodeStepArgs a = {};
    a.layers = e->mega_layers;
    a.final_norm = e->final_norm.ptr; a.lm_head = e->lm_head.ptr;
    a.rope_cos = e->rope_cos.ptr; a.rope_sin = e->rope_sin.ptr;
    a.x = e->l_x; a.qkv = e->l_qkv; a.attn = e->l_attn; a.act = e->l_act;
    a.kv = e->kv; a.kv_layer_elems = e->kv_layer_elems();
    a.logits = logits; a.barrier = e->mega_barrier;
    a.trace = e->mega_trace; a.trace_stride = 1024;
    { static const int dbg = getenv("OVLA_MEGA_DBG") ? atoi(getenv("OVLA_MEGA_DBG")) : 0; a.dbg = dbg; }
    a.n_layers = d.llm_layers; a.M = B; a.D = D; a.I = d.llm_inter; a.H = H; a.head_dim = hd; a.vocab = d.vocab;
    a.Tmax = Tmax; a.pos = pos; a.eps = d.rms_eps;
    a.lens = lens; a.P = P;
    const int rc = decode_step_launch(a, st);
    if (rc != -2) return rc;
  }
  for (int i = 0; i < d.llm_layers; ++i) {
    LayerW& l = e->layers[i];
    OVLA_TRY(rmsnorm_launch(e->l_x, D, l.ln1.ptr, d.rms_eps, e->l_h, D, B, D, st));
    OVLA_TRY(linear(e, e->l_h, D, l.qkv_w, B, kModeBf16, e->l_qkv, 3LL * D, nullptr, nullptr, nullptr, 0, 0, 0, st));
    OVLA_TRY(decode_rope_attn_launch(e->l_qkv, 3LL * D, e->rope_cos.ptr, e->rope_sin.ptr, pos, e->k_cache(i),
                                     e->v_cache(i), B, H, hd, Tmax, e->l_attn, D, st, lens, P));
    OVLA_TRY(linear(e, e->l_attn, D, l.o_w, B, kModeBf16, e->l_x, D, nullptr, nullptr, e->l_x, D, 0, 0, st));
    OVLA_TRY(rmsnorm_launch(e->l_x, D, l.ln2.ptr, d.rms_eps, e->l_h, D, B, D, st));
    OVLA_TRY(linear(e, e->l_h, D, l.gate_up_w, B, kModeSwiGLU, e->l_act, d.llm_inter, nullptr, nullptr, nullptr, 0, 0, 0,
                    st));
    OVLA_TRY(linear(e, e->l_act, d.llm_inter, l.down_w, B, kModeBf16, e->l_x, D, nullptr, nullptr, e->l_x, D, 0, 0, st));
  }
  OVLA_TRY(rmsnorm_launch(e->l_x, D, e->final_norm.ptr, d.rms_eps, e->l_h, D, B, D, st));
  OVLA_TRY(linear(e, e->l_h, D, e->lm_head, B, kModeF32, logits, d.vocab, nullptr, nullptr, nullptr, 0, 0, 1, st));
  return 0;
}

}  // namespace

static int run_impl(OvlaEngine* e, const OvlaRunArgs* a, cudaStream_t st) {
  CUDA_TRY(cudaSetDevice(e->device));
  const OvlaDims& d = e->d;
  const int B = a->B, P = a->P, np = e->np, D = d.llm_dim;
  if (B == 0) return 0;  // empty batch: nothing to do
  if (B < 0 || B > d.max_batch) return set_error("ovla_run: batch %d exceeds max_batch %d", B, d.max_batch);
  if (P < 1) return set_error("ovla_run: prompt must hold at least the BOS token");
  const int T = np + P;
  const int n_new = a->n_new_tokens;
  if (n_new < 0 || n_new > 64) return set_error("ovla_run: n_new_tokens out of range");
  if (T + std::max(0, n_new - 1) > d.max_seq)
    return set_error("ovla_run: sequence %d + %d new tokens exceeds max_seq %d", T, n_new, d.max_seq);
  if (a->pool_len < 0 || a->pool_len > T) return set_error("ovla_run: pool_len %d out of range (T=%d)", a->pool_len, T);
  if (!a->input_ids_dev || !a->pixel_values_dev) return set_error("ovla_run: null input");
  if (a->prompt_lens_dev && a->pool_len > 0 && a->pool_len - (P - 1) < 1)
    return set_error("ovla_run: pool_len %d leaves nothing to pool for a row of length 1 (P=%d)", a->pool_len, P);
  // The KV cache is laid out for the live batch: [B, H, max_seq, hd] per layer half (capacity max_batch).

  // ---- vision towers + projector
  const bf16* px = static_cast<const bf16*>(a->pixel_values_dev);
  if (d.n_towers == 2 && B <= OvlaEngine::kTwoStreamBatch && e->two_streams) {
    // small batches are latency-bound: the two independent towers run concurrently on two streams (fork / join with
    // events, which also captures into the CUDA graph as two parallel branches)
    if (!e->side_stream) {
      CUDA_TRY(cudaStreamCreateWithFlags(&e->side_stream, cudaStreamNonBlocking));
      CUDA_TRY(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
      CUDA_TRY(cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming));
    }
    CUDA_TRY(cudaEventRecord(e->ev_fork, st));
    CUDA_TRY(cudaStreamWaitEvent(e->side_stream, e->ev_fork, 0));
    OVLA_TRY(run_tower(e, 0, px, B, e->vb[0], st));
    OVLA_TRY(run_tower(e, 1, px, B, e->vb[1], e->side_stream));
    CUDA_TRY(cudaEventRecord(e->ev_join, e->side_stream));
    CUDA_TRY(cudaStreamWaitEvent(st, e->ev_join, 0));
  } else {
    for (int t = 0; t < d.n_towers; ++t) OVLA_TRY(run_tower(e, t, px, B, e->vb[0], st));
  }
  const int Mp = B * np, Dv = e->vision_dim;
  if (a->patches_out_dev)
    CUDA_TRY(cudaMemcpyAsync(a->patches_out_dev, e->p_cat, sizeof(bf16) * Mp * Dv, cudaMemcpyDeviceToDevice, st));
  const bf16* proj_out;
  if (e->n_proj == 3) {
    OVLA_TRY(linear(e, e->p_cat, Dv, e->pj_w[0], Mp, kModeBf16, e->p_1, 4LL * Dv, e->pj_b[0].ptr, nullptr, nullptr, 0, 1, 0, st));
    OVLA_TRY(linear(e, e->p_1, 4LL * Dv, e->pj_w[1], Mp, kModeBf16, e->p_2, D, e->pj_b[1].ptr, nullptr, nullptr, 0, 1, 0, st));
    OVLA_TRY(linear(e, e->p_2, D, e->pj_w[2], Mp, kModeBf16, e->p_3, D, e->pj_b[2].ptr, nullptr, nullptr, 0, 0, 0, st));
    proj_out = e->p_3;
  } else {
    OVLA_TRY(linear(e, e->p_cat, Dv, e->pj_w[0], Mp, kModeBf16, e->p_1, D, e->pj_b[0].ptr, nullptr, nullptr, 0, 1, 0, st));
    OVLA_TRY(linear(e, e->p_1, D, e->pj_w[1], Mp, kModeBf16, e->p_2, D, e->pj_b[1].ptr, nullptr, nullptr, 0, 0, 0, st));
    proj_out = e->p_2;
  }
  if (a->projector_out_dev)
    CUDA_TRY(cudaMemcpyAsync(a->projector_out_dev, proj_out, sizeof(bf16) * Mp * D, cudaMemcpyDeviceToDevice, st));

  // ---- splice + prefill
  OVLA_TRY(embed_splice_launch(a->input_ids_dev, B, P, e->embed.ptr, d.vocab, proj_out, np, D, e->l_x, e->err_flag, st));
  OVLA_TRY(run_prefill(e, B, T, a, st));
  if (a->pool_len > 0 && a->pooled_out_dev)
    CUDA_TRY(cudaMemcpyAsync(a->pooled_out_dev, e->pooled, sizeof(float) * (d.llm_layers + 1LL) * B * D,
                             cudaMemcpyDeviceToDevice, st));

  // ---- greedy decode: lm_head on the last prefill position only, argmax, then cached single-token steps
  for (int s = 0; s < n_new; ++s) {
    float* lg = a->step_logits_out_dev ? a->step_logits_out_dev + 1LL * s * B * d.vocab : e->logits;
    if (s == 0) {
      // HF: logits = lm_head(h) in bf16, then .float()  => fp32 storage of bf16-rounded values
      const bf16* last = e->l_h + 1LL * (T - 1) * D;
      long long ld_last = 1LL * T * D;
      if (a->prompt_lens_dev) {  // ragged: each row's own last real position, gathered into a dense [B, D] block
        OVLA_TRY(gather_last_rows_launch(e->l_h, 1LL * T * D, D, T - 1, a->prompt_lens_dev, P, B, D, e->l_attn,
                                         e->err_flag, st));
        last = e->l_attn;
        ld_last = D;
      }
      OVLA_TRY(linear(e, last, ld_last, e->lm_head, B, kModeF32, lg, d.vocab, nullptr, nullptr, nullptr, 0, 0, 1, st));
    } else {
      OVLA_TRY(embed_splice_launch(e->tokens + 1LL * (s - 1) * B, B, 1, e->embed.ptr, d.vocab, nullptr, 0, D, e->l_x,
                                   e->err_flag, st));
      OVLA_TRY(run_decode_step(e, B, T + s - 1, lg, a->prompt_lens_dev, P, st));
    }
    OVLA_TRY(argmax_launch(lg, d.vocab, B, d.vocab, e->tokens + 1LL * s * B, st));
  }
  if (n_new > 0 && a->tokens_out_dev) {  // [n_new, B] -> [B, n_new]
    for (int s = 0; s < n_new; ++s)
      CUDA_TRY(cudaMemcpy2DAsync(a->tokens_out_dev + s, sizeof(long long) * n_new, e->tokens + 1LL * s * B,
                                 sizeof(long long), sizeof(long long), B, cudaMemcpyDeviceToDevice, st));
  }
  return 0;
}

// Small batches are launch-bound (~2000 kernels per pass): the whole pass is captured once per distinct argument set
// into a CUDA graph on the engine's own stream and replayed afterwards.  The first call with a new key runs eagerly
// (it also performs the one-time cudaFuncSetAttribute calls), the second captures, later calls replay.
// passes that may be replayed from a CUDA graph (small batches; never while the per-kernel profiler records events)
static bool graph_eligible(const OvlaEngine* e, const OvlaRunArgs* a) {
  return a->B > 0 && a->B <= e->graph_max_batch && !prof_enabled();
}

extern "C" int ovla_run(OvlaEngine* e, const OvlaRunArgs* a, void* stream) {
  if (!e || !a) return set_error("ovla_run: null argument");
  cudaStream_t user = static_cast<cudaStream_t>(stream);
  if (!graph_eligible(e, a)) return run_impl(e, a, user);
  CUDA_TRY(cudaSetDevice(e->device));
  GraphKey key;
  memset(&key, 0, sizeof(key));  // field-wise copy keeps the padding bytes zero for the memcmp ordering
  key.a.input_ids_dev = a->input_ids_dev; key.a.pixel_values_dev = a->pixel_values_dev;
  key.a.B = a->B; key.a.P = a->P; key.a.pool_len = a->pool_len; key.a.pool_mode = a->pool_mode;
  key.a.n_new_tokens = a->n_new_tokens; key.a.pooled_out_dev = a->pooled_out_dev;
  key.a.tokens_out_dev = a->tokens_out_dev; key.a.step_logits_out_dev = a->step_logits_out_dev;
  key.a.hidden_out_dev = a->hidden_out_dev; key.a.projector_out_dev = a->projector_out_dev;
  key.a.patches_out_dev = a->patches_out_dev; key.a.prompt_lens_dev = a->prompt_lens_dev;
  auto it = e->graphs.find(key);
  if (it == e->graphs.end()) {
    if (e->graphs.size() >= 16) {  // bounded cache: drop everything (keys are few in steady state)
      for (auto& kv : e->graphs)
        if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
      e->graphs.clear();
    }
    e->graphs.emplace(key, GraphEntry{});
    return run_impl(e, a, user);
  }
  GraphEntry& g = it->second;
  if (!g.exec) {
    if (!e->own_stream) {
      CUDA_TRY(cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking));
      CUDA_TRY(cudaEventCreateWithFlags(&e->ev_in, cudaEventDisableTiming));
      CUDA_TRY(cudaEventCreateWithFlags(&e->ev_out, cudaEventDisableTiming));
    }
    const long long before = launch_count();
    cudaGraph_t graph = nullptr;
    CUDA_TRY(cudaStreamBeginCapture(e->own_stream, cudaStreamCaptureModeThreadLocal));
    const int rc = run_impl(e, a, e->own_stream);
    const cudaError_t ce = cudaStreamEndCapture(e->own_stream, &graph);
    if (rc != 0 || ce != cudaSuccess || !graph) {
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      e->graphs.erase(it);
      if (rc != 0) return -1;
      return run_impl(e, a, user);  // capture unsupported here: stay eager
    }
    g.launches = launch_count() - before;
    count_launch(-static_cast<int>(g.launches));  // capture recorded the kernels, it did not run them
    const cudaError_t ie = cudaGraphInstantiate(&g.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) {
      g.exec = nullptr;
      cudaGetLastError();
      e->graphs.erase(it);
      return run_impl(e, a, user);
    }
  }
  CUDA_TRY(cudaEventRecord(e->ev_in, user));
  CUDA_TRY(cudaStreamWaitEvent(e->own_stream, e->ev_in, 0));
  CUDA_TRY(cudaGraphLaunch(g.exec, e->own_stream));
  ++e->graph_replays;
  CUDA_TRY(cudaEventRecord(e->ev_out, e->own_stream));
  CUDA_TRY(cudaStreamWaitEvent(user, e->ev_out, 0));
  count_launch(static_cast<int>(g.launches));
  return 0;
}

extern "C" int ovla_run_host(OvlaEngine* e, const long long* ids_host, const void* px_host, int B, int P, int pool_len,
                             int pool_mode, int n_new, float* pooled_host, long long* tokens_host, const int* lens_host,
                             void* stream) {
  if (!e) return set_error("ovla_run_host: null engine");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUDA_TRY(cudaSetDevice(e->device));
  const OvlaDims& d = e->d;
  if (B == 0) return 0;
  if (B < 0 || B > d.max_batch) return set_error("ovla_run_host: batch %d exceeds max_batch %d", B, d.max_batch);
  if (P < 1 || P > e->max_P) return set_error("ovla_run_host: prompt length %d out of range [1,%d]", P, e->max_P);
  if (!ids_host || !px_host) return set_error("ovla_run_host: null input");
  const long long px_elems = 1LL * B * 3 * d.n_towers * d.image_size * d.image_size;
  CUDA_TRY(cudaMemcpyAsync(e->in_ids, ids_host, sizeof(long long) * B * P, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(e->in_px, px_host, sizeof(bf16) * px_elems, cudaMemcpyHostToDevice, st));
  OvlaRunArgs a = {};
  if (lens_host) {
    for (int b = 0; b < B; ++b)
      if (lens_host[b] < 1 || lens_host[b] > P)
        return set_error("ovla_run_host: prompt length %d of row %d outside [1, %d]", lens_host[b], b, P);
    CUDA_TRY(cudaMemcpyAsync(e->in_lens, lens_host, sizeof(int) * B, cudaMemcpyHostToDevice, st));
    a.prompt_lens_dev = e->in_lens;
  }
  a.input_ids_dev = e->in_ids;
  a.pixel_values_dev = e->in_px;
  a.B = B;
  a.P = P;
  a.pool_len = pooled_host ? pool_len : 0;
  a.pool_mode = pool_mode;
  a.n_new_tokens = tokens_host ? n_new : 0;
  a.tokens_out_dev = e->out_tokens;
  // eager passes (no CUDA graph: B above the graph limit or graphs disabled) stream the pooled states out layer by layer
  // (only into page-locked memory: an async copy into pageable memory blocks the host until the pooling kernel it waits
  // for has run, which would serialise kernel submission behind the device)
  bool host_pinned = false;
  if (pooled_host) {
    cudaPointerAttributes pa = {};
    if (cudaPointerGetAttributes(&pa, pooled_host) == cudaSuccess) host_pinned = pa.type == cudaMemoryTypeHost;
    else cudaGetLastError();
  }
  const bool stream_pooled = a.pool_len > 0 && host_pinned && !graph_eligible(e, &a);
  if (stream_pooled) {
    if (!e->copy_stream) {
      CUDA_TRY(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
      CUDA_TRY(cudaEventCreateWithFlags(&e->ev_pool, cudaEventDisableTiming));
      CUDA_TRY(cudaEventCreateWithFlags(&e->ev_copied, cudaEventDisableTiming));
    }
    e->pooled_host_async = pooled_host;
  }
  const int rc_run = ovla_run(e, &a, st);
  e->pooled_host_async = nullptr;
  if (rc_run) return rc_run;
  if (stream_pooled) {
    CUDA_TRY(cudaEventRecord(e->ev_copied, e->copy_stream));
    CUDA_TRY(cudaStreamWaitEvent(st, e->ev_copied, 0));
  } else if (a.pool_len > 0) {
    CUDA_TRY(cudaMemcpyAsync(pooled_host, e->pooled, sizeof(float) * (d.llm_layers + 1LL) * B * d.llm_dim,
                             cudaMemcpyDeviceToHost, st));
  }
  if (a.n_new_tokens > 0)
    CUDA_TRY(cudaMemcpyAsync(tokens_host, e->out_tokens, sizeof(long long) * B * n_new, cudaMemcpyDeviceToHost, st));
  int err = 0;
  CUDA_TRY(cudaMemcpyAsync(&err, e->err_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  if (err) {
    cudaMemset(e->err_flag, 0, sizeof(int));
    if (err == 2) return set_error("ovla_run_host: a prompt length lies outside [1, %d]", P);
    return set_error("ovla_run_host: input_ids contain a token id outside [0, %d)", d.vocab);
  }
  return 0;
}
