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

#define OVLA_ABI_VERSION 1

/* ------------------------------------------------------------------ library */
int ovla_abi_version(void);
const char* ovla_last_error(void);
/* number of CUDA kernels this library has launched since the last reset (bench.py "gpu_launches") */
long long ovla_launch_count(void);
void ovla_reset_launch_count(void);

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
} OvlaGemmEpilogue;

enum { OVLA_GEMM_BF16 = 0, OVLA_GEMM_SWIGLU = 1, OVLA_GEMM_F32OUT = 2 };
enum { OVLA_KIND_BF16 = 0, OVLA_KIND_TF32 = 1 };

/* mode: OVLA_GEMM_*; kind: OVLA_KIND_* (TF32 only with F32OUT); tile_n/cta_group: 0,0 = heuristic.
 * OVLA_GEMM_SWIGLU expects W rows interleaved [32 gate rows | 32 up rows] (see ovla_interleave_gate_up)
 * and writes out[M, N/2] = silu(gate) * up (LlamaMLP, transformers modeling_llama.py).              */
int ovla_gemm(const void* a_dev, long long lda, const void* w_dev, long long ldw, int M, int N, int K, int mode,
              int kind, void* out_dev, long long ldo, const OvlaGemmEpilogue* epi, int tile_n, int cta_group,
              void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OVLA_B200_H_ */
