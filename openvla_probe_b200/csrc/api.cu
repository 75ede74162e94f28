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

extern "C" {

int ovla_abi_version(void) { return OVLA_ABI_VERSION; }
const char* ovla_last_error(void) { return last_error(); }
long long ovla_launch_count(void) { return launch_count(); }
void ovla_reset_launch_count(void) { reset_launch_count(); }

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
  }
  return gemm_launch(a_dev, lda, w_dev, ldw, M, N, K, mode, kind, e, tile_n, cta_group, num_sms(),
                     static_cast<cudaStream_t>(stream));
}

}  // extern "C"
