// denoiser_tc.cu — tcgen05/TMEM bf16 engine of the denoiser (placeholder until the engine lands:
// creation reports "unsupported" so callers fall back to asking for PSTL_PRECISION_FP32 explicitly).
#include "mlp_common.cuh"

int pstl_tc_create(pstl_denoiser* d) {
  (void)d;
  pstl_set_error("PSTL_PRECISION_BF16: tcgen05 engine not built in this revision");
  return PSTL_ERR_UNSUPPORTED;
}
void pstl_tc_destroy(pstl_denoiser* d) { (void)d; }
int pstl_tc_sample(pstl_denoiser*, const float*, int, const float*, float*, int, const float*, int, const float*,
                   unsigned long long, unsigned long long, float, float, int, int, float*, int, int, cudaStream_t) {
  pstl_set_error("tcgen05 engine not built");
  return PSTL_ERR_UNSUPPORTED;
}
