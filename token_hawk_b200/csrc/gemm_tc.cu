// gemm_tc.cu -- batched-prompt GEMM on tcgen05 tensor cores (placeholder until the kernel lands).
#include "common.cuh"

extern "C" int thk_gemm_f16_tc(thk_ctx*, const float*, const uint16_t*, float*, int64_t, int64_t, int64_t) {
    thk_set_error("thk_gemm_f16_tc: tensor-core prefill GEMM not built yet");
    return THK_E_UNSUPPORTED;
}
