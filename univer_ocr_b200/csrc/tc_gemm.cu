// tcgen05 (5th-gen tensor core) TF32 GEMM paths.  Until the kernels land the seams answer
// UOCR_ERR_UNSUPPORTED and the FP32 SGEMM of gemm.cu runs.
#include "gemm_common.cuh"

namespace uocr {

int fc_fwd_fast(int, const float*, const float*, float*, int64_t, int64_t, int64_t, int, float,
                cudaStream_t) {
    return UOCR_ERR_UNSUPPORTED;
}

int fc_bwd_fast(int, const float*, const float*, const float*, float*, float*, int64_t, int64_t,
                int64_t, int, cudaStream_t) {
    return UOCR_ERR_UNSUPPORTED;
}

}  // namespace uocr
