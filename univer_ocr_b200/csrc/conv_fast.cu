// Shape-specialised convolution kernels for the my_model geometries.
// Until a geometry has a tuned kernel the entry points answer UOCR_ERR_UNSUPPORTED and the
// general FP32 path of conv.cu runs.
#include "conv_common.cuh"

namespace uocr {

int conv_fwd_fast(const ConvGeom&, int, const float*, const float*, const float*, float*, int, float,
                  cudaStream_t) {
    return UOCR_ERR_UNSUPPORTED;
}

int conv_dgrad_fast(const ConvGeom&, int, const float*, const float*, float*, cudaStream_t) {
    return UOCR_ERR_UNSUPPORTED;
}

size_t conv_wgrad_fast_workspace(const ConvGeom&, int) { return 0; }

int conv_wgrad_fast(const ConvGeom&, int, const float*, const float*, float*, float*, int, float*,
                    cudaStream_t) {
    return UOCR_ERR_UNSUPPORTED;
}

}  // namespace uocr
