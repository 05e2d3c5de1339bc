// GEMM argument block + the seam between the FP32 SGEMM (gemm.cu) and the tcgen05 TF32 GEMM
// (tc_gemm.cu).
#pragma once
#include "common.cuh"

namespace uocr {

struct GemmArgs {
    const float* A; int64_t lda;
    const float* B; int64_t ldb;
    float* C; int64_t ldc;
    int64_t M, N, K;
    const float* bias;      // length N, added before the activation (may be NULL)
    int act; float alpha;
    int accumulate;         // C += result instead of C = result
    int64_t a_ones_m;       // row m of op(A) that reads as all ones (-1: none)
};

int sgemm_fp32(const GemmArgs& p, bool ta, bool tb, int splitk, cudaStream_t st);

// tensor-core paths: UOCR_ERR_UNSUPPORTED when math_mode / shape has no tcgen05 kernel
int fc_fwd_fast(int math_mode, const float* x, const float* w, const float* w_kmajor /* may be NULL */, float* y,
                int64_t batch, int64_t n_in, int64_t n_out, int act, float alpha, cudaStream_t st);
// FullyConnected (hidden width 128) + activation + FullyConnected as one tensor-core kernel (tc_gemm.cu: fc_chain2_kernel)
int fc_chain2_fwd_fast(int math_mode, const float* x, const float* w1, const float* w1t, const float* w2, const float* w2t,
                       float* y, int64_t batch, int64_t k1, int64_t n1, int64_t n2, int act1, float alpha1, cudaStream_t st);
int fc_window_fwd_fast(int math_mode, const float* x, const float* w, const float* w_kmajor, float* y, int64_t n,
                       int64_t wd, int64_t c, int width, int64_t n_out, int act, float alpha, cudaStream_t st);
int weights_to_kmajor(const float* w, float* wt, int64_t k_rows, int64_t n_cols, cudaStream_t st);
int fc_bwd_fast(int math_mode, const float* x, const float* w, const float* dy, float* dx, float* dw,
                int64_t batch, int64_t n_in, int64_t n_out, int accumulate, cudaStream_t st);

}  // namespace uocr
