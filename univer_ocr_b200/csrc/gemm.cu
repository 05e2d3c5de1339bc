// FullyConnected forward / backward on the FP32 FFMA path (nn/layers/layers.py:335-347):
//   y  = [x, 1] . W           dx = dy . W[:-1]^T           dW += [x, 1]^T . dy
// One register-tiled SGEMM (64x64x16 CTA tile, 4x4 per thread) serves all three through
// transposed accessors; the "ones" column of [x, 1] is synthesised, never materialised (the
// reference concatenates a ones column on every call, layers.py:336).
// The TF32 tensor-core variant lives in tc_gemm.cu; this is the check-mode reference for it.
#include "common.cuh"
#include "gemm_common.cuh"

namespace uocr {

constexpr int BM = 64, BN = 64, BK = 16, PADT = 4;

template <bool TA, bool TB>
__global__ void __launch_bounds__(256) sgemm_kernel(GemmArgs p) {
    __shared__ __align__(16) float As[BK][BM + PADT];
    __shared__ __align__(16) float Bs[BK][BN + PADT];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
    const int64_t kchunk = (p.K + gridDim.z - 1) / gridDim.z;
    const int64_t kb = (int64_t)blockIdx.z * kchunk;
    const int64_t ke = min(p.K, kb + kchunk);

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int64_t k0 = kb; k0 < ke; k0 += BK) {
#pragma unroll
        for (int r = 0; r < (BM * BK) / 256; ++r) {
            const int e = r * 256 + tid;
            int mm, kk;
            if (TA) { kk = e / BM; mm = e % BM; } else { mm = e / BK; kk = e % BK; }
            const int64_t m = m0 + mm, k = k0 + kk;
            float v = 0.f;
            if (m < p.M && k < ke) {
                if (m == p.a_ones_m) v = 1.f;
                else v = TA ? __ldg(p.A + k * p.lda + m) : __ldg(p.A + m * p.lda + k);
            }
            As[kk][mm] = v;
        }
#pragma unroll
        for (int r = 0; r < (BN * BK) / 256; ++r) {
            const int e = r * 256 + tid;
            int nn, kk;
            if (TB) { nn = e / BK; kk = e % BK; } else { kk = e / BN; nn = e % BN; }
            const int64_t n = n0 + nn, k = k0 + kk;
            float v = 0.f;
            if (n < p.N && k < ke) v = TB ? __ldg(p.B + n * p.ldb + k) : __ldg(p.B + k * p.ldb + n);
            Bs[kk][nn] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t m = m0 + ty * 4 + i;
        if (m >= p.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t n = n0 + tx * 4 + j;
            if (n >= p.N) continue;
            float* c = p.C + m * p.ldc + n;
            if (gridDim.z > 1) {
                atomicAdd(c, acc[i][j]);                       // split-K: C pre-initialised
            } else {
                float v = acc[i][j] + (p.bias ? __ldg(p.bias + n) : 0.f);
                v = apply_act(v, p.act, p.alpha);
                *c = p.accumulate ? *c + v : v;
            }
        }
    }
}

int sgemm_fp32(const GemmArgs& p, bool ta, bool tb, int splitk, cudaStream_t st) {
    dim3 grid((unsigned)ceil_div(p.N, BN), (unsigned)ceil_div(p.M, BM), (unsigned)splitk);
    if (splitk > 1 && !p.accumulate) {
        // atomics need a zeroed destination when the caller asked for overwrite semantics
        if (p.ldc == p.N) {
            cudaError_t e = cudaMemsetAsync(p.C, 0, sizeof(float) * p.M * p.N, st);
            if (e != cudaSuccess) { set_error("memset: %s", cudaGetErrorString(e)); return UOCR_ERR_CUDA; }
        } else {
            set_error("split-K overwrite needs a dense C");
            return UOCR_ERR_INVALID;
        }
    }
    if (ta && !tb) sgemm_kernel<true, false><<<grid, 256, 0, st>>>(p);
    else if (!ta && tb) sgemm_kernel<false, true><<<grid, 256, 0, st>>>(p);
    else if (!ta && !tb) sgemm_kernel<false, false><<<grid, 256, 0, st>>>(p);
    else sgemm_kernel<true, true><<<grid, 256, 0, st>>>(p);
    UOCR_LAUNCHED("sgemm");
    return UOCR_OK;
}

}  // namespace uocr

using namespace uocr;

extern "C" {

int uocr_weights_to_kmajor(const float* w, float* wt, int64_t k_rows, int64_t n_cols, void* stream) {
    UOCR_REQUIRE(w && wt, "NULL pointer");
    UOCR_REQUIRE(k_rows > 0 && n_cols > 0, "non-positive dimension");
    return weights_to_kmajor(w, wt, k_rows, n_cols, as_stream(stream));
}

static int fc_fwd_impl(const float* x, const float* w, const float* w_kmajor, float* y, int64_t batch, int64_t n_in,
                       int64_t n_out, int act, float alpha, int math_mode, void* stream);

int uocr_window_fc_fwd(const float* x, const float* w, const float* w_kmajor, float* y, int64_t n, int64_t wd,
                       int64_t c, int32_t width, int64_t n_out, int act, float alpha, int math_mode, void* stream) {
    UOCR_REQUIRE(x && w && y, "NULL pointer");
    UOCR_REQUIRE(n > 0 && wd > 0 && c > 0 && width > 0 && n_out > 0, "non-positive dimension");
    UOCR_REQUIRE(wd >= width, "Input width must be >= than output width");
    UOCR_REQUIRE(act >= UOCR_ACT_NONE && act <= UOCR_ACT_SIGMOID, "unknown activation %d", act);
    cudaStream_t st = as_stream(stream);
    int rc = fc_window_fwd_fast(math_mode, x, w, w_kmajor, y, n, wd, c, width, n_out, act, alpha, st);
    if (rc != UOCR_ERR_UNSUPPORTED) return rc;
    // any other geometry / math mode: materialise the windows in scratch, then the ordinary FullyConnected
    Scratch win(st);
    rc = win.alloc(sizeof(float) * (size_t)n * wd * width * c);
    if (rc) return rc;
    rc = uocr_window_batch_fwd(x, (float*)win.ptr, n, 1, wd, c, width, stream);
    if (rc) return rc;
    return fc_fwd_impl((const float*)win.ptr, w, w_kmajor, y, n * wd, (int64_t)width * c, n_out, act, alpha, math_mode, stream);
}

int uocr_fc_chain2_fwd(const float* x, const float* w1, const float* w1_kmajor, const float* w2, const float* w2_kmajor,
                       float* y, int64_t batch, int64_t n_in, int64_t n_hidden, int64_t n_out, int act1, float alpha1,
                       int math_mode, void* stream) {
    UOCR_REQUIRE(x && w1 && w2 && y, "NULL pointer");
    UOCR_REQUIRE(batch > 0 && n_in > 0 && n_hidden > 0 && n_out > 0, "non-positive dimension");
    UOCR_REQUIRE(act1 >= UOCR_ACT_NONE && act1 <= UOCR_ACT_SIGMOID, "unknown activation %d", act1);
    cudaStream_t st = as_stream(stream);
    int rc = fc_chain2_fwd_fast(math_mode, x, w1, w1_kmajor, w2, w2_kmajor, y, batch, n_in, n_hidden, n_out, act1, alpha1, st);
    if (rc != UOCR_ERR_UNSUPPORTED) return rc;
    Scratch hidden(st);                                   // any other geometry / FP32 mode: the two layers one after the other
    rc = hidden.alloc(sizeof(float) * (size_t)batch * n_hidden);
    if (rc) return rc;
    rc = fc_fwd_impl(x, w1, w1_kmajor, (float*)hidden.ptr, batch, n_in, n_hidden, act1, alpha1, math_mode, stream);
    if (rc) return rc;
    return fc_fwd_impl((const float*)hidden.ptr, w2, w2_kmajor, y, batch, n_hidden, n_out, UOCR_ACT_NONE, 0.f, math_mode, stream);
}

int uocr_fc_fwd(const float* x, const float* w, float* y, int64_t batch, int64_t n_in, int64_t n_out,
                int act, float alpha, int math_mode, void* stream) {
    return fc_fwd_impl(x, w, nullptr, y, batch, n_in, n_out, act, alpha, math_mode, stream);
}

int uocr_fc_fwd_kmajor(const float* x, const float* w, const float* w_kmajor, float* y, int64_t batch, int64_t n_in,
                       int64_t n_out, int act, float alpha, int math_mode, void* stream) {
    return fc_fwd_impl(x, w, w_kmajor, y, batch, n_in, n_out, act, alpha, math_mode, stream);
}

static int fc_fwd_impl(const float* x, const float* w, const float* w_kmajor, float* y, int64_t batch, int64_t n_in,
                       int64_t n_out, int act, float alpha, int math_mode, void* stream) {
    UOCR_REQUIRE(x && w && y, "NULL pointer");
    UOCR_REQUIRE(batch > 0 && n_in > 0 && n_out > 0, "non-positive dimension");
    UOCR_REQUIRE(act >= UOCR_ACT_NONE && act <= UOCR_ACT_SIGMOID, "unknown activation %d", act);
    cudaStream_t st = as_stream(stream);
    int rc = fc_fwd_fast(math_mode, x, w, w_kmajor, y, batch, n_in, n_out, act, alpha, st);
    if (rc != UOCR_ERR_UNSUPPORTED) return rc;
    GemmArgs p{};
    p.A = x; p.lda = n_in; p.B = w; p.ldb = n_out; p.C = y; p.ldc = n_out;
    p.M = batch; p.N = n_out; p.K = n_in;
    p.bias = w + n_in * n_out; p.act = act; p.alpha = alpha; p.accumulate = 0; p.a_ones_m = -1;
    return sgemm_fp32(p, false, false, 1, st);
}

int uocr_fc_bwd(const float* x, const float* w, const float* dy, float* dx, float* dw, int64_t batch,
                int64_t n_in, int64_t n_out, int accumulate, int math_mode, void* stream) {
    UOCR_REQUIRE(x && w && dy && dw, "NULL pointer");
    UOCR_REQUIRE(batch > 0 && n_in > 0 && n_out > 0, "non-positive dimension");
    cudaStream_t st = as_stream(stream);
    int rc = fc_bwd_fast(math_mode, x, w, dy, dx, dw, batch, n_in, n_out, accumulate, st);
    if (rc != UOCR_ERR_UNSUPPORTED) return rc;
    if (dx) {
        GemmArgs p{};
        p.A = dy; p.lda = n_out; p.B = w; p.ldb = n_out; p.C = dx; p.ldc = n_in;
        p.M = batch; p.N = n_in; p.K = n_out; p.a_ones_m = -1;
        rc = sgemm_fp32(p, false, true, 1, st);
        if (rc) return rc;
    }
    GemmArgs q{};
    q.A = x; q.lda = n_in; q.B = dy; q.ldb = n_out; q.C = dw; q.ldc = n_out;
    q.M = n_in + 1; q.N = n_out; q.K = batch; q.accumulate = accumulate; q.a_ones_m = n_in;
    const int64_t tiles = ceil_div(q.M, BM) * ceil_div(q.N, BN);
    int64_t splitk = (148 * 2 + tiles - 1) / tiles;
    const int64_t max_split = ceil_div(batch, 8 * BK);
    if (splitk > max_split) splitk = max_split;
    if (splitk < 1) splitk = 1;
    return sgemm_fp32(q, true, false, (int)splitk, st);
}

}  // extern "C"
