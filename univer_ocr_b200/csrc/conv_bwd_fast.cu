// Shape-specialised BACKWARD kernels for my_model's small-channel convolutions
// (nn/layers/convolutional.py:101-145 / :197-288 are the reference semantics).
//
//   dgrad, stride 1 : dX = conv_fwd(dy, flip(w)^T, padding k-1-p) -- the tuned forward stencils
//                     are reused on a flipped / transposed copy of the (tiny) weight tensor
//   dgrad, strided  : gather form with the tap parity resolved at compile time
//                     (thread = PX consecutive input pixels x all input channels)
//   wgrad           : thread = strip of PX output columns x R output rows, ALL taps of one
//                     (ci-chunk, co-chunk) accumulated in registers while streaming down the rows;
//                     one block-level reduction per CTA at the end (warp shuffles + smem), partials
//                     to workspace[chunk][element][cta], summed by a finalize kernel -- deterministic,
//                     no atomics; db is accumulated alongside (padding_value border included in dW,
//                     as the reference's saved padded input does).
#include "conv_common.cuh"

namespace uocr {

template <int CIV> struct BVec;
template <> struct BVec<1> {
    float v[1];
    __device__ __forceinline__ void load(const float* p) { v[0] = __ldg(p); }
    __device__ __forceinline__ void fill(float f) { v[0] = f; }
};
template <> struct BVec<2> {
    float v[2];
    __device__ __forceinline__ void load(const float* p) {
        const float2 t = __ldg(reinterpret_cast<const float2*>(p));
        v[0] = t.x; v[1] = t.y;
    }
    __device__ __forceinline__ void fill(float f) { v[0] = v[1] = f; }
};
template <> struct BVec<4> {
    float v[4];
    __device__ __forceinline__ void load(const float* p) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    __device__ __forceinline__ void fill(float f) { v[0] = v[1] = v[2] = v[3] = f; }
};

// ------------------------------------------------------------------ weight flip + transpose
// wt[ky'][kx'][co][ci] = w[kh-1-ky'][kw-1-kx'][ci][co]
__global__ void __launch_bounds__(256) flip_transpose_kernel(const float* __restrict__ w, float* __restrict__ wt,
                                                             int kh, int kw, int cin, int cout) {
    const int total = kh * kw * cin * cout;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < total; i += gridDim.x * 256) {
        const int ci = i % cin;
        const int co = (i / cin) % cout;
        const int kx = (i / (cin * cout)) % kw;
        const int ky = i / (cin * cout * kw);
        wt[i] = w[(((kh - 1 - ky) * kw + (kw - 1 - kx)) * cin + ci) * cout + co];
    }
}

// ------------------------------------------------------------------ strided dgrad (COUT <= 4)
template <int KH, int KW, int SH, int SW, int PW, int CIN, int COUT, int PX>
__global__ void __launch_bounds__(256) conv_small_dgrad_kernel(ConvGeom g, const float* __restrict__ dy,
                                                               const float* __restrict__ w,
                                                               float* __restrict__ dx) {
    static_assert(PX % SW == 0, "strip start must keep the tap parity compile-time");
    __shared__ float s_w[KH * KW * CIN * COUT];              // [ky][kx][ci][co]
    for (int i = threadIdx.x; i < KH * KW * CIN * COUT; i += 256) s_w[i] = w[i];
    __syncthreads();
    // dy columns touched by this strip: ox = ix0 / SW + off, off in [OFF_LO, OFF_HI]
    constexpr int OFF_LO = (PW - (KW - 1)) >= 0 ? (PW - (KW - 1)) / SW : -(((KW - 1) - PW + SW - 1) / SW);
    constexpr int OFF_HI = (PX - 1 + PW) / SW;
    constexpr int NDY = OFF_HI - OFF_LO + 1;

    const int strips = (g.w + PX - 1) / PX;
    const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int xs = (int)(idx % strips);
    const int iy = (int)((idx / strips) % g.h);
    const int64_t n = idx / ((int64_t)strips * g.h);
    if (n >= g.n) return;
    const int ix0 = xs * PX;
    const int oxb = ix0 / SW + OFF_LO;

    float acc[PX][CIN];
#pragma unroll
    for (int p = 0; p < PX; ++p)
#pragma unroll
        for (int c = 0; c < CIN; ++c) acc[p][c] = 0.f;

#pragma unroll
    for (int ky = 0; ky < KH; ++ky) {
        const int ty = iy + g.ph - ky;
        if (ty < 0 || ty % SH != 0) continue;
        const int oy = ty / SH;
        if (oy >= g.ho) continue;
        const float* drow = dy + ((n * g.ho + oy) * (int64_t)g.wo) * COUT;
        BVec<COUT> seg[NDY];
#pragma unroll
        for (int j = 0; j < NDY; ++j) {
            const int ox = oxb + j;
            if (ox >= 0 && ox < g.wo) seg[j].load(drow + (int64_t)ox * COUT);
            else seg[j].fill(0.f);
        }
#pragma unroll
        for (int kx = 0; kx < KW; ++kx) {
#pragma unroll
            for (int p = 0; p < PX; ++p) {
                const int q = p + PW - kx;                       // compile-time after unrolling
                if (q % SW != 0) continue;
                const int j = (q >= 0 ? q / SW : -((-q) / SW)) - OFF_LO;
#pragma unroll
                for (int c = 0; c < CIN; ++c)
#pragma unroll
                    for (int o = 0; o < COUT; ++o)
                        acc[p][c] = fmaf(seg[j].v[o], s_w[((ky * KW + kx) * CIN + c) * COUT + o], acc[p][c]);
            }
        }
    }
    float* dp = dx + ((n * g.h + iy) * (int64_t)g.w + ix0) * CIN;
#pragma unroll
    for (int p = 0; p < PX; ++p) {
        if (ix0 + p >= g.w) break;
#pragma unroll
        for (int c = 0; c < CIN; ++c) dp[p * CIN + c] = acc[p][c];
    }
}

template <int KH, int KW, int SH, int SW, int PW, int CIN, int COUT, int PX>
static int launch_small_dgrad(const ConvGeom& g, const float* dy, const float* w, float* dx, cudaStream_t st) {
    const int strips = (g.w + PX - 1) / PX;
    const int64_t blocks = ceil_div((int64_t)g.n * g.h * strips, 256);
    if (blocks > 0x7fffffff) return UOCR_ERR_UNSUPPORTED;
    conv_small_dgrad_kernel<KH, KW, SH, SW, PW, CIN, COUT, PX><<<(unsigned)blocks, 256, 0, st>>>(g, dy, w, dx);
    UOCR_LAUNCHED("conv_small_dgrad");
    return UOCR_OK;
}

int conv_dgrad_fast(const ConvGeom& g, int math_mode, const float* dy, const float* w, float* dx,
                    cudaStream_t st) {
    if (math_mode == UOCR_MATH_TF32) {
        const int rc = conv_dgrad_tc(g, dy, w, dx, st);
        if (rc != UOCR_ERR_UNSUPPORTED) return rc;
    }
    if (g.sh == 1 && g.sw == 1 && g.kh - 1 - g.ph >= 0 && g.kw - 1 - g.pw >= 0 &&
        (int64_t)g.kh * g.kw * g.cin * g.cout <= 4096) {
        // stride 1: forward stencil on the flipped / transposed weights
        Scratch wt(st);
        int rc = wt.alloc(sizeof(float) * g.kh * g.kw * g.cin * g.cout);
        if (rc) return rc;
        flip_transpose_kernel<<<4, 256, 0, st>>>(w, (float*)wt.ptr, g.kh, g.kw, g.cin, g.cout);
        UOCR_LAUNCHED("flip_transpose");
        ConvGeom t = g;
        t.h = g.ho; t.w = g.wo; t.cin = g.cout; t.cout = g.cin;
        t.ph = g.kh - 1 - g.ph; t.pw = g.kw - 1 - g.pw;
        t.ho = g.h; t.wo = g.w; t.padding_value = 0.f; t.bias = 0; t.ups = 1;
        rc = conv_fwd_fast(t, 1, UOCR_MATH_FP32, dy, (const float*)wt.ptr, nullptr, dx, UOCR_ACT_NONE, 0.f, st);
        if (rc == UOCR_ERR_UNSUPPORTED)
            rc = conv_fwd_general(t, dy, (const float*)wt.ptr, nullptr, dx, UOCR_ACT_NONE, 0.f, st);
        return rc;
    }
#define UOCR_DG(KH_, KW_, SH_, SW_, PW_, CIN_, COUT_, PX_)                                             \
    if (g.kh == KH_ && g.kw == KW_ && g.sh == SH_ && g.sw == SW_ && g.pw == PW_ && g.cin == CIN_ &&     \
        g.cout == COUT_)                                                                                \
        return launch_small_dgrad<KH_, KW_, SH_, SW_, PW_, CIN_, COUT_, PX_>(g, dy, w, dx, st);
    if (((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0) {
        UOCR_DG(5, 5, 2, 2, 2, 1, 1, 8)      // Paragraph down_*
        UOCR_DG(5, 5, 2, 2, 2, 1, 4, 8)      // Line down_1
        UOCR_DG(5, 5, 2, 2, 2, 4, 4, 4)      // Line down_2
    }
#undef UOCR_DG
    return UOCR_ERR_UNSUPPORTED;
}

// ------------------------------------------------------------------ wgrad
template <int KH, int KW, int SH, int SW, int CIN, int CIV, int COT, int PX, int R>
__global__ void __launch_bounds__(256) conv_small_wgrad_kernel(ConvGeom g, const float* __restrict__ x,
                                                               const float* __restrict__ dy,
                                                               float* __restrict__ ws, int nblk) {
    constexpr int NACC = KH * KW * CIV * COT;
    constexpr int NOUT = NACC + COT;                         // + db partials
    constexpr int NIN = (PX - 1) * SW + KW;
    __shared__ float red[8][NOUT];
    const int cochunks = g.cout / COT;
    const int cic = blockIdx.y / cochunks, coc = blockIdx.y % cochunks;
    const int ci0 = cic * CIV, co0 = coc * COT;

    const int strips = (g.wo + PX - 1) / PX;
    const int rgroups = (g.ho + R - 1) / R;
    const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int xs = (int)(idx % strips);
    const int rg = (int)((idx / strips) % rgroups);
    const int64_t n = idx / ((int64_t)strips * rgroups);
    const bool live = n < g.n;

    float acc[KH][KW][CIV][COT];
    float dbacc[COT];
#pragma unroll
    for (int a = 0; a < KH; ++a)
#pragma unroll
        for (int b = 0; b < KW; ++b)
#pragma unroll
            for (int c = 0; c < CIV; ++c)
#pragma unroll
                for (int d = 0; d < COT; ++d) acc[a][b][c][d] = 0.f;
#pragma unroll
    for (int d = 0; d < COT; ++d) dbacc[d] = 0.f;

    if (live) {
        const int ox0 = xs * PX;
        const int ix0 = ox0 * SW - g.pw;
        const int oy_end = min(g.ho, (rg + 1) * R);
        for (int oy = rg * R; oy < oy_end; ++oy) {
            float dyv[PX][COT];
            const float* drow = dy + ((n * g.ho + oy) * (int64_t)g.wo + ox0) * g.cout + co0;
#pragma unroll
            for (int p = 0; p < PX; ++p) {
                if (ox0 + p < g.wo) {
                    BVec<COT> t;
                    t.load(drow + (int64_t)p * g.cout);
#pragma unroll
                    for (int d = 0; d < COT; ++d) dyv[p][d] = t.v[d];
                } else {
#pragma unroll
                    for (int d = 0; d < COT; ++d) dyv[p][d] = 0.f;
                }
#pragma unroll
                for (int d = 0; d < COT; ++d) dbacc[d] += dyv[p][d];
            }
#pragma unroll
            for (int ky = 0; ky < KH; ++ky) {
                const int iy = oy * SH + ky - g.ph;
                const bool yin = iy >= 0 && iy < g.h;
                const float* xrow = x + ((n * g.h + (yin ? iy : 0)) * (int64_t)g.w) * CIN + ci0;
                BVec<CIV> xin[NIN];
#pragma unroll
                for (int j = 0; j < NIN; ++j) {
                    const int ix = ix0 + j;
                    if (yin && ix >= 0 && ix < g.w) xin[j].load(xrow + (int64_t)ix * CIN);
                    else xin[j].fill(g.padding_value);
                }
#pragma unroll
                for (int kx = 0; kx < KW; ++kx)
#pragma unroll
                    for (int p = 0; p < PX; ++p)
#pragma unroll
                        for (int c = 0; c < CIV; ++c)
#pragma unroll
                            for (int d = 0; d < COT; ++d)
                                acc[ky][kx][c][d] = fmaf(xin[p * SW + kx].v[c], dyv[p][d], acc[ky][kx][c][d]);
            }
        }
    }

    // ---- CTA reduction: shuffle within warps, smem across the 8 warps
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int a = 0; a < KH; ++a)
#pragma unroll
        for (int b = 0; b < KW; ++b)
#pragma unroll
            for (int c = 0; c < CIV; ++c)
#pragma unroll
                for (int d = 0; d < COT; ++d) {
                    const float s = warp_sum(acc[a][b][c][d]);
                    if (lane == 0) red[wid][((a * KW + b) * CIV + c) * COT + d] = s;
                }
#pragma unroll
    for (int d = 0; d < COT; ++d) {
        const float s = warp_sum(dbacc[d]);
        if (lane == 0) red[wid][NACC + d] = s;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < NOUT; e += 256) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += red[k][e];
        ws[((int64_t)blockIdx.y * NOUT + e) * nblk + blockIdx.x] = s;
    }
}

// one CTA per output element: sums its nblk partials and scatters into dw / db
__global__ void __launch_bounds__(256) conv_small_wgrad_finalize_kernel(
    const float* __restrict__ ws, int nblk, int kh, int kw, int cin, int civ, int cot, int cout, int bias,
    float* __restrict__ dw, float* __restrict__ db, int accumulate) {
    __shared__ float red[8];
    const int nacc = kh * kw * civ * cot, nout = nacc + cot;
    const int chunk = blockIdx.x / nout, e = blockIdx.x % nout;
    const int cochunks = cout / cot;
    const int cic = chunk / cochunks, coc = chunk % cochunks;
    const float* src = ws + (int64_t)blockIdx.x * nblk;
    float s = 0.f;
    for (int i = threadIdx.x; i < nblk; i += 256) s += src[i];
    s = block_sum(s, red);
    if (threadIdx.x != 0) return;
    if (e < nacc) {
        const int d = e % cot, c = (e / cot) % civ, tap = e / (cot * civ);
        float* dst = dw + ((int64_t)tap * cin + cic * civ + c) * cout + coc * cot + d;
        *dst = accumulate ? *dst + s : s;
    } else if (cic == 0) {
        float* dst = db + coc * cot + (e - nacc);
        const float v = bias ? s : 0.f;
        *dst = accumulate ? *dst + v : v;
    }
}

struct SmallWgradPlan {
    bool ok = false;
    int kh, kw, civ, cot, px, r;
};

static SmallWgradPlan plan_small_wgrad(const ConvGeom& g) {
    SmallWgradPlan p;
    auto set = [&](int civ, int cot, int px, int r) {
        p.ok = true; p.kh = g.kh; p.kw = g.kw; p.civ = civ; p.cot = cot; p.px = px; p.r = r;
    };
    const bool k33 = g.kh == 3 && g.kw == 3 && g.sh == 1 && g.sw == 1;
    const bool k55s1 = g.kh == 5 && g.kw == 5 && g.sh == 1 && g.sw == 1;
    const bool k55s2 = g.kh == 5 && g.kw == 5 && g.sh == 2 && g.sw == 2;
    const bool k53 = g.kh == 5 && g.kw == 3 && g.sh == 2 && g.sw == 1;
    if (k33 && g.cin == 1 && g.cout % 4 == 0) set(1, 4, 4, 16);
    else if (k33 && g.cin % 4 == 0 && g.cout == 1) set(4, 1, 4, 16);
    else if (k55s1 && g.cin == 1 && g.cout == 1) set(1, 1, 8, 16);
    else if (k55s2 && g.cin == 1 && g.cout == 1) set(1, 1, 4, 16);
    else if (k55s2 && g.cin == 1 && g.cout % 2 == 0) set(1, 2, 4, 8);
    else if (k55s2 && g.cin == 4) set(4, 1, 4, 8);
    else if (k55s1 && g.cin == 4) set(4, 1, 4, 8);
    else if (k53 && g.cin == 1 && g.cout % 4 == 0) set(1, 4, 4, 7);
    return p;
}

static void small_wgrad_dims(const ConvGeom& g, const SmallWgradPlan& p, int* nblk, int* chunks, int* nout) {
    const int strips = (g.wo + p.px - 1) / p.px, rgroups = (g.ho + p.r - 1) / p.r;
    *nblk = (int)ceil_div((int64_t)g.n * strips * rgroups, 256);
    *chunks = (g.cin / p.civ) * (g.cout / p.cot);
    *nout = p.kh * p.kw * p.civ * p.cot + p.cot;
}

size_t conv_wgrad_fast_workspace(const ConvGeom& g, int) {
    const SmallWgradPlan p = plan_small_wgrad(g);
    if (!p.ok) return 0;
    int nblk, chunks, nout;
    small_wgrad_dims(g, p, &nblk, &chunks, &nout);
    return sizeof(float) * (size_t)nblk * chunks * nout;
}

template <int KH, int KW, int SH, int SW, int CIN, int CIV, int COT, int PX, int R>
static int launch_small_wgrad(const ConvGeom& g, const float* x, const float* dy, float* ws, int nblk, int chunks,
                              cudaStream_t st) {
    conv_small_wgrad_kernel<KH, KW, SH, SW, CIN, CIV, COT, PX, R><<<dim3(nblk, chunks), 256, 0, st>>>(g, x, dy, ws, nblk);
    UOCR_LAUNCHED("conv_small_wgrad");
    return UOCR_OK;
}

int conv_wgrad_fast(const ConvGeom& g, int math_mode, const float* x, const float* dy, float* dw, float* db,
                    int accumulate, float* ws, cudaStream_t st) {
    if (math_mode == UOCR_MATH_TF32) {
        const int rc = conv_wgrad_tc(g, x, dy, dw, db, accumulate, st);
        if (rc != UOCR_ERR_UNSUPPORTED) return rc;
    }
    const SmallWgradPlan p = plan_small_wgrad(g);
    if (!p.ok || !ws) return UOCR_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy)) & 15) return UOCR_ERR_UNSUPPORTED;
    int nblk, chunks, nout;
    small_wgrad_dims(g, p, &nblk, &chunks, &nout);
    if (chunks > 65535) return UOCR_ERR_UNSUPPORTED;
    int rc = UOCR_ERR_UNSUPPORTED;
    const bool k33 = g.kh == 3;
    if (k33 && g.cin == 1) rc = launch_small_wgrad<3, 3, 1, 1, 1, 1, 4, 4, 16>(g, x, dy, ws, nblk, chunks, st);
    else if (k33 && g.cin == 16) rc = launch_small_wgrad<3, 3, 1, 1, 16, 4, 1, 4, 16>(g, x, dy, ws, nblk, chunks, st);
    else if (k33 && g.cin == 4) rc = launch_small_wgrad<3, 3, 1, 1, 4, 4, 1, 4, 16>(g, x, dy, ws, nblk, chunks, st);
    else if (g.kw == 5 && g.sh == 1 && g.cin == 1) rc = launch_small_wgrad<5, 5, 1, 1, 1, 1, 1, 8, 16>(g, x, dy, ws, nblk, chunks, st);
    else if (g.kw == 5 && g.sh == 2 && g.cin == 1 && p.cot == 1) rc = launch_small_wgrad<5, 5, 2, 2, 1, 1, 1, 4, 16>(g, x, dy, ws, nblk, chunks, st);
    else if (g.kw == 5 && g.sh == 2 && g.cin == 1 && p.cot == 2) rc = launch_small_wgrad<5, 5, 2, 2, 1, 1, 2, 4, 8>(g, x, dy, ws, nblk, chunks, st);
    else if (g.kw == 5 && g.sh == 2 && g.cin == 4) rc = launch_small_wgrad<5, 5, 2, 2, 4, 4, 1, 4, 8>(g, x, dy, ws, nblk, chunks, st);
    else if (g.kw == 5 && g.sh == 1 && g.cin == 4) rc = launch_small_wgrad<5, 5, 1, 1, 4, 4, 1, 4, 8>(g, x, dy, ws, nblk, chunks, st);
    else if (g.kw == 3 && g.kh == 5 && g.cin == 1) rc = launch_small_wgrad<5, 3, 2, 1, 1, 1, 4, 4, 7>(g, x, dy, ws, nblk, chunks, st);
    if (rc) return rc;
    conv_small_wgrad_finalize_kernel<<<chunks * nout, 256, 0, st>>>(ws, nblk, p.kh, p.kw, g.cin, p.civ, p.cot, g.cout,
                                                                   g.bias, dw, db, accumulate);
    UOCR_LAUNCHED("conv_small_wgrad_finalize");
    return UOCR_OK;
}

}  // namespace uocr
