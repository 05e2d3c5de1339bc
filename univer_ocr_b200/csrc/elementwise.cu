// Bandwidth-class kernels of libuocr: array helpers, activations, Upsample2D, MaxPool2D,
// Conv2DToBatchedFixedWidthed, PredToText hit mask.  All are pure streaming kernels: 128-bit
// vectorised, grid sized to a few waves over the 148 SMs, no shared memory (no reuse).
#include "common.cuh"

namespace uocr {

// ---------------------------------------------------------------- generic elementwise driver
// out[i] = f(a[i], b[i]).  Vector path: 4 x float4 in flight per thread (loads first, then
// math, then stores) so each CTA keeps 16 KB of loads outstanding.
template <int NIN, class F>
__global__ void __launch_bounds__(kThreads) ew_vec4_kernel(float4* out, const float4* a,
                                                           const float4* b, int64_t n4, F f) {
    // no __restrict__: `out` may alias an input (in-place `+=`), each element is read then
    // written by the same thread only
    constexpr int U = 4;
    const int64_t tile = (int64_t)kThreads * U;
    for (int64_t base = (int64_t)blockIdx.x * tile; base < n4; base += (int64_t)gridDim.x * tile) {
        float4 va[U], vb[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = base + (int64_t)u * kThreads + threadIdx.x;
            if (i < n4) {
                va[u] = a[i];
                if (NIN > 1) vb[u] = b[i];
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = base + (int64_t)u * kThreads + threadIdx.x;
            if (i < n4) {
                float4 r;
                r.x = f(va[u].x, NIN > 1 ? vb[u].x : 0.f);
                r.y = f(va[u].y, NIN > 1 ? vb[u].y : 0.f);
                r.z = f(va[u].z, NIN > 1 ? vb[u].z : 0.f);
                r.w = f(va[u].w, NIN > 1 ? vb[u].w : 0.f);
                out[i] = r;
            }
        }
    }
}

template <int NIN, class F>
__global__ void __launch_bounds__(kThreads) ew_scalar_kernel(float* out, const float* a,
                                                             const float* b, int64_t n, F f) {
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * kThreads)
        out[i] = f(a[i], NIN > 1 ? b[i] : 0.f);
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <int NIN, class F>
static int launch_ew(const char* name, float* out, const float* a, const float* b, int64_t n, F f,
                     cudaStream_t st) {
    if (n <= 0) return UOCR_OK;
    const bool vec = aligned16(out) && aligned16(a) && (NIN < 2 || aligned16(b));
    const int64_t n4 = vec ? n / 4 : 0;
    if (n4 > 0) {
        ew_vec4_kernel<NIN, F><<<ew_grid(n4, 4), kThreads, 0, st>>>(
            reinterpret_cast<float4*>(out), reinterpret_cast<const float4*>(a),
            reinterpret_cast<const float4*>(b), n4, f);
        UOCR_LAUNCHED(name);
    }
    const int64_t done = n4 * 4;
    if (done < n) {
        ew_scalar_kernel<NIN, F><<<ew_grid(n - done, 1), kThreads, 0, st>>>(
            out + done, a + done, NIN > 1 ? b + done : nullptr, n - done, f);
        UOCR_LAUNCHED(name);
    }
    return UOCR_OK;
}

struct LeakyFwd {
    float alpha;
    __device__ float operator()(float x, float) const { return x >= 0.f ? x : alpha * x; }
};
struct LeakyBwd {   // a = x (saved input), b = dy
    float alpha;
    __device__ float operator()(float x, float dy) const { return x >= 0.f ? dy : (x < 0.f ? alpha * dy : 0.f * dy); }
};
struct SigmoidFwd {
    __device__ float operator()(float x, float) const { return 1.f / (1.f + expf(-x)); }
};
struct SigmoidBwd {   // grad * e / (e + 1)^2, e = exp(-x)   (layers.py:413-415)
    __device__ float operator()(float x, float dy) const {
        const float e = expf(-x);
        const float d = e + 1.f;
        return dy * e / (d * d);
    }
};
struct ActBwdFromOut {   // a = y (activation output), b = dy
    int act;
    float alpha;
    __device__ float operator()(float y, float dy) const {
        if (act == UOCR_ACT_LEAKY) return y >= 0.f ? dy : alpha * dy;
        if (act == UOCR_ACT_SIGMOID) return dy * y * (1.f - y);
        return dy;
    }
};
struct Axpby {   // a = x, b = y(old)
    float ca, cb;
    __device__ float operator()(float x, float y) const { return ca * x + cb * y; }
};
struct Axpy1 {   // y = y + a*x with cb == 1 (exact `+=`)
    float ca;
    __device__ float operator()(float x, float y) const { return fmaf(ca, x, y); }
};
struct ScaleOnly {
    float ca;
    __device__ float operator()(float x, float) const { return ca * x; }
};
struct Mul {
    __device__ float operator()(float x, float y) const { return x * y; }
};

// ---------------------------------------------------------------- fill / convert
__global__ void __launch_bounds__(kThreads) fill_kernel(float* __restrict__ dst, float v, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * kThreads)
        dst[i] = v;
}

template <typename TO, typename TI>
__global__ void __launch_bounds__(kThreads) convert_kernel(TO* __restrict__ dst,
                                                           const TI* __restrict__ src, float scale,
                                                           int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * kThreads)
        dst[i] = static_cast<TO>(src[i]) * static_cast<TO>(scale);
}

// uint8 image planes -> float32 pixel / divisor (the reference's `encode_layers`: PNG planes / 255 in float64,
// train_data_generator.py:24-37).  dst must equal float32(float64(u) / divisor): a multiplication by 1/255 is off by
// one ulp for 126 of the 256 values, an IEEE division is exact for all of them (tests/test_gpu_round2.py), so the
// 256 quotients are formed once per CTA with __fdiv_rn and looked up: 16 pixels per thread, one 128-bit load and four
// 128-bit stores.
__global__ void __launch_bounds__(kThreads) u8_div_kernel(float* __restrict__ dst, const uint8_t* __restrict__ src,
                                                          float divisor, int64_t n) {
    __shared__ float lut[256];
    lut[threadIdx.x] = __fdiv_rn((float)threadIdx.x, divisor);            // kThreads == 256
    __syncthreads();
    const int64_t n16 = n / 16;
    const uint4* src16 = reinterpret_cast<const uint4*>(src);
    float4* dst4 = reinterpret_cast<float4*>(dst);
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n16; i += (int64_t)gridDim.x * kThreads) {
        const uint4 v = __ldg(src16 + i);
        const uint32_t words[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t wd = words[k];
            dst4[i * 4 + k] = make_float4(lut[wd & 255u], lut[(wd >> 8) & 255u], lut[(wd >> 16) & 255u], lut[wd >> 24]);
        }
    }
    for (int64_t i = n16 * 16 + (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads)
        dst[i] = lut[src[i]];
}

__global__ void __launch_bounds__(kThreads) nan_flag_kernel(const float* __restrict__ x, int64_t n,
                                                            int32_t* flag) {
    bool bad = false;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * kThreads)
        bad |= isnan(x[i]);
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

// one CTA, float64 accumulation: deterministic; used for parameter-sized arrays only
__global__ void __launch_bounds__(1024) sum_kernel(const float* __restrict__ x, int64_t n,
                                                   float* out, int accumulate) {
    __shared__ double part[32];
    double acc = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += 1024) acc += (double)x[i];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double r = part[threadIdx.x];
        r = warp_sum(r);
        if (threadIdx.x == 0) out[0] = (accumulate ? out[0] : 0.f) + (float)r;
    }
}

__global__ void __launch_bounds__(kThreads) copy2d_kernel(float* __restrict__ dst, int64_t dpitch,
                                                          const float* __restrict__ src,
                                                          int64_t spitch, int64_t rows, int64_t cols) {
    const int64_t total = rows * cols;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * kThreads) {
        const int64_t r = i / cols, c = i - r * cols;
        dst[r * dpitch + c] = src[r * spitch + c];
    }
}

__global__ void __launch_bounds__(kThreads) pad_hw_kernel(float* __restrict__ dst,
                                                          const float* __restrict__ src, int64_t n,
                                                          int64_t h, int64_t w, int64_t c, int64_t top,
                                                          int64_t left, int64_t ho, int64_t wo,
                                                          float value) {
    const int64_t total = n * ho * wo * c;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * kThreads) {
        int64_t t = i;
        const int64_t ch = t % c; t /= c;
        const int64_t x = t % wo; t /= wo;
        const int64_t y = t % ho; t /= ho;
        const int64_t sy = y - top, sx = x - left;
        float v = value;
        if (sy >= 0 && sy < h && sx >= 0 && sx < w) v = src[((t * h + sy) * w + sx) * c + ch];
        dst[i] = v;
    }
}

// ---------------------------------------------------------------- Upsample2D
// forward: one thread per OUTPUT element (coalesced writes; each input element is re-read
// sx times by neighbouring threads and sy times by another row -> L1/L2 hits)
__global__ void __launch_bounds__(kThreads) upsample_fwd_kernel(const float* __restrict__ x,
                                                                float* __restrict__ y, int64_t n,
                                                                int64_t h, int64_t w, int64_t c, int sy,
                                                                int sx) {
    const int64_t wo = w * sx, ho = h * sy;
    const int64_t total = n * ho * wo * c;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * kThreads) {
        int64_t t = i;
        const int64_t ch = t % c; t /= c;
        const int64_t ox = t % wo; t /= wo;
        const int64_t oy = t % ho; t /= ho;
        y[i] = x[((t * h + oy / sy) * w + ox / sx) * c + ch];
    }
}

// backward: one thread per INPUT-sized element, sums its sy x sx block of dy
__global__ void __launch_bounds__(kThreads) upsample_bwd_kernel(const float* __restrict__ dy,
                                                                float* __restrict__ dx, int64_t n,
                                                                int64_t h, int64_t w, int64_t c, int sy,
                                                                int sx) {
    const int64_t wo = w * sx, ho = h * sy;
    const int64_t total = n * h * w * c;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * kThreads) {
        int64_t t = i;
        const int64_t ch = t % c; t /= c;
        const int64_t ix = t % w; t /= w;
        const int64_t iy = t % h; t /= h;
        float acc = 0.f;
        for (int a = 0; a < sy; ++a)
            for (int b = 0; b < sx; ++b)
                acc += dy[((t * ho + iy * sy + a) * wo + ix * sx + b) * c + ch];
        dx[i] = acc;
    }
}

// ---------------------------------------------------------------- MaxPool2D (CPU semantics)
__global__ void __launch_bounds__(kThreads) maxpool_fwd_kernel(
    const float* __restrict__ x, float* __restrict__ y, uint8_t* __restrict__ mask, int64_t n,
    int64_t h, int64_t w, int64_t c, int kh, int kw, int ph, int pw, int sh, int sw, int64_t ho,
    int64_t wo) {
    const int64_t hp = h + 2 * ph, wp = w + 2 * pw;
    const int64_t total = n * ho * wo * c;
    const int64_t mh = kh * ho, mw = kw * wo;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * kThreads) {
        int64_t t = i;
        const int64_t ch = t % c; t /= c;
        const int64_t ox = t % wo; t /= wo;
        const int64_t oy = t % ho; t /= ho;
        const int64_t b = t;
        float m = -INFINITY;
        bool any = false;
        for (int ky = 0; ky < kh; ++ky) {
            const int64_t py = oy * sh + ky;
            if (py >= hp) break;                         // overhang beyond the padded array
            for (int kx = 0; kx < kw; ++kx) {
                const int64_t px = ox * sw + kx;
                if (px >= wp) break;
                const int64_t iy = py - ph, ix = px - pw;
                const float v = (iy >= 0 && iy < h && ix >= 0 && ix < w)
                                    ? x[((b * h + iy) * w + ix) * c + ch] : 0.f;   // zero padding tap
                m = any ? fmaxf(m, v) : v;
                any = true;
            }
        }
        y[i] = any ? m : 0.f;
        for (int ky = 0; ky < kh; ++ky) {
            const int64_t py = oy * sh + ky;
            for (int kx = 0; kx < kw; ++kx) {
                const int64_t px = ox * sw + kx;
                uint8_t hit = 0;
                if (py < hp && px < wp) {
                    const int64_t iy = py - ph, ix = px - pw;
                    const float v = (iy >= 0 && iy < h && ix >= 0 && ix < w)
                                        ? x[((b * h + iy) * w + ix) * c + ch] : 0.f;
                    hit = (v == m) ? 1 : 0;
                }
                mask[((b * mh + oy * kh + ky) * mw + ox * kw + kx) * c + ch] = hit;
            }
        }
    }
}

__global__ void __launch_bounds__(kThreads) maxpool_bwd_kernel(
    const float* __restrict__ dy, const uint8_t* __restrict__ mask, float* __restrict__ dx, int64_t n,
    int64_t h, int64_t w, int64_t c, int kh, int kw, int ph, int pw, int sh, int sw, int64_t ho,
    int64_t wo) {
    const int64_t total = n * h * w * c;
    const int64_t mh = kh * ho, mw = kw * wo;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * kThreads) {
        int64_t t = i;
        const int64_t ch = t % c; t /= c;
        const int64_t ix = t % w; t /= w;
        const int64_t iy = t % h; t /= h;
        const int64_t b = t;
        const int64_t py = iy + ph, px = ix + pw;          // coordinates in the padded array
        float acc = 0.f;
        for (int ky = 0; ky < kh; ++ky) {
            const int64_t ty = py - ky;
            if (ty < 0 || ty % sh != 0) continue;
            const int64_t oy = ty / sh;
            if (oy >= ho) continue;
            for (int kx = 0; kx < kw; ++kx) {
                const int64_t tx = px - kx;
                if (tx < 0 || tx % sw != 0) continue;
                const int64_t ox = tx / sw;
                if (ox >= wo) continue;
                const uint8_t* win = mask + ((b * mh + oy * kh) * mw + ox * kw) * c + ch;
                if (!win[((int64_t)ky * mw + kx) * c]) continue;
                int cnt = 0;
                for (int a = 0; a < kh; ++a)
                    for (int d = 0; d < kw; ++d) cnt += win[((int64_t)a * mw + d) * c];
                acc += dy[((b * ho + oy) * wo + ox) * c + ch] / (float)cnt;
            }
        }
        dx[i] = acc;
    }
}

// ---------------------------------------------------------------- Conv2DToBatchedFixedWidthed
template <typename V>
__global__ void __launch_bounds__(kThreads) window_fwd_kernel(const V* __restrict__ x,
                                                              V* __restrict__ y, int64_t n, int64_t h,
                                                              int64_t w, int64_t cv, int width) {
    const int hw = width / 2;
    const int64_t total = n * w * h * width * cv;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * kThreads) {
        int64_t t = i;
        const int64_t ch = t % cv; t /= cv;
        const int64_t k = t % width; t /= width;
        const int64_t row = t % h; t /= h;
        const int64_t wi = t % w; t /= w;
        const int64_t col = wi + k - hw;
        V v{};
        if (col >= 0 && col < w) v = x[((t * h + row) * w + col) * cv + ch];
        y[i] = v;
    }
}

__device__ __forceinline__ void vadd(float& a, const float& b) { a += b; }
__device__ __forceinline__ void vadd(float4& a, const float4& b) {
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
}

template <typename V>
__global__ void __launch_bounds__(kThreads) window_bwd_kernel(const V* __restrict__ dy,
                                                              V* __restrict__ dx, int64_t n, int64_t h,
                                                              int64_t w, int64_t cv, int width) {
    const int hw = width / 2;
    const int64_t total = n * h * w * cv;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * kThreads) {
        int64_t t = i;
        const int64_t ch = t % cv; t /= cv;
        const int64_t col = t % w; t /= w;
        const int64_t row = t % h; t /= h;
        V acc{};
        for (int k = 0; k < width; ++k) {
            const int64_t wi = col - k + hw;
            if (wi >= 0 && wi < w) vadd(acc, dy[(((t * w + wi) * h + row) * width + k) * cv + ch]);
        }
        dx[i] = acc;
    }
}

// ---------------------------------------------------------------- PredToText hit mask
// one warp per row
__global__ void __launch_bounds__(kThreads) row_max_hits_kernel(const float* __restrict__ pred,
                                                                uint8_t* __restrict__ hits,
                                                                int64_t rows, int64_t cols) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (kThreads / 32);
    for (int64_t r = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); r < rows; r += warps) {
        const float* p = pred + r * cols;
        float m = -INFINITY;
        for (int64_t j = lane; j < cols; j += 32) m = fmaxf(m, p[j]);
        m = warp_max(m);
        for (int64_t j = lane; j < cols; j += 32)
            hits[r * cols + j] = (p[j] == m && m != 0.f) ? 1 : 0;
    }
}

// ---------------------------------------------------------------- mean/max threshold masks
// mask[n, p, c] = x[n, p, c] > 0.5 * (mean_p x[n, :, c] + max_p x[n, :, c]): the binarisation the
// crop stages apply to every predicted map (interpreter/interpreter.py:437-438, 549).  Per image
// kThrBlocks CTAs reduce (sum in float64, max) per channel into a fixed-order partial table, one
// small kernel folds the partials (deterministic), the third compares.
constexpr int kThrBlocks = 64;
constexpr int kThrMaxC = 8;

__global__ void __launch_bounds__(kThreads) threshold_partials_kernel(const float* __restrict__ x,
                                                                      double* __restrict__ part,
                                                                      int64_t hw, int c) {
    __shared__ double s_sum[kThreads / 32][kThrMaxC];
    __shared__ float s_max[kThreads / 32][kThrMaxC];
    const int64_t n = blockIdx.y;
    const float* xi = x + n * hw * c;
    double sum[kThrMaxC];
    float mx[kThrMaxC];
#pragma unroll
    for (int k = 0; k < kThrMaxC; ++k) { sum[k] = 0.0; mx[k] = -INFINITY; }
    // a thread always sees the same channel phase: stride is a multiple of c
    const int64_t total = hw * c;
    const int64_t stride = (int64_t)gridDim.x * kThreads * c;
    for (int64_t base = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * c; base < total; base += stride) {
#pragma unroll
        for (int k = 0; k < kThrMaxC; ++k)
            if (k < c) {
                const float v = xi[base + k];
                sum[k] += (double)v;
                mx[k] = fmaxf(mx[k], v);
            }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < kThrMaxC; ++k)
        if (k < c) {
            double s = sum[k];
            float m = mx[k];
            for (int o = 16; o; o >>= 1) {
                s += __shfl_xor_sync(0xffffffffu, s, o);
                m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            }
            if (lane == 0) { s_sum[warp][k] = s; s_max[warp][k] = m; }
        }
    __syncthreads();
    if (threadIdx.x < c) {
        double s = 0.0;
        float m = -INFINITY;
        for (int w = 0; w < kThreads / 32; ++w) { s += s_sum[w][threadIdx.x]; m = fmaxf(m, s_max[w][threadIdx.x]); }
        double* out = part + ((n * gridDim.x + blockIdx.x) * c + threadIdx.x) * 2;
        out[0] = s;
        out[1] = (double)m;
    }
}

__global__ void threshold_finalize_kernel(const double* __restrict__ part, double* __restrict__ thr,
                                          int64_t groups, int blocks, int c, int64_t hw, int mean_only) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // g = n * c + channel
    if (g >= groups) return;
    const int64_t n = g / c, ch = g % c;
    double s = 0.0, m = -INFINITY;
    for (int b = 0; b < blocks; ++b) {
        const double* p = part + ((n * blocks + b) * c + ch) * 2;
        s += p[0];
        m = fmax(m, p[1]);
    }
    thr[g] = mean_only ? s / (double)hw : 0.5 * (s / (double)hw + m);
}

__global__ void __launch_bounds__(kThreads) threshold_apply_kernel(const float* __restrict__ x,
                                                                   const double* __restrict__ thr,
                                                                   uint8_t* __restrict__ mask,
                                                                   int64_t total, int64_t per_image, int c) {
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * kThreads) {
        const int64_t n = i / per_image;
        mask[i] = (double)x[i] > thr[n * c + (i % c)] ? 1 : 0;
    }
}

// ---------------------------------------------------------------- fast paths: tiled pooling / x2 upsampling
// The general kernels above take any geometry with 64-bit index chains and one scalar per thread (maxpool backward:
// 3 % of HBM peak).  The shapes that matter -- non-overlapping windows (kernel == stride, no padding, floor mode) and
// Upsample2D(2) -- are pure streaming: one thread per window / source cell and V channels (128-, 64- or 32-bit
// accesses), 32-bit indices, every byte touched once.
template <int V> struct VecT;
template <> struct VecT<1> { typedef float F; typedef uint8_t B; };
template <> struct VecT<2> { typedef float2 F; typedef uchar2 B; };
template <> struct VecT<4> { typedef float4 F; typedef uchar4 B; };

template <int V, int K>
__global__ void __launch_bounds__(kThreads) maxpool_tile_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                                    uint8_t* __restrict__ mask, int n, int h, int w,
                                                                    int c, int ho, int wo) {
    typedef typename VecT<V>::F F;
    typedef typename VecT<V>::B B;
    const int cg = c / V;
    const int64_t total = (int64_t)n * ho * wo * cg;
    const int64_t mw = (int64_t)K * wo;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
        const int g = (int)(i % cg);
        int64_t t = i / cg;
        const int ox = (int)(t % wo); t /= wo;
        const int oy = (int)(t % ho);
        const int64_t b = t / ho;
        float v[K][K][V], m[V];
#pragma unroll
        for (int ky = 0; ky < K; ++ky)
#pragma unroll
            for (int kx = 0; kx < K; ++kx) {
                const F f = *reinterpret_cast<const F*>(x + ((b * h + oy * K + ky) * w + ox * K + kx) * c + g * V);
                const float* fp = reinterpret_cast<const float*>(&f);
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    v[ky][kx][j] = fp[j];
                    m[j] = (ky == 0 && kx == 0) ? fp[j] : fmaxf(m[j], fp[j]);       // same tap order as the general kernel
                }
            }
        F out;
        float* op = reinterpret_cast<float*>(&out);
#pragma unroll
        for (int j = 0; j < V; ++j) op[j] = m[j];
        *reinterpret_cast<F*>(y + i * V) = out;
#pragma unroll
        for (int ky = 0; ky < K; ++ky)
#pragma unroll
            for (int kx = 0; kx < K; ++kx) {
                B hit;
                uint8_t* hp = reinterpret_cast<uint8_t*>(&hit);
#pragma unroll
                for (int j = 0; j < V; ++j) hp[j] = v[ky][kx][j] == m[j] ? 1 : 0;
                *reinterpret_cast<B*>(mask + ((b * (K * ho) + oy * K + ky) * mw + ox * K + kx) * c + g * V) = hit;
            }
    }
}

// dx of a tiled pooling: every input position belongs to exactly one window.  Rows / columns beyond ho*K, wo*K (floor
// mode leftovers) are zeroed by the caller.
template <int V, int K>
__global__ void __launch_bounds__(kThreads) maxpool_tile_bwd_kernel(const float* __restrict__ dy,
                                                                    const uint8_t* __restrict__ mask,
                                                                    float* __restrict__ dx, int n, int h, int w, int c,
                                                                    int ho, int wo) {
    typedef typename VecT<V>::F F;
    typedef typename VecT<V>::B B;
    const int cg = c / V;
    const int64_t total = (int64_t)n * ho * wo * cg;
    const int64_t mw = (int64_t)K * wo;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
        const int g = (int)(i % cg);
        int64_t t = i / cg;
        const int ox = (int)(t % wo); t /= wo;
        const int oy = (int)(t % ho);
        const int64_t b = t / ho;
        uint8_t hit[K][K][V];
        int cnt[V];
#pragma unroll
        for (int j = 0; j < V; ++j) cnt[j] = 0;
#pragma unroll
        for (int ky = 0; ky < K; ++ky)
#pragma unroll
            for (int kx = 0; kx < K; ++kx) {
                const B mb = *reinterpret_cast<const B*>(mask + ((b * (K * ho) + oy * K + ky) * mw + ox * K + kx) * c + g * V);
                const uint8_t* mp = reinterpret_cast<const uint8_t*>(&mb);
#pragma unroll
                for (int j = 0; j < V; ++j) { hit[ky][kx][j] = mp[j]; cnt[j] += mp[j]; }
            }
        const F gv = *reinterpret_cast<const F*>(dy + i * V);
        const float* gp = reinterpret_cast<const float*>(&gv);
        float share[V];
#pragma unroll
        for (int j = 0; j < V; ++j) share[j] = cnt[j] ? gp[j] / (float)cnt[j] : 0.f;
#pragma unroll
        for (int ky = 0; ky < K; ++ky)
#pragma unroll
            for (int kx = 0; kx < K; ++kx) {
                F out;
                float* op = reinterpret_cast<float*>(&out);
#pragma unroll
                for (int j = 0; j < V; ++j) op[j] = hit[ky][kx][j] ? share[j] : 0.f;
                *reinterpret_cast<F*>(dx + ((b * h + oy * K + ky) * w + ox * K + kx) * c + g * V) = out;
            }
    }
}

// Upsample2D(2), C % 4 == 0: thread = one source pixel x 4 channels -> the same 128-bit word to 2 x 2 output pixels
__global__ void __launch_bounds__(kThreads) upsample2_c4_fwd_kernel(const float4* __restrict__ x, float4* __restrict__ y,
                                                                    int64_t total, int h, int w, int cg) {
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
        const int g = (int)(i % cg);
        int64_t t = i / cg;
        const int ix = (int)(t % w); t /= w;
        const int iy = (int)(t % h);
        const int64_t b = t / h;
        const float4 v = x[i];
        float4* o = y + ((b * 2 * h + 2 * iy) * (2 * w) + 2 * ix) * cg + g;
        o[0] = v; o[cg] = v;
        o += (int64_t)2 * w * cg;
        o[0] = v; o[cg] = v;
    }
}
// Upsample2D(2), C == 1, W even: thread = two source pixels -> (a, a, b, b) as one 128-bit store in each of two rows
__global__ void __launch_bounds__(kThreads) upsample2_c1_fwd_kernel(const float2* __restrict__ x, float4* __restrict__ y,
                                                                    int64_t total, int h, int w2) {
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
        const int jx = (int)(i % w2);
        int64_t t = i / w2;
        const int iy = (int)(t % h);
        const int64_t b = t / h;
        const float2 v = x[i];
        const float4 o = make_float4(v.x, v.x, v.y, v.y);
        float4* dst = y + ((b * 2 * h + 2 * iy) * w2 + jx);
        dst[0] = o;
        dst[w2] = o;
    }
}
// backward of the above: sums of 2 x 2 blocks
__global__ void __launch_bounds__(kThreads) upsample2_c4_bwd_kernel(const float4* __restrict__ dy, float4* __restrict__ dx,
                                                                    int64_t total, int h, int w, int cg) {
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
        const int g = (int)(i % cg);
        int64_t t = i / cg;
        const int ix = (int)(t % w); t /= w;
        const int iy = (int)(t % h);
        const int64_t b = t / h;
        const float4* s = dy + ((b * 2 * h + 2 * iy) * (2 * w) + 2 * ix) * cg + g;
        const float4 a = s[0], c2 = s[cg];
        s += (int64_t)2 * w * cg;
        const float4 d = s[0], e = s[cg];
        // the general kernel's order: (0,0) + (0,1) + (1,0) + (1,1)
        dx[i] = make_float4(((a.x + c2.x) + d.x) + e.x, ((a.y + c2.y) + d.y) + e.y, ((a.z + c2.z) + d.z) + e.z,
                            ((a.w + c2.w) + d.w) + e.w);
    }
}
__global__ void __launch_bounds__(kThreads) upsample2_c1_bwd_kernel(const float4* __restrict__ dy, float2* __restrict__ dx,
                                                                    int64_t total, int h, int w2) {
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
        const int jx = (int)(i % w2);
        int64_t t = i / w2;
        const int iy = (int)(t % h);
        const int64_t b = t / h;
        const float4* s = dy + ((b * 2 * h + 2 * iy) * w2 + jx);
        const float4 r0 = s[0], r1 = s[w2];
        dx[i] = make_float2(((r0.x + r0.y) + r1.x) + r1.y, ((r0.z + r0.w) + r1.z) + r1.w);
    }
}

template <int K>
static bool maxpool_tile_launch(bool fwd, const float* a, float* out, uint8_t* mask_w, const uint8_t* mask_r, int64_t n,
                                int64_t h, int64_t w, int64_t c, int64_t ho, int64_t wo, cudaStream_t st) {
    const uintptr_t al = reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(out);
    const uintptr_t alm = reinterpret_cast<uintptr_t>(fwd ? mask_w : mask_r);
    const int v = (c % 4 == 0 && al % 16 == 0 && alm % 4 == 0) ? 4 : (c % 2 == 0 && al % 8 == 0 && alm % 2 == 0) ? 2 : 1;
    const int64_t total = n * ho * wo * (c / v);
    const int grid = ew_grid(total, 1);
#define UOCR_POOL_CASE(V)                                                                                            \
    if (fwd) maxpool_tile_fwd_kernel<V, K><<<grid, kThreads, 0, st>>>(a, out, mask_w, (int)n, (int)h, (int)w, (int)c,  \
                                                                       (int)ho, (int)wo);                              \
    else maxpool_tile_bwd_kernel<V, K><<<grid, kThreads, 0, st>>>(a, mask_r, out, (int)n, (int)h, (int)w, (int)c,      \
                                                                   (int)ho, (int)wo)
    if (v == 4) { UOCR_POOL_CASE(4); } else if (v == 2) { UOCR_POOL_CASE(2); } else { UOCR_POOL_CASE(1); }
#undef UOCR_POOL_CASE
    return true;
}

// non-overlapping windows entirely inside the image: kernel == stride in {2, 3}, no padding, no ceil-mode overhang
static bool maxpool_is_tiled(int64_t h, int64_t w, int kh, int kw, int ph, int pw, int sh, int sw, int64_t ho, int64_t wo) {
    return kh == kw && (kh == 2 || kh == 3) && sh == kh && sw == kw && ph == 0 && pw == 0 && ho * kh <= h && wo * kw <= w &&
           h < (1 << 30) && w < (1 << 30);
}

}  // namespace uocr

using namespace uocr;

extern "C" {

int uocr_fill_f32(float* dst, float value, int64_t n, void* stream) {
    if (n <= 0) return UOCR_OK;
    UOCR_REQUIRE(dst, "dst is NULL");
    fill_kernel<<<ew_grid(n, 4), kThreads, 0, as_stream(stream)>>>(dst, value, n);
    UOCR_LAUNCHED("fill");
    return UOCR_OK;
}

int uocr_f64_to_f32(float* dst, const double* src, int64_t n, void* stream) {
    if (n <= 0) return UOCR_OK;
    UOCR_REQUIRE(dst && src, "NULL pointer");
    convert_kernel<float, double><<<ew_grid(n, 4), kThreads, 0, as_stream(stream)>>>(dst, src, 1.f, n);
    UOCR_LAUNCHED("f64_to_f32");
    return UOCR_OK;
}

int uocr_f32_to_f64(double* dst, const float* src, int64_t n, void* stream) {
    if (n <= 0) return UOCR_OK;
    UOCR_REQUIRE(dst && src, "NULL pointer");
    convert_kernel<double, float><<<ew_grid(n, 4), kThreads, 0, as_stream(stream)>>>(dst, src, 1.f, n);
    UOCR_LAUNCHED("f32_to_f64");
    return UOCR_OK;
}

int uocr_u8_to_f32(float* dst, const uint8_t* src, float scale, int64_t n, void* stream) {
    if (n <= 0) return UOCR_OK;
    UOCR_REQUIRE(dst && src, "NULL pointer");
    convert_kernel<float, uint8_t><<<ew_grid(n, 4), kThreads, 0, as_stream(stream)>>>(dst, src, scale, n);
    UOCR_LAUNCHED("u8_to_f32");
    return UOCR_OK;
}

int uocr_u8_div_f32(float* dst, const uint8_t* src, float divisor, int64_t n, void* stream) {
    if (n <= 0) return UOCR_OK;
    UOCR_REQUIRE(dst && src, "NULL pointer");
    UOCR_REQUIRE(divisor != 0.f, "divisor is 0");
    UOCR_REQUIRE((reinterpret_cast<uintptr_t>(dst) & 15) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0,
                 "dst and src must be 16-byte aligned");
    static_assert(kThreads == 256, "u8_div_kernel builds its 256-entry table with one thread per entry");
    u8_div_kernel<<<ew_grid(n, 16), kThreads, 0, as_stream(stream)>>>(dst, src, divisor, n);
    UOCR_LAUNCHED("u8_div_f32");
    return UOCR_OK;
}

int uocr_axpby_f32(float* y, const float* x, float a, float b, int64_t n, void* stream) {
    if (n <= 0) return UOCR_OK;
    UOCR_REQUIRE(y && x, "NULL pointer");
    if (b == 0.f) return launch_ew<1>("scale", y, x, nullptr, n, ScaleOnly{a}, as_stream(stream));
    if (b == 1.f) return launch_ew<2>("axpy", y, x, y, n, Axpy1{a}, as_stream(stream));
    return launch_ew<2>("axpby", y, x, y, n, Axpby{a, b}, as_stream(stream));
}

int uocr_mul_f32(float* out, const float* a, const float* b, int64_t n, void* stream) {
    if (n <= 0) return UOCR_OK;
    UOCR_REQUIRE(out && a && b, "NULL pointer");
    return launch_ew<2>("mul", out, a, b, n, Mul{}, as_stream(stream));
}

int uocr_nan_flag_f32(const float* x, int64_t n, int32_t* flag, void* stream) {
    if (n <= 0) return UOCR_OK;
    UOCR_REQUIRE(x && flag, "NULL pointer");
    nan_flag_kernel<<<ew_grid(n, 4), kThreads, 0, as_stream(stream)>>>(x, n, flag);
    UOCR_LAUNCHED("nan_flag");
    return UOCR_OK;
}

int uocr_sum_f32(const float* x, int64_t n, float* out, int accumulate, void* stream) {
    UOCR_REQUIRE(out && (x || n <= 0), "NULL pointer");
    sum_kernel<<<1, 1024, 0, as_stream(stream)>>>(x, n < 0 ? 0 : n, out, accumulate);
    UOCR_LAUNCHED("sum");
    return UOCR_OK;
}

int uocr_copy2d_f32(float* dst, int64_t dst_pitch, const float* src, int64_t src_pitch, int64_t rows,
                    int64_t cols, void* stream) {
    if (rows <= 0 || cols <= 0) return UOCR_OK;
    UOCR_REQUIRE(dst && src, "NULL pointer");
    UOCR_REQUIRE(dst_pitch >= cols && src_pitch >= cols, "pitch smaller than cols");
    copy2d_kernel<<<ew_grid(rows * cols, 4), kThreads, 0, as_stream(stream)>>>(dst, dst_pitch, src,
                                                                               src_pitch, rows, cols);
    UOCR_LAUNCHED("copy2d");
    return UOCR_OK;
}

int uocr_pad_hw_f32(float* dst, const float* src, int64_t n, int64_t h, int64_t w, int64_t c,
                    int64_t top, int64_t bottom, int64_t left, int64_t right, float value,
                    void* stream) {
    UOCR_REQUIRE(dst && src, "NULL pointer");
    UOCR_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0, "non-positive dimension");
    UOCR_REQUIRE(top >= 0 && bottom >= 0 && left >= 0 && right >= 0, "negative padding");
    const int64_t ho = h + top + bottom, wo = w + left + right;
    pad_hw_kernel<<<ew_grid(n * ho * wo * c, 4), kThreads, 0, as_stream(stream)>>>(
        dst, src, n, h, w, c, top, left, ho, wo, value);
    UOCR_LAUNCHED("pad_hw");
    return UOCR_OK;
}

int uocr_leaky_relu_fwd(const float* x, float* y, int64_t n, float alpha, void* stream) {
    if (n <= 0) return UOCR_OK;
    UOCR_REQUIRE(x && y, "NULL pointer");
    return launch_ew<1>("leaky_relu_fwd", y, x, nullptr, n, LeakyFwd{alpha}, as_stream(stream));
}

int uocr_leaky_relu_bwd(const float* x, const float* dy, float* dx, int64_t n, float alpha,
                        void* stream) {
    if (n <= 0) return UOCR_OK;
    UOCR_REQUIRE(x && dy && dx, "NULL pointer");
    return launch_ew<2>("leaky_relu_bwd", dx, x, dy, n, LeakyBwd{alpha}, as_stream(stream));
}

int uocr_sigmoid_fwd(const float* x, float* y, int64_t n, void* stream) {
    if (n <= 0) return UOCR_OK;
    UOCR_REQUIRE(x && y, "NULL pointer");
    return launch_ew<1>("sigmoid_fwd", y, x, nullptr, n, SigmoidFwd{}, as_stream(stream));
}

int uocr_sigmoid_bwd(const float* x, const float* dy, float* dx, int64_t n, void* stream) {
    if (n <= 0) return UOCR_OK;
    UOCR_REQUIRE(x && dy && dx, "NULL pointer");
    return launch_ew<2>("sigmoid_bwd", dx, x, dy, n, SigmoidBwd{}, as_stream(stream));
}

int uocr_act_bwd_from_output(const float* y, const float* dy, float* dx, int64_t n, int act,
                             float alpha, void* stream) {
    if (n <= 0) return UOCR_OK;
    UOCR_REQUIRE(y && dy && dx, "NULL pointer");
    UOCR_REQUIRE(act != UOCR_ACT_LEAKY || alpha > 0.f,
                 "leaky backward from the output needs alpha > 0 (sign(y) == sign(x))");
    return launch_ew<2>("act_bwd_from_output", dx, y, dy, n, ActBwdFromOut{act, alpha},
                        as_stream(stream));
}

int uocr_upsample2d_fwd(const float* x, float* y, int64_t n, int64_t h, int64_t w, int64_t c,
                        int32_t sy, int32_t sx, void* stream) {
    UOCR_REQUIRE(x && y, "NULL pointer");
    UOCR_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0 && sy > 0 && sx > 0, "non-positive dimension");
    const bool al = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
    if (sy == 2 && sx == 2 && al && c % 4 == 0 && h < (1 << 30) && w < (1 << 30)) {
        const int64_t total = n * h * w * (c / 4);
        upsample2_c4_fwd_kernel<<<ew_grid(total, 1), kThreads, 0, as_stream(stream)>>>(
            reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(y), total, (int)h, (int)w, (int)(c / 4));
    } else if (sy == 2 && sx == 2 && al && c == 1 && w % 2 == 0 && h < (1 << 30) && w < (1 << 30)) {
        const int64_t total = n * h * (w / 2);
        upsample2_c1_fwd_kernel<<<ew_grid(total, 1), kThreads, 0, as_stream(stream)>>>(
            reinterpret_cast<const float2*>(x), reinterpret_cast<float4*>(y), total, (int)h, (int)(w / 2));
    } else {
        upsample_fwd_kernel<<<ew_grid(n * h * w * c * sy * sx, 4), kThreads, 0, as_stream(stream)>>>(
            x, y, n, h, w, c, sy, sx);
    }
    UOCR_LAUNCHED("upsample2d_fwd");
    return UOCR_OK;
}

int uocr_upsample2d_bwd(const float* dy, float* dx, int64_t n, int64_t h, int64_t w, int64_t c,
                        int32_t sy, int32_t sx, void* stream) {
    UOCR_REQUIRE(dy && dx, "NULL pointer");
    UOCR_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0 && sy > 0 && sx > 0, "non-positive dimension");
    const bool al = ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0;
    if (sy == 2 && sx == 2 && al && c % 4 == 0 && h < (1 << 30) && w < (1 << 30)) {
        const int64_t total = n * h * w * (c / 4);
        upsample2_c4_bwd_kernel<<<ew_grid(total, 1), kThreads, 0, as_stream(stream)>>>(
            reinterpret_cast<const float4*>(dy), reinterpret_cast<float4*>(dx), total, (int)h, (int)w, (int)(c / 4));
    } else if (sy == 2 && sx == 2 && al && c == 1 && w % 2 == 0 && h < (1 << 30) && w < (1 << 30)) {
        const int64_t total = n * h * (w / 2);
        upsample2_c1_bwd_kernel<<<ew_grid(total, 1), kThreads, 0, as_stream(stream)>>>(
            reinterpret_cast<const float4*>(dy), reinterpret_cast<float2*>(dx), total, (int)h, (int)(w / 2));
    } else {
        upsample_bwd_kernel<<<ew_grid(n * h * w * c, 2), kThreads, 0, as_stream(stream)>>>(dy, dx, n, h, w, c, sy, sx);
    }
    UOCR_LAUNCHED("upsample2d_bwd");
    return UOCR_OK;
}

int uocr_maxpool2d_out_hw(int64_t h, int64_t w, int32_t kh, int32_t kw, int32_t ph, int32_t pw,
                          int32_t sh, int32_t sw, int32_t ceil_mode, int64_t* ho, int64_t* wo) {
    UOCR_REQUIRE(ho && wo, "NULL pointer");
    UOCR_REQUIRE(kh > 0 && kw > 0 && sh > 0 && sw > 0 && ph >= 0 && pw >= 0, "bad pooling geometry");
    const int64_t nh = h + 2 * ph - kh, nw = w + 2 * pw - kw;
    UOCR_REQUIRE(nh >= 0 && nw >= 0, "kernel larger than padded input");
    *ho = (ceil_mode ? (nh + sh - 1) / sh : nh / sh) + 1;
    *wo = (ceil_mode ? (nw + sw - 1) / sw : nw / sw) + 1;
    return UOCR_OK;
}

int uocr_maxpool2d_fwd(const float* x, float* y, uint8_t* mask, int64_t n, int64_t h, int64_t w,
                       int64_t c, int32_t kh, int32_t kw, int32_t ph, int32_t pw, int32_t sh,
                       int32_t sw, int64_t ho, int64_t wo, void* stream) {
    UOCR_REQUIRE(x && y && mask, "NULL pointer");
    UOCR_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0 && ho > 0 && wo > 0, "non-positive dimension");
    UOCR_REQUIRE(kh > 0 && kw > 0 && sh > 0 && sw > 0 && ph >= 0 && pw >= 0, "bad pooling geometry");
    if (maxpool_is_tiled(h, w, kh, kw, ph, pw, sh, sw, ho, wo)) {
        if (kh == 2) maxpool_tile_launch<2>(true, x, y, mask, nullptr, n, h, w, c, ho, wo, as_stream(stream));
        else maxpool_tile_launch<3>(true, x, y, mask, nullptr, n, h, w, c, ho, wo, as_stream(stream));
    } else {
        maxpool_fwd_kernel<<<ew_grid(n * ho * wo * c, 1), kThreads, 0, as_stream(stream)>>>(
            x, y, mask, n, h, w, c, kh, kw, ph, pw, sh, sw, ho, wo);
    }
    UOCR_LAUNCHED("maxpool2d_fwd");
    return UOCR_OK;
}

int uocr_maxpool2d_bwd(const float* dy, const uint8_t* mask, float* dx, int64_t n, int64_t h,
                       int64_t w, int64_t c, int32_t kh, int32_t kw, int32_t ph, int32_t pw,
                       int32_t sh, int32_t sw, int64_t ho, int64_t wo, void* stream) {
    UOCR_REQUIRE(dy && mask && dx, "NULL pointer");
    UOCR_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0 && ho > 0 && wo > 0, "non-positive dimension");
    UOCR_REQUIRE(kh > 0 && kw > 0 && sh > 0 && sw > 0 && ph >= 0 && pw >= 0, "bad pooling geometry");
    if (maxpool_is_tiled(h, w, kh, kw, ph, pw, sh, sw, ho, wo)) {
        if (ho * kh != h || wo * kw != w)                   // floor-mode leftovers belong to no window
            UOCR_CUDA(cudaMemsetAsync(dx, 0, sizeof(float) * n * h * w * c, as_stream(stream)));
        if (kh == 2) maxpool_tile_launch<2>(false, dy, dx, nullptr, mask, n, h, w, c, ho, wo, as_stream(stream));
        else maxpool_tile_launch<3>(false, dy, dx, nullptr, mask, n, h, w, c, ho, wo, as_stream(stream));
    } else {
        maxpool_bwd_kernel<<<ew_grid(n * h * w * c, 1), kThreads, 0, as_stream(stream)>>>(
            dy, mask, dx, n, h, w, c, kh, kw, ph, pw, sh, sw, ho, wo);
    }
    UOCR_LAUNCHED("maxpool2d_bwd");
    return UOCR_OK;
}

int uocr_window_batch_fwd(const float* x, float* y, int64_t n, int64_t h, int64_t w, int64_t c,
                          int32_t width, void* stream) {
    UOCR_REQUIRE(x && y, "NULL pointer");
    UOCR_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0 && width > 0, "non-positive dimension");
    UOCR_REQUIRE(w >= width, "Input width must be >= than output width");
    const int64_t total = n * w * h * width * c;
    if (c % 4 == 0 && aligned16(x) && aligned16(y)) {
        window_fwd_kernel<float4><<<ew_grid(total / 4, 2), kThreads, 0, as_stream(stream)>>>(
            reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(y), n, h, w, c / 4, width);
    } else {
        window_fwd_kernel<float><<<ew_grid(total, 4), kThreads, 0, as_stream(stream)>>>(x, y, n, h, w, c,
                                                                                     width);
    }
    UOCR_LAUNCHED("window_batch_fwd");
    return UOCR_OK;
}

int uocr_window_batch_bwd(const float* dy, float* dx, int64_t n, int64_t h, int64_t w, int64_t c,
                          int32_t width, void* stream) {
    UOCR_REQUIRE(dy && dx, "NULL pointer");
    UOCR_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0 && width > 0, "non-positive dimension");
    UOCR_REQUIRE(w >= width, "Input width must be >= than output width");
    const int64_t total = n * h * w * c;
    if (c % 4 == 0 && aligned16(dy) && aligned16(dx)) {
        window_bwd_kernel<float4><<<ew_grid(total / 4, 1), kThreads, 0, as_stream(stream)>>>(
            reinterpret_cast<const float4*>(dy), reinterpret_cast<float4*>(dx), n, h, w, c / 4, width);
    } else {
        window_bwd_kernel<float><<<ew_grid(total, 2), kThreads, 0, as_stream(stream)>>>(dy, dx, n, h, w,
                                                                                     c, width);
    }
    UOCR_LAUNCHED("window_batch_bwd");
    return UOCR_OK;
}

int uocr_row_max_hits(const float* pred, uint8_t* hits, int64_t rows, int64_t cols, void* stream) {
    if (rows <= 0 || cols <= 0) return UOCR_OK;
    UOCR_REQUIRE(pred && hits, "NULL pointer");
    int64_t blocks = ceil_div(rows, kThreads / 32);
    if (blocks > 148 * 8) blocks = 148 * 8;
    row_max_hits_kernel<<<(int)blocks, kThreads, 0, as_stream(stream)>>>(pred, hits, rows, cols);
    UOCR_LAUNCHED("row_max_hits");
    return UOCR_OK;
}

int uocr_threshold_mask_workspace(int64_t n, int64_t c, size_t* bytes) {
    UOCR_REQUIRE(bytes, "bytes is NULL");
    *bytes = (n <= 0 || c <= 0) ? 0 : (size_t)(n * kThrBlocks * c * 2 + n * c) * sizeof(double);
    return UOCR_OK;
}

static int threshold_mask_impl(const float* x, uint8_t* mask, int64_t n, int64_t hw, int64_t c, void* workspace,
                               int mean_only, void* stream) {
    if (n <= 0 || hw <= 0 || c <= 0) return UOCR_OK;
    UOCR_REQUIRE(x && mask && workspace, "NULL pointer");
    UOCR_REQUIRE(c <= kThrMaxC, "threshold_mask supports at most 8 channels");
    UOCR_REQUIRE(n <= 65535, "threshold_mask: at most 65535 images per call");
    double* part = static_cast<double*>(workspace);
    double* thr = part + n * kThrBlocks * c * 2;
    threshold_partials_kernel<<<dim3(kThrBlocks, (unsigned)n), kThreads, 0, as_stream(stream)>>>(x, part, hw,
                                                                                               (int)c);
    UOCR_LAUNCHED("threshold_partials");
    const int64_t groups = n * c;
    threshold_finalize_kernel<<<(int)ceil_div(groups, 128), 128, 0, as_stream(stream)>>>(part, thr, groups,
                                                                                       kThrBlocks, (int)c, hw, mean_only);
    UOCR_LAUNCHED("threshold_finalize");
    const int64_t total = n * hw * c;
    threshold_apply_kernel<<<ew_grid(total, 4), kThreads, 0, as_stream(stream)>>>(x, thr, mask, total, hw * c,
                                                                                (int)c);
    UOCR_LAUNCHED("threshold_apply");
    return UOCR_OK;
}

int uocr_threshold_mask(const float* x, uint8_t* mask, int64_t n, int64_t hw, int64_t c, void* workspace,
                        void* stream) {
    return threshold_mask_impl(x, mask, n, hw, c, workspace, 0, stream);
}

int uocr_above_mean_mask(const float* x, uint8_t* mask, int64_t n, int64_t hw, int64_t c, void* workspace,
                         void* stream) {
    return threshold_mask_impl(x, mask, n, hw, c, workspace, 1, stream);
}

}  // extern "C"
