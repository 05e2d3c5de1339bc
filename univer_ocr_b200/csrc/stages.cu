// Device-side pieces of the reference's crop stages (web_app/components/interpreter/interpreter.py:234-523):
//   uocr_crop_masked_f32 / uocr_crop_label_mask   `(image * mask)[:, region_y, region_x, :]`, `mask[:, region_y, region_x, :]`
//                                                 with mask = (labels == l) of label_layer             (:303-309, 511)
//   uocr_zoom_nearest_f32                         ndimage.zoom(image, (1, zf, zf, 1), order=0) + the zero padding to
//                                                 `minimal_width`                                       (:513-521)
//   uocr_rotate_f32 / uocr_rotate_nearest_u8      ndimage.rotate(array, angle, axes=(2, 1), order=1 | 0, reshape=True)
//                                                 (rotate_array, :188-192)
//   uocr_mask_bbox                                ndimage.find_objects(mask)[0] of a boolean array    (:230, 303, 341)
// Selection and resampling work: one thread per output element, coordinates in double and in SciPy's order of
// operations (ni_interpolation.c: NI_ZoomShift, NI_GeometricTransform), every product and sum rounded separately
// (__dmul_rn / __dadd_rn: no FMA contraction), so that nearest-neighbour ties and the linear weights come out bit for
// bit as SciPy's (oracle/np_stages.py is the restatement these kernels follow; it is pinned against SciPy itself).
#include "common.cuh"

namespace uocr {
namespace {

__global__ void __launch_bounds__(256) crop_masked_kernel(const float* __restrict__ image, const int32_t* __restrict__ labels,
                                                          int32_t label, float* __restrict__ out, int64_t total, int h,
                                                          int w, int c, int y0, int x0, int ch, int cw) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        const int k = (int)(i % c);
        int64_t r = i / c;
        const int x = (int)(r % cw); r /= cw;
        const int y = (int)(r % ch);
        const int64_t n = r / ch;
        const int64_t px = (n * h + (y0 + y)) * w + (x0 + x);
        const float v = image[px * c + k];
        // image * mask: a product (not a select), so that a negative value under a False mask gives -0.0 like NumPy's
        out[i] = labels ? v * (labels[px] == label ? 1.f : 0.f) : v;
    }
}

__global__ void __launch_bounds__(256) crop_label_mask_kernel(const int32_t* __restrict__ labels, int32_t label,
                                                              uint8_t* __restrict__ out, int64_t total, int h, int w,
                                                              int y0, int x0, int ch, int cw) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        const int x = (int)(i % cw);
        const int64_t r = i / cw;
        const int y = (int)(r % ch);
        const int64_t n = r / ch;
        out[i] = labels[(n * h + (y0 + y)) * w + (x0 + x)] == label ? 1 : 0;
    }
}

// NI_ZoomShift, order 0, mode constant: output index j reads input floor(c + 0.5), c = j * ratio; 0 where c > len - 1
__device__ __forceinline__ int zoom_index(int j, double ratio, int len) {
    const double c = __dmul_rn((double)j, ratio);
    if (c < 0.0 || c > (double)(len - 1)) return -1;
    return (int)floor(__dadd_rn(c, 0.5));
}

__global__ void __launch_bounds__(256) zoom_nearest_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                           int64_t total, int h, int w, int c, int oh, int ow, int owp,
                                                           double ry, double rx) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        const int k = (int)(i % c);
        int64_t r = i / c;
        const int x = (int)(r % owp); r /= owp;
        const int y = (int)(r % oh);
        const int64_t n = r / oh;
        float v = 0.f;
        if (x < ow) {
            const int iy = zoom_index(y, ry, h), ix = zoom_index(x, rx, w);
            if (iy >= 0 && ix >= 0) v = src[((n * h + iy) * w + ix) * c + k];
        }
        dst[i] = v;
    }
}

struct RotateGeom { double m00, m01, m10, m11, off0, off1; };

// NI_GeometricTransform: input coordinate = (offset + oy * m_0) + ox * m_1; outside [0, len - 1]: the constant 0
__device__ __forceinline__ bool rotate_coords(const RotateGeom& g, int oy, int ox, int h, int w, double& cy, double& cx) {
    cy = __dadd_rn(__dadd_rn(g.off0, __dmul_rn((double)oy, g.m00)), __dmul_rn((double)ox, g.m01));
    cx = __dadd_rn(__dadd_rn(g.off1, __dmul_rn((double)oy, g.m10)), __dmul_rn((double)ox, g.m11));
    return cy >= 0.0 && cy <= (double)(h - 1) && cx >= 0.0 && cx <= (double)(w - 1);
}

__device__ __forceinline__ int mirror_next(int idx, int len) {      // the neighbour past the last sample (its weight is 0)
    return len <= 1 ? 0 : (idx >= len ? 2 * len - 2 - idx : idx);
}

template <typename T, int ORDER>
__global__ void __launch_bounds__(256) rotate_kernel(const T* __restrict__ src, T* __restrict__ dst, int64_t total, int h,
                                                     int w, int c, int oh, int ow, const RotateGeom g) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        const int k = (int)(i % c);
        int64_t r = i / c;
        const int ox = (int)(r % ow); r /= ow;
        const int oy = (int)(r % oh);
        const int64_t n = r / oh;
        double cy, cx;
        T v = (T)0;
        if (rotate_coords(g, oy, ox, h, w, cy, cx)) {
            const T* im = src + n * h * w * c + k;
            if (ORDER == 0) {
                const int iy = (int)floor(__dadd_rn(cy, 0.5)), ix = (int)floor(__dadd_rn(cx, 0.5));
                v = im[((int64_t)min(iy, h - 1) * w + min(ix, w - 1)) * c];
            } else {
                const double fy = floor(cy), fx = floor(cx);
                const double ty = __dsub_rn(cy, fy), tx = __dsub_rn(cx, fx);
                const int y0 = (int)fy, x0 = (int)fx;
                const int ys[2] = {y0, mirror_next(y0 + 1, h)}, xs[2] = {x0, mirror_next(x0 + 1, w)};
                const double wy[2] = {__dsub_rn(1.0, ty), ty}, wx[2] = {__dsub_rn(1.0, tx), tx};
                double t = 0.0;
#pragma unroll
                for (int a = 0; a < 2; ++a)
#pragma unroll
                    for (int b = 0; b < 2; ++b) {
                        double coeff = (double)im[((int64_t)ys[a] * w + xs[b]) * c];
                        coeff = __dmul_rn(coeff, wy[a]);
                        coeff = __dmul_rn(coeff, wx[b]);
                        t = __dadd_rn(t, coeff);
                    }
                v = (T)t;
            }
        }
        dst[i] = v;
    }
}

// Rows spanned by the nearest-rotated mask for up to two rotations at once, without materialising them: what
// FindObjectHeightInRotated._func (interpreter.py:229-232) extracts from rotate_array(mask, angle, good_rotation=False),
// for the two probe angles of one step of the ternary search (:318-333).
struct SpanGeoms { RotateGeom g[2]; int oh[2], ow[2]; };

__global__ void row_span_init_kernel(int32_t* __restrict__ spans, int k) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 2 * k) spans[i] = (i & 1) ? -1 : 0x7fffffff;
}

__global__ void __launch_bounds__(256) rotated_row_span_kernel(const uint8_t* __restrict__ mask, int32_t* __restrict__ spans,
                                                               int n, int h, int w, int c, const SpanGeoms sg) {
    const int k = blockIdx.y;
    const RotateGeom g = sg.g[k];
    const int oh = sg.oh[k], ow = sg.ow[k];
    const int64_t total = (int64_t)n * oh * ow * c;
    int y0 = 0x7fffffff, y1 = -1;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        const int ch = (int)(i % c);
        int64_t r = i / c;
        const int ox = (int)(r % ow); r /= ow;
        const int oy = (int)(r % oh);
        const int64_t img = r / oh;
        double cy, cx;
        if (rotate_coords(g, oy, ox, h, w, cy, cx)) {
            const int iy = min((int)floor(__dadd_rn(cy, 0.5)), h - 1), ix = min((int)floor(__dadd_rn(cx, 0.5)), w - 1);
            if (mask[((img * h + iy) * w + ix) * c + ch]) { y0 = min(y0, oy); y1 = max(y1, oy); }
        }
    }
    for (int o = 16; o; o >>= 1) {
        y0 = min(y0, __shfl_xor_sync(0xffffffffu, y0, o));
        y1 = max(y1, __shfl_xor_sync(0xffffffffu, y1, o));
    }
    if ((threadIdx.x & 31) == 0 && y1 >= 0) { atomicMin(spans + 2 * k, y0); atomicMax(spans + 2 * k + 1, y1); }
}

__global__ void mask_bbox_init_kernel(int32_t* __restrict__ box) {
    box[0] = 0x7fffffff; box[1] = -1; box[2] = 0x7fffffff; box[3] = -1;
}

__global__ void __launch_bounds__(256) mask_bbox_kernel(const uint8_t* __restrict__ mask, int32_t* __restrict__ box,
                                                        int64_t total, int h, int w, int c) {
    int y0 = 0x7fffffff, y1 = -1, x0 = 0x7fffffff, x1 = -1;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        if (mask[i]) {
            const int64_t px = i / c;
            const int x = (int)(px % w), y = (int)((px / w) % h);
            y0 = min(y0, y); y1 = max(y1, y); x0 = min(x0, x); x1 = max(x1, x);
        }
    }
    for (int o = 16; o; o >>= 1) {
        y0 = min(y0, __shfl_xor_sync(0xffffffffu, y0, o)); y1 = max(y1, __shfl_xor_sync(0xffffffffu, y1, o));
        x0 = min(x0, __shfl_xor_sync(0xffffffffu, x0, o)); x1 = max(x1, __shfl_xor_sync(0xffffffffu, x1, o));
    }
    if ((threadIdx.x & 31) == 0 && y1 >= 0) {
        atomicMin(box + 0, y0); atomicMax(box + 1, y1); atomicMin(box + 2, x0); atomicMax(box + 3, x1);
    }
}

__global__ void __launch_bounds__(256) channel_slice_u8_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                               int64_t positions, int c, int k) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < positions; i += (int64_t)gridDim.x * 256)
        dst[i] = src[i * c + k];
}

__global__ void __launch_bounds__(256) channel_planes_u8_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                                int64_t positions, int c) {
    const int64_t total = positions * c;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        const int64_t k = i / positions, p = i - k * positions;             // writes are contiguous per plane
        dst[i] = src[p * c + k];
    }
}

inline int stage_grid(int64_t total) {
    const int64_t blocks = ceil_div(total, 256);
    return (int)(blocks < 148 * 16 ? (blocks > 0 ? blocks : 1) : 148 * 16);
}

inline bool fits32(int64_t v) { return v >= 0 && v <= 0x7fffffff; }

}  // namespace
}  // namespace uocr

using namespace uocr;

extern "C" {

int uocr_crop_masked_f32(const float* image, const int32_t* labels, int32_t label, float* out, int64_t n, int64_t h,
                         int64_t w, int64_t c, int64_t y0, int64_t x0, int64_t ch, int64_t cw, void* stream) {
    UOCR_REQUIRE(image && out, "NULL pointer");
    UOCR_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0 && fits32(h) && fits32(w) && fits32(c), "bad dimension");
    UOCR_REQUIRE(y0 >= 0 && x0 >= 0 && ch >= 0 && cw >= 0 && y0 + ch <= h && x0 + cw <= w, "region outside the image");
    const int64_t total = n * ch * cw * c;
    if (total == 0) return UOCR_OK;
    crop_masked_kernel<<<stage_grid(total), 256, 0, as_stream(stream)>>>(image, labels, label, out, total, (int)h, (int)w,
                                                                        (int)c, (int)y0, (int)x0, (int)ch, (int)cw);
    UOCR_LAUNCHED("crop_masked");
    return UOCR_OK;
}

int uocr_crop_label_mask(const int32_t* labels, int32_t label, uint8_t* out, int64_t n, int64_t h, int64_t w, int64_t y0,
                         int64_t x0, int64_t ch, int64_t cw, void* stream) {
    UOCR_REQUIRE(labels && out, "NULL pointer");
    UOCR_REQUIRE(n > 0 && h > 0 && w > 0 && fits32(h) && fits32(w), "bad dimension");
    UOCR_REQUIRE(y0 >= 0 && x0 >= 0 && ch >= 0 && cw >= 0 && y0 + ch <= h && x0 + cw <= w, "region outside the image");
    const int64_t total = n * ch * cw;
    if (total == 0) return UOCR_OK;
    crop_label_mask_kernel<<<stage_grid(total), 256, 0, as_stream(stream)>>>(labels, label, out, total, (int)h, (int)w,
                                                                            (int)y0, (int)x0, (int)ch, (int)cw);
    UOCR_LAUNCHED("crop_label_mask");
    return UOCR_OK;
}

int uocr_zoom_nearest_f32(const float* src, float* dst, int64_t n, int64_t h, int64_t w, int64_t c, int64_t out_h,
                          int64_t out_w, int64_t out_wp, void* stream) {
    UOCR_REQUIRE(src && dst, "NULL pointer");
    UOCR_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0 && fits32(h) && fits32(w) && fits32(c), "bad dimension");
    UOCR_REQUIRE(out_h >= 0 && out_w >= 0 && out_wp >= out_w && fits32(out_h) && fits32(out_wp), "bad output shape");
    const int64_t total = n * out_h * out_wp * c;
    if (total == 0) return UOCR_OK;
    // zoom = (in - 1) / (out - 1), 1 where out == 1 (scipy/ndimage/_interpolation.py: zoom_div / zoom_nominator)
    const double ry = out_h > 1 ? (double)(h - 1) / (double)(out_h - 1) : 1.0;
    const double rx = out_w > 1 ? (double)(w - 1) / (double)(out_w - 1) : 1.0;
    zoom_nearest_kernel<<<stage_grid(total), 256, 0, as_stream(stream)>>>(src, dst, total, (int)h, (int)w, (int)c,
                                                                         (int)out_h, (int)out_w, (int)out_wp, ry, rx);
    UOCR_LAUNCHED("zoom_nearest");
    return UOCR_OK;
}

static int rotate_check(const void* src, const void* dst, int64_t n, int64_t h, int64_t w, int64_t c, int64_t out_h,
                        int64_t out_w, const double* matrix, const double* offset) {
    UOCR_REQUIRE(src && dst && matrix && offset, "NULL pointer");
    UOCR_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0 && fits32(h) && fits32(w) && fits32(c), "bad dimension");
    UOCR_REQUIRE(out_h >= 0 && out_w >= 0 && fits32(out_h) && fits32(out_w), "bad output shape");
    return UOCR_OK;
}

int uocr_rotate_f32(const float* src, float* dst, int64_t n, int64_t h, int64_t w, int64_t c, int64_t out_h,
                    int64_t out_w, const double* matrix, const double* offset, int order, void* stream) {
    int rc = rotate_check(src, dst, n, h, w, c, out_h, out_w, matrix, offset);
    if (rc != UOCR_OK) return rc;
    UOCR_REQUIRE(order == 0 || order == 1, "spline order %d is not supported (0: nearest, 1: linear)", order);
    const int64_t total = n * out_h * out_w * c;
    if (total == 0) return UOCR_OK;
    const RotateGeom g{matrix[0], matrix[1], matrix[2], matrix[3], offset[0], offset[1]};
    if (order == 0)
        rotate_kernel<float, 0><<<stage_grid(total), 256, 0, as_stream(stream)>>>(src, dst, total, (int)h, (int)w, (int)c,
                                                                                 (int)out_h, (int)out_w, g);
    else
        rotate_kernel<float, 1><<<stage_grid(total), 256, 0, as_stream(stream)>>>(src, dst, total, (int)h, (int)w, (int)c,
                                                                                 (int)out_h, (int)out_w, g);
    UOCR_LAUNCHED("rotate");
    return UOCR_OK;
}

int uocr_rotate_nearest_u8(const uint8_t* src, uint8_t* dst, int64_t n, int64_t h, int64_t w, int64_t c, int64_t out_h,
                           int64_t out_w, const double* matrix, const double* offset, void* stream) {
    int rc = rotate_check(src, dst, n, h, w, c, out_h, out_w, matrix, offset);
    if (rc != UOCR_OK) return rc;
    const int64_t total = n * out_h * out_w * c;
    if (total == 0) return UOCR_OK;
    const RotateGeom g{matrix[0], matrix[1], matrix[2], matrix[3], offset[0], offset[1]};
    rotate_kernel<uint8_t, 0><<<stage_grid(total), 256, 0, as_stream(stream)>>>(src, dst, total, (int)h, (int)w, (int)c,
                                                                               (int)out_h, (int)out_w, g);
    UOCR_LAUNCHED("rotate_nearest_u8");
    return UOCR_OK;
}

int uocr_row_spans_reset(int32_t* spans, int64_t count, void* stream) {
    UOCR_REQUIRE(spans, "NULL pointer");
    UOCR_REQUIRE(count > 0 && count <= (1 << 20), "bad count");
    row_span_init_kernel<<<(unsigned)ceil_div(2 * count, 256), 256, 0, as_stream(stream)>>>(spans, (int)count);
    UOCR_LAUNCHED("row_span_init");
    return UOCR_OK;
}

int uocr_rotated_row_spans(const uint8_t* mask, int32_t* spans, int64_t n, int64_t h, int64_t w, int64_t c, int count,
                           const double* matrices, const double* offsets, const int64_t* out_shapes, int reset,
                           void* stream) {
    UOCR_REQUIRE(mask && spans && matrices && offsets && out_shapes, "NULL pointer");
    UOCR_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0 && fits32(h) && fits32(w) && fits32(c), "bad dimension");
    UOCR_REQUIRE(count == 1 || count == 2, "1 or 2 rotations per call, got %d", count);
    SpanGeoms sg{};
    int64_t most = 0;
    for (int k = 0; k < count; ++k) {
        UOCR_REQUIRE(out_shapes[2 * k] >= 0 && out_shapes[2 * k + 1] >= 0 && fits32(out_shapes[2 * k]) &&
                     fits32(out_shapes[2 * k + 1]), "bad output shape");
        sg.g[k] = RotateGeom{matrices[4 * k], matrices[4 * k + 1], matrices[4 * k + 2], matrices[4 * k + 3],
                             offsets[2 * k], offsets[2 * k + 1]};
        sg.oh[k] = (int)out_shapes[2 * k]; sg.ow[k] = (int)out_shapes[2 * k + 1];
        const int64_t total = n * out_shapes[2 * k] * out_shapes[2 * k + 1] * c;
        if (total > most) most = total;
    }
    if (reset) {
        row_span_init_kernel<<<1, 256, 0, as_stream(stream)>>>(spans, count);
        UOCR_LAUNCHED("row_span_init");
    }
    if (most == 0) return UOCR_OK;
    rotated_row_span_kernel<<<dim3((unsigned)stage_grid(most), (unsigned)count), 256, 0, as_stream(stream)>>>(
        mask, spans, (int)n, (int)h, (int)w, (int)c, sg);
    UOCR_LAUNCHED("rotated_row_span");
    return UOCR_OK;
}

int uocr_channel_slice_u8(const uint8_t* src, uint8_t* dst, int64_t positions, int64_t c, int64_t k, void* stream) {
    UOCR_REQUIRE(src && dst, "NULL pointer");
    UOCR_REQUIRE(positions >= 0 && c > 0 && k >= 0 && k < c && fits32(c), "bad dimension");
    if (positions == 0) return UOCR_OK;
    channel_slice_u8_kernel<<<stage_grid(positions), 256, 0, as_stream(stream)>>>(src, dst, positions, (int)c, (int)k);
    UOCR_LAUNCHED("channel_slice_u8");
    return UOCR_OK;
}

int uocr_channel_planes_u8(const uint8_t* src, uint8_t* dst, int64_t positions, int64_t c, void* stream) {
    UOCR_REQUIRE(src && dst, "NULL pointer");
    UOCR_REQUIRE(positions >= 0 && c > 0 && fits32(c), "bad dimension");
    if (positions == 0) return UOCR_OK;
    channel_planes_u8_kernel<<<stage_grid(positions * c), 256, 0, as_stream(stream)>>>(src, dst, positions, (int)c);
    UOCR_LAUNCHED("channel_planes_u8");
    return UOCR_OK;
}

int uocr_mask_bbox(const uint8_t* mask, int32_t* box, int64_t n, int64_t h, int64_t w, int64_t c, void* stream) {
    UOCR_REQUIRE(mask && box, "NULL pointer");
    UOCR_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0 && fits32(h) && fits32(w) && fits32(c), "bad dimension");
    mask_bbox_init_kernel<<<1, 1, 0, as_stream(stream)>>>(box);
    UOCR_LAUNCHED("mask_bbox_init");
    const int64_t total = n * h * w * c;
    mask_bbox_kernel<<<stage_grid(total), 256, 0, as_stream(stream)>>>(mask, box, total, (int)h, (int)w, (int)c);
    UOCR_LAUNCHED("mask_bbox");
    return UOCR_OK;
}

}  // extern "C"
