// Monochrome conv pair on the tensor cores, both convolutions, operands resident in TENSOR MEMORY.
//
//   y = act2(conv3x3(act1(conv3x3(x, w1) + b1), w2) + b2)        1 -> 16 -> 1 channels, padding 1, stride 1
//   replaces: make_monochrome's conv_1 -> leaky_relu_1 -> conv_2 -> sigmoid (my_model/model.py:119-122), i.e.
//   Convolutional2D._forward (convolutional.py:62-99) twice + LeakyRelu/Sigmoid._forward (layers.py:390-415).
//
// The pair costs 288 FMA per pixel; on the CUDA cores (conv_fast.cu) that is FFMA-issue bound (0.44 ms for 64
// tiles of 496 x 736, FMA pipe 62 % busy).  The obvious implicit GEMM (M = pixels, K = taps x channels, N = Cout)
// wastes the tensor core because Cout = 1.  This kernel turns BOTH convolutions into GEMMs whose N dimension is
// full, by contracting over what is small and shifting afterwards:
//
//   GEMM 1   H[p, c]  = sum_t X9[p, t] . w1[t, c]          M = 128 hidden pixels, K = 16 (9 taps of the 3 x 3
//                                                          window of x around p, a constant 1 that carries b1,
//                                                          6 zeros), N = 16 hidden channels
//   act1     A2[p, c] = max(H, alpha H)                     (LeakyReLU, in registers)
//   GEMM 2   Z[p, t]  = sum_c A2[p, c] . w2[t, c]          K = 16 channels, N = 16 (9 taps used): the
//                                                          contribution of hidden pixel p to its 9 neighbours
//   shift    y[q]     = act2(b2 + sum_t Z[q + off(t), t])   9 adds per pixel instead of 144 FMA
//
// Four tcgen05.mma (128 x 16 x 8, kind::tf32, 8 cycles each) per 128 hidden pixels replace 128 x 288 FFMA.  The A
// operands never touch shared memory: lane = hidden pixel, so each thread writes the 3 x 3 window it already holds
// in registers into its own TMEM lane with tcgen05.st, GEMM 1 reads A from TMEM and leaves H in TMEM, the thread
// reads its H row back with tcgen05.ld, applies act1, stores A2 in place, GEMM 2 reads it from TMEM and writes Z
// over the dead X9 columns (32 TMEM columns per hidden row in flight).  B operands (w1 + b1, w2: 2 x 1 KB) sit in
// shared memory in the K-major no-swizzle layout.
//
// Geometry: a warp owns a strip of 32 hidden columns (30 output columns + 1 halo column each side) and streams
// down a band of RB output rows; the vertical part of the shift is a rolling set of 3 x 3 partial sums in
// registers, the horizontal part two warp shuffles per output pixel.  The four warps of a CTA (4 strips, together
// the M = 128 rows of one MMA) advance in lock step, two hidden rows per pipeline step:
//   st X9 -> [barA] -> MMA1 -> [barD1] -> ld H, act, st A2 -> [barA2] -> MMA2 -> [barD2] -> ld Z, shift, store
// Each hop costs ~500 cycles (measured, tools/tmem_probe.cu), so latency is hidden by occupancy: 8 CTAs per SM x
// 64 TMEM columns = all 512 columns, 16 hidden rows in flight per SM.
//
// Numerics: x and the weights are rounded to TF32 (cvt.rna); H is consumed as TF32 by truncation inside the
// tensor core, made unbiased by scaling w1 and b1 with (1 + 2^-11) when the B operand is built (act1 is positively
// homogeneous); accumulation is FP32.  Tolerance in tests: 1e-3 of the output range (TF32 mode).
#include "conv_common.cuh"
#include "tc_ptx.cuh"

namespace uocr {

constexpr int PM_OUT = 30;            // output columns per warp strip
constexpr int PM_THREADS = 128;       // 4 warps = the 128 rows of one UMMA
constexpr int PM_TMEM_COLS = 64;      // 2 hidden rows in flight x (16 columns X / Z + 16 columns H / A2)

struct PairTcParams {
    const float* x; const float* w1; const float* b1; const float* w2; const float* b2; float* y;
    int H, W;
    int strips, bands, rb;            // strips per image row, bands per image, output rows per band
    int64_t items;                    // n * bands * strips
    int steps;                        // ceil((rb + 2) / 2)
    float alpha1, alpha2;
};

// Per-thread state.  The x window is a ring of 4 image rows; slot s holds {x[., hc-1], x[., hc], x[., hc+1], one}
// as one register quad that tcgen05.st.x4 writes to TMEM columns 4s..4s+3 without any register shuffling: the K
// index of GEMM 1 is (ring slot, column), and the B operand is stored in 4 row-rotated variants (one per ring
// phase) instead of rotating the registers.  `one` (1 for a hidden pixel inside the image, else 0) multiplies b1
// in the variant's row 3 and zero weights elsewhere.
struct PairTcState {
    uint32_t xw[4][4];
    float acc[3][3];                  // rolling partial sums: [output row % 3][kx]
};

// ring slot <- x[xr, hc - 1 .. hc + 1] (xrow = &x[xr, hc]); zero outside the image / for idle lanes
__device__ __forceinline__ void pm_load_row(const float* __restrict__ xrow, bool rowok, bool c0, bool c1, bool c2,
                                            uint32_t* dst) {
    dst[0] = (rowok && c0) ? __float_as_uint(__ldg(xrow - 1)) : 0u;
    dst[1] = (rowok && c1) ? __float_as_uint(__ldg(xrow)) : 0u;
    dst[2] = (rowok && c2) ? __float_as_uint(__ldg(xrow + 1)) : 0u;
}
// FP32 -> TF32 round to nearest: the tensor core ignores the low 13 mantissa bits, so adding half an ulp is enough
// (volatile: keeps the add where it is written -- hoisted next to the prefetching load it would stall the warp)
__device__ __forceinline__ void pm_round_row(uint32_t* q) {
    asm volatile("add.u32 %0, %0, 0x1000;\n\tadd.u32 %1, %1, 0x1000;\n\tadd.u32 %2, %2, 0x1000;"
                 : "+r"(q[0]), "+r"(q[1]), "+r"(q[2]));
}
__device__ __forceinline__ void tc_st4_nowait(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
__device__ __forceinline__ void tc_st16_zero(uint32_t taddr) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
        ::"r"(taddr), "r"(0u) : "memory");
}
__device__ __forceinline__ uint32_t pm_tmem_base(uint32_t slot_addr) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(slot_addr));
    return v;
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// One pipeline step = hidden rows i0 = 2k and i0 + 1 of the band.  PX = i0 % 4 (x ring phase), PH = i0 % 3
// (partial-sum ring phase) are compile-time so that every register array index is static.
template <int PX, int PH, bool LEAKY, bool SIGMOID>
__device__ __forceinline__ void pm_step(const PairTcParams& p, PairTcState& s, int k, const float* __restrict__& xnext,
                                        float* __restrict__& ynext, int hr0, int nrows, bool c0, bool c1, bool c2,
                                        bool store_lane, bool warp0, uint32_t tlane, uint32_t bars, uint32_t sb1,
                                        uint32_t sb2, uint32_t stm, float bias2) {
    const uint32_t par = (uint32_t)(k & 1);
    const int i0 = 2 * k;
    // ---------------- phase A: the x ring (3 x 3 windows of two hidden rows) -> TMEM, GEMM 1's A operand
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int hr = hr0 + i0 + r;
        const uint32_t tx = tlane + (uint32_t)(32 * r);
        if (hr >= 0 && hr < p.H) {                        // warp-uniform
#pragma unroll
            for (int q = 0; q < 4; ++q) tc_st4_nowait(tx + 4 * q, s.xw[q]);
        } else {                                          // hidden rows outside the image are conv_2's zero padding
            tc_st16_zero(tx);
        }
    }
    tc_wait_st();
    tc_fence_before();
    mbar_arrive(bars + 0);
    // Prefetch the two x rows the NEXT step adds to the window (x rows hr0 + i0 + 3, + 4), straight into the ring
    // slots that died with this step's stores; their latency hides behind this step's waits; rounded in phase C.
    {
        const int xr = hr0 + i0 + 3;
        pm_load_row(xnext, xr >= 0 && xr < p.H, c0, c1, c2, s.xw[PX % 4]);
        pm_load_row(xnext + p.W, xr + 1 >= 0 && xr + 1 < p.H, c0, c1, c2, s.xw[(PX + 1) % 4]);
        xnext += 2 * p.W;
    }
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
    if (warp0) {                                          // warp 0: tlane == TMEM base (lane field 0)
        mbar_wait(bars + 0, par);
        tc_fence_after();
        if (elect_one()) {
            const uint32_t tb = pm_tmem_base(stm);        // re-read: keeps the MMA operand addresses out of the
#pragma unroll                                            // loop-carried register set
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int m = 0; m < 2; ++m)
                    tc_mma_tf32_ts(tb + (uint32_t)(32 * r + 16), tb + (uint32_t)(32 * r + 8 * m),
                                   make_kmajor_nosw_desc(sb1 + (uint32_t)(((PX + r) % 4) * 1024 + 2 * m * 256), 256, 128),
                                   idesc, m);
            tc_commit(bars + 8);
        }
        __syncwarp();
    }
    // ---------------- phase B: H -> act1 -> A2 (in place)
    mbar_wait(bars + 8, par);
    tc_fence_after();
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        uint32_t h[16];
        tc_ld16_nowait(tlane + (uint32_t)(32 * r + 16), h);
        tc_wait_ld();
        if (LEAKY) {
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                const float v = __uint_as_float(h[c]);
                h[c] = __float_as_uint(fmaxf(v, v * p.alpha1));
            }
        }
        tc_st16_nowait(tlane + (uint32_t)(32 * r + 16), h);
    }
    tc_wait_st();
    tc_fence_before();
    mbar_arrive(bars + 16);
    if (warp0) {
        mbar_wait(bars + 16, par);
        tc_fence_after();
        if (elect_one()) {
            const uint32_t tb = pm_tmem_base(stm);
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int m = 0; m < 2; ++m)
                    tc_mma_tf32_ts(tb + (uint32_t)(32 * r), tb + (uint32_t)(32 * r + 16 + 8 * m),
                                   make_kmajor_nosw_desc(sb2 + (uint32_t)(2 * m) * 256u, 256, 128), idesc, m);
            tc_commit(bars + 24);
        }
        __syncwarp();
    }
    // ---------------- phase C: Z -> rolling shift-and-add -> output rows
    mbar_wait(bars + 24, par);
    tc_fence_after();
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int ph = (PH + r) % 3;
        uint32_t z[9];
        tc_ld8_nowait(tlane + (uint32_t)(32 * r), z);
        tc_ld1_nowait(tlane + (uint32_t)(32 * r + 8), z + 8);
        tc_wait_ld();
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            s.acc[(ph + 2) % 3][kx] = __uint_as_float(z[kx]);                 // ky = 0: opens output row i + 2
            s.acc[(ph + 1) % 3][kx] += __uint_as_float(z[3 + kx]);            // ky = 1: output row i + 1
            s.acc[ph][kx] += __uint_as_float(z[6 + kx]);                      // ky = 2: completes output row i
        }
        // y[oy, ox] = b2 + Z-sum of hidden columns ox - 1 (kx = 0), ox (kx = 1), ox + 1 (kx = 2)
        const float left = __shfl_up_sync(0xffffffffu, s.acc[ph][0], 1);
        const float right = __shfl_down_sync(0xffffffffu, s.acc[ph][2], 1);
        float v = s.acc[ph][1] + left + right + bias2;
        if (SIGMOID) v = __fdividef(1.f, 1.f + __expf(-v));
        else if (p.alpha2 != 1.f) v = v >= 0.f ? v : v * p.alpha2;            // LeakyReLU; alpha2 == 1: no activation
        const int orow = i0 + r - 2;                      // output row of the band completed by hidden row i
        if (store_lane && orow >= 0 && orow < nrows) *ynext = v;
        if (orow >= 0) ynext += p.W;
    }
    pm_round_row(s.xw[PX % 4]);
    pm_round_row(s.xw[(PX + 1) % 4]);
}

template <bool LEAKY, bool SIGMOID, int OCC>
__global__ void __launch_bounds__(PM_THREADS, OCC) conv3x3_pair_tmem_kernel(const PairTcParams p) {
    __shared__ __align__(128) float s_b1[4 * 256];      // 4 ring phases x chunk kq (= ring slot): 16 rows (n) x 4 floats
    __shared__ __align__(128) float s_b2[256];          // chunk kq: 16 rows (n = tap) x 4 floats (k = channel 4 kq + kk)
    __shared__ __align__(8) uint64_t s_bar[4];
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    for (int i = tid; i < 4 * 256; i += PM_THREADS) {
        // B1 variant phi: B1[n = channel][k = 4 slot + b]: ring slot holds window row a = (slot - phi) mod 4
        const int phi = i >> 8, slot = (i >> 6) & 3, n = (i & 63) >> 2, b = i & 3;
        const int a = (slot - phi) & 3;
        float v = 0.f;
        if (a < 3 && b < 3) v = __ldg(p.w1 + (a * 3 + b) * 16 + n);
        if (slot == 0 && b == 3) v = __ldg(p.b1 + n);
        // scaled so that the tensor core's truncation of H to TF32 is unbiased
        s_b1[i] = round_tf32(v * (1.f + 1.f / 2048.f));
    }
    for (int i = tid; i < 256; i += PM_THREADS) {
        const int kq = i >> 6, n = (i & 63) >> 2, k = kq * 4 + (i & 3);
        s_b2[i] = n < 9 ? round_tf32(__ldg(p.w2 + n * 16 + k)) : 0.f;
    }
    if (tid == 0) {
        mbar_init(smem_u32(&s_bar[0]), PM_THREADS);       // X stored by every thread
        mbar_init(smem_u32(&s_bar[1]), 1);                // MMA1 committed
        mbar_init(smem_u32(&s_bar[2]), PM_THREADS);       // A2 stored
        mbar_init(smem_u32(&s_bar[3]), 1);                // MMA2 committed
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // s_b1 / s_b2 are read by the async proxy
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(&s_tmem)), "r"((uint32_t)PM_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tlane = s_tmem + ((uint32_t)(warp * 32) << 16);

    // work item of this warp: (image, band, strip); warps past the end run on zeros and store nothing
    int64_t item = (int64_t)blockIdx.x * 4 + warp;
    const bool active = item < p.items;
    if (!active) item = p.items - 1;
    const int strip = (int)(item % p.strips);
    const int64_t rest = item / p.strips;
    const int band = (int)(rest % p.bands);
    const int64_t img = rest / p.bands;
    const int hc = strip * PM_OUT - 1 + lane;           // hidden column of this lane
    const bool c1 = active && hc >= 0 && hc < p.W;      // hidden pixels outside the image are zero padding
    const bool c0 = c1 && hc - 1 >= 0, c2 = c1 && hc + 1 < p.W;
    const bool store_lane = c1 && lane >= 1 && lane <= PM_OUT;
    const int hr0 = band * p.rb - 1;                    // first hidden row of the band
    const int nrows = min(p.H - band * p.rb, p.rb);     // output rows of the band
    const float bias2 = __ldg(p.b2);
    // &x[first x row of the band = hr0 - 1, hc] (never dereferenced outside the image), &y[band's first row, hc]
    const float* xnext = p.x + (img * p.H + (hr0 - 1)) * p.W + hc;
    float* ynext = p.y + (img * p.H + (hr0 + 1)) * p.W + hc;

    PairTcState s;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) s.acc[a][b] = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int xr = hr0 - 1 + a;
        pm_load_row(xnext, xr >= 0 && xr < p.H, c0, c1, c2, s.xw[a]);
        pm_round_row(s.xw[a]);
        s.xw[a][3] = c1 ? __float_as_uint(1.f) : 0u;
        xnext += p.W;
    }
    const uint32_t bars = smem_u32(&s_bar[0]);
    const uint32_t sb1 = smem_u32(s_b1), sb2 = smem_u32(s_b2);
    const bool warp0 = warp == 0;

#define PM_STEP(PX, PH, K) \
    pm_step<PX, PH, LEAKY, SIGMOID>(p, s, K, xnext, ynext, hr0, nrows, c0, c1, c2, store_lane, warp0, tlane, bars, sb1, \
                                    sb2, smem_u32(&s_tmem), bias2)
#pragma unroll 1
    for (int k = 0; k < p.steps; k += 6) {          // (2k) % 4 and (2k) % 3 are static inside the body
        PM_STEP(0, 0, k);
        if (k + 1 < p.steps) PM_STEP(2, 2, k + 1);
        if (k + 2 < p.steps) PM_STEP(0, 1, k + 2);
        if (k + 3 < p.steps) PM_STEP(2, 0, k + 3);
        if (k + 4 < p.steps) PM_STEP(0, 2, k + 4);
        if (k + 5 < p.steps) PM_STEP(2, 1, k + 5);
    }
#undef PM_STEP

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"
                     ::"r"(s_tmem), "r"((uint32_t)PM_TMEM_COLS) : "memory");
    }
}

static int env_int_pm(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

int conv3x3_pair_tmem(const float* x, const float* w1, const float* b1, const float* w2, const float* b2, float* y,
                      int64_t n, int64_t h, int64_t w, int c1, int act1, float alpha1, int act2, float alpha2,
                      cudaStream_t st) {
    if (c1 != 16) return UOCR_ERR_UNSUPPORTED;
    const bool leaky = act1 == UOCR_ACT_LEAKY;
    if (!(act1 == UOCR_ACT_NONE || (leaky && alpha1 >= 0.f && alpha1 <= 1.f))) return UOCR_ERR_UNSUPPORTED;
    static const int rb_env = env_int_pm("UOCR_PAIR_RB", 30);
    PairTcParams p{};
    p.x = x; p.w1 = w1; p.b1 = b1; p.w2 = w2; p.b2 = b2; p.y = y;
    p.H = (int)h; p.W = (int)w;
    p.rb = (int)(h < rb_env ? h : rb_env);
    p.strips = (int)ceil_div(w, PM_OUT);
    p.bands = (int)ceil_div(h, p.rb);
    p.items = n * p.bands * p.strips;
    p.steps = (p.rb + 2 + 1) / 2;
    p.alpha1 = alpha1;
    p.alpha2 = act2 == UOCR_ACT_LEAKY ? alpha2 : 1.f;
    const int64_t ctas = ceil_div(p.items, 4);
    if (ctas > 0x7fffffff) return UOCR_ERR_UNSUPPORTED;
    const unsigned grid = (unsigned)ctas;
    // CTAs per SM (register cap 64 / 72 / 80 per thread): 8 fills all 512 TMEM columns but spills
    static const int occ = env_int_pm("UOCR_PAIR_OCC", 7);
#define PM_LAUNCH(LK, SG)                                                                           \
    do {                                                                                            \
        if (occ >= 8) conv3x3_pair_tmem_kernel<LK, SG, 8><<<grid, PM_THREADS, 0, st>>>(p);          \
        else if (occ == 7) conv3x3_pair_tmem_kernel<LK, SG, 7><<<grid, PM_THREADS, 0, st>>>(p);     \
        else conv3x3_pair_tmem_kernel<LK, SG, 6><<<grid, PM_THREADS, 0, st>>>(p);                   \
    } while (0)
    if (act2 == UOCR_ACT_SIGMOID) {
        if (leaky) PM_LAUNCH(true, true); else PM_LAUNCH(false, true);
    } else {
        if (leaky) PM_LAUNCH(true, false); else PM_LAUNCH(false, false);
    }
#undef PM_LAUNCH
    UOCR_LAUNCHED("conv3x3_pair_tmem");
    return UOCR_OK;
}

}  // namespace uocr
