// Monochrome conv pair, third tensor-core formulation: the 1 -> 16 convolution reads its A operand STRAIGHT FROM THE
// IMAGE ROWS in shared memory (no im2col, no per-pixel register work), warp-specialised and persistent.
//
//   y = act2(conv3x3(act1(conv3x3(x, w1) + b1), w2) + b2)        1 -> 16 -> 1 channels, padding 1, stride 1
//   replaces: make_monochrome's conv_1 -> leaky_relu_1 -> conv_2 -> sigmoid (my_model/model.py:119-122), i.e.
//   Convolutional2D._forward (convolutional.py:62-99) twice + LeakyRelu/Sigmoid._forward (layers.py:390-415).
//
// conv3x3_pair_tmem_kernel (conv_pair_tc.cu) builds the 3 x 3 window of every hidden pixel in registers and stores it
// into tensor memory: ~125 SASS instructions per pixel and thread, 0.137 ms for 64 tiles (ncu: issue-bound on the CUDA
// cores, tensor pipe 20 %).  Here the CUDA cores only do what no tensor core can: LeakyReLU and the final shift-add.
//
// GEMM 1 (space-to-depth along x).  A shared-memory matrix descriptor is address arithmetic: row m of the A operand
// starts SBO/8 = 16 bytes after row m - 1.  With a row of the image lying contiguously in shared memory, "row m" of A
// is therefore the 12 floats x[4m .. 4m + 11] -- rows OVERLAP, nothing is copied.  Position m owns the four hidden
// pixels 4m + c0 + phase (phase = 0..3, c0 = 3 or 5 chosen per tile so that TMA box origins stay 16-byte aligned),
// whose 3-wide windows all lie inside those 12 floats:
//     H[m, (phase, c)] = sum_{ky, f} x[row + ky - 1, 4m + f] . B1[(ky, f), (phase, c)]        M = 128, N = 64, K = 36 + 4
// B1 holds w1[ky, kx, c] at f = c0 + phase + kx - 1 and zeros elsewhere; the bias rides on a constant chunk {1,0,0,0}
// that every row reads (the descriptor's K stride, LBO, may point anywhere).  K = 40 = five tcgen05.mma 128x64x8:
// three take chunks 0-1 of one image row each, one pairs chunk 2 of rows 0 and 1 (LBO = row pitch of the ring), one
// pairs chunk 2 of row 2 with the constant chunk.  The M = 128 rows are two independent x tiles of 64 positions
// (62 valid: 248 hidden pixels, 246 outputs), so three tiles cover the 736-pixel page tile with 1 % waste.
// ACT     four warps read H (tcgen05.ld, lane = position, 64 columns), apply LeakyReLU, zero hidden pixels that lie
//         outside the image (conv_2's zero padding) and store A2 in place.
// GEMM 2  as in conv_pair_tc.cu: Z[m, (phase, tap)] = sum_c A2[m, (phase, c)] . w2[tap, c]   (A from TMEM, 8 MMAs
//         128x16x8 per row), the nine taps stay in N and are combined afterwards:
// EPI     four warps keep rolling vertical sums (3 output rows x 4 pixels x 3 kx per thread), exchange the two
//         horizontal neighbours by warp shuffle (and through shared memory across the one warp boundary inside a
//         tile), add b2, apply act2 and store.
// One persistent CTA per SM: TMA warp (image rows -> 8-slot ring, zero fill outside the image = conv_1's padding,
// rounded to TF32 by the TMA unit), MMA warp, 4 ACT warps, 4 EPI warps; 4 hidden rows in flight in the 512 TMEM columns
// (64 H/A2 + 64 Z each); mbarrier pipeline full/empty (ring), h_full, a2_full, z_full, z_empty.
//
// Numerics: as conv_pair_tc.cu -- x and weights rounded to TF32, H consumed as TF32 by truncation inside the tensor
// core and made unbiased by scaling w1 and b1 with (1 + 2^-11); FP32 accumulation.  Tolerance in tests: 1e-3.
#include <cuda.h>

#include "conv_common.cuh"
#include "tc_ptx.cuh"

namespace uocr {

int make_tmap_plain_tf32(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                         const uint64_t* strides_bytes, const uint32_t* box);      // tc_gemm.cu

namespace {

constexpr int PR_THREADS = 672;          // warps 0-7 ACT (two per TMEM lane quarter), 8-15 EPI (two per quarter), 16 TMA, 17-20 MMA (one per TMEM slot)
constexpr int PR_RING = 8;               // image rows in the shared-memory ring (+ 1 mirror slot)
constexpr int PR_SLOT = 2048;            // bytes per ring slot: 2 tiles x 256 floats
constexpr int PR_OUT = 246;              // output columns per tile
constexpr int PR_TSLOTS = 4;             // hidden rows in flight in tensor memory
constexpr int PR_LAG = 2;                // GEMM 2 of hidden row r is issued after GEMM 1 of row r + PR_LAG

struct PairRowsParams {
    const float* w1; const float* b1; const float* w2; const float* b2; float* y;
    int N, H, W;
    int xt, nb, rb, npair;               // x tiles, bands, rows per band, image pairs
    int total;                           // work units = nb * xt * npair
    float alpha1, alpha2;
};

struct __align__(128) PairRowsSmem {
    float ring[(PR_RING + 1) * PR_SLOT / 4];   // slot s: tile 0 floats [0, 256), tile 1 floats [256, 512)
    float ones[128 * 4];                       // {1, 0, 0, 0} per A row: the K chunk that carries b1
    float b1[2][10 * 64 * 4];                  // per c0 variant: chunk kc (16 B) of row n = (phase, c) at kc * 1024 + n * 16
    float b2[4][256];                          // per phase: chunk kq of row n at kq * 256 + n * 16; row n = tap pr_tap(phase, n)
    float xch[2][2][2];                        // [EPI type][tile][exchange parity]
    uint64_t full[PR_RING], empty[PR_RING];
    uint64_t hfull[PR_TSLOTS], a2full[PR_TSLOTS], zfull[PR_TSLOTS], zempty[PR_TSLOTS];
    uint64_t drain;                            // the MMA warp's last commit: nothing asynchronous outlives the CTA
    uint32_t tmem;
};

struct WorkUnit {
    int y0, rows;        // first output row of the band, output rows
    int t, n0, n1;       // x tile, the two images (n1 may be >= N: idle half)
    int S, c0;           // image x of ring float 0; hidden pixel of position m, phase f: x = S + 4 m + c0 + f
};

__device__ __forceinline__ WorkUnit pr_decode(const PairRowsParams& p, int w) {
    WorkUnit u;
    const int pi = w % p.npair;
    const int rest = w / p.npair;
    u.t = rest % p.xt;
    const int b = rest / p.xt;
    u.y0 = b * p.rb;
    u.rows = min(p.rb, p.H - u.y0);
    u.n0 = 2 * pi;
    u.n1 = 2 * pi + 1;
    // hidden pixel 0 of the tile is image x = PR_OUT t - 1; ring float 0 must sit at a multiple of 4 (TMA box origin)
    u.c0 = (u.t & 1) ? 5 : 3;
    u.S = PR_OUT * u.t - 1 - u.c0;
    return u;
}

// plain try_wait spin.  (A suspend-time hint turns the failed probe into NANOSLEEP.SYNCS, whose wake-up costs ~1 us:
// fine where 16 rows per SM hide it (conv_pair_tc.cu), fatal here where a row's stages are chained -- measured 1715
// cycles per row with the hint.)
__device__ __forceinline__ void pr_wait(uint32_t bar, uint32_t parity) { mbar_wait(bar, parity); }

__device__ __forceinline__ void pr_ld16(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    tc_ld16_nowait(taddr, r);
}
__device__ __forceinline__ void pr_st16(uint32_t taddr, const float* v) {
    tc_st16_nowait(taddr, reinterpret_cast<const uint32_t*>(v));
}

__device__ __forceinline__ bool pr_elect() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// ------------------------------------------------------------------ TMA warp (convergent; one elected lane issues)
__device__ void pr_producer(const PairRowsParams& p, const CUtensorMap* map, PairRowsSmem& sm) {
    const uint32_t ring = smem_u32(sm.ring);
    uint32_t ut = 0;                                       // input rows issued so far (all work units)
    for (int w = blockIdx.x; w < p.total; w += gridDim.x) {
        const WorkUnit u = pr_decode(p, w);
        const int nin = u.rows + 4;                        // image rows y0 - 2 .. y0 + rows + 1
        for (int r = 0; r < nin; ++r, ++ut) {
            const uint32_t slot = ut % PR_RING;
            pr_wait(smem_u32(&sm.empty[slot]), ((ut / PR_RING) & 1u) ^ 1u);
            if (pr_elect()) {
                const uint32_t bar = smem_u32(&sm.full[slot]);
                const bool mirror = slot == 0;             // also at the end of the ring: slots (RING - 1, RING) are adjacent
                mbar_arrive_expect_tx(bar, mirror ? 2u * PR_SLOT : (uint32_t)PR_SLOT);
                const int y = u.y0 - 2 + r;
                const uint32_t dst = ring + slot * PR_SLOT;
                tma_load_3d(dst, map, bar, u.S, y, u.n0);          // out-of-image rows / columns / images: zeros
                tma_load_3d(dst + 1024, map, bar, u.S, y, u.n1);
                if (mirror) {
                    const uint32_t dst2 = ring + PR_RING * PR_SLOT;
                    tma_load_3d(dst2, map, bar, u.S, y, u.n0);
                    tma_load_3d(dst2 + 1024, map, bar, u.S, y, u.n1);
                }
            }
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------ MMA warps (convergent; one elected lane issues)
// FOUR issuing warps, one per tensor-memory slot: warp j owns the hidden rows whose running index is j mod 4, i.e.
// exactly the rows that use TMEM slot j, and issues both GEMMs of those rows.  A single issuer was the pipeline's
// pacemaker: ~120 mostly dependent scalar / uniform instructions per row (waits, descriptor arithmetic, 13 MMAs,
// 3 commits) = ~1100 cycles per row on one warp, 4x the tensor time (ncu: every other warp waiting on h_full).
// Hazards inside a slot (GEMM 1 of the next row overwrites the H / A2 columns GEMM 2 of the previous row reads) are
// ordered by the issuing thread's program order; hazards across slots do not exist.  An image row of the ring is
// read by up to three hidden rows, issued by three different warps, so its `empty` barrier counts three arrivals:
// every hidden row commits (or, out of the image, plainly arrives) on the three rows it covers, and the first / last
// hidden row of a unit add the arrivals of the readers that do not exist.
__device__ __forceinline__ void pr_conv2(PairRowsSmem& sm, uint32_t tmem, int j, uint32_t kv, uint64_t db2_0,
                                         uint64_t db2_1, uint32_t id16) {
    const uint32_t par = (kv / PR_TSLOTS) & 1u;
    pr_wait(smem_u32(&sm.a2full[j]), par);
    pr_wait(smem_u32(&sm.zempty[j]), par ^ 1u);
    tc_fence_after();
    if (pr_elect()) {
        const uint32_t base = tmem + 128u * (uint32_t)j;
#pragma unroll
        for (int f = 0; f < 4; ++f) {
            tc_mma_tf32_ts(base + 64u + 16u * f, base + 16u * f, db2_0 + 64u * f, id16, 0);
            tc_mma_tf32_ts(base + 64u + 16u * f, base + 16u * f + 8u, db2_1 + 64u * f, id16, 1);
        }
        tc_commit(smem_u32(&sm.zfull[j]));
    }
    __syncwarp();
}

__device__ void pr_mma(const PairRowsParams& p, PairRowsSmem& sm, uint32_t tmem, int j) {
    const uint32_t ring = smem_u32(sm.ring), ones = smem_u32(sm.ones), sb2 = smem_u32(sm.b2);
    // instruction descriptors: D FP32, A / B TF32, K-major both, M = 128, N = 64 / 16
    const uint32_t id64 = (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t id16 = (1u << 4) | (2u << 7) | (2u << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t da_row = make_kmajor_nosw_desc(0, 16, 128);         // + start: chunks 0, 1 of one image row
    const uint64_t da_pair = make_kmajor_nosw_desc(0, PR_SLOT, 128);   // + start: chunk 2 of two adjacent ring slots
    const uint64_t da_bias = make_kmajor_nosw_desc(0, 0, 128);         // + start + LBO: chunk 2 of a row, then `ones`
    // + 64 f: phase f's copy of w2 (1 KB apart = 64 16-byte units), rows in that phase's tap order (pr_tap)
    const uint64_t db2_0 = make_kmajor_nosw_desc(sb2, 256, 128), db2_1 = make_kmajor_nosw_desc(sb2 + 512, 256, 128);
    const uint32_t d1 = tmem + 128u * (uint32_t)j;
    uint32_t ut = 0, kv = 0;                               // image rows before this unit; valid hidden rows so far
    bool pending = false;                                  // GEMM 2 of this warp's previous row not issued yet
    uint32_t pkv = 0;
    for (int w = blockIdx.x; w < p.total; w += gridDim.x) {
        const WorkUnit u = pr_decode(p, w);
        const uint32_t sb1 = smem_u32(sm.b1[u.c0 == 5 ? 1 : 0]);
        const uint64_t db_ky0 = make_kmajor_nosw_desc(sb1 + 0 * 1024, 1024, 128);
        const uint64_t db_ky1 = make_kmajor_nosw_desc(sb1 + 3 * 1024, 1024, 128);
        const uint64_t db_ky2 = make_kmajor_nosw_desc(sb1 + 6 * 1024, 1024, 128);
        const uint64_t db_c2 = make_kmajor_nosw_desc(sb1 + 2 * 1024, 3 * 1024, 128);
        const uint64_t db_bias = make_kmajor_nosw_desc(sb1 + 8 * 1024, 1024, 128);
        const int nh = u.rows + 2;                         // hidden rows y0 - 1 .. y0 + rows
        // rows of the unit inside the image: [r_lo, r_hi); the (at most two) rows outside belong to warp 0
        const int r_lo = u.y0 == 0 ? 1 : 0, r_hi = min(nh, p.H - u.y0 + 1);
        // this warp's rows: valid row r has running index kv + (r - r_lo); the first one with index = j mod 4, then
        // every fourth; warp 0 also takes the row above (r = 0) / below (r = r_hi) the image where the unit has one
        const int below = (j == 0 && r_hi < nh) ? r_hi : nh;
        int first = r_lo + (int)((uint32_t)(j - (int)(kv & 3u)) & 3u);
        if (first >= r_hi) first = below;
        int r = (j == 0 && r_lo == 1) ? 0 : first;
        while (r < nh) {
            const bool valid = r >= r_lo && r < r_hi;
            const uint32_t kvr = kv + (uint32_t)(r - r_lo);
            int r_next = r < r_lo ? first : (valid ? r + 4 : nh);
            if (valid && r_next >= r_hi) r_next = below;
            if (pending) {                                 // same slot: before GEMM 1 overwrites its H / A2 columns
                pr_conv2(sm, tmem, j, pkv, db2_0, db2_1, id16);
                pending = false;
            }
            const uint32_t u0 = ut + (uint32_t)r;
            const uint32_t s0 = u0 % PR_RING, s1 = (u0 + 1u) % PR_RING, s2 = (u0 + 2u) % PR_RING;
            const uint32_t e0 = smem_u32(&sm.empty[s0]), e1 = smem_u32(&sm.empty[s1]), e2 = smem_u32(&sm.empty[s2]);
            pr_wait(smem_u32(&sm.full[s0]), (u0 / PR_RING) & 1u);
            pr_wait(smem_u32(&sm.full[s1]), ((u0 + 1u) / PR_RING) & 1u);
            pr_wait(smem_u32(&sm.full[s2]), ((u0 + 2u) / PR_RING) & 1u);
            tc_fence_after();
            if (pr_elect()) {
                if (valid) {
                    const uint64_t a0 = (uint64_t)((ring + s0 * PR_SLOT) >> 4);
                    const uint64_t a1 = (uint64_t)((ring + s1 * PR_SLOT) >> 4);
                    const uint64_t a2 = (uint64_t)((ring + s2 * PR_SLOT) >> 4);
                    tc_mma_tf32(d1, da_row + a0, db_ky0, id64, 0);
                    tc_mma_tf32(d1, da_row + a1, db_ky1, id64, 1);
                    tc_mma_tf32(d1, da_row + a2, db_ky2, id64, 1);
                    // chunk 2 of rows 0 and 1 (physically adjacent slots: the ring's last slot is mirrored behind it)
                    tc_mma_tf32(d1, da_pair + a0 + 2u, db_c2, id64, 1);
                    // chunk 2 of row 2 + the constant chunk that carries b1 (LBO = distance to `ones`)
                    tc_mma_tf32(d1, da_bias + (a2 + 2u) + ((uint64_t)((ones >> 4) - (uint32_t)(a2 + 2u)) << 16), db_bias,
                                id64, 1);
                    tc_commit(smem_u32(&sm.hfull[j]));
                    tc_commit(e0);
                    tc_commit(e1);
                    tc_commit(e2);
                } else {                                   // out of the image: nothing reads the rows, release them
                    mbar_arrive(e0);
                    mbar_arrive(e1);
                    mbar_arrive(e2);
                }
                // readers that do not exist: image row 0 of a unit has one reader, row 1 two; likewise at the end
                if (r == 0) { mbar_arrive(e0); mbar_arrive(e0); mbar_arrive(e1); }
                if (r == nh - 1) { mbar_arrive(e1); mbar_arrive(e2); mbar_arrive(e2); }
            }
            __syncwarp();
            if (valid) { pending = true; pkv = kvr; }
            r = r_next;
        }
        kv += (uint32_t)(r_hi - r_lo);
        ut += (uint32_t)(nh + 2);
    }
    if (pending) pr_conv2(sm, tmem, j, pkv, db2_0, db2_1, id16);
    // a CTA must not exit (and hand its shared memory to the next CTA) while a bulk copy into it or a tcgen05.commit
    // arrival on one of its barriers is still in flight (that was an intermittent "unspecified launch failure"): every
    // image row has been waited for above; the last commit of every issuing warp is waited for here
    if (pr_elect()) tc_commit(smem_u32(&sm.drain));
    __syncwarp();
    pr_wait(smem_u32(&sm.drain), 0);
}

// Column order of Z inside a phase's 16-column block (= row order of that phase's B2): the taps EPI type A needs from
// the phase sit at columns 0.., those of type B at columns 8.., each kx-major (kx, then ky).  -> ky * 3 + kx or -1.
//   type A computes output pixels 0, 1 of a position:  y0 = V0[-1] + V1[0] + V2[1],  y1 = V0[0] + V1[1] + V2[2]
//   type B computes output pixels 2, 3:                y2 = V0[1] + V1[2] + V2[3],   y3 = V0[2] + V1[3] + V2[4]
//   (Vkx[h] = vertical sum of hidden pixel h's contributions with horizontal tap kx; h = -1 / 4: neighbouring position)
__device__ __forceinline__ int pr_tap(int f, int n) {
    int kx = -1, ky = 0;
    if (f == 0)      { if (n < 6) { kx = n / 3; ky = n % 3; } else if (n >= 8 && n < 11) { kx = 2; ky = n - 8; } }
    else if (f == 1) { if (n < 6) { kx = 1 + n / 3; ky = n % 3; } else if (n >= 8 && n < 11) { kx = 0; ky = n - 8; } }
    else if (f == 2) { if (n < 3) { kx = 2; ky = n; } else if (n >= 8 && n < 14) { kx = (n - 8) / 3; ky = (n - 8) % 3; } }
    else             { if (n < 3) { kx = 0; ky = n; } else if (n >= 8 && n < 14) { kx = 1 + (n - 8) / 3; ky = (n - 8) % 3; } }
    return kx < 0 ? -1 : ky * 3 + kx;
}

// ------------------------------------------------------------------ ACT warps: H -> LeakyReLU -> A2 (in place)
// two warps per TMEM lane quarter: type 0 owns phases 0, 1 (columns 0..31 of the row's block), type 1 phases 2, 3
template <bool LEAKY>
__device__ void pr_act(const PairRowsParams& p, PairRowsSmem& sm, uint32_t tmem, int q, int type, int lane) {
    const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16) + 32u * (uint32_t)type;
    const int j = (q & 1) * 32 + lane;                     // position inside the tile
    const uint32_t hbar = smem_u32(&sm.hfull[0]), abar = smem_u32(&sm.a2full[0]);
    const uint32_t alpha_bits = __float_as_uint(p.alpha1);
    uint32_t k = 0;
    for (int w = blockIdx.x; w < p.total; w += gridDim.x) {
        const WorkUnit u = pr_decode(p, w);
        const int x0 = u.S + u.c0 + 4 * j + 2 * type;      // image x of this thread's first hidden pixel
        // hidden pixels outside the image are conv_2's zero padding
        const bool in0 = x0 >= 0 && x0 < p.W, in1 = x0 + 1 >= 0 && x0 + 1 < p.W;
        const bool edge = !(in0 && in1);
        // hidden rows of the unit inside the image: i = y0 - 1 + r in [0, H)
        const int r_lo = u.y0 == 0 ? 1 : 0, r_hi = min(u.rows + 2, p.H - u.y0 + 1);
        for (int r = r_lo; r < r_hi; ++r, ++k) {
            const uint32_t s = k & 3u;
            pr_wait(hbar + 8u * s, (k >> 2) & 1u);
            tc_fence_after();
            uint32_t h[32];
            tc_ld16_nowait(tl + 128u * s, h);
            tc_ld16_nowait(tl + 128u * s + 16u, h + 16);
            tc_wait_ld();
            if (LEAKY) {                                   // max(h, alpha h): packed multiply (FMUL2) + FMNMX
#pragma unroll
                for (int c = 0; c < 32; c += 2) {
                    uint32_t t0, t1;
                    asm("{\n\t.reg .b64 x, y, r;\n\tmov.b64 x, {%2, %3};\n\tmov.b64 y, {%4, %4};\n\t"
                        "mul.rn.f32x2 r, x, y;\n\tmov.b64 {%0, %1}, r;\n\t}"
                        : "=r"(t0), "=r"(t1) : "r"(h[c]), "r"(h[c + 1]), "r"(alpha_bits));
                    h[c] = __float_as_uint(fmaxf(__uint_as_float(h[c]), __uint_as_float(t0)));
                    h[c + 1] = __float_as_uint(fmaxf(__uint_as_float(h[c + 1]), __uint_as_float(t1)));
                }
            }
            if (edge) {                                    // rare: the two threads at the image's left / right border
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    if (!in0) h[c] = 0u;
                    if (!in1) h[16 + c] = 0u;
                }
            }
            tc_st16_nowait(tl + 128u * s, h);
            tc_st16_nowait(tl + 128u * s + 16u, h + 16);
            tc_wait_st();
            tc_fence_before();
            mbar_arrive(abar + 8u * s);
        }
    }
}

// ------------------------------------------------------------------ EPI warps: Z -> shift-add -> act2 -> y
// two warps per TMEM lane quarter (see pr_tap): per thread 6 vertical sums x 3 output rows in flight
struct EpiState {
    float acc[3][6];     // [output row % 3][v]; type A: v = V0[0] V1[0] V1[1] V2[1] V2[2] V0[3]; type B: V2[0] V0[1] V0[2] V1[2] V1[3] V2[3]
};

__device__ __forceinline__ void pr_ld4(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}

struct EpiCtx {
    uint32_t tl;         // TMEM address of this warp's lanes, column 64 (Z) of slot 0
    uint32_t zfull, zempty, xch;     // shared-memory addresses: barriers of slot 0, this warp pair's exchange buffer
    float bias2, alpha2;
    int barrier_id;
    bool send, recv, st0, st1;       // exchange role of this lane; which of the thread's two pixels are outputs
};

__device__ __forceinline__ float pr_sigmoid(float v) {     // 1 / (1 + 2^(-v log2 e)): FMUL, EX2, FADD, RCP
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
    return r;
}

// hidden row r of the unit (ring phase PH = r % 3, compile time).  `valid`: the row lies inside the image and its Z
// tile is in tensor memory; otherwise it contributes nothing.  Completes output row r - 2 of the band (`emit`).
template <int PH, bool SIGMOID, int TYPE>
__device__ __forceinline__ void pr_epi_row(const EpiCtx& c, EpiState& st, uint32_t& k, uint32_t& xcount, bool valid,
                                           bool emit, float* __restrict__ yrow) {
    if (valid) {
        float za[8], zb[8], zc[4], zd[4];                  // z[v][ky] in the order of EpiState::acc
        const uint32_t s = k & 3u;
        pr_wait(c.zfull + 8u * s, (k >> 2) & 1u);
        tc_fence_after();
        const uint32_t zt = c.tl + 128u * s;
        if (TYPE == 0) {
            tc_ld8_nowait(zt + 0u, reinterpret_cast<uint32_t*>(za));          // phase 0: V0[0], V1[0]
            tc_ld8_nowait(zt + 16u, reinterpret_cast<uint32_t*>(zb));         // phase 1: V1[1], V2[1]
            pr_ld4(zt + 32u, zc);                                             // phase 2: V2[2]
            pr_ld4(zt + 48u, zd);                                             // phase 3: V0[3]
        } else {
            pr_ld4(zt + 8u, zc);                                              // phase 0: V2[0]
            pr_ld4(zt + 24u, zd);                                             // phase 1: V0[1]
            tc_ld8_nowait(zt + 40u, reinterpret_cast<uint32_t*>(za));         // phase 2: V0[2], V1[2]
            tc_ld8_nowait(zt + 56u, reinterpret_cast<uint32_t*>(zb));         // phase 3: V1[3], V2[3]
        }
        tc_wait_ld();
        tc_fence_before();
        mbar_arrive(c.zempty + 8u * s);
        ++k;
        // hidden row i feeds output rows i + 1 (ky = 0: opens it), i (ky = 1), i - 1 (ky = 2: completes it)
#define PR_ACC(V, SRC, OFF)                                \
        st.acc[(PH + 2) % 3][V] = SRC[OFF];                \
        st.acc[(PH + 1) % 3][V] += SRC[OFF + 1];           \
        st.acc[PH][V] += SRC[OFF + 2];
        if (TYPE == 0) {
            PR_ACC(0, za, 0) PR_ACC(1, za, 3) PR_ACC(2, zb, 0) PR_ACC(3, zb, 3) PR_ACC(4, zc, 0) PR_ACC(5, zd, 0)
        } else {
            PR_ACC(0, zc, 0) PR_ACC(1, zd, 0) PR_ACC(2, za, 0) PR_ACC(3, za, 3) PR_ACC(4, zb, 0) PR_ACC(5, zb, 3)
        }
#undef PR_ACC
    } else {
#pragma unroll
        for (int v = 0; v < 6; ++v) st.acc[(PH + 2) % 3][v] = 0.f;
    }
    if (!emit) return;                                     // warp-uniform
    // the one value from the neighbouring position: by shuffle, across the warp boundary inside a tile (positions
    // 31 | 32) through shared memory.  Type A needs V0[3] of the position to the left, type B V2[0] of the one to the right.
    const uint32_t slot = c.xch + 4u * (xcount++ & 1u);
    float nb;
    if (TYPE == 0) nb = __shfl_up_sync(0xffffffffu, st.acc[PH][5], 1);
    else nb = __shfl_down_sync(0xffffffffu, st.acc[PH][0], 1);
    if (c.send) asm volatile("st.shared.f32 [%0], %1;" ::"r"(slot), "f"(TYPE == 0 ? st.acc[PH][5] : st.acc[PH][0]) : "memory");
    asm volatile("bar.sync %0, 64;" ::"r"(c.barrier_id) : "memory");
    if (c.recv) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(nb) : "r"(slot) : "memory");
    float v0, v1;
    if (TYPE == 0) {
        v0 = (nb + st.acc[PH][1]) + (st.acc[PH][3] + c.bias2);                 // V0[-1] + V1[0] + V2[1]
        v1 = (st.acc[PH][0] + st.acc[PH][2]) + (st.acc[PH][4] + c.bias2);      // V0[0] + V1[1] + V2[2]
    } else {
        v0 = (st.acc[PH][1] + st.acc[PH][3]) + (st.acc[PH][5] + c.bias2);      // V0[1] + V1[2] + V2[3]
        v1 = (st.acc[PH][2] + st.acc[PH][4]) + (nb + c.bias2);                 // V0[2] + V1[3] + V2[4]
    }
    if (SIGMOID) {
        v0 = pr_sigmoid(v0);
        v1 = pr_sigmoid(v1);
    } else if (c.alpha2 != 1.f) {
        v0 = v0 >= 0.f ? v0 : v0 * c.alpha2;
        v1 = v1 >= 0.f ? v1 : v1 * c.alpha2;
    }
    if (c.st0) yrow[0] = v0;
    if (c.st1) yrow[1] = v1;
}

template <bool SIGMOID, int TYPE>
__device__ void pr_epi(const PairRowsParams& p, PairRowsSmem& sm, uint32_t tmem, int q, int lane) {
    const int tile = q >> 1;
    const int j = (q & 1) * 32 + lane;
    EpiCtx c;
    c.tl = tmem + ((uint32_t)(q * 32) << 16) + 64u;
    c.zfull = smem_u32(&sm.zfull[0]);
    c.zempty = smem_u32(&sm.zempty[0]);
    c.xch = smem_u32(&sm.xch[TYPE][tile][0]);
    c.bias2 = __ldg(p.b2);
    c.alpha2 = p.alpha2;
    c.barrier_id = 1 + 2 * TYPE + tile;
    // type A: the left warp's lane 31 sends V0[3] to the right warp's lane 0; type B: the other way round with V2[0]
    c.send = TYPE == 0 ? ((q & 1) == 0 && lane == 31) : ((q & 1) == 1 && lane == 0);
    c.recv = TYPE == 0 ? ((q & 1) == 1 && lane == 0) : ((q & 1) == 0 && lane == 31);
    uint32_t k = 0, xcount = 0;
    for (int w = blockIdx.x; w < p.total; w += gridDim.x) {
        const WorkUnit u = pr_decode(p, w);
        const int n = tile ? u.n1 : u.n0;
        const int x0 = u.S + u.c0 + 4 * j + 2 * TYPE;      // image x of this thread's first output pixel
        // hidden pixel hp of the tile is an output iff 1 <= hp <= PR_OUT, inside the image, real image
        const int hp = 4 * j + 2 * TYPE;
        c.st0 = n < p.N && hp >= 1 && hp <= PR_OUT && x0 < p.W;
        c.st1 = n < p.N && hp + 1 >= 1 && hp + 1 <= PR_OUT && x0 + 1 < p.W;
        EpiState st;
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int v = 0; v < 6; ++v) st.acc[a][v] = 0.f;
        const int nh = u.rows + 2;
        const int r_lo = u.y0 == 0 ? 1 : 0, r_hi = min(nh, p.H - u.y0 + 1);   // hidden rows inside the image
        // &y[n, y0 - 2, x0]: row r completes output row y0 + r - 2 (never dereferenced where st0 / st1 are false)
        float* yrow = p.y + ((int64_t)(n < p.N ? n : 0) * p.H + (u.y0 - 2)) * (int64_t)p.W + x0;
#define PR_ROW(PH, R)                                                                                              \
        pr_epi_row<PH, SIGMOID, TYPE>(c, st, k, xcount, (R) >= r_lo && (R) < r_hi, (R) >= 2, yrow);                \
        yrow += p.W;
        for (int r = 0; r < nh; r += 3) {
            PR_ROW(0, r)
            if (r + 1 < nh) { PR_ROW(1, r + 1) }
            if (r + 2 < nh) { PR_ROW(2, r + 2) }
        }
#undef PR_ROW
    }
}

template <bool LEAKY, bool SIGMOID>
__global__ void __launch_bounds__(PR_THREADS, 1) conv3x3_pair_rows_kernel(const __grid_constant__ CUtensorMap map_x,
                                                                          const PairRowsParams p) {
    __shared__ PairRowsSmem sm;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);

    // ---- operands: B1 (two c0 variants), B2, the constant chunk
    for (int i = tid; i < 2 * 10 * 64 * 4; i += PR_THREADS) {
        const int var = i / 2560, rest = i % 2560;
        const int kc = rest / 256, n = (rest % 256) >> 2, e = rest & 3;        // chunk, row n = (phase, c), element
        const int c0 = var ? 5 : 3, f = n >> 4, c = n & 15;
        float v = 0.f;
        if (kc < 9) {
            const int ky = kc / 3, fr = 4 * (kc % 3) + e;                       // float 4 m + fr of image row ky
            const int kx = fr - (c0 + f) + 1;
            if (kx >= 0 && kx < 3) v = __ldg(p.w1 + (ky * 3 + kx) * 16 + c);
        } else if (e == 0) {
            v = __ldg(p.b1 + c);
        }
        // scaled so that the tensor core's truncation of H to TF32 (GEMM 2's A operand) is unbiased
        sm.b1[var][kc * 256 + n * 4 + e] = round_tf32(v * (1.f + 1.f / 2048.f));
    }
    for (int i = tid; i < 4 * 256; i += PR_THREADS) {
        const int f = i >> 8, rest = i & 255;
        const int kq = rest >> 6, n = (rest & 63) >> 2, kk = rest & 3;
        const int tap = pr_tap(f, n);
        sm.b2[f][rest] = tap >= 0 ? round_tf32(__ldg(p.w2 + tap * 16 + kq * 4 + kk)) : 0.f;
    }
    for (int i = tid; i < 512; i += PR_THREADS) sm.ones[i] = (i & 3) == 0 ? 1.f : 0.f;
    if (tid == 0) {
        for (int s = 0; s < PR_RING; ++s) {
            mbar_init(smem_u32(&sm.full[s]), 1);
            mbar_init(smem_u32(&sm.empty[s]), 3);           // the three hidden rows that read an image row
        }
        for (int s = 0; s < PR_TSLOTS; ++s) {
            mbar_init(smem_u32(&sm.hfull[s]), 1);
            mbar_init(smem_u32(&sm.a2full[s]), 256);
            mbar_init(smem_u32(&sm.zfull[s]), 1);
            mbar_init(smem_u32(&sm.zempty[s]), 256);
        }
        mbar_init(smem_u32(&sm.drain), PR_TSLOTS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // operands are read by the async proxy
    if (warp == 17) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(&sm.tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem;

    if (warp == 16) {
        pr_producer(p, &map_x, sm);
    } else if (warp >= 17) {
        pr_mma(p, sm, tmem, warp - 17);
    } else if (warp < 8) {
        pr_act<LEAKY>(p, sm, tmem, warp & 3, warp >> 2, lane);
    } else if (warp < 12) {
        pr_epi<SIGMOID, 0>(p, sm, tmem, warp & 3, lane);
    } else {
        pr_epi<SIGMOID, 1>(p, sm, tmem, warp & 3, lane);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 17) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

}  // namespace

int conv3x3_pair_rows(const float* x, const float* w1, const float* b1, const float* w2, const float* b2, float* y,
                      int64_t n, int64_t h, int64_t w, int c1, int act1, float alpha1, int act2, float alpha2,
                      cudaStream_t st) {
    if (c1 != 16 || w % 4 || w < 8 || n > 0x3fffffff) return UOCR_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(x) & 15) != 0) return UOCR_ERR_UNSUPPORTED;
    const bool leaky = act1 == UOCR_ACT_LEAKY;
    if (!(act1 == UOCR_ACT_NONE || (leaky && alpha1 >= 0.f && alpha1 <= 1.f))) return UOCR_ERR_UNSUPPORTED;
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    PairRowsParams p{};
    p.w1 = w1; p.b1 = b1; p.w2 = w2; p.b2 = b2; p.y = y;
    p.N = (int)n; p.H = (int)h; p.W = (int)w;
    p.xt = (int)ceil_div(w, PR_OUT);
    p.npair = (int)ceil_div(n, 2);
    // bands: every unit costs rows + 2 hidden rows; choose the band count that minimises waves x (rows + 2)
    int best_nb = 1;
    double best_cost = 1e300;
    for (int nb = 1; nb <= 32 && nb <= h; ++nb) {
        const int64_t rb = ceil_div(h, nb);
        const int64_t units = (int64_t)ceil_div(h, rb) * p.xt * p.npair;
        const double cost = (double)ceil_div(units, sms) * (double)(rb + 2 + 6);      // + 6: pipeline fill per unit
        if (cost < best_cost) { best_cost = cost; best_nb = nb; }
    }
    p.rb = (int)ceil_div(h, best_nb);
    p.nb = (int)ceil_div(h, p.rb);
    const int64_t total = (int64_t)p.nb * p.xt * p.npair;
    if (total > 0x7fffffff) return UOCR_ERR_UNSUPPORTED;
    p.total = (int)total;
    p.alpha1 = alpha1;
    p.alpha2 = act2 == UOCR_ACT_LEAKY ? alpha2 : 1.f;
    CUtensorMap map;
    const uint64_t dims[3] = {(uint64_t)w, (uint64_t)h, (uint64_t)n};
    const uint64_t strides[2] = {(uint64_t)w * 4, (uint64_t)w * h * 4};
    const uint32_t box[3] = {256, 1, 1};
    int rc = make_tmap_plain_tf32(&map, x, 3, dims, strides, box);
    if (rc) return rc;
    unsigned grid = (unsigned)(total < sms ? total : sms);
    // tests: fewer persistent CTAs than work units (UOCR_PAIR_ROWS_GRID), to exercise the multi-unit bookkeeping on small inputs
    { const char* e = getenv("UOCR_PAIR_ROWS_GRID"); if (e && atoi(e) > 0 && (unsigned)atoi(e) < grid) grid = (unsigned)atoi(e); }
    if (act2 == UOCR_ACT_SIGMOID) {
        if (leaky) conv3x3_pair_rows_kernel<true, true><<<grid, PR_THREADS, 0, st>>>(map, p);
        else conv3x3_pair_rows_kernel<false, true><<<grid, PR_THREADS, 0, st>>>(map, p);
    } else {
        if (leaky) conv3x3_pair_rows_kernel<true, false><<<grid, PR_THREADS, 0, st>>>(map, p);
        else conv3x3_pair_rows_kernel<false, false><<<grid, PR_THREADS, 0, st>>>(map, p);
    }
    UOCR_LAUNCHED("conv3x3_pair_rows");
    return UOCR_OK;
}

}  // namespace uocr
