// Probe for the design of the tensor-core Monochrome pair kernel (not part of libuocr):
//   1. tcgen05.ld / tcgen05.st throughput per SM (the microarchitecture notes and the programming guide disagree
//      by two orders of magnitude),
//   2. tcgen05.mma with the A operand read from TENSOR MEMORY (written there by tcgen05.st) -- correctness against
//      a host product and the round-trip latency st -> mbarrier -> mma -> commit -> ld.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_probe tools/tmem_probe.cu && ./tmem_probe
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                   "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t tmem_alloc(uint32_t* slot, uint32_t cols, int warp) {
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    return *slot;
}
__device__ __forceinline__ void tmem_free(uint32_t base, uint32_t cols, int warp) {
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols) : "memory");
    }
}

// MODE 0: ld x16 + wait each; 1: two ld x16 then wait; 2: st x16 + wait each; 3: two st then wait
template <int MODE>
__global__ void __launch_bounds__(128) bw_kernel(int iters, uint32_t* sink, long long* cycles) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    const uint32_t base = tmem_alloc(&slot, 64, warp);
    const uint32_t t0 = base + ((uint32_t)(warp * 32) << 16);
    uint32_t r[16], q[16], acc = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) { r[i] = threadIdx.x + i; q[i] = i; }
    tc_st16(t0, r); tc_st16(t0 + 16, r); tc_st16(t0 + 32, r); tc_st16(t0 + 48, r);
    tc_wait_st();
    __syncthreads();
    const long long c0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) { tc_ld16(t0 + (it & 3) * 16, r); tc_wait_ld(); acc ^= r[it & 15]; }
        if (MODE == 1) { tc_ld16(t0, r); tc_ld16(t0 + 16, q); tc_wait_ld(); acc ^= r[it & 15] ^ q[(it + 1) & 15]; }
        if (MODE == 2) { r[0] = it; tc_st16(t0 + (it & 3) * 16, r); tc_wait_st(); }
        if (MODE == 3) { r[0] = it; tc_st16(t0, r); tc_st16(t0 + 16, r); tc_wait_st(); }
    }
    __syncthreads();
    const long long c1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = c1 - c0;
    if (acc == 0x12345678) sink[0] = acc;
    tmem_free(base, 64, warp);
}

__device__ __forceinline__ uint64_t make_kmajor_nosw_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(lbo_bytes >> 4) << 16;
    d |= (uint64_t)(sbo_bytes >> 4) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// D[128 x 16] = A[128 x 16] (TMEM, written with tcgen05.st) . B[16 (n) x 16 (k)]^T (smem, no swizzle), repeated
// `iters` times through the full producer/consumer chain; reports cycles per round trip.
__global__ void __launch_bounds__(160) mma_ts_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                     float* __restrict__ D, int iters, long long* cycles) {
    __shared__ __align__(128) float s_b[4 * 64];      // chunk kq: 16 rows (n) x 4 floats (k = 4 kq ..)
    __shared__ uint64_t bars[2];
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 256; i += 160) {
        const int kq = i >> 6, n = (i & 63) >> 2, kk = i & 3;
        s_b[i] = B[n * 16 + kq * 4 + kk];
    }
    if (tid == 0) {
        mbar_init(smem_u32(&bars[0]), 128);     // A ready: every compute thread arrives
        mbar_init(smem_u32(&bars[1]), 1);       // D ready: tcgen05.commit
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const uint32_t base = tmem_alloc(&slot, 64, warp);
    const long long c0 = clock64();
    if (warp < 4) {
        const uint32_t t0 = base + ((uint32_t)(warp * 32) << 16);
        uint32_t a[16], d[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) a[k] = __float_as_uint(A[tid * 16 + k]);
        for (int it = 0; it < iters; ++it) {
            tc_st16(t0, a);
            tc_wait_st();
            tc_fence_before();
            mbar_arrive(smem_u32(&bars[0]));
            mbar_wait(smem_u32(&bars[1]), it & 1);
            tc_fence_after();
            tc_ld16(t0 + 32, d);
            tc_wait_ld();
        }
#pragma unroll
        for (int n = 0; n < 16; ++n) D[tid * 16 + n] = __uint_as_float(d[n]);
    } else if (lane == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t sb = smem_u32(s_b);
        for (int it = 0; it < iters; ++it) {
            mbar_wait(smem_u32(&bars[0]), it & 1);
            tc_fence_after();
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                const uint64_t db = make_kmajor_nosw_desc(sb + (uint32_t)(2 * m) * 256u, 256, 128);
                tc_mma_tf32_ts(base + 32, base + (uint32_t)(m * 8), db, idesc, m > 0 ? 1u : 0u);
            }
            tc_commit(smem_u32(&bars[1]));
        }
    }
    __syncthreads();
    const long long c1 = clock64();
    if (tid == 0) cycles[0] = c1 - c0;
    tmem_free(base, 64, warp);
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("device %s, %d SMs, clock %d kHz\n", prop.name, sms, prop.clockRate);
    uint32_t* sink; long long* cyc;
    CK(cudaMalloc(&sink, 64)); CK(cudaMalloc(&cyc, 64));
    const int iters = 20000;
    for (int mode = 0; mode < 4; ++mode) {
        for (int per_sm : {1, 2, 4, 8}) {
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                if (mode == 0) bw_kernel<0><<<sms * per_sm, 128>>>(iters, sink, cyc);
                if (mode == 1) bw_kernel<1><<<sms * per_sm, 128>>>(iters, sink, cyc);
                if (mode == 2) bw_kernel<2><<<sms * per_sm, 128>>>(iters, sink, cyc);
                if (mode == 3) bw_kernel<3><<<sms * per_sm, 128>>>(iters, sink, cyc);
                cudaEventRecord(e1);
                CK(cudaEventSynchronize(e1));
            }
            long long c; CK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
            const double per_it = (mode & 1) ? 2.0 : 1.0;
            const double bytes_per_sm = (double)iters * per_it * 128 * 64 * per_sm;   // 128 lanes x 16 cols x 4 B
            printf("%s x16%s, %d CTA/SM (x4 warps): %.1f cycles/iter/CTA, %.1f B/clk/SM\n", mode < 2 ? "ld" : "st",
                   (mode & 1) ? " x2 in flight" : "", per_sm, (double)c / iters, bytes_per_sm / (double)c);
        }
    }
    // ---- MMA with A from TMEM
    std::vector<float> hA(128 * 16), hB(16 * 16), hD(128 * 16), ref(128 * 16);
    auto tf32 = [](float v) { uint32_t u; memcpy(&u, &v, 4); u = (u + 0x1000) & 0xffffe000u; float r; memcpy(&r, &u, 4); return r; };
    srand(1);
    for (auto& v : hA) v = tf32((float)rand() / RAND_MAX - 0.5f);
    for (auto& v : hB) v = tf32((float)rand() / RAND_MAX - 0.5f);
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 16; ++n) {
            double s = 0;
            for (int k = 0; k < 16; ++k) s += (double)hA[m * 16 + k] * hB[n * 16 + k];
            ref[m * 16 + n] = (float)s;
        }
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, hA.size() * 4)); CK(cudaMalloc(&dB, hB.size() * 4)); CK(cudaMalloc(&dD, hD.size() * 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice));
    for (int it : {1, 2, 1000}) {
        CK(cudaMemset(dD, 0, hD.size() * 4));
        mma_ts_kernel<<<1, 160>>>(dA, dB, dD, it, cyc);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
        long long c; CK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
        double err = 0;
        for (size_t i = 0; i < hD.size(); ++i) err = fmax(err, fabs((double)hD[i] - ref[i]));
        printf("mma A-from-TMEM iters=%d: max |err| = %.3e (D[0]=%f ref %f, D[17]=%f ref %f), %.1f cycles per round trip\n",
               it, err, hD[0], ref[0], hD[17], ref[17], (double)c / it);
    }
    return 0;
}
