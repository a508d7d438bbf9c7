// Probe + micro-benchmark of the CTA-pair form of tcgen05.mma (cta_group::2, M = 256, K = 16, bf16, no-swizzle
// K-major operands in the conv_rs layout): (1) operand placement check -- CTA r of the pair holds rows
// [128 r, 128 r + 128) of A and rows [N/2 r, N/2 r + N/2) of B in ITS shared memory at the same offsets, D rows
// [128 r, +128) land in ITS tensor memory -- against a CPU product; (2) issue cost per MMA as a function of N
// (the single-CTA form costs 32 + N/4 cycles for N <= 128: tools/umma_bench.cu; the pair form should read
// A (4 KB) + half of B per SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_2cta_test umma_2cta_test.cu && ./umma_2cta_test
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWL:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra WD;\n\tbra WL;\n\tWD:\n\t}" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// kind::f16 instruction descriptor: D fp32, A/B bf16, K-major, N >> 3 at bit 17, M >> 4 at bit 24 (M = 256 for the pair)
__device__ __forceinline__ uint32_t make_idesc(int n, int m) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma2(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
                 "l"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void commit2(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
                 : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// A: [256][16] bf16 row-major (global), B: [N][16] bf16 row-major, D: [256][N] fp32.
// shared layout (both operands): [kg = 2][rows][8 elements] -> LBO = rows * 16 B, SBO = 128 B
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
pair_kernel(const __nv_bfloat16 *A, const __nv_bfloat16 *B, float *D, int N, int iters, long long *cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_rank();
    const int NH = N / 2;
    __nv_bfloat16 *sA = reinterpret_cast<__nv_bfloat16 *>(smem);           // [2][128][8]
    __nv_bfloat16 *sB = reinterpret_cast<__nv_bfloat16 *>(smem + 8192);    // [2][NH][8]
    for (int i = tid; i < 2 * 128 * 8; i += 128) {
        const int e = i & 7, r = (i >> 3) & 127, kg = i >> 10;
        sA[i] = A[(size_t)(rank * 128 + r) * 16 + kg * 8 + e];
    }
    for (int i = tid; i < 2 * NH * 8; i += 128) {
        const int e = i & 7, r = (i >> 3) % NH, kg = (i >> 3) / NH;
        sB[i] = B[(size_t)(rank * NH + r) * 16 + kg * 8 + e];
    }
    if (tid == 0) {
        mbar_init(smem_u32(&bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    const uint64_t ad = make_desc(smem_u32(sA), 128 * 16, 128), bd = make_desc(smem_u32(sB), NH * 16, 128);
    const uint32_t idesc = make_idesc(N, 256);
    long long dt = 0;
    for (int rep = 0; rep < 2; ++rep) {
        const long long t0 = clock64();
        if (rank == 0 && warp == 0) {
            if (elect_one()) {
                const int nacc = 512 / N;
                for (int i = 0; i < iters; i += 4) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) umma2(tmem + (uint32_t)(((i + u) % nacc) * N), ad, bd, idesc, (i + u) >= nacc ? 1u : 0u);
                }
                commit2(smem_u32(&bar), 3);
            }
            __syncwarp();
        }
        mbar_wait(smem_u32(&bar), rep & 1);
        dt = clock64() - t0;
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // read accumulator 0: warp w reads lanes [32 w, +32); every accumulator got iters / nacc identical products
    for (int cb = 0; cb < N; cb += 16) {
        uint32_t v[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
              "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)cb));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (blockIdx.x < 2)
            for (int q = 0; q < 16; ++q) D[(size_t)(rank * 128 + warp * 32 + lane) * N + cb + q] = __uint_as_float(v[q]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    if (tid == 0) cycles[blockIdx.x] = dt;
}

int main() {
    const int Ns[] = {32, 48, 64, 80, 96, 128, 192, 256};
    __nv_bfloat16 *dA, *dB;
    float *dD;
    long long *dc;
    cudaMalloc(&dA, 256 * 16 * 2);
    cudaMalloc(&dB, 256 * 16 * 2);
    cudaMalloc(&dD, 256 * 256 * 4);
    cudaMalloc(&dc, 148 * 8);
    static float hA[256 * 16], hB[256 * 16];
    static __nv_bfloat16 bA[256 * 16], bB[256 * 16];
    srand(1);
    for (int i = 0; i < 256 * 16; ++i) {
        bA[i] = __float2bfloat16((float)(rand() % 17 - 8) / 8.f);
        bB[i] = __float2bfloat16((float)(rand() % 13 - 6) / 4.f);
        hA[i] = __bfloat162float(bA[i]);
        hB[i] = __bfloat162float(bB[i]);
    }
    cudaMemcpy(dA, bA, sizeof(bA), cudaMemcpyHostToDevice);
    cudaMemcpy(dB, bB, sizeof(bB), cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int N : Ns) {
        const int iters = 512;
        const int nacc = 512 / N;
        cudaMemset(dD, 0, 256 * 256 * 4);
        pair_kernel<<<148, 128, 32 * 1024>>>(dA, dB, dD, N, iters, dc);
        cudaError_t e = cudaDeviceSynchronize();
        static float hD[256 * 256];
        long long hc[148];
        cudaMemcpy(hD, dD, 256 * N * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(hc, dc, sizeof(hc), cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < 148; ++i) mx = hc[i] > mx ? hc[i] : mx;
        // accumulator 0 received the products of iterations 0, nacc, 2 nacc, ... (the first overwrites): count of them, two reps
        int cnt = 0;
        for (int i = 0; i < iters; ++i)
            if (i % nacc == 0) ++cnt;
        double maxerr = 0.0;
        for (int m = 0; m < 256; ++m)
            for (int n = 0; n < N; ++n) {
                double ref = 0.0;
                for (int k = 0; k < 16; ++k) ref += (double)hA[m * 16 + k] * hB[n * 16 + k];
                maxerr = fmax(maxerr, fabs(hD[m * N + n] - ref * cnt));
            }
        printf("pair M=256 N=%3d: %s  %.1f cycles/MMA  (%.0f MAC/cyc/SM)  max |D - ref| = %.3g\n", N, cudaGetErrorString(e), (double)mx / iters,
               128.0 * N * 16 * iters / (double)mx, maxerr);
    }
    return 0;
}
