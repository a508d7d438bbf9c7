// Correctness probe for the planned tcgen05 weight-gradient GEMM: one tcgen05.mma (M = 128, N = 32,
// K = 16, bf16, fp32 accumulate) with BOTH operands MN-major, no swizzle, laid out the way a staged activation-plane tile
// is: for every group of 8 M (or N) elements, consecutive K rows are consecutive 16-byte rows (8 K rows = one 128-byte core
// matrix).  Tries both assignments of (leading, stride) byte offsets and prints which one reproduces A * B [verified on
// B200: LBO = 8-row step, SBO = group step], and -- the open point of the planned weight-gradient kernel, NOT YET RUN -- an
// A start address shifted by one 16-byte K row (A is stored with 24 K rows per group so that rows 1..16 exist).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_mn_test umma_mn_test.cu && ./umma_mn_test
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWL:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra WD;\n\tbra WL;\n\tWD:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// kind::f16: D fp32 (bit 4), A / B bf16 (bits 7, 10), A / B MN-major (bits 15, 16), N >> 3 at 17, M >> 4 at 24
__device__ __forceinline__ uint32_t make_idesc_mn(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}

constexpr int M = 128, N = 32, K = 16, KA = 24;  // KA: K rows stored per A group (rows beyond 16 are only used by the shifted probe)

// A is [M][KA]; the MMA reads K rows shift .. shift + 15
__global__ void __launch_bounds__(128, 1) probe(const __nv_bfloat16 *A, const __nv_bfloat16 *B, float *D, int lbo, int sbo_a, int sbo_b,
                                                int swap, int shift) {
    __shared__ __align__(1024) uint8_t sA[M / 8 * KA * 16];
    __shared__ __align__(1024) uint8_t sB[N / 8 * 256];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x;
    for (int i = tid; i < M * KA; i += 128) {  // A[m][k] -> group m / 8, K row k, element m % 8
        const int m = i / KA, k = i % KA;
        *reinterpret_cast<__nv_bfloat16 *>(sA + (m / 8) * KA * 16 + k * 16 + (m % 8) * 2) = A[i];
    }
    for (int i = tid; i < K * N; i += 128) {  // B[k][n] -> group n / 8, K row k, element n % 8
        const int k = i / N, n = i % N;
        *reinterpret_cast<__nv_bfloat16 *>(sB + (n / 8) * 256 + k * 16 + (n % 8) * 2) = B[i];
    }
    if (tid == 0) {
        mbar_init(smem_u32(&bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(32u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if (tid == 0) {
        const uint64_t da = swap ? make_desc(smem_u32(sA) + 16 * shift, sbo_a, lbo) : make_desc(smem_u32(sA) + 16 * shift, lbo, sbo_a);
        const uint64_t db = swap ? make_desc(smem_u32(sB), sbo_b, lbo) : make_desc(smem_u32(sB), lbo, sbo_b);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 1;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem), "l"(da),
                     "l"(db), "r"(make_idesc_mn(N))
                     : "memory");  // predicate false = overwrite the accumulator
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    mbar_wait(smem_u32(&bar), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int warp = tid >> 5, lane = tid & 31;
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                       "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; ++j) D[(warp * 32 + lane) * N + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32u) : "memory");
}

int main() {
    static __nv_bfloat16 hA[M * KA], hB[K * N];
    static float ref[M * N], hD[M * N];
    srand(1);
    for (int i = 0; i < M * KA; ++i) hA[i] = __float2bfloat16((float)(rand() % 17 - 8) / 8.f);
    for (int i = 0; i < K * N; ++i) hB[i] = __float2bfloat16((float)(rand() % 13 - 6) / 4.f);
    __nv_bfloat16 *dA, *dB;
    float *dD;
    cudaMalloc(&dA, sizeof(hA));
    cudaMalloc(&dB, sizeof(hB));
    cudaMalloc(&dD, sizeof(hD));
    cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
    struct {
        const char *name;
        int swap, shift;
    } variants[] = {{"LBO = 8-row step (128), SBO = group step", 0, 0},
                    {"LBO and SBO swapped", 1, 0},
                    {"LBO = 128, SBO = group step, A start + 16 bytes (K rows 1..16)", 0, 1},
                    {"LBO = 128, SBO = group step, A start + 48 bytes (K rows 3..18)", 0, 3}};
    for (auto &v : variants) {
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < N; ++n) {
                float s = 0.f;
                for (int k = 0; k < K; ++k) s += __bfloat162float(hA[m * KA + k + v.shift]) * __bfloat162float(hB[k * N + n]);
                ref[m * N + n] = s;
            }
        cudaMemset(dD, 0xff, sizeof(hD));
        probe<<<1, 128>>>(dA, dB, dD, 128, KA * 16, 256, v.swap, v.shift);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost);
        double mx = 0.0;
        for (int i = 0; i < M * N; ++i) mx = fmax(mx, fabs((double)hD[i] - ref[i]));
        printf("MN-major A and B, no swizzle, %s: %s, max |D - A*B| = %g %s\n", v.name, cudaGetErrorString(e), mx, mx < 1e-3 ? "<-- matches" : "");
        if (e != cudaSuccess) break;
    }
    return 0;
}
