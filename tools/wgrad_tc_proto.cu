// Prototype of the inner loop of the planned tcgen05 weight-gradient GEMM, self-checking.
// STATUS: verified on B200 (profiles/r1_d_wgrad_tc_proto.log: relative error 1.4e-5 against a float64 sum).  It extends
// tools/umma_mn_test.cu (one MMA, MN-major operand recipe) to the real loop shape:
//   dW[ci][co] = sum_p E[ci][p] * dY[co][p],   M = 128 input channels, N = 32 output channels, K = P pixels,
// operands read from "planes" exactly as the library stores activations / dL/dy: for every 8-channel group, pixel p is
// the 16-byte row p (bf16 hi plane set and bf16 lo plane set), three MMAs per product (hi*hi + hi*lo + lo*hi, fp32
// accumulate in TMEM).  A stage of S = 128 pixels is copied group by group into shared memory (the box TMA would deliver:
// group stride S * 16 bytes = SBO, 8-pixel core matrices 128 bytes apart = LBO) and consumed by 8 K-steps of 16 pixels,
// the descriptor start address advancing by 256 bytes per step.  One CTA; the product kernel adds TMA, a stage ring,
// the per-sample rstd / shift epilogue and the tap coordinate shift.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o wgrad_tc_proto wgrad_tc_proto.cu && ./wgrad_tc_proto
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
__device__ __forceinline__ uint32_t make_idesc_mn(int n) {  // kind::f16, D fp32, A / B bf16, both MN-major (bits 15, 16)
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b),
                 "r"(idesc), "r"(acc)
                 : "memory");
}

constexpr int CI = 128, CO = 32, S = 128;  // channels of the tile, pixels per stage

// planes: [hi | lo][C / 8][P][8] bf16
__global__ void __launch_bounds__(128, 1) proto(const __nv_bfloat16 *E, const __nv_bfloat16 *DY, float *D, int P) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *sEh = smem, *sEl = sEh + CI / 8 * S * 16, *sYh = sEl + CI / 8 * S * 16, *sYl = sYh + CO / 8 * S * 16;
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(smem_u32(&bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(32u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    const uint32_t idesc = make_idesc_mn(CO);
    const uint4 *gE = reinterpret_cast<const uint4 *>(E), *gY = reinterpret_cast<const uint4 *>(DY);
    const size_t e_lo = (size_t)CI / 8 * P, y_lo = (size_t)CO / 8 * P;  // uint4 (pixel) offset of the lo plane set
    uint32_t phase = 0, first = 1;
    for (int p0 = 0; p0 < P; p0 += S) {
        // stage copy: group g, pixel p0 + r -> shared row r of group g (what one TMA box {8 ch, S pixels, groups} delivers)
        for (int i = tid; i < CI / 8 * S; i += 128) {
            const int g = i / S, r = i - g * S;
            const bool ok = p0 + r < P;
            const uint4 z = make_uint4(0u, 0u, 0u, 0u);
            reinterpret_cast<uint4 *>(sEh)[i] = ok ? gE[(size_t)g * P + p0 + r] : z;
            reinterpret_cast<uint4 *>(sEl)[i] = ok ? gE[e_lo + (size_t)g * P + p0 + r] : z;
        }
        for (int i = tid; i < CO / 8 * S; i += 128) {
            const int g = i / S, r = i - g * S;
            const bool ok = p0 + r < P;
            const uint4 z = make_uint4(0u, 0u, 0u, 0u);
            reinterpret_cast<uint4 *>(sYh)[i] = ok ? gY[(size_t)g * P + p0 + r] : z;
            reinterpret_cast<uint4 *>(sYl)[i] = ok ? gY[y_lo + (size_t)g * P + p0 + r] : z;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor core
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int ks = 0; ks < S / 16; ++ks) {
                const uint32_t off = ks * 256;  // 16 pixels further along K
                const uint64_t aH = make_desc(smem_u32(sEh) + off, 128, S * 16), aL = make_desc(smem_u32(sEl) + off, 128, S * 16);
                const uint64_t bH = make_desc(smem_u32(sYh) + off, 128, S * 16), bL = make_desc(smem_u32(sYl) + off, 128, S * 16);
                umma(tmem, aH, bH, idesc, first ? 0u : 1u);
                first = 0;
                umma(tmem, aH, bL, idesc, 1u);
                umma(tmem, aL, bH, idesc, 1u);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
        mbar_wait(smem_u32(&bar), phase);  // single buffer: the stage may be overwritten once its MMAs have completed
        phase ^= 1;
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int warp = tid >> 5, lane = tid & 31;
    for (int c0 = 0; c0 < CO; c0 += 16) {
        uint32_t v[16];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                       "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; ++j) D[(warp * 32 + lane) * CO + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32u) : "memory");
}

static void to_planes(const float *x, int C, int P, __nv_bfloat16 *out) {  // x [C][P] -> [hi | lo][C / 8][P][8]
    for (int c = 0; c < C; ++c)
        for (int p = 0; p < P; ++p) {
            const float v = x[(size_t)c * P + p];
            const __nv_bfloat16 h = __float2bfloat16(v);
            const __nv_bfloat16 l = __float2bfloat16(v - __bfloat162float(h));
            const size_t o = ((size_t)(c / 8) * P + p) * 8 + c % 8;
            out[o] = h;
            out[(size_t)C * P + o] = l;
        }
}

int main() {
    const int P = 4000;  // not a multiple of the stage: the tail stage is zero filled
    float *x = (float *)malloc(sizeof(float) * CI * P), *y = (float *)malloc(sizeof(float) * CO * P);
    srand(3);
    for (int i = 0; i < CI * P; ++i) x[i] = (float)rand() / RAND_MAX - 0.5f;
    for (int i = 0; i < CO * P; ++i) y[i] = (float)rand() / RAND_MAX - 0.5f;
    __nv_bfloat16 *hE = (__nv_bfloat16 *)malloc(2 * sizeof(__nv_bfloat16) * CI * P), *hY = (__nv_bfloat16 *)malloc(2 * sizeof(__nv_bfloat16) * CO * P);
    to_planes(x, CI, P, hE);
    to_planes(y, CO, P, hY);
    double *ref = (double *)malloc(sizeof(double) * CI * CO), nrm = 0.0;
    for (int ci = 0; ci < CI; ++ci)
        for (int co = 0; co < CO; ++co) {
            double s = 0.0;
            for (int p = 0; p < P; ++p) s += (double)x[(size_t)ci * P + p] * (double)y[(size_t)co * P + p];
            ref[ci * CO + co] = s;
            nrm += s * s;
        }
    __nv_bfloat16 *dE, *dY;
    float *dD, *hD = (float *)malloc(sizeof(float) * CI * CO);
    cudaMalloc(&dE, 2 * sizeof(__nv_bfloat16) * CI * P);
    cudaMalloc(&dY, 2 * sizeof(__nv_bfloat16) * CO * P);
    cudaMalloc(&dD, sizeof(float) * CI * CO);
    cudaMemcpy(dE, hE, 2 * sizeof(__nv_bfloat16) * CI * P, cudaMemcpyHostToDevice);
    cudaMemcpy(dY, hY, 2 * sizeof(__nv_bfloat16) * CO * P, cudaMemcpyHostToDevice);
    const int smem = 2 * CI / 8 * S * 16 + 2 * CO / 8 * S * 16;
    cudaFuncSetAttribute(proto, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    proto<<<1, 128, smem>>>(dE, dY, dD, P);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(hD, dD, sizeof(float) * CI * CO, cudaMemcpyDeviceToHost);
    double err = 0.0;
    for (int i = 0; i < CI * CO; ++i) err += ((double)hD[i] - ref[i]) * ((double)hD[i] - ref[i]);
    printf("wgrad_tc_proto: %s, ||dW - ref|| / ||ref|| = %.3e (bf16 hi/lo split, 3 MMAs per product: expect ~1e-5) %s\n", cudaGetErrorString(e),
           sqrt(err / nrm), sqrt(err / nrm) < 1e-4 ? "OK" : "MISMATCH");
    return 0;
}
