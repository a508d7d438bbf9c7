// Micro-benchmark 2: the exact MMA issue pattern of conv_tc.cu (bf16x3 merged-N): per tap, G wide MMAs
// (A_hi x [W_hi|W_lo], width 2N) then G narrow ones (A_lo x W_hi, width N) into G accumulators, 9 taps with
// different A shifts and B tiles, repeated.  Timing only.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWL:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra WD;\n\tbra WL;\n\tWD:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ uint32_t make_idesc(int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24); }
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
struct Cfg { int N, G, x3, Wr, reps, same_shape, commit_every; };
template <int G, int X3>
__global__ void __launch_bounds__(128, 1) bench(Cfg c, long long *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x;
    for (int i = tid; i < 200 * 1024 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    if (tid == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    long long dt = 0;
    if (tid < 32) {
        constexpr uint32_t kHi = (128u >> 4) | (1u << 14);
        const int cw = X3 ? 2 * c.N : c.N;
        const uint32_t PL = 40960;  // plane stride (A): planes hi k0,k1 then lo k0,k1
        const uint32_t a_lo0 = (smem_u32(smem) >> 4) | ((PL >> 4) << 16);
        const uint32_t b_lo0 = (smem_u32(smem + 168 * 1024) >> 4) | ((((uint32_t)cw * 16) >> 4) << 16);
        const uint32_t idw = make_idesc(c.same_shape ? c.N : cw), idn = make_idesc(c.N);
        int parity = 0;
        for (int rep = 0; rep < 2; ++rep) {
            long long t0 = clock64();
            if (elect_one()) {
                for (int r = 0; r < c.reps; ++r) {
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        const uint32_t shift = (uint32_t)((tap / 3) * c.Wr + tap % 3);
                        const uint32_t alo = a_lo0 + shift;
                        const uint64_t bd = ((uint64_t)kHi << 32) | (uint64_t)(b_lo0 + (uint32_t)(tap * 2 * cw));
#pragma unroll
                        for (int gt = 0; gt < G; ++gt) umma(tmem + gt * cw, ((uint64_t)kHi << 32) | (uint64_t)(alo + gt * 128), bd, idw);
                        if (X3) {
#pragma unroll
                            for (int gt = 0; gt < G; ++gt) umma(tmem + gt * cw, ((uint64_t)kHi << 32) | (uint64_t)(alo + (2 * PL >> 4) + gt * 128), bd, idn);
                        }
                    }
                    if (c.commit_every) { commit(smem_u32(&bar)); }
                }
                if (!c.commit_every) commit(smem_u32(&bar));
            }
            __syncwarp();
            if (c.commit_every) { for (int r = 0; r < c.reps; ++r) { mbar_wait(smem_u32(&bar), parity); parity ^= 1; } }
            else { mbar_wait(smem_u32(&bar), parity); parity ^= 1; }
            dt = clock64() - t0;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    if (tid == 0) out[blockIdx.x] = dt;
}
template <int G, int X3>
void run(const char *name, Cfg c, long long *d) {
    cudaFuncSetAttribute(bench<G, X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
    bench<G, X3><<<148, 128, 208 * 1024>>>(c, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, d, 148 * 8, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    const int nm = c.reps * 9 * G * (X3 ? 2 : 1);
    printf("%-44s %s %7.1f cycles/MMA  (%d MMAs)\n", name, cudaGetErrorString(e), (double)mx / nm, nm); fflush(stdout);
}
int main() { setvbuf(stdout, NULL, _IONBF, 0);
    long long *d;
    cudaMalloc(&d, 148 * 8);
    run<4, 1>("x3 N=32 G=4 Wr=34", {32, 4, 1, 34, 16, 0, 0}, d);
    run<4, 1>("x3 N=32 G=4 Wr=34 same-shape(N,N)", {32, 4, 1, 34, 16, 1, 0}, d);
    run<2, 1>("x3 N=32 G=2 Wr=34", {32, 2, 1, 34, 16, 0, 0}, d);
    run<1, 1>("x3 N=32 G=1 Wr=34", {32, 1, 1, 34, 16, 0, 0}, d);
    run<1, 1>("x3 N=32 G=1 Wr=34 same-shape", {32, 1, 1, 34, 16, 1, 0}, d);
    run<4, 0>("bf16 N=32 G=4 Wr=34", {32, 4, 0, 34, 16, 0, 0}, d);
    run<1, 0>("bf16 N=32 G=1 Wr=34", {32, 1, 0, 34, 16, 0, 0}, d);
    run<2, 1>("x3 N=64 G=2 Wr=5", {64, 2, 1, 5, 16, 0, 0}, d);
    run<2, 1>("x3 N=64 G=2 Wr=5 same-shape", {64, 2, 1, 5, 16, 1, 0}, d);
    run<1, 1>("x3 N=64 G=1 Wr=1", {64, 1, 1, 1, 16, 0, 0}, d);
    run<2, 0>("bf16 N=64 G=2 Wr=5", {64, 2, 0, 5, 16, 0, 0}, d);
    run<1, 0>("bf16 N=64 G=1 Wr=1", {64, 1, 0, 1, 16, 0, 0}, d);
    run<4, 1>("x3 N=32 G=4 commit per 72", {32, 4, 1, 34, 16, 0, 1}, d);
    run<2, 1>("x3 N=64 G=2 commit per 36", {64, 2, 1, 5, 16, 0, 1}, d);
    return 0;
}
