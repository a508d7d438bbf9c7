// Pointwise (1x1) convolutions of the TCN bottleneck on tensor cores (model.py:556-561: DepthwiseSeparableConv =
// depthwise k3 -> PReLU -> gLN -> pointwise).  The pointwise conv consumes gLN(p), and
// gLN is an affine with ONE (mean, rstd) per sample, so
//     W gLN(p) = rstd * (W diag(gamma)) p  +  W beta - mean * rstd * (W gamma)
// is a plain GEMM over the raw PReLU outputs p with sample-independent weights; mean / rstd enter in the epilogue.
// tcn_wprep_kernel builds, once per forward for all 2*R*X pointwise convs, the bf16 hi/lo weight images
// (W diag(gamma), UMMA K-major order) and the two vectors W beta, W gamma; tcn_pw_kernel is a TMA-fed tcgen05 GEMM
// [128 frames x C] x [C x Nt] per CTA (A = the depthwise kernel's bf16 hi/lo planes, straight from HBM/L2) whose
// epilogue applies rstd and the bias, adds the residual (model.py:549), stores the fp32 channels-last state and
// accumulates the InstanceNorm1d statistics of the next block (model.py:530).
#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>

#include "conv.cuh"
#include "tcn.cuh"
#include "umma.cuh"

namespace miso {
namespace {

constexpr int kPwThreads = 10 * 32;  // warp 0: TMA, warp 1: MMA issuer, warps 2..9: epilogue
constexpr int kPwEpi0 = 2;
constexpr int kPwEpiThreads = 256;
constexpr int kPwSmemLimit = 227 * 1024;

struct PwGeom {
    int Nt, NH, nunit, kper, nchunk, nsp;
    int a_set, w_unit, w_off, stage, nstage;
    int off_red, off_vec, off_stage, smem_total, tmem_cols;
};

struct PwArgs {
    PwGeom g;
    const __nv_bfloat16 *wimg;  // [NH][nunit][hi|lo][kg][Nt][8]
    const float *wbeta, *wgamma;  // [C]
    const double *gln_sums;       // [B][2]
    double gln_inv_n;
    float gln_eps;
    float *out;          // fp32 channels-last [B][T][C]
    const float *resid;  // same layout, or null
    double *out_sums;    // [B][C][2] or null
    int B, T, C, t_tiles;
    int out_planes, out_ctot, use_lo;  // bf16 hi/lo plane output [B][hi|lo][out_ctot/8][T][8]
    size_t out_lo_off;
    int plain;  // no gLN epilogue (rstd 1, no bias): the data gradient GEMMs
};

template <int SPLIT>
__global__ void __launch_bounds__(kPwThreads, 1)
tcn_pw_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo, const PwArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const PwGeom &g = a.g;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Nt = g.Nt;
    constexpr int NPROD = SPLIT == 3 ? 3 : 1;
    const int mt = blockIdx.x, nh = blockIdx.y;
    const int b = mt / a.t_tiles, t0 = (mt - b * a.t_tiles) * 128;

    const uint32_t bar_full = smem_u32(smem), bar_empty = smem_u32(smem + 64), bar_done = smem_u32(smem + 128);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + 160);
    float *red = reinterpret_cast<float *>(smem + g.off_red);   // [4 lane quarters][Nt][2]
    float *vec = reinterpret_cast<float *>(smem + g.off_vec);   // [Nt] bias
    const uint32_t s_stage = smem_u32(smem + g.off_stage);

    if (tid == 0) {
        for (int s = 0; s < g.nstage; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        mbar_init(bar_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)g.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();
    pdl_wait();  // the set-up above overlapped the depthwise kernel's tail

    if (warp == 0) {
        if (elect_one()) {
            const __nv_bfloat16 *wsrc = a.wimg + (size_t)nh * g.nunit * (g.w_unit / 2);
            int s = 0, ph = 0;
            bool primed = false;
            for (int c = 0; c < g.nchunk; ++c) {
                if (primed) mbar_wait(bar_empty + 8 * s, (uint32_t)ph);
                const int nu = min(g.kper, g.nunit - g.kper * c);
                const uint32_t full = bar_full + 8 * s;
                mbar_expect_tx(full, (uint32_t)(g.nsp * g.a_set + nu * g.w_unit));
                const uint32_t sa = s_stage + (uint32_t)(s * g.stage);
                for (int sp = 0; sp < g.nsp; ++sp)
                    tma_load_4d(sa + (uint32_t)(sp * g.a_set), sp == 0 ? &tm_hi : &tm_lo, full, 0, t0, 2 * g.kper * c, b);
                bulk_load(sa + (uint32_t)g.w_off, wsrc + (size_t)c * g.kper * (g.w_unit / 2), (uint32_t)(nu * g.w_unit), full);
                if (++s == g.nstage) {
                    s = 0;
                    if (primed) ph ^= 1;
                    primed = true;
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            const uint32_t idesc = make_idesc(Nt);
            constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);
            const uint32_t a_lo0 = (s_stage >> 4) | ((2048u >> 4) << 16);                            // planes of 128 rows x 16 B
            const uint32_t b_lo0 = ((s_stage + (uint32_t)g.w_off) >> 4) | ((uint32_t)Nt << 16);      // leading offset Nt * 16 B
            const uint32_t lo_split = (uint32_t)g.a_set >> 4, b_spstep = (uint32_t)(2 * Nt);
            const uint32_t a_kstep = 4096u >> 4, b_kstep = (uint32_t)g.w_unit >> 4;
            int s = 0, ph = 0;
            for (int c = 0; c < g.nchunk; ++c) {
                mbar_wait(bar_full + 8 * s, (uint32_t)ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t st4 = (uint32_t)(s * g.stage) >> 4;
                const int nu = min(g.kper, g.nunit - g.kper * c);
#pragma unroll 1
                for (int ks = 0; ks < nu; ++ks) {
                    const uint32_t a_ks = a_lo0 + st4 + (uint32_t)ks * a_kstep;
                    const uint32_t b_ks = b_lo0 + st4 + (uint32_t)ks * b_kstep;
#pragma unroll
                    for (int pr = 0; pr < NPROD; ++pr) {  // a_hi w_hi, a_hi w_lo, a_lo w_hi
                        const uint32_t alo = a_ks + (pr == 2 ? lo_split : 0u);
                        const uint32_t blo = b_ks + (pr == 1 ? b_spstep : 0u);
                        umma_bf16(tmem_base, ((uint64_t)kDescHi << 32) | (uint64_t)alo, ((uint64_t)kDescHi << 32) | (uint64_t)blo, idesc,
                                  (c == 0 && ks == 0 && pr == 0) ? 0u : 1u);
                    }
                }
                umma_commit(bar_empty + 8 * s);
                if (c == g.nchunk - 1) umma_commit(bar_done);
                if (++s == g.nstage) {
                    s = 0;
                    ph ^= 1;
                }
            }
        }
    } else {
        // ---------------------------------------------------------------- epilogue
        const int quad = warp & 3, half = (warp - kPwEpi0) >> 2;
        const int et = tid - kPwEpi0 * 32;
        // gLN statistics of this sample (model.py:628-631)
        float rstd = 1.f, mr = 0.f;
        if (!a.plain) {
            const double mean = stat_get(a.gln_sums + (size_t)b * 2) * a.gln_inv_n;
            double var = stat_get(a.gln_sums + (size_t)b * 2 + 1) * a.gln_inv_n - mean * mean;
            if (var < 0.0) var = 0.0;
            const double rstd_d = rsqrt(var + (double)a.gln_eps);
            rstd = (float)rstd_d;
            mr = (float)(mean * rstd_d);
        }
        const int co0 = nh * Nt;
        for (int i = et; i < Nt; i += kPwEpiThreads) vec[i] = a.plain ? 0.f : __ldg(a.wbeta + co0 + i) - mr * __ldg(a.wgamma + co0 + i);
        asm volatile("bar.sync 1, %0;" ::"n"(kPwEpiThreads));
        mbar_wait(bar_done, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // The accumulator comes out of tensor memory one FRAME per lane, where a warp-wide 16-byte access to the fp32
        // channels-last state touches 32 different lines (measured: the epilogue was most of the kernel).  Each 32-column group
        // is turned through a per-warp staging tile (the pipeline stages are idle by now) so that a quarter warp handles the 128
        // contiguous bytes of one frame: residual loads, stores and the statistics all run in that layout.
        float *stg = reinterpret_cast<float *>(smem + g.off_stage) + (warp - kPwEpi0) * (32 * 36);  // [32 frames][32 + 4 floats]
        const int cq = lane & 7, rl = lane >> 3;
        const int ncol = Nt / 2;  // columns of this warp's half
        for (int cb = half * ncol; cb < (half + 1) * ncol; cb += 32) {
            uint32_t v0[16], v1[16];
            float4 rr[8];
            tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)cb, v0);
            tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(cb + 16), v1);
#pragma unroll
            for (int i = 0; i < 8; ++i) {  // residual: frames rl + 4 i of this warp's 32, channel quad cq
                const int tt = t0 + quad * 32 + rl + 4 * i;
                rr[i] = (a.resid && tt < a.T) ? *reinterpret_cast<const float4 *>(a.resid + ((size_t)b * a.T + tt) * a.C + co0 + cb + 4 * cq)
                                              : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            __syncwarp();  // the previous group's readers are done with the tile
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
                *reinterpret_cast<uint4 *>(stg + lane * 36 + j) = make_uint4(v0[j], v0[j + 1], v0[j + 2], v0[j + 3]);
                *reinterpret_cast<uint4 *>(stg + lane * 36 + 16 + j) = make_uint4(v1[j], v1[j + 1], v1[j + 2], v1[j + 3]);
            }
            __syncwarp();
            const float4 b4 = *reinterpret_cast<const float4 *>(vec + cb + 4 * cq);
            float ps[4] = {0.f, 0.f, 0.f, 0.f}, pq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int rw = rl + 4 * i, tt = t0 + quad * 32 + rw;
                const float4 ac = *reinterpret_cast<const float4 *>(stg + rw * 36 + 4 * cq);
                float4 y;
                y.x = fmaf(ac.x, rstd, b4.x) + rr[i].x;
                y.y = fmaf(ac.y, rstd, b4.y) + rr[i].y;
                y.z = fmaf(ac.z, rstd, b4.z) + rr[i].z;
                y.w = fmaf(ac.w, rstd, b4.w) + rr[i].w;
                if (tt < a.T) {
                    if (!a.out_planes) {
                        *reinterpret_cast<float4 *>(a.out + ((size_t)b * a.T + tt) * a.C + co0 + cb + 4 * cq) = y;
                    } else {  // four channels of a frame = 8 bytes of the frame's 16 in an 8-channel plane group
                        const int c = co0 + cb + 4 * cq;
                        __nv_bfloat16 *p = reinterpret_cast<__nv_bfloat16 *>(a.out) + (size_t)b * 2 * a.out_ctot * a.T + ((size_t)(c >> 3) * a.T + tt) * 8 + (c & 7);
                        const uint32_t h0 = pack_bf16x2(y.x, y.y), h1 = pack_bf16x2(y.z, y.w);
                        *reinterpret_cast<uint2 *>(p) = make_uint2(h0, h1);
                        if (a.use_lo)
                            *reinterpret_cast<uint2 *>(reinterpret_cast<char *>(p) + a.out_lo_off) =
                                make_uint2(pack_bf16x2(y.x - bf16_lo(h0), y.y - bf16_hi(h0)), pack_bf16x2(y.z - bf16_lo(h1), y.w - bf16_hi(h1)));
                    }
                    ps[0] += y.x, ps[1] += y.y, ps[2] += y.z, ps[3] += y.w;
                    pq[0] = fmaf(y.x, y.x, pq[0]), pq[1] = fmaf(y.y, y.y, pq[1]), pq[2] = fmaf(y.z, y.z, pq[2]), pq[3] = fmaf(y.w, y.w, pq[3]);
                }
            }
            if (a.out_sums) {  // the four quarter warps hold the same channels: fixed-order butterfly
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    ps[j] += __shfl_xor_sync(0xffffffffu, ps[j], 8);
                    pq[j] += __shfl_xor_sync(0xffffffffu, pq[j], 8);
                    ps[j] += __shfl_xor_sync(0xffffffffu, ps[j], 16);
                    pq[j] += __shfl_xor_sync(0xffffffffu, pq[j], 16);
                }
                if (rl == 0) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        red[((size_t)quad * Nt + cb + 4 * cq + j) * 2] = ps[j];
                        red[((size_t)quad * Nt + cb + 4 * cq + j) * 2 + 1] = pq[j];
                    }
                }
            }
        }
        if (a.out_sums) {
            asm volatile("bar.sync 1, %0;" ::"n"(kPwEpiThreads));
            for (int c = et; c < Nt; c += kPwEpiThreads) {  // the four lane quarters in a fixed order
                double s8 = 0.0, q8 = 0.0;
#pragma unroll
                for (int w4 = 0; w4 < 4; ++w4) {
                    s8 += (double)red[((size_t)w4 * Nt + c) * 2];
                    q8 += (double)red[((size_t)w4 * Nt + c) * 2 + 1];
                }
                double *dst = a.out_sums + ((size_t)b * a.C + co0 + c) * 2;
                stat_add(dst, s8);
                stat_add(dst + 1, q8);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)g.tmem_cols) : "memory");
    }
}

// weight images and bias vectors of all pointwise convs of the TCN: block (unit or vector part, half)
// transposed: the images of the DATA GRADIENT GEMMs dq = dy W^T (K = the forward's output channels, no gLN factors, no vectors)
__global__ void __launch_bounds__(256) tcn_wprep_kernel(const TcnPwTable tab, __nv_bfloat16 *wimg, float *wvec, int C, int cpad, int Nt,
                                                        int nsp, int transposed) {
    const int h = blockIdx.y;
    const float *W = tab.w[h], *gamma = tab.gamma[h], *beta = tab.beta[h];
    const int nunit = C / 16, NH = C / Nt;
    const size_t unit_elems = (size_t)nsp * 2 * Nt * 8;
    if ((int)blockIdx.x < nunit) {
        const int unit = blockIdx.x;
        __shared__ float gm[16];
        if (threadIdx.x < 16) gm[threadIdx.x] = transposed ? 1.f : gamma[unit * 16 + threadIdx.x];
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
            const int co = i % C, kg = i / C;
            float hv[8], lv[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int ci = unit * 16 + kg * 8 + e;
                const float v = (transposed ? W[(size_t)co * cpad + ci] : W[(size_t)ci * cpad + co]) * gm[kg * 8 + e];
                hv[e] = bf16_round(v);
                lv[e] = v - hv[e];
            }
            const int nhalf = co / Nt, n = co - nhalf * Nt;
            __nv_bfloat16 *dst = wimg + (((size_t)h * NH + nhalf) * nunit + unit) * unit_elems + ((size_t)kg * Nt + n) * 8;
            *reinterpret_cast<uint4 *>(dst) =
                make_uint4(pack_bf16x2(hv[0], hv[1]), pack_bf16x2(hv[2], hv[3]), pack_bf16x2(hv[4], hv[5]), pack_bf16x2(hv[6], hv[7]));
            if (nsp == 2)
                *reinterpret_cast<uint4 *>(dst + (size_t)2 * Nt * 8) =
                    make_uint4(pack_bf16x2(lv[0], lv[1]), pack_bf16x2(lv[2], lv[3]), pack_bf16x2(lv[4], lv[5]), pack_bf16x2(lv[6], lv[7]));
        }
    } else {
        // W beta and W gamma: one thread per output channel, fixed summation order
        const int co = (blockIdx.x - nunit) * blockDim.x + threadIdx.x;
        if (co < C && !transposed) {
            float sb = 0.f, sg = 0.f;
            for (int ci = 0; ci < C; ++ci) {
                const float w = W[(size_t)ci * cpad + co];
                sb = fmaf(w, beta[ci], sb);
                sg = fmaf(w, gamma[ci], sg);
            }
            wvec[((size_t)h * 2) * C + co] = sb;
            wvec[((size_t)h * 2 + 1) * C + co] = sg;
        }
    }
}


// ------------------------------------------------------------------------------------------------
// The whole TCN (model.py:486-567) in one launch.  Every statistic of the TCN is per SAMPLE (InstanceNorm1d over the frames
// of a channel, gLN over all channels and frames), so a thread-block cluster per sample -- one CTA per 128 frames -- can run
// the 2 R X half-blocks back to back with cluster barriers where the launch-per-kernel formulation needed grid boundaries:
//   per half-block:  [cluster barrier: the input state and its InstanceNorm1d sums are complete]
//     warps 2..9   v = ELU(IN1d(u)), y = dilated depthwise k3 (v), p = PReLU(y): 32-channel K chunks of the tile's [128 x C]
//                  A operand written straight into shared memory as bf16 hi / lo in UMMA K-major order (the state is read from
//                  L2: no activation plane round trip); gLN sums of p -> global fixed-point accumulators
//     warp 0       streams the pointwise conv's weight image (tcn_wprep_kernel) through a ring of bulk copies
//     warp 1       tcgen05 MMAs [128 x C] x [C x C] into tensor memory as the chunks arrive
//                  [cluster barrier: the sample's gLN sums are complete]
//     warps 2..9   epilogue: rstd * acc + (W beta - mean rstd W gamma) + residual -> next state (fp32 channels-last, or the
//                  decoder's bf16 planes for the last half) and the InstanceNorm1d sums of the next half
constexpr int kFuThreads = 18 * 32;  // warp 0: weight stream, warp 1: MMA issuer, warps 2..17: depthwise half, then epilogue
constexpr int kFuEpi0 = 2;
constexpr int kFuEpiThreads = 512;
constexpr int kFuSub = kFuEpiThreads / 128;  // warps per tensor-memory lane quarter
constexpr int kFuRing = 3;

struct FuGeom {
    int Nt, NH, nunit, nchunk, nsp;
    int a_set, a_stage, w_unit, w_stage;
    int off_sc, off_vec, off_red, off_a, off_w, smem_total, tmem_cols;
};

__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ double stat_get_cg(const double *p) {  // written by other CTAs of this launch: read through L2
    return (double)__ldcg(reinterpret_cast<const long long *>(p)) * (1.0 / kStatScale);
}
__device__ __forceinline__ float fu_elu(float x) {  // ELU through ex2.approx (absolute error ~1e-7), branch-free
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fminf(x, 0.f) * 1.4426950408889634f));
    return fmaxf(x, e - 1.f);
}

template <int SPLIT>
__global__ void __launch_bounds__(kFuThreads, 1) tcn_fused_kernel(const __grid_constant__ TcnFusedArgs a, const FuGeom g) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int C = a.C, T = a.T, Nt = g.Nt;
    constexpr int NPROD = SPLIT == 3 ? 3 : 1;
    const int t_tiles = (T + 127) / 128;
    const int b = blockIdx.x / t_tiles, t0 = (blockIdx.x - b * t_tiles) * 128;

    const uint32_t bar_afull = smem_u32(smem), bar_aempty = smem_u32(smem + 32), bar_wfull = smem_u32(smem + 64), bar_wempty = smem_u32(smem + 96);
    const uint32_t bar_acc = smem_u32(smem + 128);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + 160);
    float *redg = reinterpret_cast<float *>(smem + 192);        // [16 warps][2] (the bytes up to off_sc = 512)
    float *sc = reinterpret_cast<float *>(smem + g.off_sc);     // [C] scale, [C] shift, [3][C] taps
    float *sf = sc + C, *wt = sc + 2 * C;
    float *vec = reinterpret_cast<float *>(smem + g.off_vec);   // [C] epilogue bias
    float *red = reinterpret_cast<float *>(smem + g.off_red);   // [4 lane quarters][C][2]
    const uint32_t s_a = smem_u32(smem + g.off_a), s_w = smem_u32(smem + g.off_w);

    if (tid == 0) {
        for (int i = 0; i < kFuRing; ++i) {
            mbar_init(bar_afull + 8 * i, kFuEpiThreads / 32);
            mbar_init(bar_aempty + 8 * i, 1);
            mbar_init(bar_wfull + 8 * i, 1);
            mbar_init(bar_wempty + 8 * i, 1);
        }
        mbar_init(bar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)g.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    auto stamp = [&](int h, int i) {
        if (a.trace && blockIdx.x == 0) a.trace[h * 16 + i] = clock64();
    };
    for (int h = 0; h < a.nhalf; ++h) {
        const TcnFusedHalf &H = a.h[h];
        __syncwarp();
        if (h > 0) cluster_sync_all();  // the previous half's state and InstanceNorm1d sums are complete for the whole sample
        if (warp == 0) {
            // ------------------------------------------------------------ weight stream
            if (elect_one()) {
                const __nv_bfloat16 *wsrc = reinterpret_cast<const __nv_bfloat16 *>(a.wimg) + (size_t)h * C * C * g.nsp;
                for (int c = 0; c < g.nchunk; ++c) {
                    const int item = h * g.nchunk + c, slot = item % kFuRing, use = item / kFuRing;
                    if (use > 0) mbar_wait(bar_wempty + 8 * slot, (uint32_t)((use - 1) & 1));
                    const uint32_t full = bar_wfull + 8 * slot;
                    mbar_expect_tx(full, (uint32_t)(g.NH * 2 * g.w_unit));
                    for (int nh = 0; nh < g.NH; ++nh)
                        bulk_load(s_w + (uint32_t)(slot * g.w_stage + nh * 2 * g.w_unit), wsrc + ((size_t)nh * g.nunit + 2 * c) * (g.w_unit / 2),
                                  (uint32_t)(2 * g.w_unit), full);
                }
            }
            __syncwarp();
        } else if (warp == 1) {
            // ------------------------------------------------------------ MMA issuer
            if (elect_one()) {
                const uint32_t idesc = make_idesc(Nt);
                constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);
                for (int c = 0; c < g.nchunk; ++c) {
                    const int item = h * g.nchunk + c, slot = item % kFuRing;
                    const uint32_t ph = (uint32_t)((item / kFuRing) & 1);
                    if (c == 0) stamp(h, 10);
                    mbar_wait(bar_wfull + 8 * slot, ph);
                    if (c == 0) stamp(h, 11);
                    mbar_wait(bar_afull + 8 * slot, ph);
                    if (c == 0) stamp(h, 12);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a0 = (((s_a + (uint32_t)(slot * g.a_stage)) & 0x3FFFFu) >> 4) | ((2048u >> 4) << 16);      // K groups of 128 rows x 16 B
                    const uint32_t b0 = (((s_w + (uint32_t)(slot * g.w_stage)) & 0x3FFFFu) >> 4) | ((uint32_t)Nt << 16);      // leading offset Nt * 16 B
#pragma unroll 1
                    for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
                        for (int pr = 0; pr < NPROD; ++pr) {  // a_hi w_hi, a_hi w_lo, a_lo w_hi
                            const uint32_t alo = a0 + (uint32_t)(ks * 4096 + (pr == 2 ? g.a_set : 0)) / 16u;
                            for (int nh = 0; nh < g.NH; ++nh) {
                                const uint32_t blo = b0 + (uint32_t)((nh * 2 + ks) * g.w_unit) / 16u + (pr == 1 ? (uint32_t)(2 * Nt) : 0u);
                                umma_bf16(tmem_base + (uint32_t)(nh * Nt), ((uint64_t)kDescHi << 32) | (uint64_t)alo, ((uint64_t)kDescHi << 32) | (uint64_t)blo,
                                          idesc, (c == 0 && ks == 0 && pr == 0) ? 0u : 1u);
                            }
                        }
                    }
                    umma_commit(bar_aempty + 8 * slot);
                    umma_commit(bar_wempty + 8 * slot);
                    if (c == g.nchunk - 1) umma_commit(bar_acc);
                }
                stamp(h, 13);
            }
            __syncwarp();
        } else {
            // ------------------------------------------------------------ depthwise half: the A operand, chunk by chunk
            const int e = tid - kFuEpi0 * 32;
            if (e == 0) stamp(h, 0);
            const float al = __ldg(H.alpha);
            // Thread mapping: lanes run along the CHANNELS (a quarter warp reads the 128 contiguous bytes of a frame's 32-channel
            // chunk, a warp-wide load covers four whole frames); a thread owns channel quad cq of the chunk in the frames
            // rq, rq + 64 of the tile.  (One frame per lane -- the layout of the A operand -- costs 32 separate
            // L2 requests per load instruction: measured 9 k cycles per chunk.)
            const int cq = e & 7, rq = e >> 3, dil = H.dil;
            const float *ub = H.u + (size_t)b * T * C + 4 * cq;
            constexpr int NR = 128 * 8 / kFuEpiThreads;  // frames per thread
            int roff[NR][3];  // element offsets of the three tap rows, clamped into the tensor; tap validity in tapm
            unsigned tapm = 0;
#pragma unroll
            for (int i = 0; i < NR; ++i) {
                const int t = t0 + rq + (128 / NR) * i;
                const bool tv = t < T, mv = tv && t - dil >= 0, pv = tv && t + dil < T;
                roff[i][0] = (mv ? t - dil : 0) * C;
                roff[i][1] = (tv ? t : 0) * C;
                roff[i][2] = (pv ? t + dil : 0) * C;
                tapm |= (mv ? 1u : 0u) << (3 * i) | (tv ? 2u : 0u) << (3 * i) | (pv ? 4u : 0u) << (3 * i);
            }
            float s = 0.f, q = 0.f;
            // the state comes from L2 (written by other CTAs of this launch): the twelve 16-byte loads of the NEXT chunk are in
            // flight while this chunk is computed
            float4 cur[3 * NR], nxt[3 * NR];
            auto fetch = [&](float4 (&L)[3 * NR], int c) {
#pragma unroll
                for (int i = 0; i < NR; ++i)
#pragma unroll
                    for (int k = 0; k < 3; ++k) L[i * 3 + k] = __ldcg(reinterpret_cast<const float4 *>(ub + roff[i][k] + c * 32));
            };
            fetch(cur, 0);
            // (the first chunk's loads are in flight during the set-up below)
            for (int c = e; c < C; c += kFuEpiThreads) {
                const double *us = H.u_sums + ((size_t)b * C + c) * 2;
                const float2 af = affine_from_sums(stat_get_cg(us), stat_get_cg(us + 1), a.in_inv_n, (double)a.in_eps);
                sc[c] = af.x;
                sf[c] = af.y;
                wt[c] = __ldg(H.dw + c * 3);
                wt[C + c] = __ldg(H.dw + c * 3 + 1);
                wt[2 * C + c] = __ldg(H.dw + c * 3 + 2);
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kFuEpiThreads));
            if (e == 0) stamp(h, 1);
            // A operand [unit][8-channel K group][frame][8 ch] (bf16): this thread's four channels are 8 bytes of a frame's 16
            const uint32_t a_thr = (uint32_t)((cq >> 2) * 4096 + ((cq >> 1) & 1) * 2048 + (cq & 1) * 8);
            for (int c = 0; c < g.nchunk; ++c) {
                const int item = h * g.nchunk + c, slot = item % kFuRing, use = item / kFuRing;
                if (c + 1 < g.nchunk) fetch(nxt, c + 1);
                if (use > 0) mbar_wait(bar_aempty + 8 * slot, (uint32_t)((use - 1) & 1));
                const int cc = c * 32 + 4 * cq;
                const float4 s4 = *reinterpret_cast<const float4 *>(sc + cc), f4 = *reinterpret_cast<const float4 *>(sf + cc);
                float4 w4[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) w4[k] = *reinterpret_cast<const float4 *>(wt + k * C + cc);
                uint8_t *dst = smem + g.off_a + slot * g.a_stage + a_thr;
#pragma unroll
                for (int i = 0; i < NR; ++i) {
                    float y[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int k = 0; k < 3; ++k) {  // y += w_k * ELU(IN(u[t + (k - 1) dil])); branch-free: a dropped tap has weight 0
                        const float m = (float)((tapm >> (3 * i + k)) & 1u);
                        const float4 u = cur[i * 3 + k];
                        y[0] = fmaf(w4[k].x * m, fu_elu(fmaf(u.x, s4.x, f4.x)), y[0]);
                        y[1] = fmaf(w4[k].y * m, fu_elu(fmaf(u.y, s4.y, f4.y)), y[1]);
                        y[2] = fmaf(w4[k].z * m, fu_elu(fmaf(u.z, s4.z, f4.z)), y[2]);
                        y[3] = fmaf(w4[k].w * m, fu_elu(fmaf(u.w, s4.w, f4.w)), y[3]);
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        y[j] = y[j] > 0.f ? y[j] : al * y[j];
                        s += y[j];
                        q = fmaf(y[j], y[j], q);
                    }
                    const uint32_t h0 = pack_bf16x2(y[0], y[1]), h1 = pack_bf16x2(y[2], y[3]);
                    uint8_t *d = dst + (rq + (128 / NR) * i) * 16;
                    *reinterpret_cast<uint2 *>(d) = make_uint2(h0, h1);
                    if (SPLIT == 3)
                        *reinterpret_cast<uint2 *>(d + g.a_set) =
                            make_uint2(pack_bf16x2(y[0] - bf16_lo(h0), y[1] - bf16_hi(h0)), pack_bf16x2(y[2] - bf16_lo(h1), y[3] - bf16_hi(h1)));
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> the tensor core's async-proxy reads
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_afull + 8 * slot);
#pragma unroll
                for (int i = 0; i < 3 * NR; ++i) cur[i] = nxt[i];
            }
            if (e == 0) stamp(h, 2);
            // gLN sums of this tile (rows past T contributed zeros)
            s = warp_sum(s);
            q = warp_sum(q);
            if (lane == 0) {
                redg[(warp - kFuEpi0) * 2] = s;
                redg[(warp - kFuEpi0) * 2 + 1] = q;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kFuEpiThreads));
            if (e == 0) {
                double ds = 0.0, dq = 0.0;
                for (int w8 = 0; w8 < kFuEpiThreads / 32; ++w8) {
                    ds += (double)redg[w8 * 2];
                    dq += (double)redg[w8 * 2 + 1];
                }
                stat_add(H.g_sums + (size_t)b * 2, ds);
                stat_add(H.g_sums + (size_t)b * 2 + 1, dq);
                stamp(h, 3);
            }
        }
        __syncwarp();
        cluster_sync_all();  // the sample's gLN sums are complete
        if (warp >= kFuEpi0) {
            // ------------------------------------------------------------ epilogue
            const int e = tid - kFuEpi0 * 32;
            const int quad = warp & 3, sub = (warp - kFuEpi0) >> 2;
            if (e == 0) stamp(h, 4);
            const double mean = stat_get_cg(H.g_sums + (size_t)b * 2) * a.gln_inv_n;
            double var = stat_get_cg(H.g_sums + (size_t)b * 2 + 1) * a.gln_inv_n - mean * mean;
            if (var < 0.0) var = 0.0;
            const double rstd_d = rsqrt(var + (double)a.gln_eps);
            const float rstd = (float)rstd_d, mr = (float)(mean * rstd_d);
            const float *wbeta = a.wvec + (size_t)h * 2 * C, *wgamma = wbeta + C;
            for (int i = e; i < C; i += kFuEpiThreads) vec[i] = __ldg(wbeta + i) - mr * __ldg(wgamma + i);
            asm volatile("bar.sync 1, %0;" ::"n"(kFuEpiThreads));
            const int r = quad * 32 + lane, t = t0 + r;
            const bool valid = t < T;
            const size_t row = ((size_t)b * T + t) * C;
            if (e == 0) stamp(h, 5);
            mbar_wait(bar_acc, (uint32_t)(h & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (e == 0) stamp(h, 6);
            if (!H.out_planes) {
                // fp32 channels-last output.  The accumulator comes out of tensor memory one FRAME per lane, where a warp-wide
                // 16-byte access would touch 32 different lines of the state: each 32-column group is turned through a per-warp
                // staging tile (the idle A ring) into the layout of the depthwise half -- a quarter warp per frame's 128 bytes.
                float *stg = reinterpret_cast<float *>(smem + g.off_a) + (warp - kFuEpi0) * (32 * 36);  // [32 frames][32 + 4 floats]
                const int cq = lane & 7, rl = lane >> 3;
                const int cbeg = sub * (C / kFuSub), cend = cbeg + C / kFuSub;
                const bool has_res = H.resid != nullptr;
                uint32_t v0[16], v1[16];
                float4 rr[8];
                // accumulator columns (tensor memory) and residual (L2) of one 32-column group
                auto issue = [&](int cb) {
                    tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)cb, v0);
                    tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(cb + 16), v1);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {  // frames rl + 4 i of this warp's 32, channel quad cq (coalesced)
                        const int tt = t0 + quad * 32 + rl + 4 * i;
                        rr[i] = (has_res && tt < T) ? __ldcg(reinterpret_cast<const float4 *>(H.resid + ((size_t)b * T + tt) * C + cb + 4 * cq))
                                                   : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                };
                for (int cb = cbeg; cb < cend; cb += 32) {
                    issue(cb);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    __syncwarp();  // the previous group's readers are done with the tile
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        *reinterpret_cast<uint4 *>(stg + lane * 36 + j) = make_uint4(v0[j], v0[j + 1], v0[j + 2], v0[j + 3]);
                        *reinterpret_cast<uint4 *>(stg + lane * 36 + 16 + j) = make_uint4(v1[j], v1[j + 1], v1[j + 2], v1[j + 3]);
                    }
                    __syncwarp();
                    const float4 b4 = *reinterpret_cast<const float4 *>(vec + cb + 4 * cq);
                    float ps[4] = {0.f, 0.f, 0.f, 0.f}, pq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int rw = rl + 4 * i, tt = t0 + quad * 32 + rw;
                        const float4 ac = *reinterpret_cast<const float4 *>(stg + rw * 36 + 4 * cq);
                        float4 y;
                        y.x = fmaf(ac.x, rstd, b4.x) + rr[i].x;
                        y.y = fmaf(ac.y, rstd, b4.y) + rr[i].y;
                        y.z = fmaf(ac.z, rstd, b4.z) + rr[i].z;
                        y.w = fmaf(ac.w, rstd, b4.w) + rr[i].w;
                        if (tt < T) {
                            *reinterpret_cast<float4 *>(reinterpret_cast<float *>(H.out) + ((size_t)b * T + tt) * C + cb + 4 * cq) = y;
                            ps[0] += y.x, ps[1] += y.y, ps[2] += y.z, ps[3] += y.w;
                            pq[0] = fmaf(y.x, y.x, pq[0]), pq[1] = fmaf(y.y, y.y, pq[1]), pq[2] = fmaf(y.z, y.z, pq[2]), pq[3] = fmaf(y.w, y.w, pq[3]);
                        }
                    }
                    if (H.out_sums) {  // the four quarter warps hold the same channels: fixed-order butterfly
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            ps[j] += __shfl_xor_sync(0xffffffffu, ps[j], 8);
                            pq[j] += __shfl_xor_sync(0xffffffffu, pq[j], 8);
                            ps[j] += __shfl_xor_sync(0xffffffffu, ps[j], 16);
                            pq[j] += __shfl_xor_sync(0xffffffffu, pq[j], 16);
                        }
                        if (rl == 0) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                red[((size_t)quad * C + cb + 4 * cq + j) * 2] = ps[j];
                                red[((size_t)quad * C + cb + 4 * cq + j) * 2 + 1] = pq[j];
                            }
                        }
                    }
                }
            } else {
            // bf16 planes output (the last half): one frame per lane IS the coalesced layout of a plane
            // 16-column chunks alternate between the two warps of a lane quarter; the accumulator / residual loads of the next
            // chunk are in flight while this one is processed
            uint32_t v[16], vn[16];
            float rr[16], rn[16];
            auto fetch = [&](uint32_t (&V)[16], float (&R)[16], int cb) {
                tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)cb, V);
                if (H.resid && valid) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        const float4 x = __ldcg(reinterpret_cast<const float4 *>(H.resid + row + cb + j));
                        R[j] = x.x, R[j + 1] = x.y, R[j + 2] = x.z, R[j + 3] = x.w;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) R[j] = 0.f;
                }
            };
            fetch(v, rr, 16 * sub);
            for (int cb = 16 * sub; cb < C; cb += 16 * kFuSub) {
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                float acc[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[j] = __uint_as_float(v[j]);
                if (cb + 16 * kFuSub < C) fetch(vn, rn, cb + 16 * kFuSub);
                float y[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) y[j] = fmaf(acc[j], rstd, vec[cb + j]) + rr[j];
                if (valid) {
                    if (!H.out_planes) {
                        float *o = reinterpret_cast<float *>(H.out) + row + cb;
#pragma unroll
                        for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4 *>(o + j) = make_float4(y[j], y[j + 1], y[j + 2], y[j + 3]);
                    } else {
                        __nv_bfloat16 *op = reinterpret_cast<__nv_bfloat16 *>(H.out) + (size_t)b * 2 * a.out_ctot * T;
#pragma unroll
                        for (int g8 = 0; g8 < 16; g8 += 8) {
                            __nv_bfloat16 *p = op + ((size_t)((cb + g8) >> 3) * T + t) * 8;
                            uint32_t hp[4], lp[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                hp[j] = pack_bf16x2(y[g8 + 2 * j], y[g8 + 2 * j + 1]);
                                lp[j] = pack_bf16x2(y[g8 + 2 * j] - bf16_lo(hp[j]), y[g8 + 2 * j + 1] - bf16_hi(hp[j]));
                            }
                            *reinterpret_cast<uint4 *>(p) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
                            if (a.use_lo) *reinterpret_cast<uint4 *>(reinterpret_cast<char *>(p) + a.out_lo_off) = make_uint4(lp[0], lp[1], lp[2], lp[3]);
                        }
                    }
                }
                if (H.out_sums) {
                    float ssum[16], ssq[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        ssum[j] = valid ? y[j] : 0.f;
                        ssq[j] = valid ? y[j] * y[j] : 0.f;
                    }
                    const float s1 = warp_reduce16(ssum, lane);
                    const float s2 = warp_reduce16(ssq, lane);
                    if ((lane & 1) == 0) {
                        red[((size_t)quad * C + cb + (lane >> 1)) * 2] = s1;
                        red[((size_t)quad * C + cb + (lane >> 1)) * 2 + 1] = s2;
                    }
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    v[j] = vn[j];
                    rr[j] = rn[j];
                }
            }
            }
            if (e == 0) stamp(h, 7);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");  // the next half's MMAs overwrite what was just read
            asm volatile("bar.sync 1, %0;" ::"n"(kFuEpiThreads));
            if (H.out_sums) {
                for (int c = e; c < C; c += kFuEpiThreads) {  // the four lane quarters in a fixed order
                    double s8 = 0.0, q8 = 0.0;
#pragma unroll
                    for (int w4 = 0; w4 < 4; ++w4) {
                        s8 += (double)red[((size_t)w4 * C + c) * 2];
                        q8 += (double)red[((size_t)w4 * C + c) * 2 + 1];
                    }
                    double *dst = H.out_sums + ((size_t)b * C + c) * 2;
                    stat_add(dst, s8);
                    stat_add(dst + 1, q8);
                }
            }
            if (e == 0) stamp(h, 8);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)g.tmem_cols) : "memory");
    }
}

typedef CUresult (*PwEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                               const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PwEncodeFn pw_get_encode() {
    static PwEncodeFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PwEncodeFn>(p);
    }
    return fn;
}

int pw_round_up(int x, int m) { return (x + m - 1) / m * m; }

bool make_pw_geom(int C, int split, PwGeom &g) {
    g = PwGeom{};
    if (C % 32) return false;
    g.nsp = split == 3 ? 2 : 1;
    g.Nt = 0;
    for (int nt : {192, 128, 256, 64})  // (each half of a tile is a whole number of the epilogue's 32-column groups)
        if (C % nt == 0) {
            g.Nt = nt;
            break;
        }
    if (!g.Nt) return false;
    g.NH = C / g.Nt;
    g.nunit = C / 16;
    g.kper = 2;
    g.nchunk = (g.nunit + g.kper - 1) / g.kper;
    g.a_set = 2 * g.kper * 2048;
    g.w_unit = g.nsp * 2 * g.Nt * 16;
    g.w_off = g.nsp * g.a_set;
    g.stage = pw_round_up(g.w_off + g.kper * g.w_unit, 1024);
    g.off_red = 256;
    g.off_vec = g.off_red + 4 * g.Nt * 2 * 4;
    g.off_stage = pw_round_up(g.off_vec + g.Nt * 4, 1024);
    g.nstage = std::min(4, (kPwSmemLimit - g.off_stage) / g.stage);
    if (g.nstage < 2) return false;
    g.smem_total = g.off_stage + g.nstage * g.stage;
    int cols = 32;
    while (cols < g.Nt) cols <<= 1;
    g.tmem_cols = cols;
    return true;
}

bool make_fu_geom(int C, int split, FuGeom &g) {
    PwGeom p;
    if (!make_pw_geom(C, split, p)) return false;
    g = FuGeom{};
    g.Nt = p.Nt;
    g.NH = p.NH;
    g.nunit = p.nunit;
    g.nsp = p.nsp;
    if (g.nunit % 2 || C > 512 || C % (32 * kFuSub)) return false;  // (the epilogue splits the channels into kFuSub runs of whole 32-column groups)
    g.nchunk = g.nunit / 2;
    g.a_set = 2 * 4096;              // two 16-channel K units of 128 rows
    g.a_stage = g.nsp * g.a_set;
    g.w_unit = p.w_unit;
    g.w_stage = g.NH * 2 * g.w_unit;
    g.off_sc = 512;
    g.off_vec = g.off_sc + 5 * C * 4;
    g.off_red = g.off_vec + C * 4;
    g.off_a = pw_round_up(g.off_red + 4 * C * 2 * 4, 1024);
    g.off_w = g.off_a + kFuRing * g.a_stage;
    g.smem_total = g.off_w + kFuRing * g.w_stage;
    int cols = 32;
    while (cols < C) cols <<= 1;
    g.tmem_cols = cols;
    return g.smem_total <= kPwSmemLimit;
}

}  // namespace

bool tcn_pw_eligible(int C) {
    static const bool off = getenv("MISO_TCN_PW") && atoi(getenv("MISO_TCN_PW")) == 0;
    PwGeom g;
    return !off && make_pw_geom(C, 3, g);
}

void tcn_pw_scratch_need(int C, int nconv, size_t *wimg_bytes, size_t *wvec_bytes) {
    *wimg_bytes = (size_t)nconv * C * C * 2 * 2;  // hi + lo bf16
    *wvec_bytes = (size_t)nconv * 2 * C * sizeof(float);
}

int tcn_pw_init() {
    static bool done_dev[64] = {};  // per device ordinal: function attributes live in the device's context
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
    bool &done = done_dev[dev & 63];
    if (done) return MISO_OK;
    e = cudaFuncSetAttribute(tcn_pw_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPwSmemLimit);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tcn_pw_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPwSmemLimit);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(tcn_pw_kernel)");
    done = true;
    return MISO_OK;
}

int launch_tcn_wprep(const TcnPwTable &tab, int nconv, int C, int cpad, int split, void *wimg, float *wvec, cudaStream_t stream, int transposed) {
    PwGeom g;
    MISO_REQUIRE(make_pw_geom(C, split, g), "tcn_pw: unsupported channel count %d", C);
    MISO_REQUIRE(nconv <= kTcnMaxPw, "tcn_pw: too many pointwise convs (%d)", nconv);
    dim3 grid(g.nunit + (C + 255) / 256, nconv);
    tcn_wprep_kernel<<<grid, 256, 0, stream>>>(tab, reinterpret_cast<__nv_bfloat16 *>(wimg), wvec, C, cpad, g.Nt, g.nsp, transposed);
    MISO_LAUNCHED("tcn_wprep_kernel");
    return MISO_OK;
}

int launch_tcn_pw(const TcnPwArgs &p, int split, cudaStream_t stream) {
    PwGeom g;
    MISO_REQUIRE(make_pw_geom(p.C, split, g), "tcn_pw: unsupported channel count %d", p.C);
    PwEncodeFn enc = pw_get_encode();
    if (!enc) {
        set_error("tcn_pw: cuTensorMapEncodeTiled is not available from the driver");
        return MISO_E_CUDA;
    }
    int rc = tcn_pw_init();
    if (rc) return rc;
    // A planes [b][hi|lo][C/8][T][8 ch]: dims in 8-byte units {2, T, C/8, B}
    CUtensorMap tm[2];
    for (int sp = 0; sp < 2; ++sp) {
        cuuint64_t dims[4] = {2, (cuuint64_t)p.T, (cuuint64_t)p.C / 8, (cuuint64_t)p.B};
        cuuint64_t strides[3] = {16, (cuuint64_t)p.T * 16, (cuuint64_t)2 * (p.C / 8) * p.T * 16};
        cuuint32_t box[4] = {2, 128, (cuuint32_t)(2 * g.kper), 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        void *addr = const_cast<char *>(reinterpret_cast<const char *>(p.planes)) + (sp ? p.lo_off : 0);
        CUresult r = enc(&tm[sp], CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, addr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("tcn_pw: cuTensorMapEncodeTiled failed (%d) for T=%d C=%d", (int)r, p.T, p.C);
            return MISO_E_CUDA;
        }
    }
    PwArgs k{};
    k.g = g;
    k.wimg = reinterpret_cast<const __nv_bfloat16 *>(p.wimg) + (size_t)p.index * p.C * p.C * g.nsp;
    k.wbeta = p.wvec + (size_t)p.index * 2 * p.C;
    k.wgamma = k.wbeta + p.C;
    k.gln_sums = p.gln_sums;
    k.gln_inv_n = p.gln_inv_n;
    k.gln_eps = p.gln_eps;
    k.out = p.out;
    k.resid = p.resid;
    k.out_sums = p.out_sums;
    k.out_planes = p.out_planes;
    k.out_ctot = p.out_ctot;
    k.use_lo = p.use_lo;
    k.out_lo_off = p.out_lo_off;
    k.plain = p.plain;
    k.B = p.B;
    k.T = p.T;
    k.C = p.C;
    k.t_tiles = (p.T + 127) / 128;
    dim3 grid(p.B * k.t_tiles, g.NH);
    prof_begin(stream);
    if (split == 3)
        MISO_CUDA(launch_pdl(tcn_pw_kernel<3>, grid, dim3(kPwThreads), (size_t)g.smem_total, stream, tm[0], tm[1], k));
    else
        MISO_CUDA(launch_pdl(tcn_pw_kernel<1>, grid, dim3(kPwThreads), (size_t)g.smem_total, stream, tm[0], tm[1], k));
    prof_end(stream, 2.0 * p.B * p.T * (double)p.C * p.C, (double)p.B * p.T * p.C * ((split == 3 ? 4.0 : 2.0) + 4.0 + (p.resid ? 4.0 : 0.0)),
             MISO_PROF_TCN, (double)p.B * k.t_tiles * g.NH * g.nunit * (split == 3 ? 3.0 : 1.0) * 2.0 * 128.0 * g.Nt * 16.0);
    MISO_LAUNCHED("tcn_pw_kernel");
    return MISO_OK;
}


bool tcn_fused_eligible(int C, int T, int nhalf) {
    // Opt-in (MISO_TCN_FUSED=1): correct (parity-tested) but not faster than the launch-per-half path on B200 -- 1.02 ms against
    // 1.03 ms for the 28 half-blocks at B = 16 (profiles/r2_tcn_fused_*): 64 CTAs cannot keep enough L2 loads of the state in
    // flight for the depthwise half (2.6-3.6 k cycles per 32-channel chunk, the same with 16 warps), and every half-block pays
    // ~11 k cycles of barrier / statistics round trips.
    static const bool on = getenv("MISO_TCN_FUSED") && atoi(getenv("MISO_TCN_FUSED")) != 0;
    FuGeom g;
    const int t_tiles = (T + 127) / 128;
    return on && nhalf <= kTcnMaxHalf && t_tiles <= 8 && make_fu_geom(C, 3, g);
}

namespace {
long long *g_fu_trace = nullptr;
}
void tcn_fused_set_trace(long long *d_buf) { g_fu_trace = d_buf; }

int launch_tcn_fused(const TcnFusedArgs &a_in, int split, cudaStream_t stream) {
    TcnFusedArgs a = a_in;
    a.trace = g_fu_trace;
    FuGeom g;
    MISO_REQUIRE(make_fu_geom(a.C, split, g) && a.nhalf <= kTcnMaxHalf, "tcn_fused: unsupported shape (C=%d, %d half-blocks)", a.C, a.nhalf);
    const int t_tiles = (a.T + 127) / 128;
    MISO_REQUIRE(t_tiles <= 8, "tcn_fused: %d frames need more than a portable cluster", a.T);
    static bool done_dev[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
    if (!done_dev[dev & 63]) {
        e = cudaFuncSetAttribute(tcn_fused_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPwSmemLimit);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(tcn_fused_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPwSmemLimit);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(tcn_fused_kernel)");
        done_dev[dev & 63] = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(a.B * t_tiles);
    cfg.blockDim = dim3(kFuThreads);
    cfg.dynamicSmemBytes = (size_t)g.smem_total;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = t_tiles;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    prof_begin(stream);
    if (split == 3)
        MISO_CUDA(cudaLaunchKernelEx(&cfg, tcn_fused_kernel<3>, a, g));
    else
        MISO_CUDA(cudaLaunchKernelEx(&cfg, tcn_fused_kernel<1>, a, g));
    prof_end(stream, 2.0 * a.nhalf * a.B * a.T * (double)a.C * a.C, (double)a.nhalf * a.B * a.T * a.C * 12.0, MISO_PROF_TCN,
             (double)a.nhalf * a.B * t_tiles * g.NH * g.nunit * (split == 3 ? 3.0 : 1.0) * 2.0 * 128.0 * g.Nt * 16.0);
    MISO_LAUNCHED("tcn_fused_kernel");
    return MISO_OK;
}

}  // namespace miso
