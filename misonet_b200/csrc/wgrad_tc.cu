// tcgen05 weight gradient of the DenseBlock convs (3x3, stride 1, pad (1,1): 90 % of the training FLOPs; the backward of
// model.py:437-482 inside trainer.py:207's loss.backward()).
//
//     dW[kt][kf][ci][co] = sum_b sum_{t,f} xn_b[t + kt - 1][f + kf - 1][ci] * dy_b[t][f][co],
//     xn = x * rstd_b[ci] + shift_b[ci] inside the tensor, 0 outside (zero padding AFTER the InstanceNorm, model.py:445)
//  =  sum_b rstd_b[ci] * G_b[kt][kf][ci][co]  +  shift_b[ci] * S_b[kt][kf][co]
//     G_b = sum over pixels of RAW x (TMA zero fill = the padding) times dy;  S_b = sum of dy over the pixels whose tap
//     source lies inside the tensor.
// The per-sample affine factors out of the pixel sum, so the GEMM runs over the raw bf16 hi/lo planes exactly as they sit
// in HBM -- no register pass (the mma.sync kernels this replaces were bound by load -> normalise -> split -> shared memory).
//
// GEMM roles (K = pixels, both operands MN-major: a plane stores consecutive pixels of 8 channels as consecutive 16-byte
// rows, which IS the MN-major no-swizzle core-matrix layout; tools/umma_mn_test.cu, tools/wgrad_tc_proto.cu):
//   A [M = 128][K = 16 pixels] = dy planes of THREE consecutive output frames stacked along M (frame r-1, r, r+1 of the
//       input frame r <-> kt = 2, 1, 0), 3 * cout <= 128 rows; the bin tap kf is a 16-byte shift of the start address;
//   B [N <= 160][K]            = x planes of input frame r (N = input channels of the chunk) plus ONE EXTRA 8-channel group
//       holding the constant 1 (a [T][F] "ones" tensor whose TMA zero fill is the in-bounds indicator): its accumulator
//       column is S_b for free;
//   D [kf = 0..2][128 lanes = (kt, co)][N columns = ci | ones] in tensor memory (3 N <= 480 columns), fp32.
// bf16x3: a_hi b_hi + a_hi b_lo + a_lo b_hi.  An MMA of width N >= 128 costs N / 2 cycles (tensor bound), so one input
// frame of 128 pixels and 152 channels is 3 kf x 8 K-steps x 3 products x 80 cycles against ~2 k cycles of TMA traffic.
//
// The same kernel covers the other 3x3 layers of the stack through a per-tap table (WtcTaps): with K running over the
// pixels of the COARSER of the two tensors,
//   conv stride 1, pad p (p = 1: DenseBlocks, p = 0: the first layer)   dy bin = x bin - kf + p     A shifted by 2 - kf pixels
//   transposed conv stride 1 (the last layer)                           dy bin = x bin + kf - p     A shifted by kf pixels
//   conv stride (1,2), pad 0 (encoder down-sampling)                    x bin = 2 dy bin + kf       B = even / odd / even + 1
//   transposed conv stride (1,2), pad 0 (decoder up-sampling)           dy bin = 2 x bin + kf       A = even / odd / even + 1
// where even / odd are two tiles loaded with a TMA element stride of 2 along the bins.  For the transposed layers the
// three stacked dy frames r-1, r, r+1 of input frame r are the taps kt = 0, 1, 2 (kt = 2, 1, 0 for the convs).
//
// A CTA owns a frame range of ONE sample (the affine is per sample) and an (input-channel chunk, output-channel chunk)
// slice; it writes its raw accumulators to a partial buffer, and wgrad_tc_reduce_kernel applies rstd / shift and sums
// the CTAs in a fixed order into the torch-layout gradient (deterministic: no atomics).
#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "bwd.cuh"
#include "conv.cuh"
#include "umma.cuh"

namespace miso {
namespace {

constexpr int kWtcThreads = 6 * 32;  // warp 0: TMA producer, warp 1: MMA issuer, warps 2..5: epilogue (128 TMEM lanes)
constexpr int kWtcSmemLimit = 227 * 1024;
constexpr int kWtcMaxStages = 4;
constexpr int kWtcMaxN = 160;

struct WtcTaps {      // per bin tap kf: which tile (even / odd set) and how many pixels its start address is shifted
    int aset[3], ashift[3], bset[3], bshift[3];
    int nAset, nBset;             // tiles per operand (2: even and odd bins, TMA element stride 2)
    int a_mul, a_org[2];          // first dy bin of a stage's tile = a_mul * h * Kpx + a_org[set]
    int b_mul, b_org[2];          // first x bin ...
};

struct WtcGeom {
    int Kpx;      // K pixels per stage (64 / 32 / 16)
    int nhalf;    // stages per frame
    int ngA;      // 8-channel groups of the output-channel chunk
    int ngE;      // 8-channel groups of the input-channel chunk
    int N;        // MMA N: 8 * (ngE + 1 ones group) rounded up to 16
    int pitchA;   // pixels per dy row in shared memory (Kpx + the largest A shift)
    int pitchB;   // pixels per x row (Kpx + the largest B shift)
    int a_box;    // bytes of one dy tile: 3 frames x ngA groups x pitchA x 16
    int a_set;    // ... rounded up to 128 (TMA destination alignment)
    int b_set;    // bytes of one x tile: N / 8 groups x pitchB x 16
    int a_sp, b_sp;  // distance between the hi and the lo tiles of an operand (all its sets)
    int off_b, stage, nstage, smem_total;
    WtcTaps tp;
};

struct WtcArgs {
    WtcGeom g;
    float *partial;   // [cta][3 kf][128][N] fp32
    int B, T;
    int nper, fper;   // CTAs per sample, frames per CTA
    int pe0, pa0;     // first x / dy plane group of the slice
    int rows;         // valid accumulator lanes (3 * output channels of the chunk)
};

struct WtcMaps {      // [hi | lo][set]
    CUtensorMap y[2][2], e[2][2], ones[2];
};

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *tm, uint32_t bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
                 "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
// kind::f16 instruction descriptor: D fp32, A / B bf16, BOTH MN-major (bits 15, 16), N >> 3 at 17, M = 128
__device__ __forceinline__ uint32_t make_idesc_mn(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}

template <int SPLIT>
__global__ void __launch_bounds__(kWtcThreads, 1) wgrad_tc_kernel(const __grid_constant__ WtcMaps tm, const WtcArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const WtcGeom &g = a.g;
    const WtcTaps &tp = g.tp;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int NSP = SPLIT == 3 ? 2 : 1;
    const uint32_t bar_full = smem_u32(smem), bar_empty = smem_u32(smem + 64), bar_done = smem_u32(smem + 128);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + 160);
    const uint32_t s_stage = smem_u32(smem + 1024);

    if (tid == 0) {
        for (int s = 0; s < g.nstage; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        mbar_init(bar_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // the x tiles' groups behind the real ones (the ones group of the lo tiles, the padding group up to N) are never loaded:
    // zero the whole stage area once
    for (int i = tid; i < g.nstage * g.stage / 16; i += kWtcThreads) reinterpret_cast<uint4 *>(smem + 1024)[i] = make_uint4(0u, 0u, 0u, 0u);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    const int cta = blockIdx.x;  // b * nper + q
    const int b = cta / a.nper, q = cta - b * a.nper;
    const int r0 = q * a.fper, r1 = min(a.T, r0 + a.fper);
    const int nst = max(0, r1 - r0) * g.nhalf;  // stages of this CTA
    const int eb = g.ngE * g.pitchB * 16;        // bytes of the x box of one tile

    if (warp == 0) {
        if (elect_one()) {
            int s = 0, ph = 0;
            bool primed = false;
            for (int it = 0; it < nst; ++it) {
                const int r = r0 + it / g.nhalf, h = it - (it / g.nhalf) * g.nhalf;
                if (primed) mbar_wait(bar_empty + 8 * s, (uint32_t)ph);
                const uint32_t full = bar_full + 8 * s;
                const uint32_t sa = s_stage + (uint32_t)(s * g.stage);
                mbar_expect_tx(full, (uint32_t)(NSP * (tp.nAset * g.a_box + tp.nBset * eb) + tp.nBset * g.pitchB * 16));
                for (int sp = 0; sp < NSP; ++sp) {
                    // dy: {8 ch, bins, groups, frames, sample} -> [frame][group][pixel]: frames r-1, r, r+1
                    for (int st = 0; st < tp.nAset; ++st)
                        tma_load_5d(sa + (uint32_t)(sp * g.a_sp + st * g.a_set), &tm.y[sp][st], full, 0, tp.a_mul * h * g.Kpx + tp.a_org[st], a.pa0, r - 1,
                                    b);
                    // x: {8 ch, bins, frames, groups, sample} -> [group][pixel]
                    for (int st = 0; st < tp.nBset; ++st)
                        tma_load_5d(sa + (uint32_t)(g.off_b + sp * g.b_sp + st * g.b_set), &tm.e[sp][st], full, 0, tp.b_mul * h * g.Kpx + tp.b_org[st], r,
                                    a.pe0, b);
                }
                // ones: {8 ch, bins, frames} -> the group behind the x groups of the hi tiles
                for (int st = 0; st < tp.nBset; ++st)
                    tma_load_3d(sa + (uint32_t)(g.off_b + st * g.b_set + eb), &tm.ones[st], full, 0, tp.b_mul * h * g.Kpx + tp.b_org[st], r);
                if (++s == g.nstage) {
                    s = 0;
                    if (primed) ph ^= 1;
                    primed = true;
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            const uint32_t idesc = make_idesc_mn(g.N);
            const uint32_t sboA = (uint32_t)(g.pitchA * 16), sboB = (uint32_t)(g.pitchB * 16);
            int s = 0, ph = 0;
            for (int it = 0; it < nst; ++it) {
                mbar_wait(bar_full + 8 * s, (uint32_t)ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa = s_stage + (uint32_t)(s * g.stage);
#pragma unroll
                for (int kf = 0; kf < 3; ++kf) {
                    const uint32_t a0 = sa + (uint32_t)(tp.aset[kf] * g.a_set + tp.ashift[kf] * 16);
                    const uint32_t b0 = sa + (uint32_t)(g.off_b + tp.bset[kf] * g.b_set + tp.bshift[kf] * 16);
                    for (int ks = 0; ks < g.Kpx / 16; ++ks) {
#pragma unroll
                        for (int pr = 0; pr < (SPLIT == 3 ? 3 : 1); ++pr) {  // a_hi b_hi, a_hi b_lo, a_lo b_hi
                            const uint32_t aaddr = a0 + (uint32_t)((pr == 2 ? g.a_sp : 0) + ks * 256);
                            const uint32_t baddr = b0 + (uint32_t)((pr == 1 ? g.b_sp : 0) + ks * 256);
                            umma_bf16(tmem + (uint32_t)(kf * g.N), make_desc(aaddr, 128, sboA), make_desc(baddr, 128, sboB), idesc,
                                      (it == 0 && ks == 0 && pr == 0) ? 0u : 1u);
                        }
                    }
                }
                umma_commit(bar_empty + 8 * s);
                if (++s == g.nstage) {
                    s = 0;
                    ph ^= 1;
                }
            }
            umma_commit(bar_done);
        }
    } else {
        // epilogue: raw accumulators -> partial[cta][kf][lane][N]
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        float *dst = a.partial + (size_t)cta * 3 * 128 * g.N;
        if (nst > 0) {
            mbar_wait(bar_done, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        for (int kf = 0; kf < 3; ++kf) {
            for (int cb = 0; cb < g.N; cb += 16) {
                uint32_t v[16];
                if (nst > 0) {
                    tmem_ld16(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(kf * g.N + cb), v);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = 0u;
                }
                if (row < a.rows) {
                    float4 *o = reinterpret_cast<float4 *>(dst + ((size_t)kf * 128 + row) * g.N + cb);
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                        o[j >> 2] = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

// dW[co][ci][kt][kf] += sum over the slice's CTAs (fixed order) of rstd_b[ci] * P[kf][(2 - kt) * co_n + co][ci] +
// shift_b[ci] * P[kf][..][ones column]
struct WtcReduceArgs {
    const float *partial;
    float *dw;
    const double *x_sums;  // [B][x_ctot][2] or null
    int x_ctot, x_coff;
    double inv_n;
    float eps;
    int B, nper, ncta;     // CTAs per slice = B * nper
    int cin, cout_real;    // layer sizes (torch layout strides)
    int ci0, ci_n, co0, co_n, N, slice;
    int rev;               // 0: frame block j of the accumulator rows is the tap kt = 2 - j (convs); 1: kt = j (transposed convs)
    int transposed;        // gradient layout: Conv2d [cout][cin][3][3] or ConvTranspose2d [cin][cout][3][3]
};
__global__ void __launch_bounds__(256) wgrad_tc_reduce_kernel(const WtcReduceArgs a) {
    extern __shared__ float2 aff[];  // [B][ci_n]
    for (int i = threadIdx.x; i < a.B * a.ci_n; i += blockDim.x) {
        const int b = i / a.ci_n, ci = i - b * a.ci_n;
        float2 af = make_float2(1.f, 0.f);
        if (a.x_sums && a.ci0 + ci < a.cin) {
            const double *sp = a.x_sums + ((size_t)b * a.x_ctot + a.x_coff + a.ci0 + ci) * 2;
            af = affine_from_sums(stat_get(sp), stat_get(sp + 1), a.inv_n, (double)a.eps);
        }
        aff[i] = af;
    }
    __syncthreads();
    // 64 outputs per block (consecutive input channels: coalesced partial reads) x 4 threads per output, each summing a
    // contiguous quarter of the CTAs with eight loads in flight; the quarters are added in a fixed order
    __shared__ float part[4][64];
    const int total = 9 * a.co_n * a.ci_n;
    const float *P = a.partial + (size_t)a.slice * a.ncta * 3 * 128 * a.N;
    const size_t cstride = (size_t)3 * 128 * a.N;
    const int lane64 = threadIdx.x & 63, q4 = threadIdx.x >> 6;
    const int per = (a.ncta + 3) / 4;
    const int c_lo = q4 * per, c_hi = min(a.ncta, c_lo + per);
    for (int o0 = blockIdx.x * 64; o0 < total; o0 += gridDim.x * 64) {
        const int o = o0 + lane64;
        float acc = 0.f;
        int ci = 0, co = 0, kf = 0, kt = 0;
        bool live = o < total;
        if (live) {
            ci = o % a.ci_n;
            int r = o / a.ci_n;
            co = r % a.co_n;
            r /= a.co_n;
            kf = r % 3;
            kt = r / 3;
            live = a.co0 + co < a.cout_real && a.ci0 + ci < a.cin;
        }
        if (live) {
            const int row = (a.rev ? kt : 2 - kt) * a.co_n + co;
            const float *p0 = P + ((size_t)kf * 128 + row) * a.N;
            int c = c_lo;
            for (; c + 8 <= c_hi; c += 8) {
                float g[8], s1[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const float *p = p0 + (size_t)(c + u) * cstride;
                    g[u] = p[ci];
                    s1[u] = p[a.ci_n];
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const float2 af = aff[((c + u) / a.nper) * a.ci_n + ci];
                    acc = fmaf(af.x, g[u], fmaf(af.y, s1[u], acc));
                }
            }
            for (; c < c_hi; ++c) {
                const float *p = p0 + (size_t)c * cstride;
                const float2 af = aff[(c / a.nper) * a.ci_n + ci];
                acc = fmaf(af.x, p[ci], fmaf(af.y, p[a.ci_n], acc));
            }
        }
        part[q4][lane64] = acc;
        __syncthreads();
        if (q4 == 0 && live) {
            const float v = ((part[0][lane64] + part[1][lane64]) + part[2][lane64]) + part[3][lane64];
            const size_t idx = a.transposed ? ((size_t)(a.ci0 + ci) * a.cout_real + a.co0 + co) : ((size_t)(a.co0 + co) * a.cin + a.ci0 + ci);
            a.dw[(idx * 3 + kt) * 3 + kf] += v;
        }
        __syncthreads();
    }
}

__global__ void wtc_ones_kernel(__nv_bfloat16 *ones, size_t npix) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < npix * 8; i += (size_t)gridDim.x * blockDim.x)
        ones[i] = __float2bfloat16((i & 7) == 0 ? 1.f : 0.f);
}

typedef CUresult (*WtcEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
WtcEncodeFn wtc_get_encode() {
    static WtcEncodeFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<WtcEncodeFn>(p);
    }
    return fn;
}

int wtc_round_up(int x, int m) { return (x + m - 1) / m * m; }

enum WtcCase { WTC_NONE = 0, WTC_CONV_S1, WTC_CONVT_S1, WTC_CONV_S2, WTC_CONVT_S2 };

int wtc_case(const WgradArgs &a) {
    if (a.KT != 3 || a.KF != 3 || a.pad_t != 1) return WTC_NONE;
    if (!a.transposed && a.stride_f == 1 && (a.pad_f == 0 || a.pad_f == 1) && a.Fout == a.Fin + 2 * a.pad_f - 2) return WTC_CONV_S1;
    if (a.transposed && a.stride_f == 1 && (a.pad_f == 0 || a.pad_f == 1) && a.Fout == a.Fin + 2 - 2 * a.pad_f) return WTC_CONVT_S1;
    if (!a.transposed && a.stride_f == 2 && a.pad_f == 0 && a.Fin >= 3 && a.Fout == (a.Fin - 3) / 2 + 1) return WTC_CONV_S2;
    if (a.transposed && a.stride_f == 2 && a.pad_f == 0 && a.Fout == (a.Fin - 1) * 2 + 3) return WTC_CONVT_S2;
    return WTC_NONE;
}

// Fk: extent of the K (pixel) grid = bins of the coarser tensor
bool make_wtc_geom(int kase, int pad_f, int Fk, int ci_n, int co_n, int split, WtcGeom &g) {
    g = WtcGeom{};
    const int nsp = split == 3 ? 2 : 1;
    g.Kpx = Fk + 1 >= 64 ? 64 : (Fk + 1 >= 32 ? 32 : 16);
    g.nhalf = (Fk + g.Kpx - 1) / g.Kpx;
    g.ngA = co_n / 8;
    g.ngE = (ci_n + 7) / 8;
    g.N = wtc_round_up(8 * (g.ngE + 1), 16);
    if (g.N > kWtcMaxN || 3 * co_n > 128) return false;
    WtcTaps &tp = g.tp;
    tp.nAset = tp.nBset = 1;
    tp.a_mul = tp.b_mul = 1;
    g.pitchA = g.pitchB = g.Kpx;
    switch (kase) {
        case WTC_CONV_S1:  // dy bin = x bin - kf + pad: the tile starts pad - 2 bins left of the stage
            tp.a_org[0] = pad_f - 2;
            for (int kf = 0; kf < 3; ++kf) tp.ashift[kf] = 2 - kf;
            g.pitchA = g.Kpx + 2;
            break;
        case WTC_CONVT_S1:  // dy bin = x bin + kf - pad
            tp.a_org[0] = -pad_f;
            for (int kf = 0; kf < 3; ++kf) tp.ashift[kf] = kf;
            g.pitchA = g.Kpx + 2;
            break;
        case WTC_CONV_S2:  // x bin = 2 dy bin + kf: even, odd, even + 1
            tp.nBset = 2;
            tp.b_mul = 2;
            tp.b_org[0] = 0;
            tp.b_org[1] = 1;
            tp.bset[1] = 1;
            tp.bshift[2] = 1;
            g.pitchB = g.Kpx + 8;  // a multiple of 8 pixels keeps every group 128-byte aligned (the ones group is a TMA target)
            break;
        case WTC_CONVT_S2:  // dy bin = 2 x bin + kf
            tp.nAset = 2;
            tp.a_mul = 2;
            tp.a_org[0] = 0;
            tp.a_org[1] = 1;
            tp.aset[1] = 1;
            tp.ashift[2] = 1;
            g.pitchA = g.Kpx + 1;
            break;
        default:
            return false;
    }
    g.a_box = 3 * g.ngA * g.pitchA * 16;
    g.a_set = wtc_round_up(g.a_box, 128);
    g.a_sp = tp.nAset * g.a_set;
    g.b_set = g.N / 8 * g.pitchB * 16;
    g.b_sp = tp.nBset * g.b_set;
    g.off_b = wtc_round_up(nsp * g.a_sp, 128);
    // an A descriptor reads 16 groups of pitchA pixels whatever 3 * ngA is: keep that inside the stage
    const int a_reach = (nsp * g.a_sp - g.a_set) + 16 * g.pitchA * 16 + 3 * 16 + (g.Kpx / 16) * 256;
    g.stage = wtc_round_up(std::max(g.off_b + nsp * g.b_sp, a_reach), 1024);
    g.nstage = std::min(kWtcMaxStages, (kWtcSmemLimit - 1024) / g.stage);
    if (g.nstage < 2) return false;
    g.smem_total = 1024 + g.nstage * g.stage;
    return true;
}

}  // namespace

bool wgrad_tc_eligible(const WgradArgs &a) {
    static const bool off = getenv("MISO_WGRAD_TC") && atoi(getenv("MISO_WGRAD_TC")) == 0;
    static const bool dense_only = getenv("MISO_WGRAD_TC") && atoi(getenv("MISO_WGRAD_TC")) == 2;  // A/B: DenseBlock convs only
    if (off || !a.dyp || !a.partial || !a.ones) return false;
    const int kase = wtc_case(a);
    if (kase == WTC_NONE) return false;
    if (dense_only && !(kase == WTC_CONV_S1 && a.pad_f == 1)) return false;
    if (a.x_layout != LAYOUT_PLANES || a.x_ctot % 8 || a.x_coff % 8 || a.cout % 8 || a.cout_real > a.cout) return false;  // cout_real < cout: zero pad channels (the network output)
    if (a.cin % 8 && a.x_coff + ((a.cin + 7) & ~7) > a.x_ctot) return false;  // a ragged channel count needs zero planes behind it
    const int Fk = kase == WTC_CONV_S2 ? a.Fout : a.Fin;
    if (Fk < 15 || a.T < 2) return false;
    return true;
}

// chunking of a layer: input-channel chunks of <= 152 channels (N <= 160 with the ones group), output-channel chunks of <= 40
static void wtc_chunks(const WgradArgs &a, int &nci, int &ci_n, int &nco, int &co_n) {
    const int cin8 = (a.cin + 7) & ~7;
    nci = (cin8 + 151) / 152;
    ci_n = wtc_round_up((cin8 + nci - 1) / nci, 8);
    nco = (a.cout + 39) / 40;
    co_n = wtc_round_up((a.cout + nco - 1) / nco, 8);
}

size_t wgrad_tc_partial_bytes(int B) {
    // at most ~2 waves of CTAs, 3 x 128 x 160 floats each
    return (size_t)(2 * 148 + 4 * std::max(B, 1)) * 3 * 128 * kWtcMaxN * sizeof(float);
}

int wgrad_tc_fill_ones(__nv_bfloat16 *ones, size_t npix, cudaStream_t st) {
    wtc_ones_kernel<<<256, 256, 0, st>>>(ones, npix);
    MISO_LAUNCHED("wtc_ones_kernel");
    return MISO_OK;
}

int launch_wgrad_tc(const WgradArgs &a, cudaStream_t st) {
    WtcEncodeFn enc = wtc_get_encode();
    if (!enc) {
        set_error("wgrad_tc: cuTensorMapEncodeTiled is not available from the driver");
        return MISO_E_CUDA;
    }
    const int split = a.use_lo ? 3 : 1;
    int nci, ci_n, nco, co_n;
    wtc_chunks(a, nci, ci_n, nco, co_n);
    const int nslice = nci * nco;
    // slices are launches of their own (one after the other), so every launch gets the whole machine: one wave of CTAs
    const int nper = std::max(1, std::min(a.T / 2, std::max(1, 148 / a.B)));
    const int fper = (a.T + nper - 1) / nper;
    const int ncta = a.B * nper;
    // a.ones: bf16 pixels [npix][8] with channel 0 = 1 (wgrad_tc_fill_ones); uniform, so any [T][F] view of it works
    __nv_bfloat16 *ones = a.ones;
    int dev = 0;
    MISO_CUDA(cudaGetDevice(&dev));
    static bool attr_done[64] = {};
    if (!attr_done[dev & 63]) {
        MISO_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWtcSmemLimit));
        MISO_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWtcSmemLimit));
        attr_done[dev & 63] = true;
    }
    const int kase = wtc_case(a);
    const int Fk = kase == WTC_CONV_S2 ? a.Fout : a.Fin;
    const int cin8 = (a.cin + 7) & ~7;
    const uint64_t T = (uint64_t)a.T, Fx = (uint64_t)a.Fin, Fy = (uint64_t)a.Fout;
    for (int ic = 0; ic < nci; ++ic)
        for (int oc = 0; oc < nco; ++oc) {
            const int ci0 = ic * ci_n, cin_c = std::min(ci_n, cin8 - ci0);
            const int co0 = oc * co_n, con_c = std::min(co_n, a.cout - co0);
            WtcGeom g;
            MISO_REQUIRE(make_wtc_geom(kase, a.pad_f, Fk, cin_c, con_c, split, g), "wgrad_tc: geometry (cin %d cout %d F %d)", cin_c, con_c, Fk);
            const int slice = ic * nco + oc;
            // a slice's GEMM and its reduction are consecutive launches of one stream: every slice reuses the same region
            MISO_REQUIRE((size_t)ncta * 3 * 128 * g.N * sizeof(float) <= a.partial_bytes, "wgrad_tc: partial buffer too small");
            WtcMaps tm;
            memset(&tm, 0, sizeof(tm));
            auto fail = [&](const char *what, CUresult r) {
                set_error("wgrad_tc: cuTensorMapEncodeTiled(%s) failed (%d)", what, (int)r);
                return MISO_E_CUDA;
            };
            for (int sp = 0; sp < 2; ++sp) {
                // x planes [B][hi|lo][x_ctot/8][T][Fx][8]: {8 ch, bins, frames, groups, samples}
                for (int st = 0; st < g.tp.nBset; ++st) {
                    const uint64_t CG = (uint64_t)a.x_ctot / 8;
                    char *base = const_cast<char *>(reinterpret_cast<const char *>(a.x)) + (sp ? CG * T * Fx * 16 : 0);
                    cuuint64_t dims[5] = {8, Fx, T, CG, (cuuint64_t)a.B};
                    cuuint64_t strides[4] = {16, Fx * 16, T * Fx * 16, 2 * CG * T * Fx * 16};
                    cuuint32_t box[5] = {8, (cuuint32_t)(g.tp.b_mul * g.pitchB), 1, (cuuint32_t)g.ngE, 1};
                    cuuint32_t es[5] = {1, (cuuint32_t)g.tp.b_mul, 1, 1, 1};
                    CUresult r = enc(&tm.e[sp][st], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                    if (r != CUDA_SUCCESS) return fail("x", r);
                }
                // dy planes [B][hi|lo][cout/8][T][Fy][8] as {8 ch, bins, groups, frames, samples}: the box lands [frame][group][pixel]
                for (int st = 0; st < g.tp.nAset; ++st) {
                    const uint64_t CG = (uint64_t)a.cout / 8;
                    char *base = reinterpret_cast<char *>(a.dyp) + (sp ? CG * T * Fy * 16 : 0);
                    cuuint64_t dims[5] = {8, Fy, CG, T, (cuuint64_t)a.B};
                    cuuint64_t strides[4] = {16, T * Fy * 16, Fy * 16, 2 * CG * T * Fy * 16};
                    cuuint32_t box[5] = {8, (cuuint32_t)(g.tp.a_mul * g.pitchA), (cuuint32_t)g.ngA, 3, 1};
                    cuuint32_t es[5] = {1, (cuuint32_t)g.tp.a_mul, 1, 1, 1};
                    CUresult r = enc(&tm.y[sp][st], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                    if (r != CUDA_SUCCESS) return fail("dy", r);
                }
            }
            for (int st = 0; st < g.tp.nBset; ++st) {
                cuuint64_t dims[3] = {8, Fx, T};
                cuuint64_t strides[2] = {16, Fx * 16};
                cuuint32_t box[3] = {8, (cuuint32_t)(g.tp.b_mul * g.pitchB), 1};
                cuuint32_t es[3] = {1, (cuuint32_t)g.tp.b_mul, 1};
                CUresult r = enc(&tm.ones[st], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, ones, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) return fail("ones", r);
            }
            WtcArgs k{};
            k.g = g;
            k.partial = a.partial;
            k.B = a.B;
            k.T = a.T;
            k.nper = nper;
            k.fper = fper;
            k.pe0 = (a.x_coff + ci0) / 8;
            k.pa0 = co0 / 8;
            k.rows = 3 * con_c;
            static const bool debug = getenv("MISO_TC_DEBUG") != nullptr;
            if (debug)
                fprintf(stderr, "wgrad_tc: case %d cin=%d cout=%d Fx=%d Fy=%d | slice %d/%d ci [%d,+%d) co [%d,+%d) N=%d Kpx=%d nhalf=%d nper=%d fper=%d nstage=%d stage=%dB\n",
                        kase, a.cin, a.cout, a.Fin, a.Fout, slice, nslice, ci0, cin_c, co0, con_c, g.N, g.Kpx, g.nhalf, nper, fper, g.nstage, g.stage);
            if (split == 3)
                wgrad_tc_kernel<3><<<dim3(ncta, 1), kWtcThreads, g.smem_total, st>>>(tm, k);
            else
                wgrad_tc_kernel<1><<<dim3(ncta, 1), kWtcThreads, g.smem_total, st>>>(tm, k);
            MISO_LAUNCHED("wgrad_tc_kernel");
            WtcReduceArgs r{};
            r.partial = a.partial;
            r.dw = a.dw;
            r.x_sums = a.x_sums;
            r.x_ctot = a.x_ctot;
            r.x_coff = a.x_coff;
            r.inv_n = a.inv_n;
            r.eps = a.eps;
            r.B = a.B;
            r.nper = nper;
            r.ncta = ncta;
            r.cin = a.cin;
            r.cout_real = a.cout_real;
            r.ci0 = ci0;
            r.ci_n = g.ngE * 8;
            r.co0 = co0;
            r.co_n = con_c;
            r.N = g.N;
            r.slice = 0;
            r.rev = (kase == WTC_CONVT_S1 || kase == WTC_CONVT_S2) ? 1 : 0;
            r.transposed = a.transposed;
            const int total = 9 * con_c * r.ci_n;
            wgrad_tc_reduce_kernel<<<std::min(8 * 148, (total + 63) / 64), 256, (size_t)a.B * r.ci_n * sizeof(float2), st>>>(r);
            MISO_LAUNCHED("wgrad_tc_reduce_kernel");
        }
    return MISO_OK;
}

}  // namespace miso
