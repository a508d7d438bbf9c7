// tcgen05 implicit-GEMM (de)convolution over bf16 hi/lo activation planes -- the tensor-core
// path of the MISO conv stack (model.py:401-482: 3x3 convs, stride (1,1)/(1,2), and their
// transposed forms).
//
// Formulation.  A CTA owns a [TT frames x TF bins] tile of one sample's output.  Per 16-channel
// chunk it TMA-loads the halo tile [(TT+2) x Wr pixels x 8 channels] of two channel planes
// straight into the canonical no-swizzle K-major UMMA layout (pixel p of plane j at byte
// j*PL + p*16).  With the tile flattened as a raster of width Wr, the im2col matrix of tap
// (kt,kf) is the SAME staged tile shifted by a constant number of rows, so the nine taps are
// nine shared-memory descriptors that differ only in their start address: global traffic is
// (1 + halo) x the input instead of 9x.  Stride-2 convs load the even and odd bins as two tiles
// (TMA element stride 2); stride-2 transposed convs are two output phases over one tile.
//
// InstanceNorm of the producer (an affine per (sample, channel), applied by the consumer with
// zero padding AFTER it, model.py:411-414) is folded into the operands: a tiny prep kernel
// scales the weights per sample (W * rstd) and tabulates the border-dependent bias
// sum_taps-in-bounds W * (-mean * rstd) for the 64 (frame-mask, bin-mask) classes, so raw
// activations go from HBM to the tensor core untouched and TMA zero fill IS the padding.
//
// Pipeline: warp 0 = TMA producer (one lane), warp 1 = tcgen05.mma issuer (one lane, M = 128,
// N = cout tile, K = 16, accumulators in TMEM), warps 2..9 = epilogue (tcgen05.ld -> border bias
// -> ELU -> (sum, sumsq) statistics for the consumers -> bf16 hi/lo split -> 16-byte plane stores).
// SPLIT = 1: bf16 operands (throughput mode).  SPLIT = 3: a_hi*w_hi + a_lo*w_hi + a_hi*w_lo with
// fp32 accumulation (parity-grade).
#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "conv.cuh"

namespace miso {
namespace {

constexpr int kThreads = 320;       // 10 warps: TMA, MMA, 8 epilogue
constexpr int kEpiThreads = 256;
constexpr int kMaxTaps = 9;
constexpr int kMaxStages = 4;
constexpr int kSmemLimit = 227 * 1024;
constexpr int kFixedSmem = 1024;    // barriers, tmem slot

struct TcGeom {
    // tap table
    int ntap;
    int shift[kMaxTaps];   // row shift of the A descriptor, in pixels
    int aset[kMaxTaps];    // which input tile (0 = even / only, 1 = odd bins)
    int acc[kMaxTaps];     // accumulator set (output phase)
    int tapk[kMaxTaps];    // kt * KF + kf of the weight slot
    int first_mask;        // bit i set: tap i is the first contribution to its accumulator set
    int nset, nacc, nsp;
    // tiling
    int Wr, TT, TF, rows, G, N, nNt, ncols;
    int t_tiles, f_tiles, ntiles, nbuf;  // nbuf = 2: the epilogue of tile k overlaps the MMAs of tile k+1
    int t_org, f_org, f_mul;   // input tile origin: t0 + t_org ; f = j0 * f_mul + f_org (+1 for the odd set)
    int ostride;               // fo = j * ostride + acc
    int nchunk, nplanes;
    // shared memory
    int PL, a_stage, w_stage, stage, nstage, box_bytes;
    int off_btab, off_red, off_stage, smem_total;
    int tmem_cols;
    int strided;               // 5-D bf16 tensor map with element stride 2 along bins
    int ntap_for(const ConvArgs &a) const { return a.KT * a.KF; }
};

struct TcArgs {
    TcGeom g;
    const __nv_bfloat16 *wimg;  // [B][nNt][nchunk][ntap][2][nsp*N][8]
    const float *btab;          // [B][nNt][64][N]
    void *out;
    double *out_sums;
    int B, T, Fin, Fout;
    int in_coff;                // channel offset of the input view (multiple of 8)
    int out_ctot, out_coff, cout;
    int out_layout;
    size_t out_lo_off;
    int use_lo;
    int KT, KF, stride_f, pad_t, pad_f, transposed, elu;
};

struct PrepArgs {
    const float *w;        // packed fp32 [ntaps][cin][cout_pad]
    const float *bias;     // [cout_pad] or null
    const double *in_sums;
    int in_ctot, in_coff, cin, cout, cout_pad;
    int norm_mode;
    float norm_eps;
    double norm_inv_n;
    __nv_bfloat16 *wimg;
    float *btab;
    int B, N, nNt, nchunk, nsp, ntap, KT, KF;
    int tapk[kMaxTaps];
};

// ------------------------------------------------------------------------------------------------
// PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *tm, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
        "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap *tm, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
        "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}

// one lane of a converged warp; the compiler then keeps UTMALDG / UTCHMMA on the uniform datapath
// (a plain `lane == 0` branch makes it wrap every such instruction in a serialising loop)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// shared-memory matrix descriptor, SWIZZLE_NONE, K-major: start address [0,14) >>4, leading byte
// offset [16,30) >>4 (between the two 8-element K groups), stride byte offset [32,46) >>4 (between
// 8-row groups), version [46,48) = 1 (sm_100), layout type [61,64) = 0.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// instruction descriptor, kind::f16: D fp32 (bit 4), A/B bf16 (bits 7, 10), K-major both, N>>3 at 17, M>>4 at 24
__device__ __forceinline__ uint32_t make_idesc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}

// 16 per-lane values -> lanes 2k and 2k+1 hold the warp total of value k (16 shuffles)
__device__ __forceinline__ float warp_reduce16(float (&v)[16], int lane) {
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const int off = 16 >> s, half = 8 >> s;
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float send = upper ? v[i] : v[i + half];
            const float keep = upper ? v[i + half] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// ------------------------------------------------------------------------------------------------
// tile id -> (sample, N tile, frame tile, bin tile); bin tiles vary fastest so that consecutive tiles of a CTA
// share the sample's weight image and bias table in L2
struct TileRef {
    int b, nt, t0, j0;
};
__device__ __forceinline__ TileRef decode_tile(const TcGeom &g, int tile) {
    TileRef r;
    const int ft = tile % g.f_tiles;
    int q = tile / g.f_tiles;
    const int tt = q % g.t_tiles;
    q /= g.t_tiles;
    r.nt = q % g.nNt;
    r.b = q / g.nNt;
    r.t0 = tt * g.TT;
    r.j0 = ft * g.TF;
    return r;
}

__device__ __forceinline__ float elu_fast(float x) {
    // ELU(alpha = 1).  exp(x) - 1 through ex2.approx: absolute error ~1e-7, far below the bf16 hi+lo
    // storage step of the activations (the fp32 FMA kernels keep expm1f)
    return x > 0.f ? x : __expf(x) - 1.f;
}

template <int SPLIT>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo, const TcArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const TcGeom &g = a.g;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = g.N;
    constexpr int WB = SPLIT == 3 ? 2 : 1;  // weight rows (and accumulator columns) per output channel
    const int cw = WB * N;                  // accumulator columns per (M tile, phase)
    const int buf_cols = g.G * g.nacc * cw; // accumulator columns per tile

    // barriers: full[s], empty[s] (operand stages); tfull[2], tempty[2] (TMEM accumulator buffers)
    const uint32_t bar_full = smem_u32(smem), bar_empty = smem_u32(smem + 64);
    const uint32_t bar_tfull = smem_u32(smem + 128), bar_tempty = smem_u32(smem + 144);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + 160);
    float *btab_s = reinterpret_cast<float *>(smem + g.off_btab);
    float *red = reinterpret_cast<float *>(smem + g.off_red);
    const uint32_t s_stage = smem_u32(smem + g.off_stage);

    if (tid == 0) {
        for (int s = 0; s < g.nstage; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_tfull + 8 * i, 1);
            mbar_init(bar_tempty + 8 * i, kEpiThreads / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)g.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    {
        // statistics scratch; and zero the operand stages once so that a half-filled last chunk
        // (cin % 16 == 8) never multiplies uninitialised shared memory (later it sees stale finite data)
        for (int i = tid; i < 2 * N; i += kThreads) red[i] = 0.f;
        if (g.nplanes & 1) {
            uint4 *z = reinterpret_cast<uint4 *>(smem + g.off_stage);
            const int n16 = g.nstage * g.stage / 16;
            for (int i = tid; i < n16; i += kThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ---------------------------------------------------------------- TMA producer
        if (elect_one()) {
            const int plane0 = a.in_coff >> 3;
            int q = 0;  // running chunk index over all tiles of this CTA
            for (int tile = blockIdx.x; tile < g.ntiles; tile += gridDim.x) {
                const TileRef tr = decode_tile(g, tile);
                const int tin = tr.t0 + g.t_org;
                const int fin = tr.j0 * g.f_mul + g.f_org;
                const __nv_bfloat16 *wsrc = a.wimg + ((size_t)tr.b * g.nNt + tr.nt) * g.nchunk * (size_t)(g.w_stage / 2);
                for (int c = 0; c < g.nchunk; ++c, ++q) {
                    const int s = q % g.nstage;
                    if (q >= g.nstage) mbar_wait(bar_empty + 8 * s, ((q / g.nstage) + 1) & 1);
                    const int np = min(2, g.nplanes - 2 * c);
                    const uint32_t full = bar_full + 8 * s;
                    mbar_expect_tx(full, (uint32_t)(g.nset * g.nsp * np * g.box_bytes + g.w_stage));
                    const uint32_t sa = s_stage + (uint32_t)(s * g.stage);
                    for (int set = 0; set < g.nset; ++set)
                        for (int sp = 0; sp < g.nsp; ++sp)
                            for (int p = 0; p < np; ++p) {
                                const uint32_t dst = sa + (uint32_t)(((set * g.nsp + sp) * 2 + p) * g.PL);
                                const CUtensorMap *tm = sp == 0 ? &tm_hi : &tm_lo;
                                if (g.strided)
                                    tma_load_5d(dst, tm, full, 0, fin + set, tin, plane0 + 2 * c + p, tr.b);
                                else
                                    tma_load_4d(dst, tm, full, 2 * fin, tin, plane0 + 2 * c + p, tr.b);
                            }
                    bulk_load(sa + (uint32_t)g.a_stage, wsrc + (size_t)c * (g.w_stage / 2), (uint32_t)g.w_stage, full);
                }
            }
        }
    } else if (warp == 1) {
        // ---------------------------------------------------------------- MMA issuer
        // SPLIT == 3: the weight image of a tap is [kg][2N rows: w_hi then w_lo][8], so
        //   D[:, 0:N] , D[:, N:2N] += A_hi * [W_hi | W_lo]   (one MMA of width 2N)
        //   D[:, 0:N]              += A_lo * W_hi             (one MMA of width N)
        // and the epilogue adds the two column halves: 2 MMAs instead of 3, A_hi read once.
        const uint32_t idesc_wide = make_idesc(cw), idesc_n = make_idesc(N);
        const uint32_t a_lbo = (uint32_t)g.PL, b_lbo = (uint32_t)cw * 16;
        const uint32_t gstep = (uint32_t)(cw * g.nacc);
        int q = 0, k = 0;
        for (int tile = blockIdx.x; tile < g.ntiles; tile += gridDim.x, ++k) {
            const int buf = g.nbuf == 2 ? (k & 1) : 0;
            // the epilogue must have drained this accumulator buffer (use u = k / nbuf of it)
            {
                const int u = k / g.nbuf;
                if (u > 0) mbar_wait(bar_tempty + 8 * buf, (u + 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            const uint32_t tbuf = tmem_base + (uint32_t)(buf * buf_cols);
            for (int c = 0; c < g.nchunk; ++c, ++q) {
                const int s = q % g.nstage;
                mbar_wait(bar_full + 8 * s, (q / g.nstage) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (elect_one()) {
                    const uint32_t abase = s_stage + (uint32_t)(s * g.stage);
                    const uint32_t bbase = abase + (uint32_t)g.a_stage;
#pragma unroll 1
                    for (int i = 0; i < g.ntap; ++i) {
                        const uint32_t a_hi = abase + (uint32_t)(g.aset[i] * g.nsp * 2 * g.PL) + (uint32_t)(g.shift[i] * 16);
                        const uint64_t bdesc = make_desc(bbase + (uint32_t)(i * 2 * cw * 16), b_lbo, 128);
                        const uint32_t acc_flag = (c == 0 && ((g.first_mask >> i) & 1)) ? 0u : 1u;
                        const uint32_t d0 = tbuf + (uint32_t)g.acc[i] * cw;
                        uint64_t ad = make_desc(a_hi, a_lbo, 128);
#pragma unroll 4
                        for (int gt = 0; gt < g.G; ++gt) umma_bf16(d0 + gt * gstep, ad + (uint64_t)(gt * 128), bdesc, idesc_wide, acc_flag);
                        if constexpr (SPLIT == 3) {
                            ad = make_desc(a_hi + (uint32_t)(2 * g.PL), a_lbo, 128);
#pragma unroll 4
                            for (int gt = 0; gt < g.G; ++gt) umma_bf16(d0 + gt * gstep, ad + (uint64_t)(gt * 128), bdesc, idesc_n, 1u);
                        }
                    }
                    umma_commit(bar_empty + 8 * s);
                    if (c == g.nchunk - 1) umma_commit(bar_tfull + 8 * buf);
                }
                __syncwarp();
            }
        }
    } else {
        // ---------------------------------------------------------------- epilogue (warps 2..9)
        const int quad = warp & 3, half = (warp - 2) >> 2;
        const int et = tid - 64;
        const int npix = a.T * a.Fout;
        const bool planes = a.out_layout == LAYOUT_PLANES;
        int prev_b = -1, prev_nt = -1, k = 0;
        for (int tile = blockIdx.x; tile < g.ntiles; tile += gridDim.x, ++k) {
            const TileRef tr = decode_tile(g, tile);
            const int b = tr.b, t0 = tr.t0, j0 = tr.j0;
            const int buf = g.nbuf == 2 ? (k & 1) : 0;
            if (b != prev_b || tr.nt != prev_nt) {
                // border-bias table of this (sample, N tile)
                asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads));
                const float4 *src = reinterpret_cast<const float4 *>(a.btab + ((size_t)b * g.nNt + tr.nt) * 64 * N);
                for (int i = et; i < 16 * N; i += kEpiThreads) reinterpret_cast<float4 *>(btab_s)[i] = __ldg(src + i);
                asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads));
                prev_b = b;
                prev_nt = tr.nt;
            }
            mbar_wait(bar_tfull + 8 * buf, (k / g.nbuf) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tbuf = tmem_base + (uint32_t)(buf * buf_cols);
            __nv_bfloat16 *out_pl = reinterpret_cast<__nv_bfloat16 *>(a.out) + (size_t)b * 2 * a.out_ctot * npix;
            float *out_cl = reinterpret_cast<float *>(a.out) + (size_t)b * npix * a.out_ctot + a.out_coff;
            const int co_base = tr.nt * N;
            for (int cb = 0; cb < N; cb += 16) {
                float ssum[16], ssq[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) ssum[q] = ssq[q] = 0.f;
                for (int gt = half; gt < g.G; gt += 2) {
                    const int R = gt * 128 + quad * 32 + lane;
                    const int tt = R / g.Wr;
                    const int jl = R - tt * g.Wr;
                    const int t = t0 + tt, j = j0 + jl;
                    const bool rowok = tt < g.TT && t < a.T && jl < g.TF;
                    int tmask = 0;
#pragma unroll
                    for (int kt = 0; kt < 3; ++kt) {
                        const int ti = a.transposed ? t + a.pad_t - kt : t + kt - a.pad_t;
                        if (kt < a.KT && ti >= 0 && ti < a.T) tmask |= 1 << kt;
                    }
                    for (int ph = 0; ph < g.nacc; ++ph) {
                        const int fo = j * g.ostride + ph;
                        const bool valid = rowok && fo < a.Fout;
                        int fmask = 0;
#pragma unroll
                        for (int kf = 0; kf < 3; ++kf) {
                            bool ok;
                            if (a.transposed) {
                                const int num = fo + a.pad_f - kf;
                                const int fi = num / a.stride_f;
                                ok = num >= 0 && fi * a.stride_f == num && fi < a.Fin;
                            } else {
                                const int fi = fo * a.stride_f + kf - a.pad_f;
                                ok = fi >= 0 && fi < a.Fin;
                            }
                            if (kf < a.KF && ok) fmask |= 1 << kf;
                        }
                        const float *bt = btab_s + (tmask * 8 + fmask) * N + cb;
                        uint32_t v[16], v2[16];
                        const uint32_t taddr = tbuf + ((uint32_t)(quad * 32) << 16) + (uint32_t)((gt * g.nacc + ph) * cw + cb);
                        tmem_ld16(taddr, v);
                        if constexpr (SPLIT == 3) tmem_ld16(taddr + (uint32_t)N, v2);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        float y[16];
#pragma unroll
                        for (int q = 0; q < 16; q += 4) {
                            const float4 bb = *reinterpret_cast<const float4 *>(bt + q);
                            y[q] = __uint_as_float(v[q]) + bb.x;
                            y[q + 1] = __uint_as_float(v[q + 1]) + bb.y;
                            y[q + 2] = __uint_as_float(v[q + 2]) + bb.z;
                            y[q + 3] = __uint_as_float(v[q + 3]) + bb.w;
                        }
                        if constexpr (SPLIT == 3) {
#pragma unroll
                            for (int q = 0; q < 16; ++q) y[q] += __uint_as_float(v2[q]);
                        }
                        if (a.elu) {
#pragma unroll
                            for (int q = 0; q < 16; ++q) y[q] = elu_fast(y[q]);
                        }
                        if (valid) {
#pragma unroll
                            for (int q = 0; q < 16; ++q) {
                                ssum[q] += y[q];
                                ssq[q] = fmaf(y[q], y[q], ssq[q]);
                            }
                            const int pix = t * a.Fout + fo;
                            if (planes) {
#pragma unroll
                                for (int g8 = 0; g8 < 16; g8 += 8) {
                                    const int co = co_base + cb + g8;
                                    if (co < a.cout) {
                                        const int ca = a.out_coff + co;
                                        __nv_bfloat16 *p = out_pl + ((size_t)(ca >> 3) * npix + pix) * 8;
                                        uint32_t hp[4];
#pragma unroll
                                        for (int q = 0; q < 4; ++q) hp[q] = pack_bf16x2(y[g8 + 2 * q], y[g8 + 2 * q + 1]);
                                        *reinterpret_cast<uint4 *>(p) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
                                        if (a.use_lo) {
                                            uint32_t lp[4];
#pragma unroll
                                            for (int q = 0; q < 4; ++q)
                                                lp[q] = pack_bf16x2(y[g8 + 2 * q] - bf16_lo(hp[q]), y[g8 + 2 * q + 1] - bf16_hi(hp[q]));
                                            *reinterpret_cast<uint4 *>(reinterpret_cast<char *>(p) + a.out_lo_off) =
                                                make_uint4(lp[0], lp[1], lp[2], lp[3]);
                                        }
                                    }
                                }
                            } else {
                                float *o = out_cl + (size_t)pix * a.out_ctot + co_base + cb;
#pragma unroll
                                for (int q = 0; q < 16; ++q)
                                    if (co_base + cb + q < a.cout) o[q] = y[q];
                            }
                        }
                    }
                }
                if (a.out_sums) {
                    const float s = warp_reduce16(ssum, lane);
                    const float q2 = warp_reduce16(ssq, lane);
                    if ((lane & 1) == 0) {
                        atomicAdd(&red[(cb + (lane >> 1)) * 2], s);
                        atomicAdd(&red[(cb + (lane >> 1)) * 2 + 1], q2);
                    }
                }
            }
            // this warp is done reading the accumulator buffer: hand it back to the MMA issuer
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
            if (a.out_sums) {
                asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads));
                if (et < N) {
                    if (co_base + et < a.cout) {
                        double *dst = a.out_sums + ((size_t)b * a.out_ctot + a.out_coff + co_base + et) * 2;
                        atomicAdd(dst, (double)red[et * 2]);
                        atomicAdd(dst + 1, (double)red[et * 2 + 1]);
                    }
                    red[et * 2] = 0.f;
                    red[et * 2 + 1] = 0.f;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads));
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

// ------------------------------------------------------------------------------------------------
// Per-sample operand preparation: blocks [0, B*nchunk) write the weight images of one (sample, chunk)
//   wimg[b][nt][chunk][tap][kg][w_hi rows | w_lo rows][8] = bf16 split of W[tap][ci][co] * rstd[b][ci]
// and blocks [B*nchunk, B*nchunk + B*nNt) write the border-bias tables of one (sample, N tile)
//   btab[b][nt][tmask*8+fmask][n] = bias[co] + sum_{kt in tmask, kf in fmask} sum_ci W[kt,kf][ci][co] * shift[b][ci]
// where x_norm = x * rstd + shift is the consumer-side InstanceNorm affine (model.py:413,430,445).
__global__ void __launch_bounds__(256) conv_tc_prep_kernel(const PrepArgs p) {
    extern __shared__ float sh[];
    const int nimg = p.B * p.nchunk;
    if ((int)blockIdx.x < nimg) {
        const int b = blockIdx.x / p.nchunk, chunk = blockIdx.x - b * p.nchunk;
        float *scale = sh;  // [16]
        if (threadIdx.x < 16) {
            const int ci = chunk * 16 + threadIdx.x;
            float sc = 0.f;
            if (ci < p.cin) {
                sc = 1.f;
                if (p.norm_mode == NORM_IN) {
                    const double *s = p.in_sums + ((size_t)b * p.in_ctot + p.in_coff + ci) * 2;
                    sc = affine_from_sums(s[0], s[1], p.norm_inv_n, (double)p.norm_eps).x;
                }
            }
            scale[threadIdx.x] = sc;
        }
        __syncthreads();
        const int per_nt = p.ntap * 2 * p.N;  // (tap, kg, n) triples per N tile
        const size_t w_stage = (size_t)p.nsp * p.ntap * 2 * p.N * 8;
        for (int i = threadIdx.x; i < p.nNt * per_nt; i += blockDim.x) {
            const int nt = i / per_nt;
            int r = i - nt * per_nt;
            const int n = r % p.N;
            r /= p.N;
            const int kg = r & 1, tap = r >> 1;
            const int co = nt * p.N + n;
            float h[8], l[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int ci = chunk * 16 + kg * 8 + e;
                float v = 0.f;
                if (ci < p.cin && co < p.cout) v = p.w[((size_t)p.tapk[tap] * p.cin + ci) * p.cout_pad + co] * scale[kg * 8 + e];
                h[e] = bf16_round(v);
                l[e] = v - h[e];
            }
            // image of one (sample, N tile, chunk): [tap][kg][nsp * N rows: w_hi rows then w_lo rows][8]
            __nv_bfloat16 *dst = p.wimg + (((size_t)b * p.nNt + nt) * p.nchunk + chunk) * w_stage +
                                 ((size_t)(tap * 2 + kg) * p.nsp * p.N + n) * 8;
            *reinterpret_cast<uint4 *>(dst) =
                make_uint4(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]), pack_bf16x2(h[4], h[5]), pack_bf16x2(h[6], h[7]));
            if (p.nsp == 2)
                *reinterpret_cast<uint4 *>(dst + (size_t)p.N * 8) =
                    make_uint4(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]), pack_bf16x2(l[4], l[5]), pack_bf16x2(l[6], l[7]));
        }
    } else {
        const int idx = blockIdx.x - nimg;
        const int b = idx / p.nNt, nt = idx - b * p.nNt;
        float *shift = sh;              // [cin]
        float *wb = sh + p.cin;         // [9][N]
        for (int ci = threadIdx.x; ci < p.cin; ci += blockDim.x) {
            float v = 0.f;
            if (p.norm_mode == NORM_IN) {
                const double *s = p.in_sums + ((size_t)b * p.in_ctot + p.in_coff + ci) * 2;
                v = affine_from_sums(s[0], s[1], p.norm_inv_n, (double)p.norm_eps).y;
            }
            shift[ci] = v;
        }
        __syncthreads();
        const int ntaps = p.KT * p.KF;
        for (int i = threadIdx.x; i < ntaps * p.N; i += blockDim.x) {
            const int k = i / p.N, n = i - k * p.N;
            const int co = nt * p.N + n;
            float acc = 0.f;
            if (co < p.cout) {
                const float *w = p.w + (size_t)k * p.cin * p.cout_pad + co;
                for (int ci = 0; ci < p.cin; ++ci) acc = fmaf(w[(size_t)ci * p.cout_pad], shift[ci], acc);
            }
            wb[i] = acc;
        }
        __syncthreads();
        float *dst = p.btab + ((size_t)b * p.nNt + nt) * 64 * p.N;
        for (int i = threadIdx.x; i < 64 * p.N; i += blockDim.x) {
            const int m = i / p.N, n = i - m * p.N;
            const int tm = m >> 3, fm = m & 7;
            const int co = nt * p.N + n;
            float v = (p.bias && co < p.cout) ? p.bias[co] : 0.f;
            for (int kt = 0; kt < p.KT; ++kt)
                for (int kf = 0; kf < p.KF; ++kf)
                    if (((tm >> kt) & 1) && ((fm >> kf) & 1)) v += wb[(kt * p.KF + kf) * p.N + n];
            dst[i] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int round_up(int x, int m) { return (x + m - 1) / m * m; }

// Tiling and shared-memory plan of one launch.  Returns false when the layer does not fit.
bool make_geom(const ConvArgs &a, int split, TcGeom &g) {
    g = TcGeom{};
    const bool s2 = a.stride_f == 2;
    g.nsp = split == 3 ? 2 : 1;
    g.nset = (!a.transposed && s2) ? 2 : 1;
    g.nacc = (a.transposed && s2) ? 2 : 1;
    g.ostride = g.nacc;
    g.strided = g.nset == 2;
    const int cout16 = round_up(a.cout, 16);
    g.N = std::min(cout16, 64);
    if (cout16 % g.N) {  // keep every N tile the same width
        g.N = 48;
        if (cout16 % 48) g.N = 32;
        if (cout16 % g.N) g.N = 16;
    }
    g.nNt = cout16 / g.N;
    g.ncols = (a.transposed && s2) ? a.Fin + 1 : a.Fout;
    g.nplanes = (a.cin + 7) / 8;
    g.nchunk = (g.nplanes + 1) / 2;
    // tap table
    const int halo_t = a.KT == 3 ? 1 : 0;
    const int extra_w = a.KF == 1 ? 0 : (s2 ? 1 : 2);
    g.ntap = 0;
    g.first_mask = 0;
    int seen_acc = 0;
    auto add = [&](int dt, int df, int set, int acc, int kt, int kf, int Wr) {
        const int i = g.ntap++;
        g.shift[i] = dt * Wr + df;
        g.aset[i] = set;
        g.acc[i] = acc;
        g.tapk[i] = kt * a.KF + kf;
        if (!((seen_acc >> acc) & 1)) {
            seen_acc |= 1 << acc;
            g.first_mask |= 1 << i;
        }
    };
    // ---- tile search: bins per tile (TF), frames per tile (TT <-> G M-tiles of 128 raster rows) and whether the
    // TMEM accumulators are double buffered, by a small cost model (cycles per tile, measured MMA issue costs:
    // 32 + Nw/4 cycles for an MMA of width Nw <= 128 -- the A tile is read from shared memory at 128 B/clk)
    const int cw = g.N * g.nsp;  // accumulator columns per (M tile, phase): bf16x3 keeps the w_hi / w_lo halves apart
    g.off_btab = kFixedSmem;
    g.off_red = g.off_btab + 64 * g.N * 4;
    g.off_stage = round_up(g.off_red + 2 * g.N * 4, 1024);
    g.w_stage = g.nsp * g.ntap_for(a) * 2 * g.N * 16;
    auto mma_cost = [](int nw) { return nw <= 128 ? 32.0 + nw / 4.0 : nw / 2.0; };
    const double cyc_mma = split == 3 ? mma_cost(2 * g.N) + mma_cost(g.N) : mma_cost(g.N);
    const int ntap_n = g.ntap_for(a);
    const int maxshift_k = a.KF == 1 ? 0 : (s2 ? 1 : 2);
    double best = 1e300;
    TcGeom bg{};
    bool found = false;
    for (int tf_cap : {16, 24, 32, 48, 64, 96, 128}) {
        if (tf_cap > 16 && tf_cap / 2 >= g.ncols) continue;
        for (int nbuf = 2; nbuf >= 1; --nbuf) {
            TcGeom c = g;
            int TF = std::min(g.ncols, tf_cap);
            c.f_tiles = (g.ncols + TF - 1) / TF;
            TF = (g.ncols + c.f_tiles - 1) / c.f_tiles;
            c.TF = TF;
            c.Wr = TF + extra_w;
            if (2 * c.Wr > 256) continue;
            const int Gmax = std::min(8, 512 / (cw * g.nacc * nbuf));
            if (Gmax < 1) continue;
            const int maxshift = a.KT == 1 ? 0 : 2 * c.Wr + maxshift_k;
            for (int G = Gmax; G >= 1; --G) {
                int TT = std::min({a.T, 128 * G / c.Wr, 254 - 2 * halo_t});
                if (TT < 1) break;
                c.t_tiles = (a.T + TT - 1) / TT;
                TT = (a.T + c.t_tiles - 1) / c.t_tiles;
                c.TT = TT;
                c.rows = TT + 2 * halo_t;
                c.G = (TT * c.Wr + 127) / 128;
                c.nbuf = nbuf;
                c.box_bytes = c.rows * c.Wr * 16;
                c.PL = round_up(std::max(c.rows * c.Wr, 128 * c.G + maxshift) * 16, 128);
                c.a_stage = c.nset * c.nsp * 2 * c.PL;
                c.stage = c.a_stage + round_up(c.w_stage, 128);
                c.nstage = std::min(kMaxStages, (kSmemLimit - c.off_stage) / c.stage);
                if (c.nstage < 2) continue;
                c.ntiles = a.B * c.nNt * c.t_tiles * c.f_tiles;
                const double main_cyc = (double)c.nchunk * ntap_n * c.G * cyc_mma;
                const double load_cyc = (double)c.nchunk * (c.nset * c.nsp * 2 * c.box_bytes + c.w_stage) / 24.0;  // ~L2 -> SM bytes/clk
                const double epi_cyc = 0.35 * c.G * 128.0 * c.N * c.nacc + 1500.0;
                const double body = std::max(main_cyc, load_cyc);
                const double tile_cyc = nbuf == 2 ? std::max(body, epi_cyc) + 500.0 : body + epi_cyc + 2500.0;
                const int ctas = std::min(c.ntiles, 148);
                const double total = std::ceil((double)c.ntiles / ctas) * tile_cyc + 6000.0;
                if (total < best) {
                    best = total;
                    bg = c;
                    found = true;
                }
            }
        }
    }
    if (!found) return false;
    {
        const int keep_ntap = g.ntap;
        g = bg;
        g.ntap = keep_ntap;
    }
    const int Wr = g.Wr;
    if (a.KT == 1 && a.KF == 1) {
        add(0, 0, 0, 0, 0, 0, Wr);
        g.t_org = 0;
        g.f_org = 0;
        g.f_mul = 1;
    } else if (!a.transposed && !s2) {  // conv, stride 1: fi = fo + kf - pad_f, ti = t + kt - 1
        for (int kt = 0; kt < 3; ++kt)
            for (int kf = 0; kf < 3; ++kf) add(kt, kf, 0, 0, kt, kf, Wr);
        g.t_org = -1;
        g.f_org = -a.pad_f;
        g.f_mul = 1;
    } else if (!a.transposed && s2) {  // conv, stride 2, pad_f 0: fi = 2 fo + kf
        for (int kt = 0; kt < 3; ++kt) {
            add(kt, 0, 0, 0, kt, 0, Wr);
            add(kt, 0, 1, 0, kt, 1, Wr);
            add(kt, 1, 0, 0, kt, 2, Wr);
        }
        g.t_org = -1;
        g.f_org = 0;
        g.f_mul = 2;
    } else if (a.transposed && !s2) {  // transposed, stride 1: fi = fo + pad_f - kf, ti = t + 1 - kt
        for (int kt = 0; kt < 3; ++kt)
            for (int kf = 0; kf < 3; ++kf) add(2 - kt, 2 - kf, 0, 0, kt, kf, Wr);
        g.t_org = -1;
        g.f_org = a.pad_f - 2;
        g.f_mul = 1;
    } else {  // transposed, stride 2, pad_f 0: fo = 2 j + ph; even: kf 0 -> fi = j, kf 2 -> fi = j - 1; odd: kf 1 -> fi = j
        for (int kt = 0; kt < 3; ++kt) {
            add(2 - kt, 1, 0, 0, kt, 0, Wr);
            add(2 - kt, 0, 0, 0, kt, 2, Wr);
            add(2 - kt, 1, 0, 1, kt, 1, Wr);
        }
        g.t_org = -1;
        g.f_org = -1;
        g.f_mul = 1;
    }
    g.smem_total = g.off_stage + g.nstage * g.stage;
    int cols = 32;
    while (cols < g.nbuf * g.G * g.nacc * cw) cols <<= 1;
    g.tmem_cols = cols;
    return cols <= 512 && g.rows <= 256;
}

int encode_maps(const ConvArgs &a, const TcGeom &g, CUtensorMap *hi, CUtensorMap *lo) {
    EncodeTiledFn enc = get_encode();
    if (!enc) {
        set_error("conv_tc: cuTensorMapEncodeTiled is not available from the driver");
        return MISO_E_CUDA;
    }
    const uint64_t CG = (uint64_t)a.in_ctot / 8, T = (uint64_t)a.T, F = (uint64_t)a.Fin;
    char *base = const_cast<char *>(reinterpret_cast<const char *>(a.in));
    for (int sp = 0; sp < 2; ++sp) {
        CUtensorMap *tm = sp == 0 ? hi : lo;
        void *addr = base + (sp ? a.in_lo_off : 0);
        CUresult r;
        if (!g.strided) {
            // [2F (8-byte units)][T][CG][B]: a tile row is one contiguous run of Wr * 16 bytes
            cuuint64_t dims[4] = {2 * F, T, CG, (cuuint64_t)a.B};
            cuuint64_t strides[3] = {F * 16, T * F * 16, 2 * CG * T * F * 16};
            cuuint32_t box[4] = {(cuuint32_t)(2 * g.Wr), (cuuint32_t)g.rows, 1, 1};
            cuuint32_t es[4] = {1, 1, 1, 1};
            r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, addr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        } else {
            // [8 ch][F][T][CG][B] with element stride 2 along F: the even / odd bins of a tile
            cuuint64_t dims[5] = {8, F, T, CG, (cuuint64_t)a.B};
            cuuint64_t strides[4] = {16, F * 16, T * F * 16, 2 * CG * T * F * 16};
            cuuint32_t box[5] = {8, (cuuint32_t)(2 * g.Wr), (cuuint32_t)g.rows, 1, 1};
            cuuint32_t es[5] = {1, 2, 1, 1, 1};
            r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, addr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        }
        if (r != CUDA_SUCCESS) {
            set_error("conv_tc: cuTensorMapEncodeTiled failed (%d) for Fin=%d T=%d ctot=%d Wr=%d rows=%d", (int)r, a.Fin, a.T, a.in_ctot,
                      g.Wr, g.rows);
            return MISO_E_CUDA;
        }
    }
    return MISO_OK;
}

}  // namespace

// one-time, capture-unsafe setup (function attributes, driver entry point); called before graph capture
int conv_tc_init() {
    static bool done = false;
    if (done) return MISO_OK;
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(conv_tc_kernel)");
    if (!get_encode()) {
        set_error("conv_tc: cuTensorMapEncodeTiled is not available from the driver");
        return MISO_E_CUDA;
    }
    done = true;
    return MISO_OK;
}

bool conv_tc_eligible(const ConvArgs &a) {
    if (!((a.KT == 3 && a.KF == 3) || (a.KT == 1 && a.KF == 1))) return false;
    if (a.KT == 3 && a.pad_t != 1) return false;
    if (a.stride_f != 1 && a.stride_f != 2) return false;
    if (a.stride_f == 2 && a.pad_f != 0) return false;
    if (a.in_layout != LAYOUT_PLANES || a.in_ctot % 8 || a.in_coff % 8) return false;
    if (a.out_layout == LAYOUT_PLANES && (a.out_ctot % 8 || a.out_coff % 8 || a.cout % 8)) return false;
    if (a.resid != nullptr || a.norm_mode == NORM_GLN) return false;
    TcGeom g;
    return make_geom(a, 3, g);
}

void conv_tc_scratch_need(const ConvArgs &a, int split, size_t *wimg_bytes, size_t *btab_bytes) {
    TcGeom g;
    *wimg_bytes = *btab_bytes = 0;
    if (!make_geom(a, split, g)) return;
    *wimg_bytes = (size_t)a.B * g.nNt * g.nchunk * g.w_stage;
    *btab_bytes = (size_t)a.B * g.nNt * 64 * g.N * sizeof(float);
}

int launch_conv_tc(const ConvArgs &a, int split, const TcScratch &scratch, cudaStream_t stream) {
    MISO_REQUIRE(split == 1 || split == 3, "conv_tc: bad split");
    TcGeom g;
    MISO_REQUIRE(make_geom(a, split, g), "conv_tc: layer does not fit the tcgen05 path (cin=%d cout=%d Fin=%d)", a.cin, a.cout, a.Fin);
    const size_t need_w = (size_t)a.B * g.nNt * g.nchunk * g.w_stage, need_b = (size_t)a.B * g.nNt * 64 * g.N * sizeof(float);
    if (need_w > scratch.wimg_bytes || need_b > scratch.btab_bytes) {
        set_error("conv_tc: scratch too small (%zu/%zu weight bytes, %zu/%zu bias bytes)", scratch.wimg_bytes, need_w,
                  scratch.btab_bytes, need_b);
        return MISO_E_WORKSPACE;
    }
    static const bool debug = getenv("MISO_TC_DEBUG") != nullptr;
    if (debug)
        fprintf(stderr,
                "conv_tc: cin=%d cout=%d Fin=%d Fout=%d s=%d tr=%d | N=%d nNt=%d TF=%d Wr=%d TT=%d G=%d nbuf=%d nstage=%d stage=%dB "
                "tiles=%d (t%d x f%d) tmem=%d smem=%d\n",
                a.cin, a.cout, a.Fin, a.Fout, a.stride_f, a.transposed, g.N, g.nNt, g.TF, g.Wr, g.TT, g.G, g.nbuf, g.nstage, g.stage,
                g.ntiles, g.t_tiles, g.f_tiles, g.tmem_cols, g.smem_total);
    CUtensorMap tm_hi, tm_lo;
    int rc = encode_maps(a, g, &tm_hi, &tm_lo);
    if (rc) return rc;

    PrepArgs p{};
    p.w = a.w;
    p.bias = a.bias;
    p.in_sums = a.in_sums;
    p.in_ctot = a.in_ctot;
    p.in_coff = a.in_coff;
    p.cin = a.cin;
    p.cout = a.cout;
    p.cout_pad = a.cout_pad;
    p.norm_mode = a.norm_mode;
    p.norm_eps = a.norm_eps;
    p.norm_inv_n = a.norm_inv_n;
    p.wimg = reinterpret_cast<__nv_bfloat16 *>(scratch.wimg);
    p.btab = scratch.btab;
    p.B = a.B;
    p.N = g.N;
    p.nNt = g.nNt;
    p.nchunk = g.nchunk;
    p.nsp = g.nsp;
    p.ntap = g.ntap;
    p.KT = a.KT;
    p.KF = a.KF;
    for (int i = 0; i < g.ntap; ++i) p.tapk[i] = g.tapk[i];
    const size_t prep_smem = std::max<size_t>(64, (size_t)(a.cin + 9 * g.N) * sizeof(float));
    prof_begin(stream);
    conv_tc_prep_kernel<<<a.B * g.nchunk + a.B * g.nNt, 256, prep_smem, stream>>>(p);
    MISO_LAUNCHED("conv_tc_prep_kernel");

    TcArgs k{};
    k.g = g;
    k.wimg = p.wimg;
    k.btab = p.btab;
    k.out = a.out;
    k.out_sums = a.out_sums;
    k.B = a.B;
    k.T = a.T;
    k.Fin = a.Fin;
    k.Fout = a.Fout;
    k.in_coff = a.in_coff;
    k.out_ctot = a.out_ctot;
    k.out_coff = a.out_coff;
    k.cout = a.cout;
    k.out_layout = a.out_layout;
    k.out_lo_off = a.out_lo_off;
    k.use_lo = a.use_lo;
    k.KT = a.KT;
    k.KF = a.KF;
    k.stride_f = a.stride_f;
    k.pad_t = a.pad_t;
    k.pad_f = a.pad_f;
    k.transposed = a.transposed;
    k.elu = a.elu;

    rc = conv_tc_init();
    if (rc) return rc;
    dim3 grid(std::min(g.ntiles, 148), 1, 1);  // persistent: one CTA per SM walks the tile list
    if (split == 3)
        conv_tc_kernel<3><<<grid, kThreads, g.smem_total, stream>>>(tm_hi, tm_lo, k);
    else
        conv_tc_kernel<1><<<grid, kThreads, g.smem_total, stream>>>(tm_hi, tm_lo, k);
    {
        const double pix = (double)a.B * a.T * (a.transposed ? a.Fin : a.Fout);
        const double flops = 2.0 * pix * a.cin * a.cout * a.KT * a.KF;
        const double bytes = (a.use_lo ? 4.0 : 2.0) * a.B * a.T * ((double)a.Fin * a.cin + (double)a.Fout * a.cout);
        prof_end(stream, flops, bytes, MISO_PROF_CONV_TC);
    }
    MISO_LAUNCHED("conv_tc_kernel");
    return MISO_OK;
}

}  // namespace miso
