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
#include "tcn.cuh"
#include "umma.cuh"

namespace miso {
namespace {

constexpr int kMmaWarps = 4;        // MMA issuers: warp w owns the M tiles gt = w, w + 4, ... (disjoint accumulators)
constexpr int kEpiWarp0 = 1 + kMmaWarps;
constexpr int kThreads = (1 + kMmaWarps + 8) * 32;  // 13 warps: TMA, 4 x MMA, 8 epilogue
constexpr int kEpiThreads = 256;
constexpr int kMaxTaps = 9;
constexpr int kMaxStages = 4;
constexpr int kSmemLimit = 227 * 1024;
constexpr int kFixedSmem = 1024;    // barriers, tmem slot
constexpr int kBiasCi = 64;         // input channels per border-bias partial block (conv_tc_prep_kernel)

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
    int nchunk, nplanes, kper, nunit;  // a stage holds kper 16-channel K units (2 planes each); nunit = total units
    // shared memory
    int PL, GS, a_stage, w_unit, w_stage, stage, nstage, box_bytes;  // PL: plane bytes; GS: bytes of one (set, split) box  // w_unit: weight image bytes of one K unit
    int off_btab, off_red, off_list, off_stage, smem_total;
    int tmem_cols;
    int strided;               // 5-D bf16 tensor map with element stride 2 along bins
    int ppb;                   // planes per TMA box: 2*kper (dense planes, one box per stage) or 1 (padded planes)
    unsigned wr_magic;         // ceil(2^32 / Wr): R / Wr == __umulhi(R, wr_magic) for the raster rows of a tile
    int ntap_for(const ConvArgs &a) const { return a.KT * a.KF; }
};

struct TcArgs {
    TcGeom g;
    const __nv_bfloat16 *wimg;  // [B][nNt][nchunk][ntap][2][nsp*N][8]
    const float *btab;          // border-bias partial sums [B][nNt][nsplit][9][N] (conv_tc_prep_kernel)
    int bias_shared;            // no per-sample normalisation: one table (all zero) for every sample
    const float *bias;          // [cout_pad] or null
    int nsplit;
    void *out;
    double *out_sums;
    int B, T, Fin, Fout;
    int in_coff;                // channel offset of the input view (multiple of 8)
    int out_ctot, out_coff, cout;
    int out_layout;
    size_t out_lo_off;
    int use_lo;
    int KT, KF, stride_f, pad_t, pad_f, transposed, elu;
    size_t wimg_bstride;        // elements between the weight images of consecutive samples (0: one shared image)
    const double *gln_sums;     // non-null: multiply the accumulator by the sample's gLN rstd (weights carry gamma only)
    double gln_inv_n;
    float gln_eps;
    const float *resid;         // optional fp32 channels-last residual (TCN, model.py:549), added before the store
    int resid_ctot, resid_coff;
    long long *trace;  // debug: clock64 event log of CTA 0 (tools/tc_trace.py), or null
};

// trace regions: [0,4096) producer, [4096,8192) MMA issuer, [8192,12288) epilogue warp 2; entries are (tag, clock)
__device__ __forceinline__ void trace_ev(long long *trace, int region, int &n, int tag) {
    if (trace && blockIdx.x == 0 && n < 2040) {
        trace[region * 4096 + 2 * n] = tag;
        trace[region * 4096 + 2 * n + 1] = clock64();
        ++n;
    }
}

struct PrepArgs {
    const float *w;        // packed fp32 [ntaps][cin][cout_pad]
    const float *bias;     // [cout_pad] or null
    const double *in_sums;
    const float *gamma, *beta;  // NORM_GLN
    int in_ctot, in_coff, cin, cout, cout_pad;
    int norm_mode;
    float norm_eps;
    double norm_inv_n;
    __nv_bfloat16 *wimg;
    float *btab;
    int B, N, nNt, nunit, nsp, ntap, KT, KF;
    int shared_w;  // gLN: one weight image for all samples (scaled by gamma only; rstd is applied in the epilogue); also
                   // without any normalisation (data gradients, the network's first conv), where the bias tables are shared too
    int nb_bias;   // samples with a border-bias table of their own: B, or 1
    int nsplit;    // channel splits of the border-bias partial sums
    int tapk[kMaxTaps];
};

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

template <int SPLIT, int VAR>
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
            mbar_init(bar_empty + 8 * s, kMmaWarps);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_tfull + 8 * i, kMmaWarps);
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
        // statistics scratch
        for (int i = tid; i < 8 * 2 * N; i += kThreads) red[i] = 0.f;  // one slot per epilogue warp: fixed summation order
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();
    pdl_wait();  // everything above overlapped the previous kernel's tail; its results are needed from here on
    // contiguous tile range per CTA: consecutive tiles share the sample (weight image, bias table) and their halos in L2
    const int tiles_per_cta = (g.ntiles + (int)gridDim.x - 1) / (int)gridDim.x;
    const int tile_lo = (int)blockIdx.x * tiles_per_cta, tile_hi = min(tile_lo + tiles_per_cta, g.ntiles);

    if (warp == 0) {
        // ---------------------------------------------------------------- TMA producer
        if (elect_one()) {
            const int plane0 = a.in_coff >> 3;
            int q = 0;  // running chunk index over all tiles of this CTA
            int ntr = 0;
            for (int tile = tile_lo; tile < tile_hi; ++tile) {
                const TileRef tr = decode_tile(g, tile);
                const int tin = tr.t0 + g.t_org;
                const int fin = tr.j0 * g.f_mul + g.f_org;
                const __nv_bfloat16 *wsrc = a.wimg + (size_t)tr.b * a.wimg_bstride + (size_t)tr.nt * g.nunit * (size_t)(g.w_unit / 2);
                for (int c = 0; c < g.nchunk; ++c, ++q) {
                    const int s = q % g.nstage;
                    trace_ev(a.trace, 0, ntr, 100 + c);
                    if (q >= g.nstage) mbar_wait(bar_empty + 8 * s, ((q / g.nstage) + 1) & 1);
                    trace_ev(a.trace, 0, ntr, 200 + c);
                    const int pbase = 2 * g.kper * c;
                    const int np = min(2 * g.kper, g.nplanes - pbase);
                    const int nu = (np + 1) >> 1;
                    const uint32_t full = bar_full + 8 * s;
                    const int nbox = g.ppb == 1 ? 2 * nu : 1;  // an odd last plane pairs with an out-of-bounds (zero) plane
                    mbar_expect_tx(full, (uint32_t)(g.nset * g.nsp * nbox * g.ppb * g.box_bytes + nu * g.w_unit));
                    const uint32_t sa = s_stage + (uint32_t)(s * g.stage);
                    // one box = all 2*kper planes of the stage (planes past the view's last one are either the next
                    // channels of the buffer -- finite, multiplied by zero weights -- or out of bounds -> zero fill)
                    for (int set = 0; set < g.nset; ++set)
                        for (int sp = 0; sp < g.nsp; ++sp) {
                            const CUtensorMap *tm = sp == 0 ? &tm_hi : &tm_lo;
                            for (int p = 0; p < nbox; ++p) {
                                const uint32_t dst = sa + (uint32_t)((set * g.nsp + sp) * g.GS + p * g.PL);
                                if (g.strided)
                                    tma_load_5d(dst, tm, full, 0, fin + set, tin, plane0 + pbase + p, tr.b);
                                else
                                    tma_load_4d(dst, tm, full, 2 * fin, tin, plane0 + pbase + p, tr.b);
                            }
                        }
                    bulk_load(sa + (uint32_t)g.a_stage, wsrc + (size_t)c * g.kper * (g.w_unit / 2), (uint32_t)(nu * g.w_unit), full);
                }
            }
        }
    } else if (warp <= kMmaWarps) {
        // ---------------------------------------------------------------- MMA issuers
        // One thread issues an MMA in tens of cycles but pays hundreds per tap for descriptor set-up and the tensor
        // pipe does not run ahead of its issuer, so the M tiles are split over kMmaWarps issuing warps (disjoint
        // accumulators: no ordering between them is needed); every warp waits for / commits to the same barriers.
        const int mw = warp - 1;
        // SPLIT == 3: the weight image of a tap is [kg][2N rows: w_hi then w_lo][8], so
        //   D[:, 0:N] , D[:, N:2N] += A_hi * [W_hi | W_lo]   (one MMA of width 2N)
        //   D[:, 0:N]              += A_lo * W_hi             (one MMA of width N)
        // and the epilogue adds the two column halves: 2 MMAs instead of 3, A_hi read once.
        const uint32_t idesc_wide = make_idesc(cw), idesc_n = make_idesc(N);
        const uint32_t gstep = (uint32_t)(cw * g.nacc);
        // Descriptor arithmetic is kept to one 32-bit add per operand: the high words (stride offset 128 B,
        // version) are constants and the low words are (address >> 4) | (leading offset >> 4) << 16, so moving
        // a descriptor by a multiple of 16 bytes is an add to its low word (addresses stay below 2^18).
        constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);
        const uint32_t a_lo0 = (s_stage >> 4) | (((uint32_t)g.PL >> 4) << 16);
        const uint32_t b_lo0 = ((s_stage + (uint32_t)g.a_stage) >> 4) | ((((uint32_t)cw * 16) >> 4) << 16);
        uint32_t tapA[kMaxTaps], tapB[kMaxTaps], tapD[kMaxTaps];
#pragma unroll
        for (int i = 0; i < kMaxTaps; ++i) {
            tapA[i] = (uint32_t)(g.aset[i] * g.nsp * g.GS + g.shift[i] * 16) >> 4;
            tapB[i] = (uint32_t)(i * 2 * cw * 16) >> 4;
            tapD[i] = (uint32_t)(g.acc[i] * cw);
        }
        const uint32_t lo_split = (uint32_t)g.GS >> 4;  // A_lo planes follow the A_hi planes of the same set
        const uint32_t a_kstep = (uint32_t)(2 * g.PL) >> 4, b_kstep = (uint32_t)g.w_unit >> 4;
        int q = 0, k = 0;
        int ntr = 0;
        for (int tile = tile_lo; tile < tile_hi; ++tile, ++k) {
            const int buf = g.nbuf == 2 ? (k & 1) : 0;
            // the epilogue must have drained this accumulator buffer (use u = k / nbuf of it)
            {
                const int u = k / g.nbuf;
                if (warp == 1 && lane == 0) trace_ev(a.trace, 1, ntr, 1000);
                if (u > 0) mbar_wait(bar_tempty + 8 * buf, (u + 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (warp == 1 && lane == 0) trace_ev(a.trace, 1, ntr, 1001);
            }
            const uint32_t tbuf = tmem_base + (uint32_t)(buf * buf_cols);
            for (int c = 0; c < g.nchunk; ++c, ++q) {
                const int s = q % g.nstage;
                if (warp == 1 && lane == 0) trace_ev(a.trace, 1, ntr, 100 + c);
                mbar_wait(bar_full + 8 * s, (q / g.nstage) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (warp == 1 && lane == 0) trace_ev(a.trace, 1, ntr, 200 + c);
                if (elect_one()) {
                    const uint32_t st4 = (uint32_t)(s * g.stage) >> 4;
                    const int nu = min(g.kper, g.nunit - g.kper * c);
                    if constexpr (VAR == 0) {
#pragma unroll 1
                        for (int ks = 0; ks < nu; ++ks) {
                            const uint32_t a_ks = a_lo0 + st4 + (uint32_t)ks * a_kstep;
                            const uint32_t b_ks = b_lo0 + st4 + (uint32_t)ks * b_kstep;
                            const bool first_unit = c == 0 && ks == 0;
#pragma unroll
                            for (int i = 0; i < kMaxTaps; ++i) {
                                if (i < g.ntap) {
                                    const uint64_t bdesc = ((uint64_t)kDescHi << 32) | (uint64_t)(b_ks + tapB[i]);
                                    const uint32_t acc_flag = (first_unit && ((g.first_mask >> i) & 1)) ? 0u : 1u;
                                    const uint32_t d0 = tbuf + tapD[i];
                                    const uint32_t alo = a_ks + tapA[i];
#pragma unroll 2
                                    for (int gt = mw; gt < g.G; gt += kMmaWarps)
                                        umma_bf16(d0 + gt * gstep, ((uint64_t)kDescHi << 32) | (uint64_t)(alo + gt * 128), bdesc, idesc_wide, acc_flag);
                                    if constexpr (SPLIT == 3) {
#pragma unroll 2
                                        for (int gt = mw; gt < g.G; gt += kMmaWarps)
                                            umma_bf16(d0 + gt * gstep, ((uint64_t)kDescHi << 32) | (uint64_t)(alo + lo_split + gt * 128), bdesc, idesc_n, 1u);
                                    }
                                }
                            }
                        }
                    } else {
                        const uint32_t abase = s_stage + (uint32_t)(s * g.stage);
                        const uint32_t bbase = abase + (uint32_t)g.a_stage;
                        const uint32_t a_lbo = (uint32_t)g.PL, b_lbo = (uint32_t)cw * 16;
#pragma unroll 1
                        for (int ks = 0; ks < nu; ++ks) {
#pragma unroll 1
                            for (int i = 0; i < g.ntap; ++i) {
                                const uint32_t a_hi = abase + (uint32_t)(g.aset[i] * g.nsp * g.GS + ks * 2 * g.PL) + (uint32_t)(g.shift[i] * 16);
                                const uint64_t bdesc = make_desc(bbase + (uint32_t)(ks * g.w_unit) + (uint32_t)(i * 2 * cw * 16), b_lbo, 128);
                                const uint32_t acc_flag = (c == 0 && ks == 0 && ((g.first_mask >> i) & 1)) ? 0u : 1u;
                                const uint32_t d0 = tbuf + (uint32_t)g.acc[i] * cw;
                                uint64_t ad = make_desc(a_hi, a_lbo, 128);
#pragma unroll 2
                                for (int gt = mw; gt < g.G; gt += kMmaWarps) umma_bf16(d0 + gt * gstep, ad + (uint64_t)(gt * 128), bdesc, idesc_wide, acc_flag);
                                if constexpr (SPLIT == 3) {
                                    ad = make_desc(a_hi + (uint32_t)g.GS, a_lbo, 128);
#pragma unroll 2
                                    for (int gt = mw; gt < g.G; gt += kMmaWarps) umma_bf16(d0 + gt * gstep, ad + (uint64_t)(gt * 128), bdesc, idesc_n, 1u);
                                }
                            }
                        }
                    }
                    umma_commit(bar_empty + 8 * s);
                    if (c == g.nchunk - 1) umma_commit(bar_tfull + 8 * buf);
                }
                __syncwarp();
                if (warp == 1 && lane == 0) trace_ev(a.trace, 1, ntr, 300 + c);
            }
        }
    } else {
        // ---------------------------------------------------------------- epilogue (warps 2..9)
        const int quad = warp & 3, half = (warp - kEpiWarp0) >> 2;
        const int et = tid - kEpiWarp0 * 32;
        const int npix = a.T * a.Fout;
        const bool planes = a.out_layout == LAYOUT_PLANES;
        int prev_b = -1, prev_nt = -1, k = 0;
        int ntr = 0;
        const bool tracer = warp == kEpiWarp0 && lane == 0;
        for (int tile = tile_lo; tile < tile_hi; ++tile, ++k) {
            const TileRef tr = decode_tile(g, tile);
            const int b = tr.b, t0 = tr.t0, j0 = tr.j0;
            if (tracer) trace_ev(a.trace, 2, ntr, 1);
            const int buf = g.nbuf == 2 ? (k & 1) : 0;
            if (b != prev_b || tr.nt != prev_nt) {
                // border-bias table of this (sample, N tile)
                // border-bias table of this (sample, N tile): add the channel splits of the prep kernel's partial sums,
                // then expand to the 64 (frame-mask, bin-mask) classes
                asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads));
                float *wb = red + 16 * N;  // [9][N] scratch behind the statistics
                const int ntaps = a.KT * a.KF;
                const float *src = a.btab + ((size_t)(a.bias_shared ? 0 : b) * g.nNt + tr.nt) * a.nsplit * 9 * N;
                for (int i = et; i < ntaps * N; i += kEpiThreads) {
                    float v = 0.f;
                    for (int sp = 0; sp < a.nsplit; ++sp) v += __ldg(src + (size_t)sp * 9 * N + i);
                    wb[i] = v;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads));
                for (int i = et; i < 64 * N; i += kEpiThreads) {
                    const int m = i / N, n = i - m * N;
                    const int tm = m >> 3, fm = m & 7;
                    const int co = tr.nt * N + n;
                    float v = (a.bias && co < a.cout) ? __ldg(a.bias + co) : 0.f;
                    for (int kt = 0; kt < a.KT; ++kt)
                        for (int kf = 0; kf < a.KF; ++kf)
                            if (((tm >> kt) & 1) && ((fm >> kf) & 1)) v += wb[(kt * a.KF + kf) * N + n];
                    btab_s[i] = v;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads));
                prev_b = b;
                prev_nt = tr.nt;
            }
            float oscale = 1.f;
            if (a.gln_sums) {  // gLN rstd of this sample (model.py:628-631); gamma is in the weights, beta / mean in the bias table
                const double mean = stat_get(a.gln_sums + (size_t)b * 2) * a.gln_inv_n;
                double var = stat_get(a.gln_sums + (size_t)b * 2 + 1) * a.gln_inv_n - mean * mean;
                if (var < 0.0) var = 0.0;
                oscale = (float)rsqrt(var + (double)a.gln_eps);
            }
            mbar_wait(bar_tfull + 8 * buf, (k / g.nbuf) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (tracer) trace_ev(a.trace, 2, ntr, 2);
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
                    const int tt = g.Wr == 1 ? R : (int)__umulhi((unsigned)R, g.wr_magic);
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
                                const int fi = num >> (a.stride_f - 1);  // stride_f is 1 or 2
                                ok = num >= 0 && (fi << (a.stride_f - 1)) == num && fi < a.Fin;
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
                        if (tracer) trace_ev(a.trace, 2, ntr, 10);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        if (tracer) trace_ev(a.trace, 2, ntr, 11);
                        float y[16];
#pragma unroll
                        for (int q = 0; q < 16; ++q) y[q] = __uint_as_float(v[q]);
                        if constexpr (SPLIT == 3) {
#pragma unroll
                            for (int q = 0; q < 16; ++q) y[q] += __uint_as_float(v2[q]);
                        }
#pragma unroll
                        for (int q = 0; q < 16; q += 4) {
                            const float4 bb = *reinterpret_cast<const float4 *>(bt + q);
                            y[q] = fmaf(y[q], oscale, bb.x);
                            y[q + 1] = fmaf(y[q + 1], oscale, bb.y);
                            y[q + 2] = fmaf(y[q + 2], oscale, bb.z);
                            y[q + 3] = fmaf(y[q + 3], oscale, bb.w);
                        }
                        if (a.resid && valid) {
                            const float *rp = a.resid + ((size_t)b * npix + (size_t)t * a.Fout + fo) * a.resid_ctot + a.resid_coff + co_base + cb;
#pragma unroll
                            for (int q = 0; q < 16; q += 4) {
                                if (co_base + cb + q < a.cout) {
                                    const float4 rr = *reinterpret_cast<const float4 *>(rp + q);
                                    y[q] += rr.x;
                                    y[q + 1] += rr.y;
                                    y[q + 2] += rr.z;
                                    y[q + 3] += rr.w;
                                }
                            }
                        }
                        if (a.elu) {
#pragma unroll
                            for (int q = 0; q < 16; ++q) y[q] = elu_fast(y[q]);
                        }
                        if (tracer) trace_ev(a.trace, 2, ntr, 12);
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
                                if (((a.out_ctot | a.out_coff | a.cout) & 3) == 0) {
#pragma unroll
                                    for (int q = 0; q < 16; q += 4)
                                        if (co_base + cb + q + 4 <= a.cout) *reinterpret_cast<float4 *>(o + q) = make_float4(y[q], y[q + 1], y[q + 2], y[q + 3]);
                                } else {  // 2-channel outputs (MISO_3's last deconv, model.py:347): 8-byte stores
#pragma unroll
                                    for (int q = 0; q < 16; q += 2)
                                        if (co_base + cb + q + 2 <= a.cout) *reinterpret_cast<float2 *>(o + q) = make_float2(y[q], y[q + 1]);
                                }
                            }
                        }
                    }
                }
                if (tracer) trace_ev(a.trace, 2, ntr, 13);
                if (a.out_sums) {
                    const float s = warp_reduce16(ssum, lane);
                    const float q2 = warp_reduce16(ssq, lane);
                    if ((lane & 1) == 0) {
                        float *slot = red + (warp - kEpiWarp0) * 2 * N;  // this warp's own slot: no atomics, no ordering
                        slot[(cb + (lane >> 1)) * 2] = s;
                        slot[(cb + (lane >> 1)) * 2 + 1] = q2;
                    }
                }
            }
            // this warp is done reading the accumulator buffer: hand it back to the MMA issuer
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (tracer) trace_ev(a.trace, 2, ntr, 3);
            if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
            if (a.out_sums) {
                asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads));
                if (et < N) {
                    if (co_base + et < a.cout) {
                        double *dst = a.out_sums + ((size_t)b * a.out_ctot + a.out_coff + co_base + et) * 2;
                        double s8 = 0.0, q8 = 0.0;
#pragma unroll
                        for (int w8 = 0; w8 < 8; ++w8) {
                            s8 += (double)red[w8 * 2 * N + et * 2];
                            q8 += (double)red[w8 * 2 * N + et * 2 + 1];
                        }
                        stat_add(dst, s8);
                        stat_add(dst + 1, q8);
                    }
                }
                asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads));
            }
            if (tracer) trace_ev(a.trace, 2, ntr, 4);
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
//   wbp[b][nt][split][k][n] = sum_{ci in split} W[k][ci][co] * shift[b][ci]   (border-bias partial sums; the conv kernel
//   expands them to btab[tmask*8+fmask][n] = bias[co] + sum_{kt in tmask, kf in fmask} sum_splits wbp)
// where x_norm = x * rstd + shift is the consumer-side InstanceNorm affine (model.py:413,430,445).
// consumer-side normalisation of input channel ci of sample b as an affine (scale, shift)
__device__ __forceinline__ float2 prep_affine(const PrepArgs &p, int b, int ci) {
    if (p.norm_mode == NORM_IN) {
        const double *s = p.in_sums + ((size_t)b * p.in_ctot + p.in_coff + ci) * 2;
        return affine_from_sums(stat_get(s), stat_get(s + 1), p.norm_inv_n, (double)p.norm_eps);
    }
    if (p.norm_mode == NORM_GLN) {
        // gLN (model.py:609-632): y = gamma (x - mean) rstd + beta with one (mean, rstd) per sample.  rstd is a per-sample
        // scalar, so the weights carry gamma only (sample independent), the epilogue multiplies the accumulator by rstd,
        // and the bias term is sum_c W (beta - gamma mean rstd)
        const double *s = p.in_sums + (size_t)b * 2;
        const double mean = stat_get(s) * p.norm_inv_n;
        double var = stat_get(s + 1) * p.norm_inv_n - mean * mean;
        if (var < 0.0) var = 0.0;
        const double r = rsqrt(var + (double)p.norm_eps);
        const double gm = (double)p.gamma[ci];
        return make_float2((float)gm, (float)((double)p.beta[ci] - gm * mean * r));
    }
    return make_float2(1.f, 0.f);
}

__global__ void __launch_bounds__(256) conv_tc_prep_kernel(const PrepArgs p) {
    extern __shared__ float sh[];
    pdl_trigger();
    pdl_wait();  // the statistics come from the previous conv, which also still reads the scratch this kernel rewrites
    const int nimg = (p.shared_w ? 1 : p.B) * p.nunit;
    if ((int)blockIdx.x < nimg) {
        const int b = blockIdx.x / p.nunit, chunk = blockIdx.x - b * p.nunit;  // chunk = 16-channel K unit
        float *scale = sh;  // [16]
        if (threadIdx.x < 16) {
            const int ci = chunk * 16 + threadIdx.x;
            scale[threadIdx.x] = ci < p.cin ? prep_affine(p, b, ci).x : 0.f;
        }
        __syncthreads();
        const int per_nt = p.ntap * 2 * p.N;  // (tap, kg, n) triples per N tile
        const size_t w_stage = (size_t)p.nsp * p.ntap * 2 * p.N * 8;
#pragma unroll 3
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
            __nv_bfloat16 *dst = p.wimg + (((size_t)b * p.nNt + nt) * p.nunit + chunk) * w_stage +
                                 ((size_t)(tap * 2 + kg) * p.nsp * p.N + n) * 8;
            *reinterpret_cast<uint4 *>(dst) =
                make_uint4(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]), pack_bf16x2(h[4], h[5]), pack_bf16x2(h[6], h[7]));
            if (p.nsp == 2)
                *reinterpret_cast<uint4 *>(dst + (size_t)p.N * 8) =
                    make_uint4(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]), pack_bf16x2(l[4], l[5]), pack_bf16x2(l[6], l[7]));
        }
    } else {
        // border-bias partial sums: block (b, nt, split) covers input channels [split*kBiasCi, +kBiasCi) and writes
        //   wbp[b][nt][split][k][n] = sum_ci W[k][ci][co] * shift[b][ci]
        // (the conv kernel adds the splits and expands them into the 64 mask classes; many small blocks keep the
        // weight reads parallel -- a single block per sample is latency bound for the 768-channel bottleneck layers)
        int idx = blockIdx.x - nimg;
        const int split = idx % p.nsplit;
        idx /= p.nsplit;
        const int b = idx / p.nNt, nt = idx - b * p.nNt;  // (b < nb_bias by the grid size)
        float *shift = sh;                 // [kBiasCi]
        float *wb = sh + kBiasCi;          // [nparts <= 8][ntaps][N]: one slot per channel part, summed in a fixed order
        const int ci0 = split * kBiasCi, nci = min(kBiasCi, p.cin - ci0);
        const int ntaps = p.KT * p.KF;
        for (int i = threadIdx.x; i < nci; i += blockDim.x) shift[i] = prep_affine(p, b, ci0 + i).y;
        const int n = threadIdx.x % p.N, part = threadIdx.x / p.N, nparts = min(8, (int)blockDim.x / p.N);
        for (int i = threadIdx.x; i < nparts * ntaps * p.N; i += blockDim.x) wb[i] = 0.f;
        __syncthreads();
        const int co = nt * p.N + n;
        if (part < nparts && co < p.cout) {
            // all taps' loads in flight at once (the weights are cold in L2 behind the previous conv's activation stream)
            float acc[kMaxTaps];
#pragma unroll
            for (int k = 0; k < kMaxTaps; ++k) acc[k] = 0.f;
#pragma unroll 2
            for (int ci = part; ci < nci; ci += nparts) {
                const float sv = shift[ci];
#pragma unroll
                for (int k = 0; k < kMaxTaps; ++k)
                    if (k < ntaps) acc[k] = fmaf(__ldg(p.w + ((size_t)k * p.cin + ci0 + ci) * p.cout_pad + co), sv, acc[k]);
            }
#pragma unroll
            for (int k = 0; k < kMaxTaps; ++k)
                if (k < ntaps) wb[(part * ntaps + k) * p.N + n] = acc[k];
        }
        __syncthreads();
        float *dst = p.btab + (((size_t)b * p.nNt + nt) * p.nsplit + split) * 9 * p.N;
        for (int i = threadIdx.x; i < ntaps * p.N; i += blockDim.x) {
            float v = 0.f;
            for (int q = 0; q < nparts; ++q) v += wb[q * ntaps * p.N + i];
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

long long *g_trace_buf = nullptr;
int g_trace_cin = 0, g_trace_fin = 0;

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
    g.nunit = (g.nplanes + 1) / 2;
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
    g.off_list = round_up(g.off_red + (16 + 9) * g.N * 4, 16);
    g.off_stage = round_up(g.off_list + kMaxTaps * 8 * 2 * 16, 1024);
    g.w_unit = g.nsp * g.ntap_for(a) * 2 * g.N * 16;
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
            for (int G = Gmax; G >= 1; --G)
              for (int kper : {1, 2, 4}) {
                if (kper > 1 && kper > g.nunit) continue;
                c.kper = kper;
                c.nchunk = (g.nunit + kper - 1) / kper;
                c.w_stage = kper * g.w_unit;
                int TT = std::min({a.T, 128 * G / c.Wr, 254 - 2 * halo_t});
                if (TT < 1) break;
                c.t_tiles = (a.T + TT - 1) / TT;
                if (c.Wr > 2) TT = (a.T + c.t_tiles - 1) / c.t_tiles;  // balance the tiles (narrow rasters keep full M tiles)
                c.TT = TT;
                c.rows = TT + 2 * halo_t;
                c.G = (TT * c.Wr + 127) / 128;
                c.nbuf = nbuf;
                c.box_bytes = c.rows * c.Wr * 16;
                static const int dense_env = getenv("MISO_TC_DENSE") ? atoi(getenv("MISO_TC_DENSE")) : -1;
                const bool dense = dense_env >= 0 ? dense_env != 0 : kper > 1;
                if (dense) {  // planes are dense: one TMA box fills all of a stage's planes
                    c.PL = c.box_bytes;
                    c.ppb = 2 * kper;
                } else {      // one box per plane, planes padded so that no descriptor reaches past its plane
                    c.PL = round_up(std::max(c.rows * c.Wr, 128 * c.G + maxshift) * 16, 128);
                    c.ppb = 1;
                }
                c.GS = round_up(2 * kper * c.PL, 128);
                c.a_stage = c.nset * c.nsp * c.GS;
                c.stage = c.a_stage + round_up(c.w_stage, 128);
                // the last plane's descriptors reach (128 G + maxshift) rows past its start: junk rows only, but mapped memory
                const int tail = std::max(0, (128 * c.G + maxshift) * 16 - c.PL) + 128;
                c.nstage = std::min(kMaxStages, (kSmemLimit - c.off_stage - tail) / c.stage);
                if (c.nstage < 2) continue;
                c.ntiles = a.B * c.nNt * c.t_tiles * c.f_tiles;
                // a stage costs its MMAs plus a fixed hand-over (barrier round trip, descriptor set-up); loads must keep up
                const double main_cyc = (double)g.nunit * ntap_n * c.G * cyc_mma + 250.0 * c.nchunk;
                const double load_cyc = (double)g.nunit * (c.nset * c.nsp * 2 * c.box_bytes + g.w_unit) / 24.0 +
                                        (c.nstage >= 3 ? 0.0 : 1500.0 * c.nchunk);  // ~L2 -> SM bytes/clk; exposed latency when shallow
                const double epi_cyc = 0.35 * c.G * 128.0 * c.N * c.nacc + 1500.0;
                const double body = std::max(main_cyc, load_cyc);
                const double prod_cyc = 300.0 * c.nchunk * (c.nset * c.nsp + 2);  // one thread issues every TMA of the tile
                const double body2 = std::max(body, prod_cyc);
                const double tile_cyc = (nbuf == 2 ? std::max(body2, epi_cyc) + 500.0 : body2 + epi_cyc + 2500.0) + 20.0 * c.nchunk;
                const int ctas = std::min(c.ntiles, 148);
                // fewer than 3 stages exposes the load latency: only when nothing deeper fits
                const double total = std::ceil((double)c.ntiles / ctas) * tile_cyc + 6000.0 + (c.nstage < 3 ? 1e12 : 0.0);
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
    {
        int maxshift = 0;
        for (int i = 0; i < g.ntap; ++i) maxshift = std::max(maxshift, g.shift[i]);
        g.wr_magic = (unsigned)((0x100000000ull + (unsigned)g.Wr - 1) / (unsigned)g.Wr);
    g.smem_total = g.off_stage + g.nstage * g.stage + std::max(0, (128 * g.G + maxshift) * 16 - g.PL) + 128;
    }
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
    // the plane dimension ends with the view ([in_coff, in_coff + cin)): planes past it are zero-filled, never read
    const uint64_t CG = (uint64_t)a.in_ctot / 8, CGv = (uint64_t)(a.in_coff + a.cin + 7) / 8, T = (uint64_t)a.T, F = (uint64_t)a.Fin;
    const cuuint32_t ppb = (cuuint32_t)g.ppb;
    char *base = const_cast<char *>(reinterpret_cast<const char *>(a.in));
    for (int sp = 0; sp < 2; ++sp) {
        CUtensorMap *tm = sp == 0 ? hi : lo;
        void *addr = base + (sp ? a.in_lo_off : 0);
        CUresult r;
        if (!g.strided) {
            // [2F (8-byte units)][T][CG][B]: a tile row is one contiguous run of Wr * 16 bytes
            cuuint64_t dims[4] = {2 * F, T, CGv, (cuuint64_t)a.B};
            cuuint64_t strides[3] = {F * 16, T * F * 16, 2 * CG * T * F * 16};
            cuuint32_t box[4] = {(cuuint32_t)(2 * g.Wr), (cuuint32_t)g.rows, ppb, 1};
            cuuint32_t es[4] = {1, 1, 1, 1};
            r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, addr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        } else {
            // [8 ch][F][T][CG][B] with element stride 2 along F: the even / odd bins of a tile
            cuuint64_t dims[5] = {8, F, T, CGv, (cuuint64_t)a.B};
            cuuint64_t strides[4] = {16, F * 16, T * F * 16, 2 * CG * T * F * 16};
            cuuint32_t box[5] = {8, (cuuint32_t)(2 * g.Wr), (cuuint32_t)g.rows, ppb, 1};
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

void conv_tc_set_trace(long long *d_buf, int cin, int fin) {
    g_trace_buf = d_buf;
    g_trace_cin = cin;
    g_trace_fin = fin;
}

// one-time, capture-unsafe setup (function attributes, driver entry point); called before graph capture
int conv_tc_init() {
    // function attributes live in the device's context: track the opt-in per device ordinal
    static bool done_dev[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
    bool &done = done_dev[dev & 63];
    if (done) return MISO_OK;
    e = cudaFuncSetAttribute(conv_tc_kernel<3, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_kernel<3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(conv_tc_kernel)");
    static const bool carve = !(getenv("MISO_CARVEOUT") && atoi(getenv("MISO_CARVEOUT")) == 0);  // see conv_rs_init
    if (carve) {
        e = cudaFuncSetAttribute(conv_tc_prep_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(conv_tc_prep_kernel, carveout)");
    }
    if (!get_encode()) {
        set_error("conv_tc: cuTensorMapEncodeTiled is not available from the driver");
        return MISO_E_CUDA;
    }
    if (int rc = conv_rs_init()) return rc;
    if (int rc = tcn_pw_init()) return rc;
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
    if (a.out_layout == LAYOUT_CL_F32 && (a.cout % 2 || a.out_ctot % 2 || a.out_coff % 2)) return false;  // float4 / float2 stores
    if (a.resid && (a.resid_ctot % 4 || a.resid_coff % 4)) return false;
    TcGeom g;
    return make_geom(a, 3, g);
}

void conv_tc_scratch_need(const ConvArgs &a, int split, size_t *wimg_bytes, size_t *btab_bytes) {
    TcGeom g;
    *wimg_bytes = *btab_bytes = 0;
    if (!make_geom(a, split, g)) return;
    *wimg_bytes = (size_t)a.B * g.nNt * g.nunit * g.w_unit;
    *btab_bytes = (size_t)a.B * g.nNt * ((a.cin + kBiasCi - 1) / kBiasCi) * 9 * g.N * sizeof(float);
    for (int sp : {1, 3}) {  // the row-streaming variant has its own weight image layout
        size_t w = 0, bt = 0;
        if (conv_rs_eligible(a, sp)) conv_rs_scratch_need(a, sp, &w, &bt);
        *wimg_bytes = std::max(*wimg_bytes, w);
        *btab_bytes = std::max(*btab_bytes, bt);
    }
}

int launch_conv_tc(const ConvArgs &a, int split, const TcScratch &scratch, cudaStream_t stream) {
    MISO_REQUIRE(split == 1 || split == 3, "conv_tc: bad split");
    if (conv_rs_eligible(a, split)) return launch_conv_rs(a, split, scratch, stream);
    TcGeom g;
    MISO_REQUIRE(make_geom(a, split, g), "conv_tc: layer does not fit the tcgen05 path (cin=%d cout=%d Fin=%d)", a.cin, a.cout, a.Fin);
    const bool shared_w = a.norm_mode != NORM_IN;
    const int nb_bias = a.norm_mode == NORM_NONE ? 1 : a.B;
    const size_t need_w = (size_t)(shared_w ? 1 : a.B) * g.nNt * g.nunit * g.w_unit, need_b = (size_t)a.B * g.nNt * ((a.cin + kBiasCi - 1) / kBiasCi) * 9 * g.N * sizeof(float);
    if (need_w > scratch.wimg_bytes || need_b > scratch.btab_bytes) {
        set_error("conv_tc: scratch too small (%zu/%zu weight bytes, %zu/%zu bias bytes)", scratch.wimg_bytes, need_w,
                  scratch.btab_bytes, need_b);
        return MISO_E_WORKSPACE;
    }
    static const bool debug = getenv("MISO_TC_DEBUG") != nullptr;
    if (debug)
        fprintf(stderr,
                "conv_tc: cin=%d cout=%d Fin=%d Fout=%d s=%d tr=%d | N=%d nNt=%d TF=%d Wr=%d TT=%d G=%d nbuf=%d kper=%d nstage=%d stage=%dB "
                "tiles=%d (t%d x f%d) tmem=%d smem=%d\n",
                a.cin, a.cout, a.Fin, a.Fout, a.stride_f, a.transposed, g.N, g.nNt, g.TF, g.Wr, g.TT, g.G, g.nbuf, g.kper, g.nstage, g.stage,
                g.ntiles, g.t_tiles, g.f_tiles, g.tmem_cols, g.smem_total);
    CUtensorMap tm_hi, tm_lo;
    int rc = encode_maps(a, g, &tm_hi, &tm_lo);
    if (rc) return rc;

    PrepArgs p{};
    p.w = a.w;
    p.bias = a.bias;
    p.in_sums = a.in_sums;
    p.gamma = a.gamma;
    p.beta = a.beta;
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
    p.nunit = g.nunit;
    p.shared_w = shared_w ? 1 : 0;
    p.nb_bias = nb_bias;
    p.nsplit = (a.cin + kBiasCi - 1) / kBiasCi;
    p.nsp = g.nsp;
    p.ntap = g.ntap;
    p.KT = a.KT;
    p.KF = a.KF;
    for (int i = 0; i < g.ntap; ++i) p.tapk[i] = g.tapk[i];
    const size_t prep_smem = (size_t)(kBiasCi + 8 * 9 * g.N) * sizeof(float);
    prof_begin(stream);
    MISO_CUDA(launch_pdl_if(pdl_level() >= 1, conv_tc_prep_kernel, dim3((shared_w ? 1 : a.B) * g.nunit + nb_bias * g.nNt * p.nsplit), dim3(256), prep_smem, stream, p));
    prof_end(stream, 0.0, (double)need_w + (double)need_b, MISO_PROF_PREP);
    MISO_LAUNCHED("conv_tc_prep_kernel");
    prof_begin(stream);

    TcArgs k{};
    k.g = g;
    k.wimg = p.wimg;
    k.btab = p.btab;
    k.bias = a.bias;
    k.nsplit = p.nsplit;
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
    k.wimg_bstride = shared_w ? 0 : (size_t)g.nNt * g.nunit * (g.w_unit / 2);
    k.gln_sums = a.norm_mode == NORM_GLN ? a.in_sums : nullptr;
    k.bias_shared = nb_bias == 1 && a.B > 1 ? 1 : 0;
    k.gln_inv_n = a.norm_inv_n;
    k.gln_eps = a.norm_eps;
    k.resid = a.resid;
    k.resid_ctot = a.resid_ctot;
    k.resid_coff = a.resid_coff;
    k.trace = (g_trace_buf && a.cin == g_trace_cin && a.Fin == g_trace_fin) ? g_trace_buf : nullptr;

    rc = conv_tc_init();
    if (rc) return rc;
    dim3 grid(std::min(g.ntiles, 148), 1, 1);  // persistent: one CTA per SM walks the tile list
    static const int var = getenv("MISO_TC_VAR") ? atoi(getenv("MISO_TC_VAR")) : 0;
    if (split == 3) {
        if (var == 0)
            MISO_CUDA(launch_pdl(conv_tc_kernel<3, 0>, grid, dim3(kThreads), (size_t)g.smem_total, stream, tm_hi, tm_lo, k));
        else
            MISO_CUDA(launch_pdl(conv_tc_kernel<3, 1>, grid, dim3(kThreads), (size_t)g.smem_total, stream, tm_hi, tm_lo, k));
    } else {
        if (var == 0)
            MISO_CUDA(launch_pdl(conv_tc_kernel<1, 0>, grid, dim3(kThreads), (size_t)g.smem_total, stream, tm_hi, tm_lo, k));
        else
            MISO_CUDA(launch_pdl(conv_tc_kernel<1, 1>, grid, dim3(kThreads), (size_t)g.smem_total, stream, tm_hi, tm_lo, k));
    }
    {
        const double pix = (double)a.B * a.T * (a.transposed ? a.Fin : a.Fout);
        const double flops = 2.0 * pix * a.cin * a.cout * a.KT * a.KF;
        const double bytes = (a.use_lo ? 4.0 : 2.0) * a.B * a.T * ((double)a.Fin * a.cin + (double)a.Fout * a.cout);
        // issued: per tile, nunit K units x ntap taps x G M tiles x (one MMA of width 2N + one of width N in bf16x3)
        const double exec = (double)g.ntiles * g.nunit * g.ntap * g.G * 2.0 * 128.0 * 16.0 * (split == 3 ? 3.0 * g.N : 1.0 * g.N);
        prof_end(stream, flops, bytes, MISO_PROF_CONV_TC, exec);
    }
    MISO_LAUNCHED("conv_tc_kernel");
    return MISO_OK;
}

}  // namespace miso
