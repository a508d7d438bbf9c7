// Row-streaming tcgen05 3x3 convolution with the three FRAME taps merged into the MMA N dimension -- the
// DenseBlock convs of the MISO conv stack (model.py:437-482: 3x3, stride 1, pad (1,1); 94 % of the FLOPs of
// MISO_1/MISO_3) at every stage (F + 1 = 256, 128 bins; 64 ... 8 bins as packed strips).
//
// Why.  An SS-mode tcgen05.mma (M = 128, K = 16) costs 32 + N/4 cycles for N <= 128: the 128 x 16 A tile is read
// from shared memory at 128 B/clk (tools/umma_bench.cu), so with N = cout = 24..32 the tensor pipe idles 60 % of
// the time (conv_tc.cu).  Here one MMA computes, for one input frame row r and one bin tap kf,
//     E[r, (kt, co)] = sum_ci X[r, f + kf - 1, ci] * W[kt, kf, ci, co]        N = 3 * cout
// i.e. the contribution of input row r to the THREE output rows r + 1 - kt at once: the A tile is read once per
// 3 taps (56 cycles per 96 columns instead of 3 x 40.8), and since out[t] = sum_kt E[t + kt - 1, kt] uses the
// same TMEM lane (bin) of three different input rows, the sum over kt is free: the accumulator of output row t
// is one Nc-column TMEM slot, slots of consecutive output rows are adjacent ([kt=2 | kt=1 | kt=0] weight order),
// and the MMA of input row r simply lands on the slots of rows r-1, r, r+1.  A CTA streams down the frames of
// its strip: every input row is loaded ONCE (no frame halo re-reads) and multiplied once.
//
// TMEM is a ring of L logical slots (+ 2 extension slots so that an MMA never wraps: an MMA that starts in slot
// L-2 / L-1 spills into slots L, L+1, which the epilogue adds to slots 0, 1).  The epilogue hands every slot it
// has drained back cleared (tcgen05.st), so every MMA accumulates and the issuing thread -- whose time between
// two MMAs is on the critical path, the pipe queues only a couple of MMAs -- has no special cases.  Tiles of G
// input rows: the epilogue of tile k (output rows that became complete) overlaps the MMAs of tile k + 1;
// 2G + 2 <= L.
//
// Epilogue (8 warps, two per tensor-memory lane quarter).  An item = one output row x 16 accumulator columns of a warp's 32
// bins.  With cin <= 64 the kernel is bound by the epilogue, not by the MMAs (profiles/r2_conv_rs_epilogue_study.log), so:
// the two warps of a quarter split the 16-column CHUNKS (each takes every row; with Nc = 32 a warp owns one chunk and keeps
// its InstanceNorm statistics in registers for a whole strip), the tensor-memory load of the next row is issued as soon as
// this row's values have left the load registers, the arithmetic is packed fp32x2 (FADD2 / FMUL2 / FFMA2) with a branch-free
// ELU, and fp32 channels-last outputs (the data gradients, the network output) go through a per-warp transposition tile so
// that four lanes cover the 64 contiguous bytes of a bin instead of one lane per bin (32 lines per access).
//
// Layout in shared memory per stage (kper 16-channel K units): A planes [hi|lo][unit][kg][row][pitch px][8 ch]
// by one TMA box per plane set, then the per-sample weight image of the units (conv_rs_prep_kernel):
// [unit][kf][hi|lo][kg][3 Nc rows][8 ch].  F + 1 = 128: rows are stored at a pitch of 128 pixels starting at
// bin -1 (TMA zero fill); the right padding of row r IS the left padding of row r + 1 (shared-pad raster).
// F + 1 = 256: two column regions of 128 bins, rows at a pitch of 130 pixels (5-D bf16 tensor map).
// F + 1 <= 64: packed strips -- an M tile is the same frame of 128 / (F + 1) frame strips of one sample (RsGeom::S).
#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>

#include "conv.cuh"
#include "umma.cuh"

namespace miso {
namespace {

constexpr int kRsThreads = 10 * 32;  // warp 0: TMA producer, warp 1: MMA issuer, warps 2..9: epilogue (two per tensor-memory lane quarter)
constexpr int kRsEpi0 = 2;
constexpr int kRsEpiThreads = 256;
constexpr int kRsEpiWarps = kRsEpiThreads / 32;
constexpr int kRsEpiW = kRsEpiWarps / 4;  // epilogue warps per lane quarter
constexpr int kRsMaxG = 8;
constexpr int kRsMaxStages = 4;
constexpr int kRsSmemLimit = 227 * 1024;
constexpr int kRsFixed = 1024;
constexpr int kRsBiasCi = 64;

struct RsGeom {
    int Nc, N3, L, G, Mr, pitch, PL, GS, nsp;
    int nplanes, nunit, kper, nchunk;
    int w_unit, w_off, stage, nstage, box_bytes;
    int off_btab, off_red, off_stg, off_stage, smem_total, tmem_cols;
    int red_stride;  // floats per epilogue warp in the statistics reduction buffer
    int map5d;
    int S, TS, DS, Wr;  // packed mode (F + 1 <= 64): an M tile holds the same row of S frame strips; strip s covers the TS frames
                        // from s * DS, DS = T / S, TS = DS + T % S (the strips overlap by T % S frames so that they tile any T
                        // at a uniform stride; the duplicated frames count for the last strip only)
    // Layer variant (round 2).  The kernel computes, for grid bin j of an M tile and the three frame taps kt,
    //     out[t][ostride * j + ooff] = sum_kt sum_{i < nkf} in[t + kt - 1][j + tboff[i]] * W[kt][twk[i]]
    // with the bin taps read as pixel shifts of the staged rows (pixel 0 of a region = bin 128 m + f_org):
    //   DenseBlock conv (stride 1, pad 1)   nkf 3, tboff -1 0 1, f_org -1, out bin = j
    //   first conv (stride 1, pad 0)        nkf 3, tboff  0 1 2, f_org  0, Fout = Fin - 2
    //   transposed stride (1,2), even bins  nkf 2, tboff 0 -1 with W[.][0], W[.][2], out bin = 2 j      (frame taps flipped by the prep kernel)
    //   transposed stride (1,2), odd bins   nkf 1, tboff 0 with W[.][1],             out bin = 2 j + 1
    //   transposed stride 1, pad 0 (last)   nkf 3, tboff 0 -1 -2, f_org -2, Fout = Fin + 2
    int nkf, tshift[3], twk[3], tboff[3];
    int f_org, Wg, Fin, Fout, ostride, ooff;
};

struct RsArgs {
    RsGeom g;
    const __nv_bfloat16 *wimg;  // [B][nunit][kf][hi|lo][kg][N3][8]
    size_t wimg_bstride;        // elements per sample (0: one image for all samples -- no per-sample normalisation)
    size_t btab_bstride;        // floats per sample of the border-bias partial sums (0: shared)
    const float *btab;          // border-bias partial sums [B][nsplit][9][Nc]
    const float *bias;
    int nsplit;
    void *out;
    double *out_sums;
    int B, T, F;
    int in_coff;
    int out_ctot, out_coff, cout;
    size_t out_lo_off;
    int use_lo, elu;
    // output-channel chunks (cout > 64: the data gradients of the DenseBlock convs, cout = the forward's cin): the kernel walks
    // (chunk, sample, region, frame) rows; chunk c covers output channels [c * Nc, c * Nc + Nc) with its own weight image and
    // border-bias sums.  a.cout is the TOTAL output channel count.
    int nch;
    size_t wimg_cstride, btab_cstride;  // elements between the weight images / bias partial sums of consecutive chunks
    int out_cl;  // fp32 channels-last output [B][T*Fout][out_ctot]: 1 = the result is ADDED to it (data gradients), 2 = stored (the
                 // network output, model.py:418-423)
    long long *trace;  // debug: clock64 event log of CTA 0 (tools/tc_trace.py), or null
};

// trace regions: [0,4096) producer, [4096,8192) MMA issuer, [8192,12288) first epilogue warp; entries are (tag, clock)
__device__ __forceinline__ void rs_trace(long long *trace, int region, int &n, int tag) {
    if (trace && blockIdx.x == 0 && n < 2040) {
        trace[region * 4096 + 2 * n] = tag;
        trace[region * 4096 + 2 * n + 1] = clock64();
        ++n;
    }
}

struct RsPrepArgs {
    const float *w;  // packed fp32 [9][cin][cout_pad]
    const double *in_sums;
    int in_ctot, in_coff, cin, cout, cout_pad;
    int norm_mode;
    float norm_eps;
    double norm_inv_n;
    __nv_bfloat16 *wimg;
    float *btab;
    int B, Nc, nunit, nsp, nsplit;
    int nb;  // sample images / bias tables built: B, or 1 without a per-sample normalisation (data gradients, the first conv)
    int flip_t;  // transposed convs read input frame t + 1 - kt: the image / bias slots of frame tap kt take W[2 - kt]
    // output-channel chunks (blockIdx.y): chunk c builds the image / bias sums of output channels [c * Nc, c * Nc + Nc)
    size_t wimg_cstride, btab_cstride;
};

// the strips of one CTA: its share [rho, rho_end) of the flattened (sample, column region, frame) row space, cut
// at (sample, region) boundaries
struct RsWalk {
    int rho, rho_end, T, Mr, B;  // T: rows per unit (frames; packed mode: frames per strip)
    int b, m, ch, t0, TS, nin;
    __device__ __forceinline__ bool next() {
        if (rho >= rho_end) return false;
        const int unit = rho / T;
        t0 = rho - unit * T;
        TS = min(T - t0, rho_end - rho);
        const int bb = unit / Mr;  // (chunk, sample)
        m = unit - bb * Mr;
        ch = bb / B;
        b = bb - ch * B;
        nin = TS + 2;  // input rows j = 0 .. TS+1 <-> frames t0-1 .. t0+TS
        rho += TS;
        return true;
    }
};
__device__ __forceinline__ RsWalk rs_walk(const RsArgs &a) {
    RsWalk w;
    const int TU = a.g.S > 1 ? a.g.TS : a.T;
    const long long R = (long long)a.nch * a.B * a.g.Mr * TU;
    w.rho = (int)(R * blockIdx.x / gridDim.x);
    w.rho_end = (int)(R * (blockIdx.x + 1) / gridDim.x);
    w.T = TU;
    w.Mr = a.g.Mr;
    w.B = a.B;
    return w;
}

// The tiles of one strip: up to G consecutive input rows each.  Packed mode: the halo row above the first frame of
// a strip is the LAST frame of the strip before it (zero for strip 0) and the halo row below the last frame is the
// FIRST frame of the strip after it (zero for the last strip), i.e. the same TMA box moved by one strip; those two
// rows are tiles of their own (kind 1 / 2).
struct RsTiles {
    int j0, Gk, kind, jcur, nin, G;
    bool top, bottom;
    __device__ __forceinline__ void init(const RsArgs &a, const RsWalk &w) {
        jcur = 0;
        nin = w.nin;
        G = a.g.G;
        top = a.g.S > 1 && w.t0 == 0;
        bottom = a.g.S > 1 && w.t0 + w.TS == a.g.TS;
    }
    __device__ __forceinline__ bool next() {
        if (jcur >= nin) return false;
        j0 = jcur;
        const int lim = bottom ? nin - 1 : nin;
        if (top && jcur == 0) {
            Gk = 1;
            kind = 1;
        } else if (jcur >= lim) {
            Gk = 1;
            kind = 2;
        } else {
            Gk = min(G, lim - jcur);
            kind = 0;
        }
        jcur += Gk;
        return true;
    }
};

// packed fp32 pairs (FADD2 / FMUL2 / FFMA2): the epilogue is instruction-issue bound
__device__ __forceinline__ void add2(float &d0, float &d1, float a0, float a1, float b0, float b1) {
    asm("{\n\t.reg .b64 x, y, z;\n\tmov.b64 x, {%2, %3};\n\tmov.b64 y, {%4, %5};\n\tadd.rn.f32x2 z, x, y;\n\tmov.b64 {%0, %1}, z;\n\t}"
        : "=f"(d0), "=f"(d1)
        : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
__device__ __forceinline__ void mul2(float &d0, float &d1, float a0, float a1, float b0, float b1) {
    asm("{\n\t.reg .b64 x, y, z;\n\tmov.b64 x, {%2, %3};\n\tmov.b64 y, {%4, %5};\n\tmul.rn.f32x2 z, x, y;\n\tmov.b64 {%0, %1}, z;\n\t}"
        : "=f"(d0), "=f"(d1)
        : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
__device__ __forceinline__ void fma2(float &d0, float &d1, float a0, float a1, float b0, float b1, float c0, float c1) {
    asm("{\n\t.reg .b64 x, y, z, w;\n\tmov.b64 x, {%2, %3};\n\tmov.b64 y, {%4, %5};\n\tmov.b64 w, {%6, %7};\n\tfma.rn.f32x2 z, x, y, w;\n\tmov.b64 {%0, %1}, z;\n\t}"
        : "=f"(d0), "=f"(d1)
        : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c0), "f"(c1));
}
__device__ __forceinline__ float ex2_ftz(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int SPLIT>
__global__ void __launch_bounds__(kRsThreads, 1)
conv_rs_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo, const RsArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const RsGeom &g = a.g;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Nc = g.Nc, L = g.L, G = g.G;
    const long long t_start = a.trace ? clock64() : 0;
    constexpr int NPROD = SPLIT == 3 ? 3 : 1;

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
            mbar_init(bar_tempty + 8 * i, kRsEpiThreads / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)g.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    {
        for (int i = tid; i < kRsEpiWarps * g.red_stride; i += kRsThreads) red[i] = 0.f;
        // zero guard behind the A planes of every stage (shared-pad raster: the pixel after the last row)
        for (int i = tid; i < g.nstage * 32; i += kRsThreads)
            reinterpret_cast<uint32_t *>(smem + g.off_stage + (i >> 5) * g.stage + g.w_off - 128)[i & 31] = 0u;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    if (warp >= kRsEpi0) {  // all accumulator slots start cleared
        const int quad = warp & 3, sub = (warp - kRsEpi0) >> 2;
        for (int col = sub * 16; col < g.tmem_cols; col += 16 * kRsEpiW) tmem_zero16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)col);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    pdl_trigger();
    pdl_wait();  // everything above overlapped the previous kernel's tail; its results are needed from here on

    if (warp == 0) {
        // ---------------------------------------------------------------- TMA producer
        if (elect_one()) {
            const int plane0 = a.in_coff >> 3;
            int s = 0, ph = 0, ntr = 0;  // stage index and the parity of the release of its previous use
            bool primed = false;         // every stage has been filled once
            RsWalk w = rs_walk(a);
            while (w.next()) {
                const __nv_bfloat16 *wsrc = a.wimg + (size_t)w.ch * a.wimg_cstride + (size_t)w.b * a.wimg_bstride;
                const int f0 = 128 * w.m + g.f_org;
                RsTiles tl;
                tl.init(a, w);
                while (tl.next()) {
                    const int j0 = tl.j0;
                    const int tin = w.t0 - 1 + j0;
                    for (int c = 0; c < g.nchunk; ++c) {
                        rs_trace(a.trace, 0, ntr, 100 + c);
                        if (primed) mbar_wait(bar_empty + 8 * s, (uint32_t)ph);
                        rs_trace(a.trace, 0, ntr, 200 + c);
                        const int nu = min(g.kper, g.nunit - g.kper * c);
                        const uint32_t full = bar_full + 8 * s;
                        mbar_expect_tx(full, (uint32_t)(g.nsp * g.box_bytes + nu * g.w_unit));
                        const uint32_t sa = s_stage + (uint32_t)(s * g.stage);
                        const int pl = plane0 + 2 * g.kper * c;
                        for (int sp = 0; sp < g.nsp; ++sp) {
                            const CUtensorMap *tm = sp == 0 ? &tm_hi : &tm_lo;
                            const uint32_t dst = sa + (uint32_t)(sp * g.GS);
                            if (g.S > 1)  // {bins, strip, frame in strip, plane, sample}
                                tma_load_5d(dst, tm, full, 2 * g.f_org, tl.kind == 1 ? -1 : (tl.kind == 2 ? 1 : 0),
                                            tl.kind == 1 ? g.DS - 1 : (tl.kind == 2 ? g.TS - g.DS : tin), pl, w.b);
                            else if (g.map5d)
                                tma_load_5d(dst, tm, full, 0, f0, tin, pl, w.b);
                            else
                                tma_load_4d(dst, tm, full, 2 * f0, tin, pl, w.b);
                        }
                        bulk_load(sa + (uint32_t)g.w_off, wsrc + (size_t)c * g.kper * (g.w_unit / 2), (uint32_t)(nu * g.w_unit), full);
                        if (++s == g.nstage) {
                            s = 0;
                            if (primed) ph ^= 1;
                            primed = true;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ---------------------------------------------------------------- MMA issuer (one elected thread)
        // The tensor pipe queues only a couple of MMAs behind the running one, so anything on the issuing thread's
        // critical path between two MMAs (row bookkeeping, barrier polls) idles the pipe.  Hence: the per-row
        // parameters of a tile are computed once per tile, every MMA accumulates (the epilogue hands the slots it has
        // drained back ZEROED, so there is no first-write special case), and the row loop is unrolled.
        if (elect_one()) {
            const uint32_t idesc0 = make_idesc(0);
            constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);  // stride offset 128 B between 8-row groups, version 1
            const uint32_t a_lo0 = (s_stage >> 4) | (((uint32_t)g.PL >> 4) << 16);
            const uint32_t b_lo0 = ((s_stage + (uint32_t)g.w_off) >> 4) | ((uint32_t)g.N3 << 16);  // leading offset N3 * 16 B
            const uint32_t lo_split = (uint32_t)g.GS >> 4;
            const uint32_t a_kstep = (uint32_t)(2 * g.PL) >> 4, b_kstep = (uint32_t)g.w_unit >> 4;
            const uint32_t b_kfstep = (uint32_t)(g.nsp * 2 * g.N3), b_spstep = (uint32_t)(2 * g.N3);
            const uint32_t row_step = (uint32_t)g.pitch;
            const int nkf = g.nkf;
            uint32_t tap_a[3], tap_b[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                tap_a[i] = (uint32_t)g.tshift[i];
                tap_b[i] = (uint32_t)g.twk[i] * b_kfstep;
            }
            int s = 0, ph = 0, k = 0, Pcur = 0, ntr = 0;  // Pcur: ring position of the slot two rows above the tile's first input row
            RsWalk w = rs_walk(a);
            while (w.next()) {
                RsTiles tl;
                tl.init(a, w);
                for (; tl.next(); ++k) {
                    const int j0 = tl.j0, Gk = tl.Gk;
                    uint32_t rd[kRsMaxG], rb[kRsMaxG], rid[kRsMaxG];  // per row: accumulator address, weight row offset, idesc
#pragma unroll
                    for (int i = 0; i < kRsMaxG; ++i) {
                        const int j = j0 + i;
                        const int glo = max(0, 2 - j), ghi = min(2, w.TS + 1 - j);  // output row o = j - 2 + g in [0, TS)
                        int P = Pcur + i;
                        if (P >= L) P -= L;
                        rd[i] = tmem_base + (uint32_t)((P + glo) * Nc);
                        rb[i] = (uint32_t)(glo * Nc);
                        rid[i] = idesc0 | ((uint32_t)(((ghi - glo + 1) * Nc) >> 3) << 17);
                    }
                    rs_trace(a.trace, 1, ntr, 1000);
                    if (k >= 2) {  // the epilogue of tile k - 2 has drained the slots this tile reuses
                        mbar_wait(bar_tempty + 8 * (k & 1), ((k >> 1) + 1) & 1);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    }
                    for (int c = 0; c < g.nchunk; ++c) {
                        rs_trace(a.trace, 1, ntr, 100 + c);
                        mbar_wait(bar_full + 8 * s, (uint32_t)ph);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        rs_trace(a.trace, 1, ntr, 200 + c);
                        const uint32_t st4 = (uint32_t)(s * g.stage) >> 4;
                        const int nu = min(g.kper, g.nunit - g.kper * c);
#pragma unroll 1
                        for (int ks = 0; ks < nu; ++ks) {
                            const uint32_t a_ks = a_lo0 + st4 + (uint32_t)ks * a_kstep;
                            const uint32_t b_ks = b_lo0 + st4 + (uint32_t)ks * b_kstep;
#pragma unroll
                            for (int i = 0; i < kRsMaxG; ++i) {
                                if (i < Gk) {
                                    const uint32_t a_i = a_ks + (uint32_t)i * row_step;
                                    const uint32_t b_i = b_ks + rb[i];
#pragma unroll
                                    for (int kf = 0; kf < 3; ++kf) {
                                        if (kf < nkf) {  // bin tap kf: a pixel shift of the staged row, its own slice of the weight image
#pragma unroll
                                            for (int pr = 0; pr < NPROD; ++pr) {  // a_hi w_hi, a_hi w_lo, a_lo w_hi
                                                const uint32_t alo = a_i + tap_a[kf] + (pr == 2 ? lo_split : 0u);
                                                const uint32_t blo = b_i + tap_b[kf] + (pr == 1 ? b_spstep : 0u);
                                                umma_bf16(rd[i], ((uint64_t)kDescHi << 32) | (uint64_t)alo, ((uint64_t)kDescHi << 32) | (uint64_t)blo, rid[i], 1u);
                                            }
                                        }
                                    }
                                }
                            }
                        }
                        umma_commit(bar_empty + 8 * s);
                        if (c == g.nchunk - 1) umma_commit(bar_tfull + 8 * (k & 1));
                        rs_trace(a.trace, 1, ntr, 300 + c);
                        if (++s == g.nstage) {
                            s = 0;
                            ph ^= 1;
                        }
                    }
                    Pcur += Gk;
                    if (Pcur >= L) Pcur -= L;
                }
            }
        }
    } else {
        // ---------------------------------------------------------------- epilogue (warps 2..9)
        const int quad = warp & 3, sub = (warp - kRsEpi0) >> 2;
        const int et = tid - kRsEpi0 * 32;
        const int npix = a.T * g.Fout;
        float *myred = red + (warp - kRsEpi0) * g.red_stride;
        float *stg = reinterpret_cast<float *>(smem + g.off_stg);  // channels-last outputs: per-warp transposition tiles
        // Work split between the kRsEpiW warps of a tensor-memory lane quarter over the (row, 16-column chunk) items of a tile:
        // if the chunks of an accumulator slot divide evenly they split the CHUNKS (each warp takes every row), else the rows.
        // A warp that owns ONE chunk (Nc = 32, most layers) keeps its statistics in registers for the whole strip; otherwise
        // the partial sums are flushed per tile and chunk.
        const int nck = Nc >> 4;
        const bool by_chunk = nck % kRsEpiW == 0;
        const bool persist = by_chunk ? nck == kRsEpiW : nck == 1;
        const int cb0 = by_chunk ? 16 * sub : 0, cbstep = by_chunk ? 16 * kRsEpiW : 16;
        const int r0 = by_chunk ? 0 : sub, rstep = by_chunk ? 1 : kRsEpiW;
        int prev_b = -1, k = 0, xo = 2 % L, ntr = 0;  // xo: ring position of the next output row to drain
        const bool tracer = warp == kRsEpi0 && lane == 0;
        RsWalk w = rs_walk(a);
        while (w.next()) {
            const int b = w.b;
            const int cbase = w.ch * Nc;  // first output channel of this chunk
            if (w.ch * a.B + b != prev_b) {
                // border-bias table of this sample: add the channel splits of the prep kernel's partial sums, then
                // expand to the 64 (frame-mask, bin-mask) classes
                asm volatile("bar.sync 1, %0;" ::"n"(kRsEpiThreads));
                float *wb = red + kRsEpiWarps * g.red_stride;  // [9][Nc]
                const float *src = a.btab + (size_t)w.ch * a.btab_cstride + (size_t)b * a.btab_bstride;
                for (int i = et; i < 9 * Nc; i += kRsEpiThreads) {
                    float v = 0.f;
                    for (int sp = 0; sp < a.nsplit; ++sp) v += __ldg(src + (size_t)sp * 9 * Nc + i);
                    wb[i] = v;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(kRsEpiThreads));
                for (int i = et; i < 64 * Nc; i += kRsEpiThreads) {
                    const int mk = i / Nc, n = i - mk * Nc;
                    const int tm = mk >> 3, fm = mk & 7;
                    float v = (a.bias && cbase + n < a.cout) ? __ldg(a.bias + cbase + n) : 0.f;
                    for (int kt = 0; kt < 3; ++kt)
                        for (int kf = 0; kf < 3; ++kf)
                            if (((tm >> kt) & 1) && ((fm >> kf) & 1)) v += wb[(kt * 3 + kf) * Nc + n];
                    btab_s[i] = v;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(kRsEpiThreads));
                prev_b = w.ch * a.B + b;
            }
            const int seg = g.S > 1 ? (quad * 32 + lane) / g.Wr : 0;  // packed mode: which strip this lane belongs to
            const int jg = g.S > 1 ? (quad * 32 + lane) - seg * g.Wr : 128 * w.m + quad * 32 + lane;  // grid bin of this lane
            const int f = g.ostride * jg + g.ooff;                                                     // its output bin
            const int tseg = seg * g.DS;
            const bool fvalid = jg < g.Wg && f < g.Fout;
            const bool last_seg = seg == g.S - 1;
            int fmask = 0;  // bit kf of the ORIGINAL weight: that bin tap reads inside the input tensor
#pragma unroll
            for (int i = 0; i < 3; ++i)
                if (i < g.nkf && jg + g.tboff[i] >= 0 && jg + g.tboff[i] < g.Fin) fmask |= 1 << g.twk[i];
            __nv_bfloat16 *out_pl = reinterpret_cast<__nv_bfloat16 *>(a.out) + (size_t)b * 2 * a.out_ctot * npix;
            const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16);
            float ssum[16], ssq[16];  // statistics of the chunk in flight
#pragma unroll
            for (int q = 0; q < 16; ++q) ssum[q] = ssq[q] = 0.f;
            // this warp's partial sums of one chunk -> its slot of the strip's reduction buffer (fixed order, no atomics)
            auto flush = [&](int cb) {
                const float s = warp_reduce16(ssum, lane);
                const float q2 = warp_reduce16(ssq, lane);
                const int slot = (persist ? 0 : cb) + (lane >> 1);  // a persistent warp's buffer holds its own chunk only
                if ((lane & 1) == 0) {
                    myred[slot * 2] += s;
                    myred[slot * 2 + 1] += q2;
                }
#pragma unroll
                for (int q = 0; q < 16; ++q) ssum[q] = ssq[q] = 0.f;
            };
            RsTiles tl;
            tl.init(a, w);
            for (; tl.next(); ++k) {
                const int j0 = tl.j0, Gk = tl.Gk;
                const int o_lo = max(0, j0 - 2), o_hi = min(w.TS, j0 + Gk - 2);
                if (tracer) rs_trace(a.trace, 2, ntr, 1);
                mbar_wait(bar_tfull + 8 * (k & 1), (k >> 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (tracer) rs_trace(a.trace, 2, ntr, 2);
                for (int cb = cb0; cb < Nc; cb += cbstep) {
                    auto slot_addr = [&](int o) {
                        int x = xo + o - o_lo;
                        if (x >= L) x -= L;
                        return tlane + (uint32_t)(x * Nc + cb);
                    };
                    // one row of the chunk: v = the 16 accumulator columns of this lane's bin, already loaded
                    auto item = [&](uint32_t (&v)[16], int o) {
                        const int t = tseg + w.t0 + o;
                        const bool valid = fvalid && (last_seg || w.t0 + o < g.DS);  // frames a strip shares with the next one belong to that one
                        int x = xo + o - o_lo;
                        if (x >= L) x -= L;
                        const uint32_t taddr = tlane + (uint32_t)(x * Nc + cb);
                        float y[16];
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                        for (int q = 0; q < 16; ++q) y[q] = __uint_as_float(v[q]);
                        if (x < 2) {  // the ring's extension slots hold the rest of rows 0 and 1
                            uint32_t v2[16];
                            tmem_ld16(taddr + (uint32_t)(L * Nc), v2);
                            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                            tmem_zero16(taddr + (uint32_t)(L * Nc));
#pragma unroll
                            for (int q = 0; q < 16; q += 2) add2(y[q], y[q + 1], y[q], y[q + 1], __uint_as_float(v2[q]), __uint_as_float(v2[q + 1]));
                        }
                        tmem_zero16(taddr);  // hand the slot back cleared: every MMA accumulates
                        {   // border bias of this (frame, bin) class: the taps that read outside the tensor drop out
                            const int tmask = 7 & ~(t == 0 ? 1 : 0) & ~(t == a.T - 1 ? 4 : 0);
                            const float *bt = btab_s + (tmask * 8 + fmask) * Nc + cb;
#pragma unroll
                            for (int q = 0; q < 16; q += 4) {
                                const float4 bb = *reinterpret_cast<const float4 *>(bt + q);
                                add2(y[q], y[q + 1], y[q], y[q + 1], bb.x, bb.y);
                                add2(y[q + 2], y[q + 3], y[q + 2], y[q + 3], bb.z, bb.w);
                            }
                        }
                        if (o + rstep < o_hi) tmem_ld16(slot_addr(o + rstep), v);
                        if (a.elu) {  // ELU(y) = max(y, exp(min(y, 0)) - 1), branch-free
#pragma unroll
                            for (int q = 0; q < 16; q += 2) {
                                float e0, e1;
                                mul2(e0, e1, fminf(y[q], 0.f), fminf(y[q + 1], 0.f), 1.4426950408889634f, 1.4426950408889634f);
                                add2(e0, e1, ex2_ftz(e0), ex2_ftz(e1), -1.f, -1.f);
                                y[q] = fmaxf(y[q], e0);
                                y[q + 1] = fmaxf(y[q + 1], e1);
                            }
                        }
                        if (valid) {
#pragma unroll
                            for (int q = 0; q < 16; q += 2) {
                                add2(ssum[q], ssum[q + 1], ssum[q], ssum[q + 1], y[q], y[q + 1]);
                                fma2(ssq[q], ssq[q + 1], y[q], y[q + 1], y[q], y[q + 1], ssq[q], ssq[q + 1]);
                            }
                            const int pix = t * g.Fout + f;
                            if (!a.out_cl)
#pragma unroll
                            for (int g8 = 0; g8 < 16; g8 += 8) {
                                const int co = cb + g8;
                                if (cbase + co < a.cout) {
                                    const int ca = a.out_coff + cbase + co;
                                    __nv_bfloat16 *p = out_pl + ((size_t)(ca >> 3) * npix + pix) * 8;
                                    uint32_t hp[4];
#pragma unroll
                                    for (int q = 0; q < 4; ++q) hp[q] = pack_bf16x2(y[g8 + 2 * q], y[g8 + 2 * q + 1]);
                                    *reinterpret_cast<uint4 *>(p) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
                                    if (a.use_lo) {
                                        uint32_t lp[4];
#pragma unroll
                                        for (int q = 0; q < 4; ++q) {
                                            float l0, l1;
                                            fma2(l0, l1, bf16_lo(hp[q]), bf16_hi(hp[q]), -1.f, -1.f, y[g8 + 2 * q], y[g8 + 2 * q + 1]);
                                            lp[q] = pack_bf16x2(l0, l1);
                                        }
                                        *reinterpret_cast<uint4 *>(reinterpret_cast<char *>(p) + a.out_lo_off) =
                                            make_uint4(lp[0], lp[1], lp[2], lp[3]);
                                    }
                                }
                            }
                        }
                        if (a.out_cl) {
                            // fp32 channels-last output (1: added to the gradient buffer, 2: stored).  One bin per lane would make
                            // every 16-byte access of the warp touch 32 different lines of the buffer; the [32 bins x 16 channels]
                            // tile is turned through shared memory so that four lanes cover the 64 contiguous bytes of a bin.
                            const unsigned vmask = __ballot_sync(0xffffffffu, valid);
                            const int pix = t * g.Fout + f;
                            float *stw = stg + (warp - kRsEpi0) * (32 * 20);
                            __syncwarp();  // the previous item's readers are done with the tile
#pragma unroll
                            for (int q = 0; q < 16; q += 4) *reinterpret_cast<float4 *>(stw + lane * 20 + q) = make_float4(y[q], y[q + 1], y[q + 2], y[q + 3]);
                            __syncwarp();
                            const int ch = cbase + cb + 4 * (lane & 3);
                            float *optr[4];
                            float4 vv[4], old[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const int p = (lane >> 2) + 8 * i;
                                const int pixp = __shfl_sync(0xffffffffu, pix, p);
                                const bool ok = ((vmask >> p) & 1u) && ch < a.cout;
                                optr[i] = ok ? reinterpret_cast<float *>(a.out) + ((size_t)b * npix + pixp) * a.out_ctot + a.out_coff + ch : nullptr;
                                vv[i] = *reinterpret_cast<const float4 *>(stw + p * 20 + 4 * (lane & 3));
                                old[i] = (optr[i] && a.out_cl == 1) ? *reinterpret_cast<const float4 *>(optr[i]) : make_float4(0.f, 0.f, 0.f, 0.f);
                            }
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                if (optr[i])
                                    *reinterpret_cast<float4 *>(optr[i]) =
                                        make_float4(vv[i].x + old[i].x, vv[i].y + old[i].y, vv[i].z + old[i].z, vv[i].w + old[i].w);
                        }
                    };
                    // rows of the tile: the accumulator load of the next row is issued as soon as this row's values have left
                    // the load registers, and is in flight while this row is processed
                    uint32_t v[16];
                    int o = o_lo + r0;
                    if (o < o_hi) tmem_ld16(slot_addr(o), v);
                    for (; o < o_hi; o += rstep) item(v, o);
                    if (!persist && a.out_sums) flush(cb);
                }
                // this warp is done with the tile's slots: hand them back to the MMA issuer
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (tracer) rs_trace(a.trace, 2, ntr, 3);
                if (lane == 0) mbar_arrive(bar_tempty + 8 * (k & 1));
                xo += max(0, o_hi - o_lo);
                if (xo >= L) xo -= L;
            }
            if (persist && a.out_sums) flush(cb0);
            if (a.out_sums) {  // strip statistics -> global fixed-point accumulators
                asm volatile("bar.sync 1, %0;" ::"n"(kRsEpiThreads));
                if (et < Nc && cbase + et < a.cout) {
                    double *dst = a.out_sums + ((size_t)b * a.out_ctot + a.out_coff + cbase + et) * 2;
                    double s8 = 0.0, q8 = 0.0;
                    const int myck = et >> 4, slot = persist ? (et & 15) : et;
                    for (int w8 = 0; w8 < kRsEpiWarps; ++w8) {  // the warps that own this channel's chunk, in a fixed order
                        const bool owns = !by_chunk || (w8 >> 2) == myck % kRsEpiW;
                        if (owns) {
                            s8 += (double)red[w8 * g.red_stride + slot * 2];
                            q8 += (double)red[w8 * g.red_stride + slot * 2 + 1];
                        }
                    }
                    stat_add(dst, s8);
                    stat_add(dst + 1, q8);
                }
                asm volatile("bar.sync 1, %0;" ::"n"(kRsEpiThreads));
                for (int i = et; i < kRsEpiWarps * g.red_stride; i += kRsEpiThreads) red[i] = 0.f;
            }
            xo += 2;  // the two slots between strips stay unused
            if (xo >= L) xo -= L;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (a.trace && tid == 0) a.trace[3 * 4096 + blockIdx.x] = clock64() - t_start;  // per-CTA duration (4th trace region)
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)g.tmem_cols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// Per-sample operand preparation.  Blocks [0, B * nunit): weight image of one (sample, 16-channel K unit)
//   wimg[b][unit][kf][hi|lo][kg][n = (2 - kt) * Nc + co][8] = bf16 split of W[kt][kf][ci][co] * rstd[b][ci]
// Blocks behind them: border-bias partial sums wbp[b][split][kt*3+kf][co] = sum_{ci in split} W * shift[b][ci]
// (x_norm = x * rstd + shift is the consumer-side InstanceNorm affine, model.py:445; padding is applied after it).
__device__ __forceinline__ float2 rs_affine(const RsPrepArgs &p, int b, int ci) {
    if (p.norm_mode == NORM_IN) {
        const double *s = p.in_sums + ((size_t)b * p.in_ctot + p.in_coff + ci) * 2;
        return affine_from_sums(stat_get(s), stat_get(s + 1), p.norm_inv_n, (double)p.norm_eps);
    }
    return make_float2(1.f, 0.f);
}

__global__ void __launch_bounds__(256) conv_rs_prep_kernel(const RsPrepArgs pa) {
    extern __shared__ float sh[];
    pdl_trigger();
    pdl_wait();  // the statistics come from the previous conv, which also still reads the scratch this kernel rewrites
    RsPrepArgs p = pa;
    {   // this block's output-channel chunk: a column window of the packed weights
        const int c0 = blockIdx.y * p.Nc;
        p.w += c0;
        p.cout = min(p.Nc, p.cout - c0);
        p.wimg += (size_t)blockIdx.y * p.wimg_cstride;
        p.btab += (size_t)blockIdx.y * p.btab_cstride;
    }
    const int nimg = p.nb * p.nunit;
    const int N3 = 3 * p.Nc;
    if ((int)blockIdx.x < nimg) {
        const int b = blockIdx.x / p.nunit, unit = blockIdx.x - b * p.nunit;
        float *scale = sh;
        if (threadIdx.x < 16) {
            const int ci = unit * 16 + threadIdx.x;
            scale[threadIdx.x] = ci < p.cin ? rs_affine(p, b, ci).x : 0.f;
        }
        __syncthreads();
        const size_t unit_elems = (size_t)3 * p.nsp * 2 * N3 * 8;
        __nv_bfloat16 *img = p.wimg + ((size_t)b * p.nunit + unit) * unit_elems;
#pragma unroll 3
        for (int i = threadIdx.x; i < 3 * 2 * N3; i += blockDim.x) {
            const int n = i % N3;
            const int r = i / N3;
            const int kg = r & 1, kf = r >> 1;
            const int ktg = n / p.Nc, co = n - ktg * p.Nc;
            const int kt = p.flip_t ? ktg : 2 - ktg;  // weight frame tap of image block ktg
            float h[8], l[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int ci = unit * 16 + kg * 8 + e;
                float v = 0.f;
                if (ci < p.cin && co < p.cout) v = p.w[((size_t)(kt * 3 + kf) * p.cin + ci) * p.cout_pad + co] * scale[kg * 8 + e];
                h[e] = bf16_round(v);
                l[e] = v - h[e];
            }
            __nv_bfloat16 *dst = img + ((size_t)((kf * p.nsp) * 2 + kg) * N3 + n) * 8;
            *reinterpret_cast<uint4 *>(dst) =
                make_uint4(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]), pack_bf16x2(h[4], h[5]), pack_bf16x2(h[6], h[7]));
            if (p.nsp == 2)
                *reinterpret_cast<uint4 *>(dst + (size_t)2 * N3 * 8) =
                    make_uint4(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]), pack_bf16x2(l[4], l[5]), pack_bf16x2(l[6], l[7]));
        }
    } else {
        int idx = blockIdx.x - nimg;
        const int split = idx % p.nsplit;
        const int b = idx / p.nsplit;
        float *shift = sh;            // [kRsBiasCi]
        float *wb = sh + kRsBiasCi;   // [nparts <= 8][9][Nc]
        const int ci0 = split * kRsBiasCi, nci = min(kRsBiasCi, p.cin - ci0);
        for (int i = threadIdx.x; i < nci; i += blockDim.x) shift[i] = rs_affine(p, b, ci0 + i).y;
        const int n = threadIdx.x % p.Nc, part = threadIdx.x / p.Nc, nparts = min(8, (int)blockDim.x / p.Nc);
        for (int i = threadIdx.x; i < nparts * 9 * p.Nc; i += blockDim.x) wb[i] = 0.f;
        __syncthreads();
        if (part < nparts && n < p.cout) {
            // all taps' loads in flight at once (the weights are cold in L2 behind the previous conv's activation stream)
            float acc[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) acc[k] = 0.f;
#pragma unroll 2
            for (int ci = part; ci < nci; ci += nparts) {
                const float sv = shift[ci];
#pragma unroll
                for (int k = 0; k < 9; ++k) acc[k] = fmaf(__ldg(p.w + ((size_t)k * p.cin + ci0 + ci) * p.cout_pad + n), sv, acc[k]);
            }
            // slot k' = (kt', kf) of the kernel's table is the conv-equivalent tap: W[2 - kt'] for a transposed conv
#pragma unroll
            for (int k = 0; k < 9; ++k) wb[(part * 9 + (p.flip_t ? (2 - k / 3) * 3 + k % 3 : k)) * p.Nc + n] = acc[k];
        }
        __syncthreads();
        float *dst = p.btab + ((size_t)b * p.nsplit + split) * 9 * p.Nc;
        for (int i = threadIdx.x; i < 9 * p.Nc; i += blockDim.x) {
            float v = 0.f;
            for (int q = 0; q < nparts; ++q) v += wb[q * 9 * p.Nc + i];
            dst[i] = v;
        }
    }
}


// ------------------------------------------------------------------------------------------------
typedef CUresult (*RsEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                               const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
RsEncodeFn rs_get_encode() {
    static RsEncodeFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<RsEncodeFn>(p);
    }
    return fn;
}

int rs_round_up(int x, int m) { return (x + m - 1) / m * m; }

// Layer variants of the row-streaming kernel (RsGeom): a transposed stride-(1,2) conv is two launches (even / odd output
// bins) over the same weight image.
enum RsKind { RS_DENSE = 0, RS_FIRST, RS_DECONV2_EVEN, RS_DECONV2_ODD, RS_DECONV1 };

int rs_kinds(const ConvArgs &a, int kinds[2]) {
    static const bool variants = !(getenv("MISO_RS_VARIANTS") && atoi(getenv("MISO_RS_VARIANTS")) == 0);
    if (a.KT != 3 || a.KF != 3 || a.pad_t != 1) return 0;
    if (!a.transposed && a.stride_f == 1 && a.pad_f == 1 && a.Fin == a.Fout) {
        kinds[0] = RS_DENSE;
        return 1;
    }
    if (!variants) return 0;
    if (!a.transposed && a.stride_f == 1 && a.pad_f == 0 && a.Fout == a.Fin - 2) {
        kinds[0] = RS_FIRST;
        return 1;
    }
    // Two launches (even / odd output bins) re-read the input and write every other 16-byte pixel: measured on B200
    // (profiles/r2_conv_rs_variants.json) 238 us against 245 for the general kernel at 127 -> 255 bins and SLOWER below
    // (157 vs 137 us at 63 -> 127, 91 vs 79, 62 vs 51), so the pair is opt-in.
    static const bool deconv2 = getenv("MISO_RS_DECONV2") && atoi(getenv("MISO_RS_DECONV2")) != 0;
    if (deconv2 && a.transposed && a.stride_f == 2 && a.pad_f == 0 && a.Fout == 2 * a.Fin + 1) {
        kinds[0] = RS_DECONV2_EVEN;
        kinds[1] = RS_DECONV2_ODD;
        return 2;
    }
    if (a.transposed && a.stride_f == 1 && a.pad_f == 0 && a.Fout == a.Fin + 2) {
        kinds[0] = RS_DECONV1;
        return 1;
    }
    return 0;
}

bool rs_shape_ok(const ConvArgs &a, int kind) {
    if (a.in_layout != LAYOUT_PLANES) return false;
    if (a.in_ctot % 8 || a.in_coff % 8) return false;
    if (a.out_layout == LAYOUT_PLANES) {
        if (a.out_ctot % 8 || a.out_coff % 8 || a.cout % 8 || a.resid) return false;
    } else if (a.resid) {
        // fp32 channels-last output as an in-place accumulation (resid == out, the data gradients): no ELU / statistics
        if (kind != RS_DENSE || a.cout % 8) return false;
        if (a.resid != reinterpret_cast<const float *>(a.out) || a.resid_ctot != a.out_ctot || a.resid_coff != a.out_coff) return false;
        if (a.out_ctot % 4 || a.out_coff % 4 || a.elu || a.out_sums) return false;
    } else {
        // fp32 channels-last output, stored (the network output)
        if (a.cout % 4 || a.out_ctot % 4 || a.out_coff % 4 || a.out_sums) return false;
    }
    if (a.norm_mode == NORM_GLN) return false;
    if (a.T < 2) return false;
    return true;
}

// reads past the Wr pixels of a row (strip) land on the first pixels of the next one: fine iff those reads are padding
// reads AND the aliased pixels are zero fill (the shared-pad raster)
bool rs_alias_ok(const RsGeom &g, int Wr) {
    for (int j = 0; j < g.Wg; ++j)
        for (int i = 0; i < g.nkf; ++i) {
            const int px = j + g.tshift[i];
            if (px < Wr) continue;
            const int src = j + g.tboff[i], alias = g.f_org + px - Wr;
            if (src >= 0 && src < g.Fin) return false;
            if (alias >= 0 && alias < g.Fin) return false;
        }
    return g.Fin - g.f_org <= Wr;  // every input bin lies inside the loaded pixels
}

bool make_rs_geom(const ConvArgs &a, int split, int kind, RsGeom &g) {
    g = RsGeom{};
    if (!rs_shape_ok(a, kind)) return false;
    g.nsp = split == 3 ? 2 : 1;
    static const int max_nc = getenv("MISO_RS_MAXNC") ? atoi(getenv("MISO_RS_MAXNC")) : 64;
    {
        // more output channels than one accumulator slot holds (the data gradients of the DenseBlock convs: cout = the
        // forward's cin): nch chunks of Nc channels, evenly sized, walked by ONE launch
        const int nch = (a.cout + max_nc - 1) / max_nc;
        g.Nc = rs_round_up((a.cout + nch - 1) / nch, 16);
        if (g.Nc > max_nc) return false;
        static const bool chunks_ok = !(getenv("MISO_RS_CHUNKS") && atoi(getenv("MISO_RS_CHUNKS")) == 0);
        if (nch > 1 && (!chunks_ok || kind != RS_DENSE || a.out_sums)) return false;
    }
    g.N3 = 3 * g.Nc;
    g.Fin = a.Fin;
    g.Fout = a.Fout;
    g.ostride = 1;
    g.ooff = 0;
    g.nkf = 3;
    for (int i = 0; i < 3; ++i) g.twk[i] = i;
    switch (kind) {
        case RS_DENSE:
            g.tboff[0] = -1, g.tboff[1] = 0, g.tboff[2] = 1;
            g.f_org = -1;
            g.Wg = a.Fin;
            break;
        case RS_FIRST:
            g.tboff[0] = 0, g.tboff[1] = 1, g.tboff[2] = 2;
            g.f_org = 0;
            g.Wg = a.Fout;
            break;
        case RS_DECONV2_EVEN:  // out[2 j] = in[j] W[.][0] + in[j - 1] W[.][2]
            g.nkf = 2;
            g.tboff[0] = 0, g.tboff[1] = -1;
            g.twk[0] = 0, g.twk[1] = 2;
            g.f_org = -1;
            g.Wg = a.Fin + 1;
            g.ostride = 2;
            break;
        case RS_DECONV2_ODD:  // out[2 j + 1] = in[j] W[.][1]
            g.nkf = 1;
            g.tboff[0] = 0;
            g.twk[0] = 1;
            g.f_org = -1;
            g.Wg = a.Fin;
            g.ostride = 2;
            g.ooff = 1;
            break;
        case RS_DECONV1:  // out[j] = sum_kf in[j - kf] W[.][kf]
            g.tboff[0] = 0, g.tboff[1] = -1, g.tboff[2] = -2;
            g.f_org = -2;
            g.Wg = a.Fin + 2;
            break;
        default:
            return false;
    }
    int max_shift = 0;
    for (int i = 0; i < g.nkf; ++i) {
        g.tshift[i] = g.tboff[i] - g.f_org;
        max_shift = std::max(max_shift, g.tshift[i]);
    }
    // M tiling of the Wg grid bins of a frame row: packed strips (Wr <= 64 pixels, S strips of one sample per M tile), one
    // region on the shared-pad raster (pitch 128), or Mr regions of 128 bins at a pitch of 128 + max_shift (5-D tensor map)
    int wr = 8;
    while (wr < std::max(g.Wg, g.Fin - g.f_org)) wr <<= 1;
    if (wr <= 64) {
        static const int packed_min = getenv("MISO_RS_PACKED_MINF") ? atoi(getenv("MISO_RS_PACKED_MINF")) : 7;
        if (!rs_alias_ok(g, wr) || a.Fin < packed_min || a.T / (128 / wr) < 4) return false;
        g.Wr = wr;
        g.Mr = 1;
        g.map5d = 0;
        g.pitch = 128;
    } else {
        g.Wr = 128;
        g.Mr = (g.Wg + 127) / 128;
        g.map5d = (g.Mr >= 2 || !rs_alias_ok(g, 128)) ? 1 : 0;
        g.pitch = g.map5d ? 128 + max_shift : 128;
    }
    g.S = 128 / g.Wr;
    g.DS = a.T / g.S;
    g.TS = g.DS + a.T % g.S;
    g.nplanes = (a.cin + 7) / 8;
    g.nunit = (g.nplanes + 1) / 2;
    g.w_unit = 3 * g.nsp * 2 * g.N3 * 16;
    g.L = std::min(512 / g.Nc - 2, 30);
    g.off_btab = kRsFixed;
    g.off_red = g.off_btab + 64 * g.Nc * 4;
    {
        const int nck = g.Nc / 16;
        const bool persist = nck % kRsEpiW == 0 ? nck == kRsEpiW : nck == 1;
        g.red_stride = persist ? 32 : 2 * g.Nc;
    }
    g.off_stg = g.off_red + (kRsEpiWarps * g.red_stride + 9 * g.Nc) * 4;
    g.off_stage = rs_round_up(g.off_stg + (a.out_layout == LAYOUT_CL_F32 ? kRsEpiWarps * 32 * 20 * 4 : 0), 1024);
    static const int g_env = getenv("MISO_RS_G") ? atoi(getenv("MISO_RS_G")) : 0;
    for (int G = std::min({(g.L - 2) / 2, kRsMaxG, g_env > 0 ? g_env : kRsMaxG}); G >= 1; --G) {
        for (int kper : {2, 1}) {
            if (kper > g.nunit) continue;
            if (split == 3 && kper == 2) continue;
            g.G = G;
            g.kper = kper;
            g.nchunk = (g.nunit + kper - 1) / kper;
            g.PL = G * g.pitch * 16;
            g.box_bytes = 2 * kper * g.PL;
            g.GS = rs_round_up(g.box_bytes, 128);
            g.w_off = g.nsp * g.GS + 128;  // 128-byte zero guard behind the planes
            g.stage = rs_round_up(g.w_off + kper * g.w_unit, 1024);
            g.nstage = std::min(kRsMaxStages, (kRsSmemLimit - g.off_stage) / g.stage);
            if (g.nstage >= 3) break;
        }
        if (g.nstage >= 3) break;
    }
    if (g.nstage < 3) return false;
    if (!g.map5d && g.GS != g.box_bytes) return false;  // the shared-pad raster needs dense plane sets
    g.smem_total = g.off_stage + g.nstage * g.stage;
    int cols = 32;
    while (cols < (g.L + 2) * g.Nc) cols <<= 1;
    g.tmem_cols = cols;
    return cols <= 512;
}
// the geometry of the first launch of a layer (the phases of a transposed stride-2 conv share everything the callers size)
bool make_rs_geom(const ConvArgs &a, int split, RsGeom &g) {
    int kinds[2];
    const int n = rs_kinds(a, kinds);
    if (n == 0) return false;
    for (int i = n - 1; i >= 0; --i)
        if (!make_rs_geom(a, split, kinds[i], g)) return false;
    return true;
}

int rs_encode_maps(const ConvArgs &a, const RsGeom &g, CUtensorMap *hi, CUtensorMap *lo) {
    RsEncodeFn enc = rs_get_encode();
    if (!enc) {
        set_error("conv_rs: cuTensorMapEncodeTiled is not available from the driver");
        return MISO_E_CUDA;
    }
    const uint64_t CG = (uint64_t)a.in_ctot / 8, CGv = (uint64_t)(a.in_coff + a.cin + 7) / 8, T = (uint64_t)a.T, F = (uint64_t)a.Fin;
    char *base = const_cast<char *>(reinterpret_cast<const char *>(a.in));
    for (int sp = 0; sp < 2; ++sp) {
        CUtensorMap *tm = sp == 0 ? hi : lo;
        void *addr = base + (sp ? a.in_lo_off : 0);
        CUresult r;
        if (g.S > 1) {
            // packed strips: {bins (8-byte units), strip, frame in strip, plane, sample}; a box holds the same G frames of
            // S consecutive strips, stored [plane][frame][strip][Wr pixels] = one M tile per frame
            cuuint64_t dims[5] = {2 * F, (cuuint64_t)g.S, (cuuint64_t)g.TS, CGv, (cuuint64_t)a.B};
            cuuint64_t strides[4] = {(cuuint64_t)g.DS * F * 16, F * 16, T * F * 16, 2 * CG * T * F * 16};
            cuuint32_t box[5] = {(cuuint32_t)(2 * g.Wr), (cuuint32_t)g.S, (cuuint32_t)g.G, (cuuint32_t)(2 * g.kper), 1};
            cuuint32_t es[5] = {1, 1, 1, 1, 1};
            r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 5, addr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        } else if (!g.map5d) {
            cuuint64_t dims[4] = {2 * F, T, CGv, (cuuint64_t)a.B};
            cuuint64_t strides[3] = {F * 16, T * F * 16, 2 * CG * T * F * 16};
            cuuint32_t box[4] = {(cuuint32_t)(2 * g.pitch), (cuuint32_t)g.G, (cuuint32_t)(2 * g.kper), 1};
            cuuint32_t es[4] = {1, 1, 1, 1};
            r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, addr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        } else {
            cuuint64_t dims[5] = {8, F, T, CGv, (cuuint64_t)a.B};
            cuuint64_t strides[4] = {16, F * 16, T * F * 16, 2 * CG * T * F * 16};
            cuuint32_t box[5] = {8, (cuuint32_t)g.pitch, (cuuint32_t)g.G, (cuuint32_t)(2 * g.kper), 1};
            cuuint32_t es[5] = {1, 1, 1, 1, 1};
            r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, addr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        }
        if (r != CUDA_SUCCESS) {
            set_error("conv_rs: cuTensorMapEncodeTiled failed (%d) for F=%d T=%d ctot=%d G=%d", (int)r, a.Fin, a.T, a.in_ctot, g.G);
            return MISO_E_CUDA;
        }
    }
    return MISO_OK;
}

long long *g_rs_trace = nullptr;
int g_rs_trace_cin = 0, g_rs_trace_fin = 0;

}  // namespace

void conv_rs_set_trace(long long *d_buf, int cin, int fin) {
    g_rs_trace = d_buf;
    g_rs_trace_cin = cin;
    g_rs_trace_fin = fin;
}

int conv_rs_init() {
    static bool done_dev[64] = {};  // per device ordinal: function attributes live in the device's context
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
    bool &done = done_dev[dev & 63];
    if (done) return MISO_OK;
    e = cudaFuncSetAttribute(conv_rs_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRsSmemLimit);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_rs_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRsSmemLimit);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(conv_rs_kernel)");
    // the preparation kernel runs between two conv kernels that take (almost) all shared memory: ask for the same carve-out so
    // that the SMs do not reconfigure their L1 / shared split twice per layer (MISO_CARVEOUT=0: the driver's default)
    static const bool carve = !(getenv("MISO_CARVEOUT") && atoi(getenv("MISO_CARVEOUT")) == 0);
    if (carve) {
        e = cudaFuncSetAttribute(conv_rs_prep_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(conv_rs_prep_kernel, carveout)");
    }
    done = true;
    return MISO_OK;
}

bool conv_rs_eligible(const ConvArgs &a, int split) {
    static const bool off = getenv("MISO_RS") && atoi(getenv("MISO_RS")) == 0;
    if (off) return false;
    RsGeom g;
    return make_rs_geom(a, split, g);
}

void conv_rs_scratch_need(const ConvArgs &a, int split, size_t *wimg_bytes, size_t *btab_bytes) {
    RsGeom g;
    *wimg_bytes = *btab_bytes = 0;
    if (!make_rs_geom(a, split, g)) return;
    const int nch = (a.cout + g.Nc - 1) / g.Nc;
    *wimg_bytes = (size_t)nch * a.B * g.nunit * g.w_unit;
    *btab_bytes = (size_t)nch * a.B * ((a.cin + kRsBiasCi - 1) / kRsBiasCi) * 9 * g.Nc * sizeof(float);
}

namespace {
long long rs_rows(const ConvArgs &a, const RsGeom &g) { return (long long)a.B * g.Mr * (g.S > 1 ? g.TS : a.T); }
unsigned rs_grid(const ConvArgs &a, const RsGeom &g) { return (unsigned)std::min<long long>(148, std::max<long long>(1, rs_rows(a, g) / 2)); }
int rs_launch(const ConvArgs &a, int split, const RsGeom &g, const __nv_bfloat16 *wimg, const float *btab, int nsplit, cudaStream_t stream) {
    CUtensorMap tm_hi, tm_lo;
    int rc = rs_encode_maps(a, g, &tm_hi, &tm_lo);
    if (rc) return rc;
    rc = conv_rs_init();
    if (rc) return rc;
    RsArgs k{};
    k.g = g;
    k.wimg = wimg;
    const int nb = a.norm_mode == NORM_IN ? a.B : 1;
    k.wimg_bstride = nb > 1 ? (size_t)g.nunit * (g.w_unit / 2) : 0;
    k.btab_bstride = nb > 1 ? (size_t)nsplit * 9 * g.Nc : 0;
    k.btab = btab;
    k.bias = a.bias;
    k.nsplit = nsplit;
    k.out = a.out;
    k.out_sums = a.out_sums;
    k.B = a.B;
    k.T = a.T;
    k.F = a.Fin;
    k.in_coff = a.in_coff;
    k.out_ctot = a.out_ctot;
    k.out_coff = a.out_coff;
    k.cout = a.cout;
    k.nch = (a.cout + g.Nc - 1) / g.Nc;
    k.wimg_cstride = (size_t)nb * g.nunit * (g.w_unit / 2);
    k.btab_cstride = (size_t)nb * nsplit * 9 * g.Nc;
    k.out_lo_off = a.out_lo_off;
    k.use_lo = a.use_lo;
    k.elu = a.elu;
    k.out_cl = a.out_layout == LAYOUT_CL_F32 ? (a.resid ? 1 : 2) : 0;
    k.trace = (g_rs_trace && a.cin == g_rs_trace_cin && a.Fin == g_rs_trace_fin) ? g_rs_trace : nullptr;
    dim3 grid(rs_grid(a, g), 1, 1);
    prof_begin(stream);
    if (split == 3)
        MISO_CUDA(launch_pdl(conv_rs_kernel<3>, grid, dim3(kRsThreads), (size_t)g.smem_total, stream, tm_hi, tm_lo, k));
    else
        MISO_CUDA(launch_pdl(conv_rs_kernel<1>, grid, dim3(kRsThreads), (size_t)g.smem_total, stream, tm_hi, tm_lo, k));
    {
        // a launch's share of the layer: nkf of the three bin taps (a transposed conv's MACs are counted per INPUT pixel)
        const double pix = (double)a.B * a.T * (a.transposed ? a.Fin : a.Fout);
        const double flops = 2.0 * pix * a.cin * a.cout * 3 * g.nkf;
        const double bytes = (a.use_lo ? 4.0 : 2.0) * a.B * a.T * ((double)a.Fin * a.cin + (double)a.Fout * a.cout) * g.nkf / 3.0;
        // issued: every input row (incl. the two halo rows of a strip: ignored here) is one M = 128 tile per column region,
        // nunit K units x 3 bin taps x (3 products in bf16x3) MMAs of width N3
        const double exec = (double)rs_rows(a, g) * g.nunit * g.nkf * (split == 3 ? 3.0 : 1.0) * 2.0 * 128.0 * g.N3 * 16.0;
        prof_end(stream, flops, bytes, MISO_PROF_CONV_RS, exec);
    }
    MISO_LAUNCHED("conv_rs_kernel");
    return MISO_OK;
}
}  // namespace

int launch_conv_rs(const ConvArgs &a, int split, const TcScratch &scratch, cudaStream_t stream) {
    int kinds[2];
    const int nkind = rs_kinds(a, kinds);
    RsGeom g;
    MISO_REQUIRE(nkind > 0 && make_rs_geom(a, split, g), "conv_rs: layer does not fit the row-streaming path (cin=%d cout=%d F=%d)", a.cin, a.cout,
                 a.Fin);
    const int nsplit = (a.cin + kRsBiasCi - 1) / kRsBiasCi;
    const int nch = (a.cout + g.Nc - 1) / g.Nc;
    const size_t need_w = (size_t)nch * a.B * g.nunit * g.w_unit, need_b = (size_t)nch * a.B * nsplit * 9 * g.Nc * sizeof(float);
    if (need_w > scratch.wimg_bytes || need_b > scratch.btab_bytes) {
        set_error("conv_rs: scratch too small (%zu/%zu weight bytes, %zu/%zu bias bytes)", scratch.wimg_bytes, need_w, scratch.btab_bytes,
                  need_b);
        return MISO_E_WORKSPACE;
    }
    int rc = conv_rs_init();
    if (rc) return rc;

    RsPrepArgs p{};
    p.w = a.w;
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
    p.Nc = g.Nc;
    p.nunit = g.nunit;
    p.nsp = g.nsp;
    p.nsplit = nsplit;
    p.flip_t = a.transposed ? 1 : 0;
    p.nb = a.norm_mode == NORM_IN ? a.B : 1;
    p.wimg_cstride = (size_t)p.nb * g.nunit * (g.w_unit / 2);
    p.btab_cstride = (size_t)p.nb * nsplit * 9 * g.Nc;
    const size_t prep_smem = (size_t)(kRsBiasCi + 8 * 9 * g.Nc) * sizeof(float);
    prof_begin(stream);
    MISO_CUDA(launch_pdl_if(pdl_level() >= 1, conv_rs_prep_kernel, dim3(p.nb * g.nunit + p.nb * nsplit, nch), dim3(256), prep_smem, stream, p));
    prof_end(stream, 0.0, (double)need_w + (double)need_b, MISO_PROF_PREP);
    MISO_LAUNCHED("conv_rs_prep_kernel");
    static const bool debug = getenv("MISO_TC_DEBUG") != nullptr;
    for (int i = 0; i < nkind; ++i) {  // one launch per phase, all over the same weight image / bias sums
        MISO_REQUIRE(make_rs_geom(a, split, kinds[i], g), "conv_rs: geometry of variant %d", kinds[i]);
        if (debug)
            fprintf(stderr, "conv_rs: kind %d cin=%d cout=%d Fin=%d Fout=%d | Wg=%d S=%d Wr=%d Nc=%d L=%d G=%d Mr=%d pitch=%d map5d=%d kper=%d nchunk=%d nstage=%d stage=%dB tmem=%d smem=%d\n",
                    kinds[i], a.cin, a.cout, a.Fin, a.Fout, g.Wg, g.S, g.Wr, g.Nc, g.L, g.G, g.Mr, g.pitch, g.map5d, g.kper, g.nchunk, g.nstage, g.stage,
                    g.tmem_cols, g.smem_total);
        rc = rs_launch(a, split, g, p.wimg, p.btab, nsplit, stream);
        if (rc) return rc;
    }
    return MISO_OK;
}

}  // namespace miso
