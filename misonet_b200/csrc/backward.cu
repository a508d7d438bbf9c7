// Backward kernels of the MISO network body (fp32 FMA; first correct training path).
// reference trainer.py:159-212 calls loss.backward() through
// model.py:76-111 (conv -> ELU -> InstanceNorm2d units model.py:401-466, TCN model.py:486-567, gLN model.py:609-632).
//
// Conventions (net.cu): activations are stored as RAW ELU outputs (bf16 hi/lo planes) plus the fixed-point (sum, sumsq)
// statistics of every (sample, channel); consumers read z = e * rstd + shift.  The gradient buffer of an activation
// buffer is fp32 channels-last and accumulates dL/dz from every consumer (data gradients of the layers that read it);
// the producing layer's backward then turns its channel range in place into dL/dy (InstanceNorm + ELU backward) and
// feeds the weight-gradient GEMM and the data-gradient conv (conv_fp32.cu with transposed weights).
#include <algorithm>
#include <cstdlib>

#include "bwd.cuh"

namespace miso {

int conv_fp32_tile_n(int cout);

namespace {

constexpr int kEw = 256;

__device__ __forceinline__ void load_e4(const __nv_bfloat16 *p, size_t lo_elems, int use_lo, float *e) {
    const uint2 h = *reinterpret_cast<const uint2 *>(p);
    e[0] = bf16_lo(h.x);
    e[1] = bf16_hi(h.x);
    e[2] = bf16_lo(h.y);
    e[3] = bf16_hi(h.y);
    if (use_lo) {
        const uint2 l = *reinterpret_cast<const uint2 *>(p + lo_elems);
        e[0] += bf16_lo(l.x);
        e[1] += bf16_hi(l.x);
        e[2] += bf16_lo(l.y);
        e[3] += bf16_hi(l.y);
    }
}

__device__ __forceinline__ void split4(const float *v, uint2 &hi, uint2 &lo) {
    const uint32_t h0 = pack_bf16x2(v[0], v[1]), h1 = pack_bf16x2(v[2], v[3]);
    hi = make_uint2(h0, h1);
    lo = make_uint2(pack_bf16x2(v[0] - bf16_lo(h0), v[1] - bf16_hi(h0)), pack_bf16x2(v[2] - bf16_lo(h1), v[3] - bf16_hi(h1)));
}

// ---- InstanceNorm2d + ELU backward ------------------------------------------------------------------------------
// pass 1: per (b, channel) sums of dz and dz * z over the pixels
__global__ void __launch_bounds__(kEw) in_bwd_reduce_kernel(const InBwdArgs a, int pix_per_cta) {
    extern __shared__ float sh[];  // [c][2]
    const int b = blockIdx.y;
    const int nc4 = a.c >> 2;
    const int lanes = kEw / nc4;
    const int c4 = threadIdx.x % nc4, lane = threadIdx.x / nc4;
    for (int i = threadIdx.x; i < 2 * a.c; i += kEw) sh[i] = 0.f;
    __syncthreads();
    if (lane < lanes) {
        const int ca = a.coff + c4 * 4;
        float2 af[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const double *s = a.sums + ((size_t)b * a.ctot + ca + q) * 2;
            af[q] = affine_from_sums(stat_get(s), stat_get(s + 1), a.inv_n, (double)a.eps);
        }
        const size_t lo = (size_t)a.ctot * a.npix;
        const __nv_bfloat16 *eb = a.e + (size_t)b * 2 * lo + (size_t)(ca >> 3) * a.npix * 8 + (ca & 7);
        const float *gb = a.g + (size_t)b * a.npix * a.ctot + ca;
        const int p0 = blockIdx.x * pix_per_cta, p1 = min(a.npix, p0 + pix_per_cta);
        float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
        for (int p = p0 + lane; p < p1; p += lanes) {
            const float4 dz = *reinterpret_cast<const float4 *>(gb + (size_t)p * a.ctot);
            float e[4];
            load_e4(eb + (size_t)p * 8, lo, a.use_lo, e);
            const float d[4] = {dz.x, dz.y, dz.z, dz.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float z = fmaf(e[q], af[q].x, af[q].y);
                s1[q] += d[q];
                s2[q] = fmaf(d[q], z, s2[q]);
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            atomicAdd(&sh[(c4 * 4 + q) * 2], s1[q]);
            atomicAdd(&sh[(c4 * 4 + q) * 2 + 1], s2[q]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * a.c; i += kEw) atomicAdd(&a.red[(size_t)b * a.c * 2 + i], (double)sh[i]);
}

// pass 2: dy = rstd * (dz - mean(dz) - z * mean(dz * z)) * ELU'(y), in place; bias gradient = sum dy
__global__ void __launch_bounds__(kEw) in_bwd_apply_kernel(const InBwdArgs a, int pix_per_cta) {
    extern __shared__ float sh[];  // [c]
    const int b = blockIdx.y;
    const int nc4 = a.c >> 2;
    const int lanes = kEw / nc4;
    const int c4 = threadIdx.x % nc4, lane = threadIdx.x / nc4;
    for (int i = threadIdx.x; i < a.c; i += kEw) sh[i] = 0.f;
    __syncthreads();
    if (lane < lanes) {
        const int ca = a.coff + c4 * 4;
        float2 af[4];
        float m1[4], m2[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            af[q] = make_float2(1.f, 0.f);
            m1[q] = m2[q] = 0.f;
            if (!a.plain) {
                const double *s = a.sums + ((size_t)b * a.ctot + ca + q) * 2;
                af[q] = affine_from_sums(stat_get(s), stat_get(s + 1), a.inv_n, (double)a.eps);
                const double *r = a.red + ((size_t)b * a.c + c4 * 4 + q) * 2;
                m1[q] = (float)(r[0] * a.inv_n);
                m2[q] = (float)(r[1] * a.inv_n);
            }
        }
        const size_t lo = (size_t)a.ctot * a.npix;
        const __nv_bfloat16 *eb = a.plain ? nullptr : a.e + (size_t)b * 2 * lo + (size_t)(ca >> 3) * a.npix * 8 + (ca & 7);
        float *gb = a.g + (size_t)b * a.npix * a.ctot + ca;
        const int p0 = blockIdx.x * pix_per_cta, p1 = min(a.npix, p0 + pix_per_cta);
        float sb[4] = {0.f, 0.f, 0.f, 0.f};
        for (int p = p0 + lane; p < p1; p += lanes) {
            const float4 dz = *reinterpret_cast<const float4 *>(gb + (size_t)p * a.ctot);
            float d[4] = {dz.x, dz.y, dz.z, dz.w};
            if (!a.plain) {
                float e[4];
                load_e4(eb + (size_t)p * 8, lo, a.use_lo, e);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float z = fmaf(e[q], af[q].x, af[q].y);
                    const float de = af[q].x * (d[q] - m1[q] - z * m2[q]);
                    d[q] = e[q] > 0.f ? de : de * (e[q] + 1.f);  // ELU'(y) = exp(y) = e + 1 for y <= 0
                }
                *reinterpret_cast<float4 *>(gb + (size_t)p * a.ctot) = make_float4(d[0], d[1], d[2], d[3]);
            }
            if (a.dyp) {
                const int cl = c4 * 4;
                __nv_bfloat16 *dp = a.dyp + (size_t)b * 2 * a.c * a.npix + ((size_t)(cl >> 3) * a.npix + p) * 8 + (cl & 7);
                uint2 hi, lo2;
                split4(d, hi, lo2);
                *reinterpret_cast<uint2 *>(dp) = hi;
                *reinterpret_cast<uint2 *>(dp + (size_t)a.c * a.npix) = lo2;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) sb[q] += d[q];
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) atomicAdd(&sh[c4 * 4 + q], sb[q]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < a.c_real; i += kEw) atomicAdd(&a.dbias[i], sh[i]);
}

// ---- weight gradient ---------------------------------------------------------------------------------------------
// GEMM per tap: dW_tap [cin x cout] = A_tap^T [cin x K] * dY [K x cout], K = B * T * Fout pixels.  A CTA owns one
// (tap, 64-channel cin tile, BN-channel cout tile) and one slice of the pixels of EVERY sample; register tile 4 x TN.
constexpr int kWgBM = 64, kWgBK = 16;

template <int TN>
__global__ void __launch_bounds__(256) wgrad_kernel(const WgradArgs a, int splits) {
    constexpr int BN = 16 * TN;
    extern __shared__ __align__(16) float wsm[];
    float *As = wsm;                                                      // [BK][BM + 4]
    float *Bs = As + kWgBK * (kWgBM + 4);                                 // [BK][BN + 4]
    float2 *aff = reinterpret_cast<float2 *>(Bs + kWgBK * (BN + 4));      // [B][BM]
    const int tid = threadIdx.x;
    const int ci_tiles = (a.cin + kWgBM - 1) / kWgBM, co_tiles = (a.cout + BN - 1) / BN;
    int tile = blockIdx.y;
    const int cot = tile % co_tiles;
    tile /= co_tiles;
    const int cit = tile % ci_tiles;
    const int tap = tile / ci_tiles;
    const int kt = tap / a.KF, kf = tap - kt * a.KF;
    const int ci0 = cit * kWgBM, co0 = cot * BN;
    const int npix = a.T * a.Fout;
    const int per = (npix + splits - 1) / splits;
    const int p0 = blockIdx.x * per, p1 = min(npix, p0 + per);

    for (int i = tid; i < a.B * kWgBM; i += 256) {
        const int b = i / kWgBM, c = ci0 + (i - b * kWgBM);
        float2 v = make_float2(1.f, 0.f);
        if (a.x_sums && c < a.cin) {
            const double *s = a.x_sums + ((size_t)b * a.x_ctot + a.x_coff + c) * 2;
            v = affine_from_sums(stat_get(s), stat_get(s + 1), a.inv_n, (double)a.eps);
        }
        aff[i] = v;
    }
    __syncthreads();

    const int lp = tid >> 4, l4 = tid & 15;
    const int nchunk = p1 > p0 ? (p1 - p0 + kWgBK - 1) / kWgBK : 0;
    const int niter = nchunk * a.B;
    const bool planes = a.x_layout == LAYOUT_PLANES;
    const size_t x_lo = (size_t)a.x_ctot * a.T * a.Fin;  // bf16 elements from hi to lo plane set

    float4 ra, rb;
    auto load = [&](int it) {
        const int b = it / nchunk;
        const int p = p0 + (it - b * nchunk) * kWgBK + lp;
        ra = make_float4(0.f, 0.f, 0.f, 0.f);
        rb = ra;
        if (p >= p1) return;
        const int t = p / a.Fout, f = p - t * a.Fout;
        const int c = ci0 + l4 * 4;
        if (c < a.cin) {
            int ti, fi;
            bool ok = true;
            if (!a.transposed) {
                ti = t + kt - a.pad_t;
                fi = f * a.stride_f + kf - a.pad_f;
            } else {
                ti = t + a.pad_t - kt;
                const int num = f + a.pad_f - kf;
                fi = num / a.stride_f;
                ok = num >= 0 && fi * a.stride_f == num;
            }
            ok = ok && ti >= 0 && ti < a.T && fi >= 0 && fi < a.Fin;
            if (ok) {
                const int ca = a.x_coff + c;
                float e[4];
                if (planes) {
                    const __nv_bfloat16 *xp = reinterpret_cast<const __nv_bfloat16 *>(a.x) + (size_t)b * 2 * x_lo +
                                              (((size_t)(ca >> 3) * a.T + ti) * a.Fin + fi) * 8 + (ca & 7);
                    load_e4(xp, x_lo, a.use_lo, e);
                } else {
                    const float4 v = *reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(a.x) +
                                                                       (((size_t)b * a.T + ti) * a.Fin + fi) * a.x_ctot + ca);
                    e[0] = v.x;
                    e[1] = v.y;
                    e[2] = v.z;
                    e[3] = v.w;
                }
                const float2 *af = aff + b * kWgBM + l4 * 4;
                ra = make_float4(fmaf(e[0], af[0].x, af[0].y), fmaf(e[1], af[1].x, af[1].y), fmaf(e[2], af[2].x, af[2].y),
                                 fmaf(e[3], af[3].x, af[3].y));
                // channels beyond cin inside the last quad cannot occur: cin % 4 == 0
            }
        }
        if (l4 < BN / 4) {
            const int co = co0 + l4 * 4;
            if (co < a.cout)
                rb = *reinterpret_cast<const float4 *>(a.dy + ((size_t)b * npix + p) * a.dy_ctot + a.dy_coff + co);
        }
    };

    float acc[4][TN];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    const int ty = tid >> 4, tx = tid & 15;

    if (niter > 0) load(0);
    for (int it = 0; it < niter; ++it) {
        *reinterpret_cast<float4 *>(As + lp * (kWgBM + 4) + l4 * 4) = ra;
        if (l4 < BN / 4) *reinterpret_cast<float4 *>(Bs + lp * (BN + 4) + l4 * 4) = rb;
        __syncthreads();
        if (it + 1 < niter) load(it + 1);
#pragma unroll
        for (int k = 0; k < kWgBK; ++k) {
            const float4 a4 = *reinterpret_cast<const float4 *>(As + k * (kWgBM + 4) + ty * 4);
            const float av[4] = {a4.x, a4.y, a4.z, a4.w};
            float bv[TN];
            if constexpr (TN == 4) {
                const float4 b4 = *reinterpret_cast<const float4 *>(Bs + k * (BN + 4) + tx * 4);
                bv[0] = b4.x;
                bv[1] = b4.y;
                bv[2] = b4.z;
                bv[3] = b4.w;
            } else {
                const float2 b2 = *reinterpret_cast<const float2 *>(Bs + k * (BN + 4) + tx * 2);
                bv[0] = b2.x;
                bv[1] = b2.y;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }

    const int taps = a.KT * a.KF;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int ci = ci0 + ty * 4 + i;
        if (ci >= a.cin) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int co = co0 + tx * TN + j;
            if (co >= a.cout_real) continue;
            const size_t idx = a.transposed ? ((size_t)ci * a.cout_real + co) * taps + tap : ((size_t)co * a.cin + ci) * taps + tap;
            atomicAdd(a.dw + idx, acc[i][j]);
        }
    }
}

// ---- weight gradient on the tensor cores (bf16 hi/lo split, three MMAs per product, fp32 accumulate) ---------------
// Same GEMM and the same tile ownership as wgrad_kernel, with the inner product on mma.sync m16n8k16: the loader
// normalises the input pixels, splits xhat and dy into bf16 hi / lo halves (hi + lo carries 16-17 mantissa bits, as the
// forward's activation planes do) and stores them [pixel][channel]; ldmatrix.trans turns those K-major rows into the
// A (cin x pixels) and B (pixels x cout) fragments.  Warps 0-3 own 16 input channels each of the first 16 pixels of
// a 32-pixel chunk, warps 4-7 the same channels of the second 16 pixels; both halves add into dW with atomics.
constexpr int kWtAff = 80;   // float2 per sample in the affine tables: 8 groups of 8 channels at a pitch of 10 (conflict-free 16-byte reads)
constexpr int kWmAP = kWgBM + 8;  // bf16 row pitch of the A tiles: 144 bytes, ldmatrix rows fall into distinct banks

__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float *c, const uint32_t *a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int BN, int SL>
__global__ void __launch_bounds__(256) wgrad_mma_kernel(const WgradArgs a, int splits) {
    constexpr int BK = 32 * SL;         // pixels per chunk: SL independent loads per thread in flight
    constexpr int BP = BN + 8;          // bf16 row pitch of the B tiles
    constexpr int NB4 = BN / 32;        // float4 of dy per thread and pixel
    extern __shared__ __align__(16) unsigned char wsm_raw[];
    __nv_bfloat16 *Ah = reinterpret_cast<__nv_bfloat16 *>(wsm_raw);   // [BK][AP]
    __nv_bfloat16 *Al = Ah + BK * kWmAP;
    __nv_bfloat16 *Bh = Al + BK * kWmAP;                              // [BK][BP]
    __nv_bfloat16 *Bl = Bh + BK * BP;
    float2 *aff = reinterpret_cast<float2 *>(Bl + BK * BP);           // [B][kWtAff] (8-channel groups at a pitch of 10)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ci_tiles = (a.cin + kWgBM - 1) / kWgBM, co_tiles = (a.cout + BN - 1) / BN;
    int tile = blockIdx.x;  // tiles (tap, cin tile, cout tile) of one pixel slice are neighbours in launch order: they share x and dy in L2
    const int cot = tile % co_tiles;
    tile /= co_tiles;
    const int cit = tile % ci_tiles;
    const int tap = tile / ci_tiles;
    const int kt = tap / a.KF, kf = tap - kt * a.KF;
    const int ci0 = cit * kWgBM, co0 = cot * BN;
    const int npix = a.T * a.Fout;
    const int per = (npix + splits - 1) / splits;
    const int p0 = blockIdx.y * per, p1 = min(npix, p0 + per);

    for (int i = tid; i < a.B * kWgBM; i += 256) {
        const int b = i / kWgBM, cl = i - b * kWgBM, c = ci0 + cl;
        float2 v = make_float2(1.f, 0.f);
        if (a.x_sums && c < a.cin) {
            const double *s = a.x_sums + ((size_t)b * a.x_ctot + a.x_coff + c) * 2;
            v = affine_from_sums(stat_get(s), stat_get(s + 1), a.inv_n, (double)a.eps);
        }
        aff[b * kWtAff + (cl >> 3) * 10 + (cl & 7)] = v;
    }
    __syncthreads();

    const int ap = tid >> 3, a8 = tid & 7;        // loaders: pixels ap + 32 sl of the chunk; A: channels 8 a8 .. + 7 (one 16-byte
    const int b4 = tid & 7;                       // plane pixel), B: channels 4 b4 (+ 32 j)
    const int nchunk = p1 > p0 ? (p1 - p0 + BK - 1) / BK : 0;
    const int niter = nchunk * a.B;
    const bool planes = a.x_layout == LAYOUT_PLANES;
    const size_t x_lo = (size_t)a.x_ctot * a.T * a.Fin;
    const int cg = ci0 + a8 * 8;

    // The loader only ISSUES the loads; normalising and splitting happen at store time one iteration later (an arithmetic
    // use right behind a load would stall the warp for the memory latency before it reaches its MMAs).
    uint4 rh[SL], rl[SL];
    float4 rb[SL][NB4];
    unsigned rvalid = 0;
    int r_sample = 0;
    auto load = [&](int it) {
        const int b = it / nchunk;
        const int pc = p0 + (it - b * nchunk) * BK;
        rvalid = 0;
        r_sample = b;
#pragma unroll
        for (int sl = 0; sl < SL; ++sl) {
            const int p = pc + ap + 32 * sl;
            rh[sl] = make_uint4(0u, 0u, 0u, 0u);
            rl[sl] = rh[sl];
#pragma unroll
            for (int j = 0; j < NB4; ++j) rb[sl][j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p >= p1) continue;
            const int t = p / a.Fout, f = p - t * a.Fout;
            int ti, fi;
            bool ok = cg < a.cin;
            if (!a.transposed) {
                ti = t + kt - a.pad_t;
                fi = f * a.stride_f + kf - a.pad_f;
            } else {
                ti = t + a.pad_t - kt;
                const int num = f + a.pad_f - kf;
                fi = num / a.stride_f;
                ok = ok && num >= 0 && fi * a.stride_f == num;
            }
            ok = ok && ti >= 0 && ti < a.T && fi >= 0 && fi < a.Fin;
            if (ok) {
                const int ca = a.x_coff + cg;
                if (planes) {
                    const __nv_bfloat16 *xp = reinterpret_cast<const __nv_bfloat16 *>(a.x) + (size_t)b * 2 * x_lo +
                                              (((size_t)(ca >> 3) * a.T + ti) * a.Fin + fi) * 8;
                    rh[sl] = *reinterpret_cast<const uint4 *>(xp);
                    if (a.use_lo) rl[sl] = *reinterpret_cast<const uint4 *>(xp + x_lo);
                } else {
                    const float *xp = reinterpret_cast<const float *>(a.x) + (((size_t)b * a.T + ti) * a.Fin + fi) * a.x_ctot + ca;
                    rh[sl] = *reinterpret_cast<const uint4 *>(xp);       // fp32 channels 0-3 (raw bits)
                    rl[sl] = *reinterpret_cast<const uint4 *>(xp + 4);   // fp32 channels 4-7
                }
                rvalid |= 1u << sl;
            }
#pragma unroll
            for (int j = 0; j < NB4; ++j) {
                const int co = co0 + b4 * 4 + 32 * j;
                if (co < a.cout) rb[sl][j] = *reinterpret_cast<const float4 *>(a.dy + ((size_t)b * npix + p) * a.dy_ctot + a.dy_coff + co);
            }
        }
    };

    constexpr int NT = BN / 8;  // n tiles of 8 channels
    float acc[NT][4];
#pragma unroll
    for (int i = 0; i < NT; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
    const int wm = warp & 3, wk = warp >> 2;
    // ldmatrix lane addresses (see the fragment layouts of mma.m16n8k16): A matrices (m 0-7 | m 8-15) x (k 0-7 | k 8-15),
    // B matrices (k 0-7 | k 8-15) x (n 0-7 | n 8-15); warps 0-3 take the even 16-pixel steps of a chunk, warps 4-7 the odd ones
    const int lr = lane & 7, lm = lane >> 3;
    const uint32_t a_off = (uint32_t)(((wk * 16 + (lm >> 1) * 8 + lr) * kWmAP + wm * 16 + (lm & 1) * 8) * 2);
    const uint32_t b_off = (uint32_t)(((wk * 16 + (lm & 1) * 8 + lr) * BP + (lm >> 1) * 8) * 2);
    const uint32_t sAh = (uint32_t)__cvta_generic_to_shared(Ah), sAl = (uint32_t)__cvta_generic_to_shared(Al);
    const uint32_t sBh = (uint32_t)__cvta_generic_to_shared(Bh), sBl = (uint32_t)__cvta_generic_to_shared(Bl);

    if (niter > 0) load(0);
    for (int it = 0; it < niter; ++it) {
        {
            float2 af[8];
            const float4 *ap4 = reinterpret_cast<const float4 *>(aff + r_sample * kWtAff + a8 * 10);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 v4 = ap4[q];
                af[2 * q] = make_float2(v4.x, v4.y);
                af[2 * q + 1] = make_float2(v4.z, v4.w);
            }
#pragma unroll
            for (int sl = 0; sl < SL; ++sl) {
                float v[8];
                const uint4 hh = rh[sl], ll = rl[sl];
                if (planes) {
                    v[0] = bf16_lo(hh.x) + bf16_lo(ll.x); v[1] = bf16_hi(hh.x) + bf16_hi(ll.x);
                    v[2] = bf16_lo(hh.y) + bf16_lo(ll.y); v[3] = bf16_hi(hh.y) + bf16_hi(ll.y);
                    v[4] = bf16_lo(hh.z) + bf16_lo(ll.z); v[5] = bf16_hi(hh.z) + bf16_hi(ll.z);
                    v[6] = bf16_lo(hh.w) + bf16_lo(ll.w); v[7] = bf16_hi(hh.w) + bf16_hi(ll.w);
                } else {
                    v[0] = __uint_as_float(hh.x); v[1] = __uint_as_float(hh.y); v[2] = __uint_as_float(hh.z); v[3] = __uint_as_float(hh.w);
                    v[4] = __uint_as_float(ll.x); v[5] = __uint_as_float(ll.y); v[6] = __uint_as_float(ll.z); v[7] = __uint_as_float(ll.w);
                }
                const bool ok = (rvalid >> sl) & 1u;
#pragma unroll
                for (int q = 0; q < 8; ++q) v[q] = ok ? fmaf(v[q], af[q].x, af[q].y) : 0.f;
                uint2 h0, l0, h1, l1;
                split4(v, h0, l0);
                split4(v + 4, h1, l1);
                const int row = ap + 32 * sl;
                *reinterpret_cast<uint4 *>(Ah + row * kWmAP + a8 * 8) = make_uint4(h0.x, h0.y, h1.x, h1.y);
                *reinterpret_cast<uint4 *>(Al + row * kWmAP + a8 * 8) = make_uint4(l0.x, l0.y, l1.x, l1.y);
#pragma unroll
                for (int j = 0; j < NB4; ++j) {
                    const float w[4] = {rb[sl][j].x, rb[sl][j].y, rb[sl][j].z, rb[sl][j].w};
                    uint2 hi, lo;
                    split4(w, hi, lo);
                    *reinterpret_cast<uint2 *>(Bh + row * BP + b4 * 4 + 32 * j) = hi;
                    *reinterpret_cast<uint2 *>(Bl + row * BP + b4 * 4 + 32 * j) = lo;
                }
            }
        }
        __syncthreads();
        if (it + 1 < niter) load(it + 1);
#pragma unroll
        for (int s2 = 0; s2 < SL; ++s2) {
            const uint32_t ka = (uint32_t)(s2 * 32 * kWmAP * 2), kb = (uint32_t)(s2 * 32 * BP * 2);
            uint32_t ah[4], al[4];
            ldsm_x4_trans(sAh + a_off + ka, ah[0], ah[1], ah[2], ah[3]);
            ldsm_x4_trans(sAl + a_off + ka, al[0], al[1], al[2], al[3]);
            uint32_t bh[NT / 2][4], bl[NT / 2][4];
#pragma unroll
            for (int n2 = 0; n2 < NT / 2; ++n2) {
                ldsm_x4_trans(sBh + b_off + kb + n2 * 32, bh[n2][0], bh[n2][1], bh[n2][2], bh[n2][3]);
                ldsm_x4_trans(sBl + b_off + kb + n2 * 32, bl[n2][0], bl[n2][1], bl[n2][2], bl[n2][3]);
            }
#pragma unroll
            for (int n2 = 0; n2 < NT / 2; ++n2) {  // independent accumulators between dependent MMAs
                mma_bf16(acc[2 * n2], ah, bh[n2][0], bh[n2][1]);
                mma_bf16(acc[2 * n2 + 1], ah, bh[n2][2], bh[n2][3]);
            }
#pragma unroll
            for (int n2 = 0; n2 < NT / 2; ++n2) {
                mma_bf16(acc[2 * n2], ah, bl[n2][0], bl[n2][1]);
                mma_bf16(acc[2 * n2 + 1], ah, bl[n2][2], bl[n2][3]);
            }
#pragma unroll
            for (int n2 = 0; n2 < NT / 2; ++n2) {
                mma_bf16(acc[2 * n2], al, bh[n2][0], bh[n2][1]);
                mma_bf16(acc[2 * n2 + 1], al, bh[n2][2], bh[n2][3]);
            }
        }
        __syncthreads();
    }

    const int taps = a.KT * a.KF;
    const int g = lane >> 2, t4 = lane & 3;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int ci = ci0 + wm * 16 + g + (q >> 1) * 8;
            const int co = co0 + nt * 8 + t4 * 2 + (q & 1);
            if (ci >= a.cin || co >= a.cout_real) continue;
            const size_t idx = a.transposed ? ((size_t)ci * a.cout_real + co) * taps + tap : ((size_t)co * a.cin + ci) * taps + tap;
            atomicAdd(a.dw + idx, acc[nt][q]);
        }
}

// ---- weight gradient of the stride-1 3x3 convs: all nine taps from one staged halo tile ----------------------------
// The per-tap kernel above re-reads (and re-normalises, re-splits) every input pixel once per tap and dy once per
// (tap, cin tile).  For the stride-1 convs (the DenseBlocks: 90 % of the FLOPs) a chunk of <= 32 output pixels -- a
// 32-bin segment of one frame, or R whole frames when the layer has fewer than 32 bins -- needs the input halo tile
// [(R + 2) frames x (W + 2) bins], and the im2col row of output pixel k for tap (kt, kf) is halo row
// base_k + kt (W + 2) + kf: the nine A operands are the same staged tile at nine row offsets (ldmatrix takes a row
// address per lane).  Twelve warps: warp w owns input channels 16 (w & 3) .. + 15 for taps 3 (w >> 2) .. + 2; dy is staged
// once per chunk.
constexpr int kWtHP = 104;  // halo rows held in shared memory (>= every base_k + 2 (W + 2) + 2, see the launcher)
constexpr int kWtThreads = 384;  // 12 warps: 4 input-channel groups x 3 tap groups
constexpr int kWtTaps = 3;       // taps per warp
constexpr int kWtSlots = 3;      // halo items (pixel, 8-channel group) per thread: 102 * 8 / 384 rounded up

struct WtapsGeom {
    int W, R, nseg, nrc;  // bins per segment, frames per chunk, segments per frame, chunk rows per sample
};

__global__ void __launch_bounds__(kWtThreads, 1) wgrad_taps_kernel(const WgradArgs a, const WtapsGeom g, int splits) {
    constexpr int BN = 32, BP = BN + 8;
    extern __shared__ __align__(16) unsigned char wsm_raw[];
    __nv_bfloat16 *Ah = reinterpret_cast<__nv_bfloat16 *>(wsm_raw);  // [HP][AP]
    __nv_bfloat16 *Al = Ah + kWtHP * kWmAP;
    __nv_bfloat16 *Bh = Al + kWtHP * kWmAP;  // [32][BP]
    __nv_bfloat16 *Bl = Bh + 32 * BP;
    float2 *aff = reinterpret_cast<float2 *>(Bl + 32 * BP);  // [B][BM]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int co_tiles = (a.cout + BN - 1) / BN;
    const int cot = blockIdx.x % co_tiles, cit = blockIdx.x / co_tiles;
    const int ci0 = cit * kWgBM, co0 = cot * BN;
    const int npix = a.T * a.Fout;
    const int W2 = g.W + 2;
    const int nhalo = (g.R + 2) * W2;
    const int per_sample = g.nrc * g.nseg;
    const int total = a.B * per_sample;
    const int per = (total + splits - 1) / splits;
    const int c_lo = blockIdx.y * per, c_hi = min(total, c_lo + per);

    for (int i = tid; i < kWtHP * kWmAP / 2; i += kWtThreads) {
        reinterpret_cast<uint32_t *>(Ah)[i] = 0u;
        reinterpret_cast<uint32_t *>(Al)[i] = 0u;
    }
    for (int i = tid; i < a.B * kWgBM; i += kWtThreads) {
        const int b = i / kWgBM, cl = i - b * kWgBM, c = ci0 + cl;
        float2 v = make_float2(1.f, 0.f);
        if (a.x_sums && c < a.cin) {
            const double *s = a.x_sums + ((size_t)b * a.x_ctot + a.x_coff + c) * 2;
            v = affine_from_sums(stat_get(s), stat_get(s + 1), a.inv_n, (double)a.eps);
        }
        aff[b * kWtAff + (cl >> 3) * 10 + (cl & 7)] = v;  // 80-byte rows per 8-channel group: conflict-free 16-byte reads
    }
    __syncthreads();

    // loader constants: halo items of this thread, and its dy pixel
    int h_row[kWtSlots], h_col[kWtSlots];
#pragma unroll
    for (int sl = 0; sl < kWtSlots; ++sl) {
        const int h = (tid + kWtThreads * sl) >> 3;
        h_row[sl] = h < nhalo ? h / W2 : -1;
        h_col[sl] = h - (h / W2) * W2;
    }
    const int g8 = tid & 7;
    const int bp = tid >> 3, b4 = tid & 7;
    const int b_r = bp / g.W, b_j = bp - b_r * g.W;
    const size_t x_lo = (size_t)a.x_ctot * a.T * a.Fin;
    const int cg = ci0 + g8 * 8;
    const bool c_ok = cg < a.cin;

    // The loader only ISSUES the global loads (raw plane pixels stay in registers across the MMA phase); normalising and
    // splitting happen when the tile is stored, one iteration later -- an arithmetic use right behind a load would stall
    // the warp for the full memory latency before it reaches its MMAs.
    uint4 rh[kWtSlots], rl[kWtSlots];
    unsigned rvalid = 0;
    int rb_sample = 0;
    float4 rb;
    auto load = [&](int chunk) {
        const int b = chunk / per_sample;
        const int rem = chunk - b * per_sample;
        const int rc = rem / g.nseg, seg = rem - rc * g.nseg;
        const int t0 = rc * g.R, f0 = seg * g.W;
        const __nv_bfloat16 *xb = reinterpret_cast<const __nv_bfloat16 *>(a.x) + (size_t)b * 2 * x_lo + (size_t)((a.x_coff + cg) >> 3) * a.T * a.Fin * 8;
        rvalid = 0;
        rb_sample = b;
#pragma unroll
        for (int sl = 0; sl < kWtSlots; ++sl) {
            const int ti = t0 - 1 + h_row[sl], fi = f0 - a.pad_f + h_col[sl];
            const bool ok = h_row[sl] >= 0 && c_ok && ti >= 0 && ti < a.T && fi >= 0 && fi < a.Fin;
            const __nv_bfloat16 *xp = xb + ((size_t)(ok ? ti : 0) * a.Fin + (ok ? fi : 0)) * 8;
            rh[sl] = make_uint4(0u, 0u, 0u, 0u);
            rl[sl] = rh[sl];
            if (ok) {
                rh[sl] = *reinterpret_cast<const uint4 *>(xp);
                if (a.use_lo) rl[sl] = *reinterpret_cast<const uint4 *>(xp + x_lo);
                rvalid |= 1u << sl;
            }
        }
        rb = make_float4(0.f, 0.f, 0.f, 0.f);
        const int t = t0 + b_r, f = f0 + b_j, co = co0 + b4 * 4;
        if (tid < 256 && b_r < g.R && t < a.T && f < a.Fout && co < a.cout)
            rb = *reinterpret_cast<const float4 *>(a.dy + ((size_t)b * npix + (size_t)t * a.Fout + f) * a.dy_ctot + a.dy_coff + co);
    };

    float acc[kWtTaps][4][4];
#pragma unroll
    for (int i = 0; i < kWtTaps; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f;
    const int wm = warp & 3, wt = warp >> 2;
    const int lr = lane & 7, lm = lane >> 3;
    // per-lane halo base row of the im2col rows this lane addresses (k = kh * 16 + (lm >> 1) * 8 + lr)
    uint32_t a_base[2];
#pragma unroll
    for (int kh = 0; kh < 2; ++kh) {
        const int k = kh * 16 + (lm >> 1) * 8 + lr;
        const int r = k / g.W, j = k - r * g.W;
        a_base[kh] = (uint32_t)(((r * W2 + j) * kWmAP + wm * 16 + (lm & 1) * 8) * 2);
    }
    const uint32_t b_off = (uint32_t)((((lm & 1) * 8 + lr) * BP + (lm >> 1) * 8) * 2);
    const uint32_t sAh = (uint32_t)__cvta_generic_to_shared(Ah), sAl = (uint32_t)__cvta_generic_to_shared(Al);
    const uint32_t sBh = (uint32_t)__cvta_generic_to_shared(Bh), sBl = (uint32_t)__cvta_generic_to_shared(Bl);

    if (c_lo < c_hi) load(c_lo);
    for (int chunk = c_lo; chunk < c_hi; ++chunk) {
        {
            float2 af[8];
            {
                const float4 *ap4 = reinterpret_cast<const float4 *>(aff + rb_sample * kWtAff + g8 * 10);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 v4 = ap4[q];
                    af[2 * q] = make_float2(v4.x, v4.y);
                    af[2 * q + 1] = make_float2(v4.z, v4.w);
                }
            }
#pragma unroll
            for (int sl = 0; sl < kWtSlots; ++sl) {
                if (h_row[sl] < 0) continue;
                const int h = (tid + kWtThreads * sl) >> 3;
                float v[8];
                const uint4 hh = rh[sl], ll = rl[sl];
                v[0] = bf16_lo(hh.x) + bf16_lo(ll.x); v[1] = bf16_hi(hh.x) + bf16_hi(ll.x);
                v[2] = bf16_lo(hh.y) + bf16_lo(ll.y); v[3] = bf16_hi(hh.y) + bf16_hi(ll.y);
                v[4] = bf16_lo(hh.z) + bf16_lo(ll.z); v[5] = bf16_hi(hh.z) + bf16_hi(ll.z);
                v[6] = bf16_lo(hh.w) + bf16_lo(ll.w); v[7] = bf16_hi(hh.w) + bf16_hi(ll.w);
                const bool ok = (rvalid >> sl) & 1u;
#pragma unroll
                for (int q = 0; q < 8; ++q) v[q] = ok ? fmaf(v[q], af[q].x, af[q].y) : 0.f;
                uint2 h0, l0, h1, l1;
                split4(v, h0, l0);
                split4(v + 4, h1, l1);
                *reinterpret_cast<uint4 *>(Ah + h * kWmAP + g8 * 8) = make_uint4(h0.x, h0.y, h1.x, h1.y);
                *reinterpret_cast<uint4 *>(Al + h * kWmAP + g8 * 8) = make_uint4(l0.x, l0.y, l1.x, l1.y);
            }
        }
        if (tid < 256) {
            const float v[4] = {rb.x, rb.y, rb.z, rb.w};
            uint2 hi, lo;
            split4(v, hi, lo);
            *reinterpret_cast<uint2 *>(Bh + bp * BP + b4 * 4) = hi;
            *reinterpret_cast<uint2 *>(Bl + bp * BP + b4 * 4) = lo;
        }
        __syncthreads();
        if (chunk + 1 < c_hi) load(chunk + 1);
#pragma unroll
        for (int kh = 0; kh < 2; ++kh) {
            uint32_t bh[2][4], bl[2][4];
#pragma unroll
            for (int n2 = 0; n2 < 2; ++n2) {
                ldsm_x4_trans(sBh + b_off + kh * 16 * BP * 2 + n2 * 32, bh[n2][0], bh[n2][1], bh[n2][2], bh[n2][3]);
                ldsm_x4_trans(sBl + b_off + kh * 16 * BP * 2 + n2 * 32, bl[n2][0], bl[n2][1], bl[n2][2], bl[n2][3]);
            }
#pragma unroll
            for (int ti = 0; ti < kWtTaps; ++ti) {
                const int tap = wt * kWtTaps + ti;
                {
                    const int kt = tap / 3, kf = tap - kt * 3;
                    const uint32_t toff = (uint32_t)((kt * W2 + kf) * kWmAP * 2);
                    uint32_t ah[4], al[4];
                    ldsm_x4_trans(sAh + a_base[kh] + toff, ah[0], ah[1], ah[2], ah[3]);
                    ldsm_x4_trans(sAl + a_base[kh] + toff, al[0], al[1], al[2], al[3]);
#pragma unroll
                    for (int n2 = 0; n2 < 2; ++n2) {  // four independent accumulators between dependent MMAs
                        mma_bf16(acc[ti][2 * n2], ah, bh[n2][0], bh[n2][1]);
                        mma_bf16(acc[ti][2 * n2 + 1], ah, bh[n2][2], bh[n2][3]);
                    }
#pragma unroll
                    for (int n2 = 0; n2 < 2; ++n2) {
                        mma_bf16(acc[ti][2 * n2], ah, bl[n2][0], bl[n2][1]);
                        mma_bf16(acc[ti][2 * n2 + 1], ah, bl[n2][2], bl[n2][3]);
                    }
#pragma unroll
                    for (int n2 = 0; n2 < 2; ++n2) {
                        mma_bf16(acc[ti][2 * n2], al, bh[n2][0], bh[n2][1]);
                        mma_bf16(acc[ti][2 * n2 + 1], al, bh[n2][2], bh[n2][3]);
                    }
                }
            }
        }
        __syncthreads();
    }

    const int gq = lane >> 2, t4 = lane & 3;
#pragma unroll
    for (int ti = 0; ti < kWtTaps; ++ti) {
        const int tap = wt * kWtTaps + ti;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int ci = ci0 + wm * 16 + gq + (q >> 1) * 8;
                const int co = co0 + nt * 8 + t4 * 2 + (q & 1);
                if (ci >= a.cin || co >= a.cout_real) continue;
                atomicAdd(a.dw + ((size_t)co * a.cin + ci) * 9 + tap, acc[ti][nt][q]);
            }
    }
}

__global__ void dgrad_pack_kernel(const float *__restrict__ src, float *__restrict__ dst, int taps, int cin, int cout,
                                  int cout_pad, int cin_pad, int flip) {
    const int64_t total = (int64_t)taps * cout * cin_pad;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int ci = (int)(i % cin_pad);
        const int64_t r = i / cin_pad;
        const int co = (int)(r % cout);
        const int tap = (int)(r / cout);
        dst[i] = ci < cin ? src[((size_t)(flip ? taps - 1 - tap : tap) * cin + ci) * cout_pad + co] : 0.f;
    }
}

// ---- TCN ---------------------------------------------------------------------------------------------------------
struct GlnStat {
    float mean, rstd;
};
__device__ __forceinline__ GlnStat gln_stat(const double *s, double inv_n, float eps) {
    const double mean = stat_get(s) * inv_n;
    double var = stat_get(s + 1) * inv_n - mean * mean;
    if (var < 0.0) var = 0.0;
    GlnStat g;
    g.mean = (float)mean;
    g.rstd = (float)rsqrt(var + (double)eps);
    return g;
}

constexpr int kTcnBwFrames = 16;
// ELU / its derivative through ex2.approx (absolute error ~1e-7): the recomputed activations feed bf16 hi/lo GEMM operands
// and gradients held to 1e-3, and expm1f / expf cost ~10x the instructions
__device__ __forceinline__ float elu_fast_bw(float x) { return x > 0.f ? x : __expf(x) - 1.f; }

// y = dwconv(ELU(IN1d(u))) (pre-PReLU) and q = gLN(PReLU(y)) (the pointwise conv's input), both fp32 [B][T][C]
__global__ void __launch_bounds__(256) tcn_recompute_kernel(const TcnBwdArgs a, float *__restrict__ Y, float *__restrict__ Q) {
    extern __shared__ float sh[];  // scale, shift, 3 taps, gamma, beta : 7 C
    const int C = a.C, T = a.T;
    float *sc = sh, *sf = sh + C, *wt = sh + 2 * C, *gm = sh + 5 * C, *bt = sh + 6 * C;
    const int b = blockIdx.y;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const double *us = a.u_sums + ((size_t)b * C + c) * 2;
        const float2 af = affine_from_sums(stat_get(us), stat_get(us + 1), a.inv_T, (double)a.in_eps);
        sc[c] = af.x;
        sf[c] = af.y;
        wt[c] = a.wdw[c * 3];
        wt[C + c] = a.wdw[c * 3 + 1];
        wt[2 * C + c] = a.wdw[c * 3 + 2];
        gm[c] = a.gamma[c];
        bt[c] = a.beta[c];
    }
    __syncthreads();
    const GlnStat gs = gln_stat(a.g_sums + (size_t)b * 2, a.gln_inv_n, a.gln_eps);
    const float al = a.alpha[0];
    const float *ub = a.u + (size_t)b * T * C;
    const int t0 = blockIdx.x * kTcnBwFrames;
    const float *__restrict__ ubr = ub;
#pragma unroll 4
    for (int i = threadIdx.x; i < kTcnBwFrames * C; i += blockDim.x) {
        const int tt = i / C, c = i - tt * C;
        const int t = t0 + tt;
        if (t >= T) break;
        auto act = [&](int tq) {
            if (tq < 0 || tq >= T) return 0.f;
            return elu_fast_bw(fmaf(__ldg(ubr + (size_t)tq * C + c), sc[c], sf[c]));
        };
        const float y = fmaf(wt[c], act(t - a.dil), fmaf(wt[C + c], act(t), wt[2 * C + c] * act(t + a.dil)));
        const float p = y > 0.f ? y : al * y;
        const size_t o = ((size_t)b * T + t) * C + c;
        Y[o] = y;
        Q[o] = fmaf(gm[c] * gs.rstd, p - gs.mean, bt[c]);
    }
}

constexpr int kTcnRows = 8;   // frames per CTA of the channel-parallel kernels
constexpr int kTcnCh = 8;     // channels per thread (128 threads): C <= 1024

// gLN backward, pass 1: per-sample sums of g = gamma * dq and g * phat; per-channel gamma / beta gradients
__global__ void __launch_bounds__(128) gln_bwd_reduce_kernel(const TcnBwdArgs a, const float *__restrict__ DQ,
                                                             const float *__restrict__ Y, double *__restrict__ gred,
                                                             float *__restrict__ dgamma, float *__restrict__ dbeta) {
    __shared__ double red[2][4];
    const int C = a.C, T = a.T, b = blockIdx.y;
    const GlnStat gs = gln_stat(a.g_sums + (size_t)b * 2, a.gln_inv_n, a.gln_eps);
    const float al = a.alpha[0];
    const int t0 = blockIdx.x * kTcnRows, t1 = min(T, t0 + kTcnRows);
    float dg[kTcnCh], db[kTcnCh], gam[kTcnCh];
#pragma unroll
    for (int k = 0; k < kTcnCh; ++k) {
        dg[k] = db[k] = 0.f;
        const int c = threadIdx.x + k * 128;
        gam[k] = c < C ? a.gamma[c] : 0.f;
    }
    float a1 = 0.f, a2 = 0.f;
#pragma unroll
    for (int ti = 0; ti < kTcnRows; ++ti) {  // constant trip count: the loads of all frames are issued up front
        const int t = t0 + ti;
        if (t >= t1) break;
        const size_t row = ((size_t)b * T + t) * C;
#pragma unroll
        for (int k = 0; k < kTcnCh; ++k) {
            const int c = threadIdx.x + k * 128;
            if (c < C) {
                const float dq = __ldg(DQ + row + c), y = __ldg(Y + row + c);
                const float p = y > 0.f ? y : al * y;
                const float ph = (p - gs.mean) * gs.rstd;
                const float g = gam[k] * dq;
                a1 += g;
                a2 = fmaf(g, ph, a2);
                dg[k] = fmaf(dq, ph, dg[k]);
                db[k] += dq;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < kTcnCh; ++k) {
        const int c = threadIdx.x + k * 128;
        if (c < C) {
            atomicAdd(dgamma + c, dg[k]);
            atomicAdd(dbeta + c, db[k]);
        }
    }
    double d1 = warp_sum((double)a1), d2 = warp_sum((double)a2);
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = d1;
        red[1][threadIdx.x >> 5] = d2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        atomicAdd(gred + (size_t)b * 2, red[0][0] + red[0][1] + red[0][2] + red[0][3]);
        atomicAdd(gred + (size_t)b * 2 + 1, red[1][0] + red[1][1] + red[1][2] + red[1][3]);
    }
}

// gLN backward, pass 2, fused with the PReLU backward: DQ becomes dL/dy in place; alpha gradient
__global__ void __launch_bounds__(128) gln_bwd_apply_kernel(const TcnBwdArgs a, float *__restrict__ DQ, const float *__restrict__ Y,
                                                            const double *__restrict__ gred, float *__restrict__ dalpha) {
    __shared__ float red[4];
    const int C = a.C, T = a.T, b = blockIdx.y;
    const GlnStat gs = gln_stat(a.g_sums + (size_t)b * 2, a.gln_inv_n, a.gln_eps);
    const float al = a.alpha[0];
    const float m1 = (float)(gred[(size_t)b * 2] * a.gln_inv_n), m2 = (float)(gred[(size_t)b * 2 + 1] * a.gln_inv_n);
    const int t0 = blockIdx.x * kTcnRows, t1 = min(T, t0 + kTcnRows);
    float da = 0.f;
    for (int c = threadIdx.x; c < C; c += 128) {
        // all frames of this channel: loads first (DQ is rewritten in place, which would otherwise order every load behind the
        // previous frame's store)
        float dqv[kTcnRows], yv[kTcnRows];
#pragma unroll
        for (int ti = 0; ti < kTcnRows; ++ti) {
            const int t = t0 + ti;
            const size_t row = ((size_t)b * T + min(t, T - 1)) * C;
            dqv[ti] = DQ[row + c];
            yv[ti] = __ldg(Y + row + c);
        }
#pragma unroll
        for (int ti = 0; ti < kTcnRows; ++ti) {
            const int t = t0 + ti;
            if (t >= t1) break;
            const size_t row = ((size_t)b * T + t) * C;
            const float dq = dqv[ti], y = yv[ti];
            const float p = y > 0.f ? y : al * y;
            const float ph = (p - gs.mean) * gs.rstd;
            const float dp = gs.rstd * (a.gamma[c] * dq - m1 - ph * m2);
            if (y > 0.f) {
                DQ[row + c] = dp;
            } else {
                DQ[row + c] = dp * al;
                da = fmaf(dp, y, da);
            }
        }
    }
    da = warp_sum(da);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = da;
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(dalpha, red[0] + red[1] + red[2] + red[3]);
}

// depthwise conv + ELU backward: DN = dL/dn (n = IN1d(u)), its per-(b,c) sums for the InstanceNorm1d backward, and
// the depthwise weight gradient
// 32 frames per CTA: every CTA ends with five atomics per channel onto [b][c] / [c] accumulators, and with 8 frames the 504
// CTAs of a B = 8, T = 500 step serialised on them (56 us per launch; the arithmetic is ~10 us).  With 128 threads walking
// three channels each, one frame at a time, the 32-frame version was a chain of ~100 dependent L2 round trips (103 us under
// ncu): one channel per thread and the loads of four frames in flight.
constexpr int kTcnRowsDw = 32;
constexpr int kDwBwdThreads = 384;  // one channel per thread at the PAPER width: the frame loop is a chain of L2 round trips,
                                    // so the parallelism has to come from the channels and from four frames' loads in flight
__global__ void __launch_bounds__(kDwBwdThreads) dw_bwd_kernel(const TcnBwdArgs a, const float *__restrict__ DY, float *__restrict__ DN,
                                                               double *__restrict__ ired, float *__restrict__ dwdw) {
    const int C = a.C, T = a.T, b = blockIdx.y, d = a.dil;
    const int t0 = blockIdx.x * kTcnRowsDw, t1 = min(T, t0 + kTcnRowsDw);
    const float *__restrict__ ub = a.u + (size_t)b * T * C;
    const float *__restrict__ dyb = DY + (size_t)b * T * C;
    for (int c = threadIdx.x; c < C; c += kDwBwdThreads) {
        const double *us = a.u_sums + ((size_t)b * C + c) * 2;
        const float2 af = affine_from_sums(stat_get(us), stat_get(us + 1), a.inv_T, (double)a.in_eps);
        const float w0 = a.wdw[c * 3], w1 = a.wdw[c * 3 + 1], w2 = a.wdw[c * 3 + 2];
        float s1 = 0.f, s2 = 0.f, g0 = 0.f, g1 = 0.f, g2 = 0.f;
        auto uat = [&](int tq) { return (tq >= 0 && tq < T) ? __ldg(ub + (size_t)tq * C + c) : 0.f; };
        auto dyat = [&](int tq) { return (tq >= 0 && tq < T) ? __ldg(dyb + (size_t)tq * C + c) : 0.f; };
        for (int tb = t0; tb < t1; tb += 4) {
            // forward: y[t] = w0 v[t-d] + w1 v[t] + w2 v[t+d]; four frames at a time: all 24 loads first
            float dy0[4], dym[4], dyp[4], u0[4], um[4], up[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int t = tb + i;
                const bool in = t < t1;
                dy0[i] = in ? dyat(t) : 0.f;
                dym[i] = in ? dyat(t - d) : 0.f;
                dyp[i] = in ? dyat(t + d) : 0.f;
                u0[i] = in ? uat(t) : 0.f;
                um[i] = in ? uat(t - d) : 0.f;
                up[i] = in ? uat(t + d) : 0.f;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int t = tb + i;
                if (t >= t1) break;
                const float dv = fmaf(w0, dyp[i], fmaf(w1, dy0[i], w2 * dym[i]));
                const float n = fmaf(u0[i], af.x, af.y);
                const float en = __expf(n);
                const float dn = n > 0.f ? dv : dv * en;
                DN[((size_t)b * T + t) * C + c] = dn;
                s1 += dn;
                s2 = fmaf(dn, n, s2);
                const float vm = (t - d >= 0) ? elu_fast_bw(fmaf(um[i], af.x, af.y)) : 0.f;
                const float vp = (t + d < T) ? elu_fast_bw(fmaf(up[i], af.x, af.y)) : 0.f;
                g0 = fmaf(dy0[i], vm, g0);
                g1 = fmaf(dy0[i], n > 0.f ? n : en - 1.f, g1);
                g2 = fmaf(dy0[i], vp, g2);
            }
        }
        atomicAdd(ired + ((size_t)b * C + c) * 2, (double)s1);
        atomicAdd(ired + ((size_t)b * C + c) * 2 + 1, (double)s2);
        atomicAdd(dwdw + c * 3, g0);
        atomicAdd(dwdw + c * 3 + 1, g1);
        atomicAdd(dwdw + c * 3 + 2, g2);
    }
}

// InstanceNorm1d backward: du = rstd * (dn - mean(dn) - n * mean(dn * n))
__global__ void __launch_bounds__(128) in1d_bwd_apply_kernel(const TcnBwdArgs a, const float *__restrict__ DN,
                                                             const double *__restrict__ ired, float *__restrict__ out,
                                                             int accumulate) {
    const int C = a.C, T = a.T, b = blockIdx.y;
    const int t0 = blockIdx.x * kTcnRows, t1 = min(T, t0 + kTcnRows);
    for (int c = threadIdx.x; c < C; c += 128) {
        const double *us = a.u_sums + ((size_t)b * C + c) * 2;
        const float2 af = affine_from_sums(stat_get(us), stat_get(us + 1), a.inv_T, (double)a.in_eps);
        const float m1 = (float)(ired[((size_t)b * C + c) * 2] * a.inv_T), m2 = (float)(ired[((size_t)b * C + c) * 2 + 1] * a.inv_T);
        float uv[kTcnRows], dnv[kTcnRows], ov[kTcnRows];  // loads of all frames first: the stores to `out` order the rest
#pragma unroll
        for (int ti = 0; ti < kTcnRows; ++ti) {
            const size_t o = ((size_t)b * T + min(t0 + ti, T - 1)) * C + c;
            uv[ti] = __ldg(a.u + o);
            dnv[ti] = __ldg(DN + o);
            ov[ti] = accumulate ? out[o] : 0.f;
        }
#pragma unroll
        for (int ti = 0; ti < kTcnRows; ++ti) {
            const int t = t0 + ti;
            if (t >= t1) break;
            const size_t o = ((size_t)b * T + t) * C + c;
            const float n = fmaf(uv[ti], af.x, af.y);
            out[o] = ov[ti] + af.x * (dnv[ti] - m1 - n * m2);
        }
    }
}

__global__ void cl_to_planes_kernel(const float *__restrict__ src, __nv_bfloat16 *__restrict__ dst, int B, int npix, int C) {
    const int64_t total = (int64_t)B * npix * (C >> 2);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % (C >> 2));
        const int64_t r = i / (C >> 2);
        const int p = (int)(r % npix), b = (int)(r / npix);
        const float4 v = *reinterpret_cast<const float4 *>(src + r * C + c4 * 4);
        const float f[4] = {v.x, v.y, v.z, v.w};
        uint2 hi, lo;
        split4(f, hi, lo);
        const int c = c4 * 4;
        __nv_bfloat16 *d = dst + (size_t)b * 2 * C * npix + ((size_t)(c >> 3) * npix + p) * 8 + (c & 7);
        *reinterpret_cast<uint2 *>(d) = hi;
        *reinterpret_cast<uint2 *>(d + (size_t)C * npix) = lo;
    }
}

__global__ void planes_accumulate_kernel(const __nv_bfloat16 *__restrict__ src, int sctot, float *__restrict__ dst, int dctot,
                                         int dcoff, int C, int B, int npix) {
    const int nc4 = C >> 2;
    const int64_t total = (int64_t)B * npix * nc4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % nc4);
        const int64_t r = i / nc4;
        const int p = (int)(r % npix), b = (int)(r / npix);
        const int c = c4 * 4;
        float e[4];
        load_e4(src + (size_t)b * 2 * sctot * npix + ((size_t)(c >> 3) * npix + p) * 8 + (c & 7), (size_t)sctot * npix, 1, e);
        float4 *d = reinterpret_cast<float4 *>(dst + r * dctot + dcoff + c);
        float4 v = *d;
        v.x += e[0];
        v.y += e[1];
        v.z += e[2];
        v.w += e[3];
        *d = v;
    }
}

__global__ void copy_channels_kernel(const float *__restrict__ src, int sctot, int scoff, float *__restrict__ dst, int dctot,
                                     int dcoff, int C, int64_t rows, int accumulate) {
    const int64_t total = rows * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / C;
        const int c = (int)(i - r * C);
        const float v = src[r * sctot + scoff + c];
        float *d = dst + r * dctot + dcoff + c;
        *d = accumulate ? *d + v : v;
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------------ launchers ----
int launch_in_bwd(const InBwdArgs &a, cudaStream_t st) {
    MISO_REQUIRE(a.c % 4 == 0 && a.coff % 4 == 0 && a.ctot % 4 == 0 && a.c >= 4 && a.c <= 4 * kEw,
                 "in_bwd: channel range (c=%d coff=%d ctot=%d) must be multiples of 4, c <= %d", a.c, a.coff, a.ctot, 4 * kEw);
    const int lanes = kEw / (a.c / 4);
    // enough CTAs to fill the machine, at least 8 pixels per lane
    int chunks = ceil_div(a.npix, lanes * 8);
    const int cap = std::max(1, ceil_div(148 * 8, a.B));
    if (chunks > cap) chunks = cap;
    const int per = ceil_div(a.npix, chunks);
    chunks = ceil_div(a.npix, per);
    dim3 grid(chunks, a.B);
    if (!a.plain) {
        MISO_CUDA(cudaMemsetAsync(a.red, 0, (size_t)a.B * a.c * 2 * sizeof(double), st));
        in_bwd_reduce_kernel<<<grid, kEw, 2 * a.c * sizeof(float), st>>>(a, per);
        MISO_LAUNCHED("in_bwd_reduce_kernel");
    }
    in_bwd_apply_kernel<<<grid, kEw, a.c * sizeof(float), st>>>(a, per);
    MISO_LAUNCHED("in_bwd_apply_kernel");
    return MISO_OK;
}

template <int TN>
static int launch_wgrad_t(const WgradArgs &a, cudaStream_t st) {
    constexpr int BN = 16 * TN;
    const int taps = a.KT * a.KF;
    const int ntile = taps * ceil_div(a.cin, kWgBM) * ceil_div(a.cout, BN);
    const int npix = a.T * a.Fout;
    int splits = ceil_div(4 * 148, ntile);
    splits = std::max(1, std::min(splits, ceil_div(npix, 4 * kWgBK)));
    const size_t smem = (size_t)(kWgBK * (kWgBM + 4) + kWgBK * (BN + 4)) * sizeof(float) + (size_t)a.B * kWgBM * sizeof(float2);
    MISO_REQUIRE(smem <= 200 * 1024, "wgrad: batch %d too large for the per-sample affine table", a.B);
    static size_t attr_set[2] = {0, 0};
    size_t &cur = attr_set[TN == 4 ? 1 : 0];
    if (smem > 48 * 1024 && smem > cur) {
        MISO_CUDA(cudaFuncSetAttribute(wgrad_kernel<TN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cur = smem;
    }
    wgrad_kernel<TN><<<dim3(splits, ntile), 256, smem, st>>>(a, splits);
    MISO_LAUNCHED("wgrad_kernel");
    return MISO_OK;
}

template <int BN, int SL>
static int launch_wgrad_mma(const WgradArgs &a, cudaStream_t st) {
    constexpr int BK = 32 * SL;
    const int taps = a.KT * a.KF;
    const int ntile = taps * ceil_div(a.cin, kWgBM) * ceil_div(a.cout, BN);
    const int npix = a.T * a.Fout;
    int splits = ceil_div(4 * 148, ntile);
    splits = std::max(1, std::min(splits, ceil_div(npix, 2 * BK)));
    const size_t smem = (size_t)(2 * BK * kWmAP + 2 * BK * (BN + 8)) * 2 + (size_t)a.B * kWtAff * sizeof(float2);
    MISO_REQUIRE(smem <= 200 * 1024, "wgrad: batch %d too large for the per-sample affine table", a.B);
    static size_t cur = 0;
    if (smem > 48 * 1024 && smem > cur) {
        MISO_CUDA(cudaFuncSetAttribute(wgrad_mma_kernel<BN, SL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cur = smem;
    }
    wgrad_mma_kernel<BN, SL><<<dim3(ntile, splits), 256, smem, st>>>(a, splits);
    MISO_LAUNCHED("wgrad_mma_kernel");
    return MISO_OK;
}

static bool wgrad_taps_ok(const WgradArgs &a) {
    return !a.transposed && a.stride_f == 1 && a.KT == 3 && a.KF == 3 && a.pad_t == 1 && (a.pad_f == 0 || a.pad_f == 1) &&
           a.x_layout == LAYOUT_PLANES && a.Fin == a.Fout + 2 - 2 * a.pad_f;
}

static int launch_wgrad_taps(const WgradArgs &a, cudaStream_t st) {
    WtapsGeom g;
    g.W = std::min(a.Fout, 32);
    g.R = std::max(1, 32 / a.Fout);
    g.nseg = ceil_div(a.Fout, g.W);
    g.nrc = ceil_div(a.T, g.R);
    // every im2col row stays inside the zero-initialised halo buffer: base_31 + 2 (W + 2) + 2 < kWtHP
    const int r31 = 31 / g.W, j31 = 31 - r31 * g.W;
    MISO_REQUIRE((g.R + 2) * (g.W + 2) <= kWtHP && (g.R + 2) * (g.W + 2) <= kWtThreads * kWtSlots / 8 && r31 * (g.W + 2) + j31 + 2 * (g.W + 2) + 2 < kWtHP,
                 "wgrad_taps: halo geometry out of range (Fout=%d)", a.Fout);
    const int ntile = ceil_div(a.cin, kWgBM) * ceil_div(a.cout, 32);
    const int total = a.B * g.nrc * g.nseg;
    int splits = std::max(1, 148 / ntile);  // one CTA per SM (register-limited), one wave: half the atomics of two waves
    splits = std::max(1, std::min(splits, ceil_div(total, 8)));
    const size_t smem = (size_t)(2 * kWtHP * kWmAP + 2 * 32 * 40) * 2 + (size_t)a.B * kWtAff * sizeof(float2);
    MISO_REQUIRE(smem <= 200 * 1024, "wgrad: batch %d too large for the per-sample affine table", a.B);
    static size_t cur = 0;
    if (smem > 48 * 1024 && smem > cur) {
        MISO_CUDA(cudaFuncSetAttribute(wgrad_taps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cur = smem;
    }
    wgrad_taps_kernel<<<dim3(ntile, splits), kWtThreads, smem, st>>>(a, g, splits);
    MISO_LAUNCHED("wgrad_taps_kernel");
    return MISO_OK;
}

int launch_wgrad(const WgradArgs &a, cudaStream_t st) {
    MISO_REQUIRE(a.cin % 4 == 0 && a.x_coff % 4 == 0 && a.x_ctot % 4 == 0, "wgrad: input channels must be multiples of 4");
    MISO_REQUIRE(a.cout % 4 == 0 && a.dy_coff % 4 == 0 && a.dy_ctot % 4 == 0,
                 "wgrad: output channels must be multiples of 4 (cout=%d)", a.cout);
    MISO_REQUIRE(a.x_layout != LAYOUT_PLANES || a.x_ctot % 8 == 0, "wgrad: plane layout needs ctot %% 8 == 0");
    MISO_REQUIRE(a.x_coff % 8 == 0 && (a.cin % 8 == 0 || (a.x_layout == LAYOUT_PLANES && a.x_coff + a.cin <= a.x_ctot && a.x_sums == nullptr)),
                 "wgrad: input channel range must be 8-aligned (cin=%d coff=%d)", a.cin, a.x_coff);
    static const bool fma = getenv("MISO_WGRAD_FMA") && atoi(getenv("MISO_WGRAD_FMA")) != 0;  // debugging: the fp32 FMA GEMM
    if (fma) return a.cout <= 32 ? launch_wgrad_t<2>(a, st) : launch_wgrad_t<4>(a, st);
    if (wgrad_tc_eligible(a)) return launch_wgrad_tc(a, st);  // tcgen05 GEMM over the raw planes (the DenseBlock convs)
    static const bool no_taps = getenv("MISO_WGRAD_TAPS") && atoi(getenv("MISO_WGRAD_TAPS")) == 0;  // debugging: per-tap kernel only
    if (!no_taps && wgrad_taps_ok(a)) return launch_wgrad_taps(a, st);
    return a.cout <= 32 ? launch_wgrad_mma<32, 4>(a, st) : launch_wgrad_mma<64, 2>(a, st);
}

int launch_dgrad_pack(const float *src, float *dst, int taps, int cin, int cout, int cout_pad, int cin_pad, int flip, cudaStream_t st) {
    const int64_t total = (int64_t)taps * cout * cin_pad;
    const int blocks = (int)std::min<int64_t>((total + 255) / 256, 4096);
    dgrad_pack_kernel<<<blocks, 256, 0, st>>>(src, dst, taps, cin, cout, cout_pad, cin_pad, flip);
    MISO_LAUNCHED("dgrad_pack_kernel");
    return MISO_OK;
}

int launch_tcn_recompute(const TcnBwdArgs &a, float *Y, float *Q, cudaStream_t st) {
    tcn_recompute_kernel<<<dim3(ceil_div(a.T, kTcnBwFrames), a.B), 256, 7 * a.C * sizeof(float), st>>>(a, Y, Q);
    MISO_LAUNCHED("tcn_recompute_kernel");
    return MISO_OK;
}

int launch_gln_bwd(const TcnBwdArgs &a, float *DQ, const float *Y, double *gred, float *dgamma, float *dbeta, float *dalpha,
                   cudaStream_t st) {
    MISO_REQUIRE(a.C <= 128 * kTcnCh, "gln_bwd: C=%d > %d", a.C, 128 * kTcnCh);
    MISO_CUDA(cudaMemsetAsync(gred, 0, (size_t)a.B * 2 * sizeof(double), st));
    dim3 grid(ceil_div(a.T, kTcnRows), a.B);
    gln_bwd_reduce_kernel<<<grid, 128, 0, st>>>(a, DQ, Y, gred, dgamma, dbeta);
    MISO_LAUNCHED("gln_bwd_reduce_kernel");
    gln_bwd_apply_kernel<<<grid, 128, 0, st>>>(a, DQ, Y, gred, dalpha);
    MISO_LAUNCHED("gln_bwd_apply_kernel");
    return MISO_OK;
}

int launch_dw_bwd(const TcnBwdArgs &a, const float *DY, float *DN, double *ired, float *dwdw, float *out, int accumulate,
                  cudaStream_t st) {
    MISO_CUDA(cudaMemsetAsync(ired, 0, (size_t)a.B * a.C * 2 * sizeof(double), st));
    dim3 grid(ceil_div(a.T, kTcnRows), a.B);
    dw_bwd_kernel<<<dim3(ceil_div(a.T, kTcnRowsDw), a.B), kDwBwdThreads, 0, st>>>(a, DY, DN, ired, dwdw);
    MISO_LAUNCHED("dw_bwd_kernel");
    in1d_bwd_apply_kernel<<<grid, 128, 0, st>>>(a, DN, ired, out, accumulate);
    MISO_LAUNCHED("in1d_bwd_apply_kernel");
    return MISO_OK;
}

int launch_cl_to_planes(const float *src, __nv_bfloat16 *dst, int B, int npix, int C, cudaStream_t st) {
    MISO_REQUIRE(C % 8 == 0, "cl_to_planes: C=%d must be a multiple of 8", C);
    const int64_t total = (int64_t)B * npix * (C >> 2);
    const int blocks = (int)std::min<int64_t>((total + 255) / 256, 4096);
    cl_to_planes_kernel<<<blocks, 256, 0, st>>>(src, dst, B, npix, C);
    MISO_LAUNCHED("cl_to_planes_kernel");
    return MISO_OK;
}

int launch_planes_accumulate(const __nv_bfloat16 *src, int sctot, float *dst, int dctot, int dcoff, int C, int B, int npix,
                             cudaStream_t st) {
    MISO_REQUIRE(C % 4 == 0 && sctot % 8 == 0 && dctot % 4 == 0 && dcoff % 4 == 0, "planes_accumulate: bad channel layout");
    const int64_t total = (int64_t)B * npix * (C >> 2);
    const int blocks = (int)std::min<int64_t>((total + 255) / 256, 148 * 32);
    planes_accumulate_kernel<<<blocks, 256, 0, st>>>(src, sctot, dst, dctot, dcoff, C, B, npix);
    MISO_LAUNCHED("planes_accumulate_kernel");
    return MISO_OK;
}

int launch_copy_channels(const float *src, int sctot, int scoff, float *dst, int dctot, int dcoff, int C, int64_t rows,
                         int accumulate, cudaStream_t st) {
    const int64_t total = rows * C;
    const int blocks = (int)std::min<int64_t>((total + 255) / 256, 4096);
    copy_channels_kernel<<<blocks, 256, 0, st>>>(src, sctot, scoff, dst, dctot, dcoff, C, rows, accumulate);
    MISO_LAUNCHED("copy_channels_kernel");
    return MISO_OK;
}

}  // namespace miso

// =========================================================================== C ABI =====
namespace miso {
namespace {

struct BwdPerms {
    signed char p[24][4];
};

// gradient of criterion.py:8-63 w.r.t. the estimate, for the winning permutation of every utterance:
// L = (1/B) sum_b sum_i [ |re_i - r.re| + |im_i - r.im| + | sqrt(re^2 + im^2 + 1e-8) - |r| | ],  r = ref[perm_b[i]]
__global__ void upit_bwd_kernel(const float2 *__restrict__ est, int64_t e_sb, int64_t e_ss, const float2 *__restrict__ ref,
                                int64_t r_sb, int64_t r_ss, const int64_t *__restrict__ idx, BwdPerms perms, int B, int S, int64_t n,
                                const float *__restrict__ gout, float2 *__restrict__ grad) {
    const int64_t total = (int64_t)B * S * n;
    const float scale = gout[0] / (float)B;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = i % n;
        const int64_t r = i / n;
        const int s = (int)(r % S), b = (int)(r / S);
        const int j = perms.p[(int)idx[b]][s];
        const float2 x = est[b * e_sb + s * e_ss + e];
        const float2 y = ref[b * r_sb + j * r_ss + e];
        const float mag = sqrtf(x.x * x.x + x.y * x.y + 1e-8f);
        const float rm = sqrtf(y.x * y.x + y.y * y.y);
        auto sgn = [](float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); };
        const float sm = sgn(mag - rm) / mag;
        grad[i] = make_float2(scale * (sgn(x.x - y.x) + sm * x.x), scale * (sgn(x.y - y.y) + sm * x.y));
    }
}

// complex gradient [B][S][n] -> the network output layout fp32 [B][n][pitch] (re of every speaker, then im; pitch = 2S
// rounded up to a multiple of 4, padding zero)
__global__ void grad_pack_kernel(const float2 *__restrict__ g, float *__restrict__ gy, int B, int S, int64_t n, int pitch) {
    const int64_t total = (int64_t)B * n;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / n, e = i - b * n;
        for (int s = 0; s < S; ++s) {
            const float2 v = g[(b * S + s) * n + e];
            gy[i * pitch + s] = v.x;
            gy[i * pitch + S + s] = v.y;
        }
        for (int c = 2 * S; c < pitch; ++c) gy[i * pitch + c] = 0.f;
    }
}

// gradient of criterion.py:121-141 (loss_Enhance): the same three L1 terms without a permutation
__global__ void enhance_bwd_kernel(const float2 *__restrict__ est, const float2 *__restrict__ ref, int64_t total, float scale_div,
                                   const float *__restrict__ gout, float2 *__restrict__ grad) {
    const float scale = gout[0] / scale_div;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const float2 x = est[i], y = ref[i];
        const float mag = sqrtf(x.x * x.x + x.y * x.y + 1e-8f);
        const float rm = sqrtf(y.x * y.x + y.y * y.y);
        auto sgn = [](float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); };
        const float sm = sgn(mag - rm) / mag;
        grad[i] = make_float2(scale * (sgn(x.x - y.x) + sm * x.x), scale * (sgn(x.y - y.y) + sm * x.y));
    }
}

}  // namespace
}  // namespace miso

extern "C" {

int miso_upit_bwd(const void *d_est, int64_t e_sb, int64_t e_ss, const void *d_ref, int64_t r_sb, int64_t r_ss,
                  const int64_t *d_perm_idx, int B, int S, int T, int F, const float *d_gout, void *d_grad, void *stream) {
    using namespace miso;
    MISO_REQUIRE(d_est && d_ref && d_perm_idx && d_gout && d_grad, "miso_upit_bwd: null argument");
    MISO_REQUIRE(S >= 1 && S <= 4 && B >= 1, "miso_upit_bwd: S=%d unsupported (1..4)", S);
    BwdPerms t;
    int id[4] = {0, 1, 2, 3}, np = 0;
    do {  // lexicographic order == itertools.permutations(range(S)) (criterion.py:49), as in align.cu
        for (int i = 0; i < 4; ++i) t.p[np][i] = (signed char)(i < S ? id[i] : 0);
        np++;
    } while (std::next_permutation(id, id + S));
    const int64_t n = (int64_t)T * F;
    const int blocks = (int)std::min<int64_t>(((int64_t)B * S * n + 255) / 256, 148 * 16);
    upit_bwd_kernel<<<blocks, 256, 0, as_stream(stream)>>>(reinterpret_cast<const float2 *>(d_est), e_sb, e_ss,
                                                          reinterpret_cast<const float2 *>(d_ref), r_sb, r_ss, d_perm_idx, t, B, S, n,
                                                          d_gout, reinterpret_cast<float2 *>(d_grad));
    MISO_LAUNCHED("upit_bwd_kernel");
    return MISO_OK;
}

int miso_grad_pack(const void *d_grad, float *d_gy, int B, int S, int T, int F, void *stream) {
    using namespace miso;
    MISO_REQUIRE(d_grad && d_gy && S >= 1, "miso_grad_pack: bad argument");
    const int64_t n = (int64_t)T * F;
    const int blocks = (int)std::min<int64_t>((B * n + 255) / 256, 148 * 16);
    grad_pack_kernel<<<blocks, 256, 0, as_stream(stream)>>>(reinterpret_cast<const float2 *>(d_grad), d_gy, B, S, n, (2 * S + 7) & ~7);
    MISO_LAUNCHED("grad_pack_kernel");
    return MISO_OK;
}

int miso_loss_enhance_bwd(const void *d_est, const void *d_ref, int B, int64_t n_per_batch, const float *d_gout, void *d_grad,
                          void *stream) {
    using namespace miso;
    MISO_REQUIRE(d_est && d_ref && d_gout && d_grad && B >= 1 && n_per_batch >= 1, "miso_loss_enhance_bwd: bad argument");
    const int64_t total = (int64_t)B * n_per_batch;
    const int blocks = (int)std::min<int64_t>((total + 255) / 256, 148 * 16);
    enhance_bwd_kernel<<<blocks, 256, 0, as_stream(stream)>>>(reinterpret_cast<const float2 *>(d_est), reinterpret_cast<const float2 *>(d_ref),
                                                             total, (float)B, d_gout, reinterpret_cast<float2 *>(d_grad));
    MISO_LAUNCHED("enhance_bwd_kernel");
    return MISO_OK;
}

}  // extern "C"
