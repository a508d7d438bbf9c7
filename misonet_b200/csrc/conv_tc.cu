// tcgen05 implicit-GEMM 3x3 stride-1 convolution over channels-last activations -- the
// tensor-core path of the MISO conv stack (the DenseBlock convs, model.py:437-482, are 94 %
// of all FLOPs).
//
// Formulation.  One sample's input is viewed as a zero-padded raster of width
// W = Fin + 2*pad_f with one zero row above and below; the output uses the same raster
// (columns >= Fout = W-2 are discarded).  For tap (kt,kf) the input needed by output raster
// position R is raster position R + kt*W + kf: the im2col matrix of a tap is the SAME staged
// tile shifted by a constant number of rows.  A CTA stages a halo tile
// [H = 128*G + 2W + 2 raster pixels] x [16 channels] once per channel chunk into shared
// memory in the canonical no-swizzle K-major UMMA layout (pixel p, 8-channel group j at
// byte p*16 + j*H*16), and the nine taps are nine shared-memory descriptors whose start
// addresses differ by (kt*W + kf)*16 bytes.  Global traffic is ~(1 + (2W+2)/(128 G)) x the
// input instead of 9x.
//
// Pipeline (one CTA = G consecutive 128-row M tiles of one sample, accumulators in TMEM):
//   warps 0..7  producers: LDG fp32 -> producer's InstanceNorm as an affine -> bf16 (or a
//               bf16 hi/lo split) -> st.shared, plus the pre-packed weight image of the chunk;
//               then the epilogue: tcgen05.ld -> bias -> ELU -> store at a channel offset ->
//               (sum, sumsq) statistics for the consumers.
//   warp  8     one elected lane issues tcgen05.mma (M=128, N=cout_pad, K=16) per tap / M tile
//               and commits to the "buffer free" mbarriers.
// SPLIT = 1: bf16 operands (throughput mode, ~1e-2 relative error).
// SPLIT = 3: bf16x3 (a_hi*w_hi + a_lo*w_hi + a_hi*w_lo, fp32 accumulate): parity-grade.
#include <cuda_bf16.h>

#include "conv.cuh"

namespace miso {
namespace {

constexpr int kProducerWarps = 8;
constexpr int kThreads = (kProducerWarps + 1) * 32;
constexpr int CK = 16;  // channels per chunk = one UMMA K step
constexpr int kMaxItems = 10;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
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

// shared-memory matrix descriptor, SWIZZLE_NONE, K-major (cute::UMMA::SmemDescriptor):
// start address [0,14) >>4, leading byte offset [16,30) >>4 (between the two 8-element K groups),
// stride byte offset [32,46) >>4 (between 8-row groups), version [46,48) = 1, layout type [61,64) = 0.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// instruction descriptor (cute::UMMA::InstrDescriptor), kind::f16: D fp32, A/B bf16, K-major both
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

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&v);
}

struct TcGeom {
    int W, Fout, G, N, H, Hp, nchunk;
    size_t off_aff, off_a, a_buf_bytes, off_b, b_buf_bytes, total;
};

__host__ __device__ inline TcGeom tc_geom(const ConvArgs &a, int split) {
    TcGeom g;
    g.W = a.Fin + 2 * a.pad_f;
    g.Fout = g.W - 2;
    g.G = a.tc_G;
    g.N = a.cout_pad16;
    g.H = 128 * g.G + 2 * g.W + 2;
    g.Hp = (g.H + 15) & ~15;
    g.nchunk = (a.cin + CK - 1) / CK;
    const int nsp = split == 3 ? 2 : 1;
    g.off_aff = 64;
    g.off_a = (g.off_aff + (size_t)((a.cin + 3) & ~3) * 8 + 127) & ~(size_t)127;
    g.a_buf_bytes = (size_t)nsp * 2 * g.Hp * 16;
    g.off_b = g.off_a + 2 * g.a_buf_bytes;
    g.b_buf_bytes = (size_t)nsp * 9 * 2 * g.N * 16;
    g.total = g.off_b + 2 * g.b_buf_bytes;
    return g;
}

template <int SPLIT>
__global__ void __launch_bounds__(kThreads, 1) conv_tc_kernel(const ConvArgs a) {
    constexpr int NSP = SPLIT == 3 ? 2 : 1;
    extern __shared__ __align__(128) uint8_t smem[];
    const TcGeom g = tc_geom(a, SPLIT);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y;
    const int R0 = blockIdx.x * 128 * g.G;
    const int W = g.W, N = g.N, Hp = g.Hp;

    // barriers: full[2] (producers -> MMA), empty[2] (MMA commit -> producers), done
    const uint32_t bar_full = smem_u32(smem), bar_empty = smem_u32(smem + 16), bar_done = smem_u32(smem + 32);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + 40);
    float2 *aff = reinterpret_cast<float2 *>(smem + g.off_aff);
    const uint32_t sA = smem_u32(smem + g.off_a), sB = smem_u32(smem + g.off_b);

    if (tid == 0) {
        mbar_init(bar_full, kProducerWarps * 32);
        mbar_init(bar_full + 8, kProducerWarps * 32);
        mbar_init(bar_empty, 1);
        mbar_init(bar_empty + 8, 1);
        mbar_init(bar_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kProducerWarps) {
        // TMEM: G accumulators of N fp32 columns (power of two >= 32)
        uint32_t cols = 32;
        while (cols < (uint32_t)(g.G * N)) cols <<= 1;
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int c = tid; c < a.cin; c += kThreads) {
        float2 v = make_float2(1.f, 0.f);
        if (a.norm_mode == NORM_IN) {
            const double *s = a.in_sums + ((size_t)b * a.in_ctot + a.in_coff + c) * 2;
            v = affine_from_sums(s[0], s[1], a.norm_inv_n, (double)a.norm_eps);
        }
        aff[c] = v;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp < kProducerWarps) {
        // ------------------------------------------------------------------ producers
        // item i of this thread: halo pixel h = (warp + 8*i)*16 + (lane & 15), channel half kc = lane >> 4
        const int kc = lane >> 4;
        int poff[kMaxItems];
        const int ngroups = Hp >> 4;
#pragma unroll
        for (int i = 0; i < kMaxItems; ++i) {
            const int grp = warp + kProducerWarps * i;
            const int h = grp * 16 + (lane & 15);
            int off = -1;
            if (grp < ngroups && h < g.H) {
                const int Q = R0 + h;
                const int row = Q / W;
                const int ti = row - 1;
                const int fi = Q - row * W - a.pad_f;
                if (ti >= 0 && ti < a.T && fi >= 0 && fi < a.Fin) off = ti * a.Fin + fi;
            }
            poff[i] = off;
        }
        const float *in_b = a.in + (size_t)b * a.T * a.Fin * a.in_ctot + a.in_coff;
        const uint4 *wimg = reinterpret_cast<const uint4 *>(a.w_tc);
        const int b_vec = (int)(g.b_buf_bytes >> 4);

        for (int c = 0; c < g.nchunk; ++c) {
            const int buf = c & 1;
            if (c >= 2) mbar_wait(bar_empty + 8 * buf, ((c >> 1) + 1) & 1);
            const int ch = c * CK + kc * 8;
            const bool ok0 = ch < a.cin, ok1 = ch + 4 < a.cin;
            float2 f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = (ch + j < a.cin) ? aff[ch + j] : make_float2(0.f, 0.f);
            uint8_t *abuf = smem + g.off_a + (size_t)buf * g.a_buf_bytes;
#pragma unroll
            for (int i = 0; i < kMaxItems; ++i) {
                const int grp = warp + kProducerWarps * i;
                if (grp < ngroups) {
                    const int h = grp * 16 + (lane & 15);
                    float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
                    float x[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                    if (poff[i] >= 0) {
                        const float *p = in_b + (size_t)poff[i] * a.in_ctot + ch;
                        if (ok0) v0 = __ldg(reinterpret_cast<const float4 *>(p));
                        if (ok1) v1 = __ldg(reinterpret_cast<const float4 *>(p + 4));
                        x[0] = fmaf(v0.x, f[0].x, f[0].y);
                        x[1] = fmaf(v0.y, f[1].x, f[1].y);
                        x[2] = fmaf(v0.z, f[2].x, f[2].y);
                        x[3] = fmaf(v0.w, f[3].x, f[3].y);
                        x[4] = fmaf(v1.x, f[4].x, f[4].y);
                        x[5] = fmaf(v1.y, f[5].x, f[5].y);
                        x[6] = fmaf(v1.z, f[6].x, f[6].y);
                        x[7] = fmaf(v1.w, f[7].x, f[7].y);
                        if (!ok0) x[0] = x[1] = x[2] = x[3] = 0.f;
                        if (!ok1) x[4] = x[5] = x[6] = x[7] = 0.f;
                    }
                    uint4 hi;
                    hi.x = pack_bf16(x[0], x[1]);
                    hi.y = pack_bf16(x[2], x[3]);
                    hi.z = pack_bf16(x[4], x[5]);
                    hi.w = pack_bf16(x[6], x[7]);
                    *reinterpret_cast<uint4 *>(abuf + ((size_t)kc * Hp + h) * 16) = hi;
                    if constexpr (NSP == 2) {
                        float r[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) r[j] = x[j] - __bfloat162float(__float2bfloat16_rn(x[j]));
                        uint4 lo;
                        lo.x = pack_bf16(r[0], r[1]);
                        lo.y = pack_bf16(r[2], r[3]);
                        lo.z = pack_bf16(r[4], r[5]);
                        lo.w = pack_bf16(r[6], r[7]);
                        *reinterpret_cast<uint4 *>(abuf + ((size_t)(2 + kc) * Hp + h) * 16) = lo;
                    }
                }
            }
            // weight image of this chunk: [NSP][9 taps][2 k-groups][N][8] bf16, already in smem order
            uint4 *bbuf = reinterpret_cast<uint4 *>(smem + g.off_b + (size_t)buf * g.b_buf_bytes);
            const uint4 *wsrc = wimg + (size_t)c * b_vec;
            for (int i = tid; i < b_vec; i += kProducerWarps * 32) bbuf[i] = __ldg(wsrc + i);
            // generic-proxy writes -> visible to the tensor core (async proxy), then signal "full"
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(bar_full + 8 * buf);
        }

        // ------------------------------------------------------------------ epilogue
        mbar_wait(bar_done, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int quad = warp & 3;
        const int npix = a.T * W;
        float *red = reinterpret_cast<float *>(smem + g.off_a);  // [N][2], A buffers are free now
        if (a.out_sums) {
            for (int i = tid; i < 2 * N; i += kProducerWarps * 32) red[i] = 0.f;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kProducerWarps * 32));
        float *out_b = a.out + (size_t)b * a.T * g.Fout * a.out_ctot + a.out_coff;
        for (int gt = warp >> 2; gt < g.G; gt += 2) {
            const int R = R0 + gt * 128 + quad * 32 + lane;
            const int t = R / W;
            const int wo = R - t * W;
            const bool valid = R < npix && wo < g.Fout;
            float *o = out_b + ((size_t)t * g.Fout + wo) * a.out_ctot;
            for (int j = 0; j < N; j += 16) {
                uint32_t v[16];
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(gt * N + j);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                float y[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    float x = __uint_as_float(v[q]) + (a.bias ? a.bias[j + q] : 0.f);
                    if (a.elu) x = elu1(x);
                    y[q] = valid ? x : 0.f;
                }
                if (valid) {
                    if (((a.out_ctot | a.out_coff) & 3) == 0 && j + 16 <= a.cout) {
#pragma unroll
                        for (int q = 0; q < 16; q += 4)
                            *reinterpret_cast<float4 *>(o + j + q) = make_float4(y[q], y[q + 1], y[q + 2], y[q + 3]);
                    } else {
#pragma unroll
                        for (int q = 0; q < 16; ++q)
                            if (j + q < a.cout) o[j + q] = y[q];
                    }
                }
                if (a.out_sums) {
#pragma unroll
                    for (int q = 0; q < 16; ++q) {
                        float s = warp_sum(y[q]);
                        float sq = warp_sum(y[q] * y[q]);
                        if (lane == 0) {
                            atomicAdd(&red[(j + q) * 2], s);
                            atomicAdd(&red[(j + q) * 2 + 1], sq);
                        }
                    }
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("bar.sync 1, %0;" ::"n"(kProducerWarps * 32));
        if (a.out_sums && tid < N && tid < a.cout) {
            double *dst = a.out_sums + ((size_t)b * a.out_ctot + a.out_coff + tid) * 2;
            atomicAdd(dst, (double)red[tid * 2]);
            atomicAdd(dst + 1, (double)red[tid * 2 + 1]);
        }
    } else {
        // ------------------------------------------------------------------ MMA issuer
        const uint32_t idesc = make_idesc(N);
        const uint32_t a_lbo = (uint32_t)Hp * 16, b_lbo = (uint32_t)N * 16;
        for (int c = 0; c < g.nchunk; ++c) {
            const int buf = c & 1;
            mbar_wait(bar_full + 8 * buf, (c >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
                const uint32_t abase = sA + (uint32_t)(buf * g.a_buf_bytes);
                const uint32_t bbase = sB + (uint32_t)(buf * g.b_buf_bytes);
#pragma unroll 1
                for (int tap = 0; tap < 9; ++tap) {
                    const int kt = tap / 3, kf = tap - kt * 3;
#pragma unroll 1
                    for (int sp = 0; sp < SPLIT; ++sp) {
                        const int asel = sp == 1 ? 1 : 0, bsel = sp == 2 ? 1 : 0;
                        const uint64_t bdesc = make_desc(bbase + (uint32_t)((bsel * 9 + tap) * 2 * N * 16), b_lbo, 128);
                        for (int gt = 0; gt < g.G; ++gt) {
                            const uint32_t aaddr = abase + (uint32_t)(asel * 2 * Hp * 16) + (uint32_t)((gt * 128 + kt * W + kf) * 16);
                            umma_bf16(tmem_base + (uint32_t)(gt * N), make_desc(aaddr, a_lbo, 128), bdesc, idesc,
                                      (c | tap | sp) != 0 ? 1u : 0u);
                        }
                    }
                }
                umma_commit(bar_empty + 8 * buf);
                if (c == g.nchunk - 1) umma_commit(bar_done);
            }
            __syncwarp();
        }
    }
    __syncthreads();
    if (warp == kProducerWarps) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t cols = 32;
        while (cols < (uint32_t)(g.G * N)) cols <<= 1;
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(cols) : "memory");
    }
}

// weight image for the tcgen05 path: [nchunk][NSP (hi, lo)][9 taps][2 k-groups][N][8] bf16.
// Deconvolutions (stride 1) are convolutions with the kernel flipped in both directions.
__global__ void pack_tc_w_kernel(const float *__restrict__ w, __nv_bfloat16 *__restrict__ img, int cout, int cin, int N,
                                 int nchunk, int nsp, int transposed) {
    const int64_t total = (int64_t)nchunk * nsp * 9 * 2 * N * 8;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int e = (int)(i & 7);
        int64_t r = i >> 3;
        int n = (int)(r % N);
        r /= N;
        int kg = (int)(r & 1);
        r >>= 1;
        int tap = (int)(r % 9);
        r /= 9;
        int sp = (int)(r % nsp);
        int c = (int)(r / nsp);
        int ci = c * CK + kg * 8 + e;
        float v = 0.f;
        if (n < cout && ci < cin) {
            if (transposed)
                v = w[((int64_t)ci * cout + n) * 9 + (8 - tap)];
            else
                v = w[((int64_t)n * cin + ci) * 9 + tap];
        }
        __nv_bfloat16 hi = __float2bfloat16_rn(v);
        img[i] = sp == 0 ? hi : __float2bfloat16_rn(v - __bfloat162float(hi));
    }
}

}  // namespace

size_t conv_tc_weight_elems(int cin, int cout_pad16, int nsp) {
    return (size_t)((cin + CK - 1) / CK) * nsp * 9 * 2 * cout_pad16 * 8;
}

int pack_conv_tc_weights(const float *d_w, void *d_img, int cout, int cin, int cout_pad16, int nsp, int transposed,
                         cudaStream_t stream) {
    const int nchunk = (cin + CK - 1) / CK;
    const size_t total = conv_tc_weight_elems(cin, cout_pad16, nsp);
    const int blocks = (int)((total + 255) / 256 > 4096 ? 4096 : (total + 255) / 256);
    pack_tc_w_kernel<<<blocks, 256, 0, stream>>>(d_w, reinterpret_cast<__nv_bfloat16 *>(d_img), cout, cin, cout_pad16, nchunk,
                                                  nsp, transposed);
    MISO_LAUNCHED("pack_tc_w_kernel");
    return MISO_OK;
}

bool conv_tc_eligible(const ConvArgs &a) {
    return a.KT == 3 && a.KF == 3 && a.stride_f == 1 && a.pad_t == 1 && a.cout_pad16 <= 64 && a.cin % 4 == 0 &&
           a.resid == nullptr && a.norm_mode != NORM_GLN && a.w_tc != nullptr;
}

int conv_tc_pick_G(const ConvArgs &a) {
    // enough CTAs to fill the machine twice, otherwise the largest G the TMEM/halo budget allows
    const int W = a.Fin + 2 * a.pad_f;
    const int tiles = ceil_div(a.T * W, 128);
    int G = 4;
    while (G > 1 && (int64_t)ceil_div(tiles, G) * a.B < 2 * 148) G >>= 1;
    while (G > 1 && G * a.cout_pad16 > 256) G >>= 1;
    return G;
}

int launch_conv_tc(const ConvArgs &a_in, int split, cudaStream_t stream) {
    ConvArgs a = a_in;
    MISO_REQUIRE(conv_tc_eligible(a), "conv_tc: layer not eligible for the tcgen05 path");
    MISO_REQUIRE(split == 1 || split == 3, "conv_tc: bad split");
    if (a.transposed) {  // stride-1 ConvTranspose2d == conv with flipped kernel and pad_f -> 2 - pad_f
        a.pad_f = 2 - a.pad_f;
        a.transposed = 0;
    }
    a.tc_G = conv_tc_pick_G(a);
    const TcGeom g = tc_geom(a, split);
    MISO_REQUIRE(g.Fout == a.Fout, "conv_tc: Fout %d inconsistent with Fin %d pad %d", a.Fout, a.Fin, a.pad_f);
    MISO_REQUIRE(ceil_div(g.Hp >> 4, kProducerWarps) <= kMaxItems, "conv_tc: halo %d too large", g.H);
    MISO_REQUIRE(g.total <= 227 * 1024, "conv_tc: shared memory %zu exceeds 227 KB", g.total);
    static bool attr_set[2] = {false, false};
    const int ai = split == 3 ? 1 : 0;
    if (!attr_set[ai]) {
        cudaError_t e = split == 3 ? cudaFuncSetAttribute(conv_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)
                                   : cudaFuncSetAttribute(conv_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(conv_tc_kernel)");
        attr_set[ai] = true;
    }
    dim3 grid(ceil_div(a.T * g.W, 128 * g.G), a.B);
    prof_begin(stream);
    if (split == 3)
        conv_tc_kernel<3><<<grid, kThreads, g.total, stream>>>(a);
    else
        conv_tc_kernel<1><<<grid, kThreads, g.total, stream>>>(a);
    {
        const double flops = 2.0 * a.B * a.T * a.Fout * a.cin * a.cout * 9.0;
        const double bytes = 4.0 * a.B * a.T * ((double)a.Fin * a.cin + (double)a.Fout * a.cout);
        prof_end(stream, flops, bytes, MISO_PROF_CONV_TC);
    }
    MISO_LAUNCHED("conv_tc_kernel");
    return MISO_OK;
}

}  // namespace miso
