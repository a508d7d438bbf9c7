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
    float *red = reinterpret_cast<float *>(smem + g.off_red);   // [8 warps][16][2]
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
        const double mean = stat_get(a.gln_sums + (size_t)b * 2) * a.gln_inv_n;
        double var = stat_get(a.gln_sums + (size_t)b * 2 + 1) * a.gln_inv_n - mean * mean;
        if (var < 0.0) var = 0.0;
        const double rstd_d = rsqrt(var + (double)a.gln_eps);
        const float rstd = (float)rstd_d;
        const float mr = (float)(mean * rstd_d);
        const int co0 = nh * Nt;
        for (int i = et; i < Nt; i += kPwEpiThreads) vec[i] = __ldg(a.wbeta + co0 + i) - mr * __ldg(a.wgamma + co0 + i);
        asm volatile("bar.sync 1, %0;" ::"n"(kPwEpiThreads));
        const int r = quad * 32 + lane, t = t0 + r;
        const bool valid = t < a.T;
        const size_t row = ((size_t)b * a.T + t) * a.C + co0;
        float *myred = red + (warp - kPwEpi0) * 32;
        const int ncol = Nt / 2;  // columns of this warp's half
        mbar_wait(bar_done, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int cb = half * ncol; cb < (half + 1) * ncol; cb += 16) {
            uint32_t v[16];
            tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)cb, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            float y[16], ssum[16], ssq[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) y[q] = fmaf(__uint_as_float(v[q]), rstd, vec[cb + q]);
            if (valid) {
                if (a.resid) {
                    const float *rp = a.resid + row + cb;
#pragma unroll
                    for (int q = 0; q < 16; q += 4) {
                        const float4 rr = *reinterpret_cast<const float4 *>(rp + q);
                        y[q] += rr.x;
                        y[q + 1] += rr.y;
                        y[q + 2] += rr.z;
                        y[q + 3] += rr.w;
                    }
                }
                float *o = a.out + row + cb;
#pragma unroll
                for (int q = 0; q < 16; q += 4) *reinterpret_cast<float4 *>(o + q) = make_float4(y[q], y[q + 1], y[q + 2], y[q + 3]);
            }
            if (a.out_sums) {
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    ssum[q] = valid ? y[q] : 0.f;
                    ssq[q] = valid ? y[q] * y[q] : 0.f;
                }
                const float s = warp_reduce16(ssum, lane);
                const float q2 = warp_reduce16(ssq, lane);
                if ((lane & 1) == 0) {
                    myred[(lane >> 1) * 2] = s;
                    myred[(lane >> 1) * 2 + 1] = q2;
                }
                // the four warps of this column half (one per lane quadrant) -> one total per channel, fixed order
                asm volatile("bar.sync %0, 128;" ::"r"(2 + half));
                const int hw = ((warp - kPwEpi0) & 3) * 32 + lane;
                if (hw < 16) {
                    double s8 = 0.0, q8 = 0.0;
#pragma unroll
                    for (int w4 = 0; w4 < 4; ++w4) {
                        s8 += (double)red[(half * 4 + w4) * 32 + hw * 2];
                        q8 += (double)red[(half * 4 + w4) * 32 + hw * 2 + 1];
                    }
                    double *dst = a.out_sums + ((size_t)b * a.C + co0 + cb + hw) * 2;
                    stat_add(dst, s8);
                    stat_add(dst + 1, q8);
                }
                asm volatile("bar.sync %0, 128;" ::"r"(2 + half));
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
__global__ void __launch_bounds__(256) tcn_wprep_kernel(const TcnPwTable tab, __nv_bfloat16 *wimg, float *wvec, int C, int cpad, int Nt,
                                                        int nsp) {
    const int h = blockIdx.y;
    const float *W = tab.w[h], *gamma = tab.gamma[h], *beta = tab.beta[h];
    const int nunit = C / 16, NH = C / Nt;
    const size_t unit_elems = (size_t)nsp * 2 * Nt * 8;
    if ((int)blockIdx.x < nunit) {
        const int unit = blockIdx.x;
        __shared__ float gm[16];
        if (threadIdx.x < 16) gm[threadIdx.x] = gamma[unit * 16 + threadIdx.x];
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
            const int co = i % C, kg = i / C;
            float hv[8], lv[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int ci = unit * 16 + kg * 8 + e;
                const float v = W[(size_t)ci * cpad + co] * gm[kg * 8 + e];
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
        if (co < C) {
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
    for (int nt : {192, 128, 256, 96, 64, 32})
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
    g.off_vec = g.off_red + 8 * 32 * 4;
    g.off_stage = pw_round_up(g.off_vec + g.Nt * 4, 1024);
    g.nstage = std::min(4, (kPwSmemLimit - g.off_stage) / g.stage);
    if (g.nstage < 2) return false;
    g.smem_total = g.off_stage + g.nstage * g.stage;
    int cols = 32;
    while (cols < g.Nt) cols <<= 1;
    g.tmem_cols = cols;
    return true;
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

int launch_tcn_wprep(const TcnPwTable &tab, int nconv, int C, int cpad, int split, void *wimg, float *wvec, cudaStream_t stream) {
    PwGeom g;
    MISO_REQUIRE(make_pw_geom(C, split, g), "tcn_pw: unsupported channel count %d", C);
    MISO_REQUIRE(nconv <= kTcnMaxPw, "tcn_pw: too many pointwise convs (%d)", nconv);
    dim3 grid(g.nunit + (C + 255) / 256, nconv);
    tcn_wprep_kernel<<<grid, 256, 0, stream>>>(tab, reinterpret_cast<__nv_bfloat16 *>(wimg), wvec, C, cpad, g.Nt, g.nsp);
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

}  // namespace miso
