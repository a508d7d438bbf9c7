// fp32 FMA implicit-GEMM (de)convolution over channels-last activations -- the
// "parity mode" of the MISO conv stack (N3/N4/N5 and the TCN pointwise convs of N6,
// SURVEY.md section 8(a); reference model.py:401-482, 556-561).
//
// GEMM view: M = output pixels of one sample (tile of 128), N = output channels
// (tile of 8/32/64), K = taps x input channels (chunks of 16 channels of one tap).
// The producer's InstanceNorm / gLN is folded into the A-operand load, the epilogue
// adds bias (+ residual), applies ELU, stores at a channel offset of the destination
// buffer (this is what removes every torch.cat of the reference, model.py:470-479,99)
// and accumulates the (sum, sumsq) statistics of the stored tensor for its consumers.
#include "conv.cuh"

namespace miso {

namespace {

constexpr int kThreads = 128;
constexpr int BM = 128;
constexpr int BK = 16;
constexpr int AS = BM + 4;  // padded row pitch of the transposed A tile

template <int TN>
__global__ void __launch_bounds__(kThreads) conv_fp32_kernel(const ConvArgs a) {
    constexpr int BN = 8 * TN;
    constexpr int B_VEC = BK * BN / 4;                          // float4 per B chunk
    constexpr int B_PER_THREAD = (B_VEC + kThreads - 1) / kThreads;
    extern __shared__ __align__(16) float smem[];
    const int cin4 = (a.cin + 3) & ~3;
    float2 *aff = reinterpret_cast<float2 *>(smem);
    float *As = smem + 2 * cin4;
    float *Bs = As + BK * AS;

    const int tid = threadIdx.x;
    const int b = blockIdx.z;
    const int n0 = blockIdx.y * BN;
    const int npix = a.T * a.Fout;

    for (int c = tid; c < a.cin; c += kThreads) {
        float2 v = make_float2(1.f, 0.f);
        if (a.norm_mode == NORM_IN) {
            const double *s = a.in_sums + ((size_t)b * a.in_ctot + a.in_coff + c) * 2;
            v = affine_from_sums(stat_get(s), stat_get(s + 1), a.norm_inv_n, (double)a.norm_eps);
        } else if (a.norm_mode == NORM_GLN) {
            const double *s = a.in_sums + (size_t)b * 2;
            double mean = stat_get(s) * a.norm_inv_n;
            double var = stat_get(s + 1) * a.norm_inv_n - mean * mean;
            if (var < 0.0) var = 0.0;
            double r = rsqrt(var + (double)a.norm_eps);
            double g = (double)a.gamma[c];
            v = make_float2((float)(g * r), (float)((double)a.beta[c] - g * mean * r));
        }
        aff[c] = v;
    }
    __syncthreads();

    // loader coordinates: 4 pixels x one float4 of channels per thread and chunk
    const int c4 = tid & 3;
    int lt[4], lf[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int pix = blockIdx.x * BM + (tid >> 2) + 32 * j;
        if (pix < npix) {
            lt[j] = pix / a.Fout;
            lf[j] = pix - lt[j] * a.Fout;
        } else {
            lt[j] = -(1 << 20);
            lf[j] = 0;
        }
    }
    const int nchunk = (a.cin + BK - 1) / BK;
    const int nk = a.KT * a.KF * nchunk;
    const float *in_b = reinterpret_cast<const float *>(a.in) + (size_t)b * a.T * a.Fin * a.in_ctot + a.in_coff;
    // plane layout: [B][hi|lo][in_ctot/8][T][Fin][8] bf16
    const __nv_bfloat16 *in_pl = reinterpret_cast<const __nv_bfloat16 *>(a.in) + (size_t)b * 2 * a.in_ctot * a.T * a.Fin;
    const bool in_planes = a.in_layout == LAYOUT_PLANES;

    float4 ra[4];
    float4 rb[B_PER_THREAD];

    auto load_chunk = [&](int it) {
        const int tap = it / nchunk;
        const int c0 = (it - tap * nchunk) * BK;
        const int kt = tap / a.KF;
        const int kf = tap - kt * a.KF;
        const int c = c0 + c4 * 4;
        const bool cok = c < a.cin;
        float2 f0 = make_float2(0.f, 0.f), f1 = f0, f2 = f0, f3 = f0;
        if (cok) {
            f0 = aff[c];
            f1 = aff[c + 1];
            f2 = aff[c + 2];
            f3 = aff[c + 3];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int ti, fi;
            bool ok = cok;
            if (!a.transposed) {
                ti = lt[j] + kt - a.pad_t;
                fi = lf[j] * a.stride_f + kf - a.pad_f;
            } else {
                ti = lt[j] + a.pad_t - kt;
                int num = lf[j] + a.pad_f - kf;
                fi = num / a.stride_f;
                ok = ok && num >= 0 && (fi * a.stride_f == num);
            }
            ok = ok && ti >= 0 && ti < a.T && fi >= 0 && fi < a.Fin;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok) {
                if (in_planes) {
                    const int ca = a.in_coff + c;
                    const __nv_bfloat16 *p = in_pl + (((size_t)(ca >> 3) * a.T + ti) * a.Fin + fi) * 8 + (ca & 7);
                    const uint2 h = __ldg(reinterpret_cast<const uint2 *>(p));
                    v = make_float4(bf16_lo(h.x), bf16_hi(h.x), bf16_lo(h.y), bf16_hi(h.y));
                    if (a.use_lo) {
                        const uint2 l = __ldg(reinterpret_cast<const uint2 *>(reinterpret_cast<const char *>(p) + a.in_lo_off));
                        v.x += bf16_lo(l.x);
                        v.y += bf16_hi(l.x);
                        v.z += bf16_lo(l.y);
                        v.w += bf16_hi(l.y);
                    }
                } else {
                    v = __ldg(reinterpret_cast<const float4 *>(in_b + ((size_t)ti * a.Fin + fi) * a.in_ctot + c));
                }
                v.x = fmaf(v.x, f0.x, f0.y);
                v.y = fmaf(v.y, f1.x, f1.y);
                v.z = fmaf(v.z, f2.x, f2.y);
                v.w = fmaf(v.w, f3.x, f3.y);
            }
            ra[j] = v;
        }
#pragma unroll
        for (int r = 0; r < B_PER_THREAD; ++r) {
            int idx = tid + kThreads * r;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (idx < B_VEC) {
                int k = idx / (BN / 4);
                int n4 = idx - k * (BN / 4);
                if (c0 + k < a.cin)
                    v = __ldg(reinterpret_cast<const float4 *>(a.w + ((size_t)tap * a.cin + c0 + k) * a.cout_pad + n0 +
                                                               n4 * 4));
            }
            rb[r] = v;
        }
    };

    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    const int ty = tid >> 3, tx = tid & 7;

    load_chunk(0);
    for (int it = 0; it < nk; ++it) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int lp = (tid >> 2) + 32 * j;
            As[(c4 * 4 + 0) * AS + lp] = ra[j].x;
            As[(c4 * 4 + 1) * AS + lp] = ra[j].y;
            As[(c4 * 4 + 2) * AS + lp] = ra[j].z;
            As[(c4 * 4 + 3) * AS + lp] = ra[j].w;
        }
#pragma unroll
        for (int r = 0; r < B_PER_THREAD; ++r) {
            int idx = tid + kThreads * r;
            if (idx < B_VEC) *reinterpret_cast<float4 *>(Bs + idx * 4) = rb[r];
        }
        __syncthreads();
        if (it + 1 < nk) load_chunk(it + 1);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float4 a0 = *reinterpret_cast<const float4 *>(As + k * AS + ty * 8);
            float4 a1 = *reinterpret_cast<const float4 *>(As + k * AS + ty * 8 + 4);
            float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float bv[TN];
            if constexpr (TN == 8) {
                float4 b0 = *reinterpret_cast<const float4 *>(Bs + k * BN + tx * 8);
                float4 b1 = *reinterpret_cast<const float4 *>(Bs + k * BN + tx * 8 + 4);
                bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w;
                bv[4] = b1.x; bv[5] = b1.y; bv[6] = b1.z; bv[7] = b1.w;
            } else if constexpr (TN == 4) {
                float4 b0 = *reinterpret_cast<const float4 *>(Bs + k * BN + tx * 4);
                bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w;
            } else {
#pragma unroll
                for (int j = 0; j < TN; ++j) bv[j] = Bs[k * BN + tx * TN + j];
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }

    // ---- epilogue -------------------------------------------------------------------
    float bias[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) bias[j] = a.bias ? a.bias[n0 + tx * TN + j] : 0.f;
    float ssum[TN], ssq[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) ssum[j] = ssq[j] = 0.f;

    float *out_b = reinterpret_cast<float *>(a.out) + (size_t)b * npix * a.out_ctot + a.out_coff;
    __nv_bfloat16 *out_pl = reinterpret_cast<__nv_bfloat16 *>(a.out) + (size_t)b * 2 * a.out_ctot * npix;
    const bool out_planes = a.out_layout == LAYOUT_PLANES;
    const float *res_b = a.resid ? a.resid + (size_t)b * npix * a.resid_ctot + a.resid_coff : nullptr;
    const int co0 = n0 + tx * TN;
    const bool vec_ok = (TN % 4 == 0) && ((a.out_ctot & 3) == 0) && ((a.out_coff & 3) == 0) && (co0 + TN <= a.cout);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int pix = blockIdx.x * BM + ty * 8 + i;
        if (pix >= npix) continue;
        float v[TN];
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            float x = acc[i][j] + bias[j];
            if (res_b && co0 + j < a.cout) x += res_b[(size_t)pix * a.resid_ctot + co0 + j];
            if (a.elu) x = elu1(x);
            v[j] = x;
            ssum[j] += x;
            ssq[j] += x * x;
        }
        if (out_planes) {
            // channels co0 .. co0+TN-1 of pixel `pix` in the [C/8][T*F][8] plane layout, hi and lo sets
#pragma unroll
            for (int j0 = 0; j0 < TN; j0 += 4) {
                const int ca = a.out_coff + co0 + j0;
                __nv_bfloat16 *p = out_pl + ((size_t)(ca >> 3) * npix + pix) * 8 + (ca & 7);
                if constexpr (TN >= 4) {
                    if (co0 + j0 + 4 <= a.cout) {
                        float h[4], l[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            h[q] = bf16_round(v[j0 + q]);
                            l[q] = v[j0 + q] - h[q];
                        }
                        *reinterpret_cast<uint2 *>(p) = make_uint2(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]));
                        if (a.use_lo)
                            *reinterpret_cast<uint2 *>(reinterpret_cast<char *>(p) + a.out_lo_off) =
                                make_uint2(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]));
                        continue;
                    }
                }
#pragma unroll
                for (int q = 0; q < (TN < 4 ? TN : 4); ++q) {
                    if (co0 + j0 + q < a.cout) {
                        const float h = bf16_round(v[j0 + q]);
                        p[q] = __float2bfloat16_rn(h);
                        if (a.use_lo)
                            *reinterpret_cast<__nv_bfloat16 *>(reinterpret_cast<char *>(p + q) + a.out_lo_off) =
                                __float2bfloat16_rn(v[j0 + q] - h);
                    }
                }
            }
            continue;
        }
        float *o = out_b + (size_t)pix * a.out_ctot + co0;
        if (vec_ok) {
#pragma unroll
            for (int j = 0; j < TN; j += 4)
                *reinterpret_cast<float4 *>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
#pragma unroll
            for (int j = 0; j < TN; ++j)
                if (co0 + j < a.cout) o[j] = v[j];
        }
    }

    if (a.out_sums) {
        float *red = As;  // [16][BN][2], free after the trailing barrier of the main loop
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            red[(ty * BN + tx * TN + j) * 2 + 0] = ssum[j];
            red[(ty * BN + tx * TN + j) * 2 + 1] = ssq[j];
        }
        __syncthreads();
        if (tid < BN && n0 + tid < a.cout) {
            double s = 0.0, q = 0.0;
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                s += (double)red[(r * BN + tid) * 2 + 0];
                q += (double)red[(r * BN + tid) * 2 + 1];
            }
            double *dst = a.out_sums + ((size_t)b * a.out_ctot + a.out_coff + n0 + tid) * 2;
            stat_add(dst, s);
            stat_add(dst + 1, q);
        }
    }
}

template <int TN>
int launch(const ConvArgs &a, cudaStream_t stream) {
    constexpr int BN = 8 * TN;
    const int cin4 = (a.cin + 3) & ~3;
    size_t smem = (size_t)(2 * cin4 + BK * AS + BK * BN) * sizeof(float);
    dim3 grid(ceil_div(a.T * a.Fout, BM), a.cout_pad / BN, a.B);
    prof_begin(stream);
    conv_fp32_kernel<TN><<<grid, kThreads, smem, stream>>>(a);
    {
        // algorithmic work: a transposed conv does one MAC per input pixel and tap
        const double pix = (double)a.B * a.T * (a.transposed ? a.Fin : a.Fout);
        const double flops = 2.0 * pix * a.cin * a.cout * a.KT * a.KF;
        const double bytes = 4.0 * a.B * a.T * ((double)a.Fin * a.cin + (double)a.Fout * a.cout);  // fp32-equivalent (hi + lo)
        prof_end(stream, flops, bytes, MISO_PROF_CONV_FP32);
    }
    MISO_LAUNCHED("conv_fp32_kernel");
    return MISO_OK;
}

}  // namespace

int conv_fp32_tile_n(int cout) { return cout <= 8 ? 8 : (cout <= 32 ? 32 : 64); }

int launch_conv_fp32(const ConvArgs &a, cudaStream_t stream) {
    MISO_REQUIRE(a.in_layout != LAYOUT_PLANES || a.in_ctot % 8 == 0, "conv: plane layout needs in_ctot %% 8 == 0");
    MISO_REQUIRE(a.out_layout != LAYOUT_PLANES || (a.out_ctot % 8 == 0 && a.out_coff % 4 == 0),
                 "conv: plane layout needs out_ctot %% 8 == 0 and out_coff %% 4 == 0");
    MISO_REQUIRE(a.cin % 4 == 0 && a.in_ctot % 4 == 0 && a.in_coff % 4 == 0,
                 "conv: input channel counts/offsets must be multiples of 4 (cin=%d ctot=%d coff=%d)", a.cin, a.in_ctot,
                 a.in_coff);
    MISO_REQUIRE(a.B > 0 && a.B <= 65535 && a.T > 0 && a.Fin > 0 && a.Fout > 0, "conv: bad shape");
    const int bn = conv_fp32_tile_n(a.cout);
    MISO_REQUIRE(a.cout_pad % bn == 0, "conv: cout_pad %d not a multiple of the N tile %d", a.cout_pad, bn);
    if (bn == 8) return launch<1>(a, stream);
    if (bn == 32) return launch<4>(a, stream);
    return launch<8>(a, stream);
}

}  // namespace miso
