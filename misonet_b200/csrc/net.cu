// MISO_1 / MISO_3 network body: packed-weight handle, workspace plan and the forward
// launch sequence (reference model.py:8-111, 282-395; layers model.py:401-632).
//
// Data layout in HBM: every conv-stack activation is a set of bf16 planes
// [B][hi|lo][Ctot/8][T][F][8] (conv.cuh, LAYOUT_PLANES).  A DenseBlock owns ONE buffer
// [x | y0 | y1 | y2 | y3] and conv k reads the channel prefix it needs (model.py:470-479
// concatenates instead); an encoder's final output is written straight into the skip half
// of the decoder buffer that will consume it (model.py:99 concatenates instead), and the
// next encoder reads it from there.  Tensors are stored as raw ELU outputs plus fp64
// (sum, sumsq) per (sample, channel); InstanceNorm is applied by whoever consumes them.
// The TCN state stays fp32 channels-last [B, T, C].
#include <algorithm>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

#include "bwd.cuh"
#include "conv.cuh"
#include "tcn.cuh"

namespace miso {

int conv_fp32_tile_n(int cout);

namespace {

constexpr float kInEps = 1e-5f;   // nn.InstanceNorm default (model.py:413,430,445,530)
constexpr float kGlnEps = 1e-8f;  // model.py:6,631

// ------------------------------------------------------------------ small kernels ----
__global__ void pack_conv_w_kernel(const float *__restrict__ w, float *__restrict__ packed, int cout, int cin, int taps,
                                   int cout_pad, int transposed) {
    int64_t total = (int64_t)taps * cin * cout_pad;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int co = (int)(i % cout_pad);
        int64_t r = i / cout_pad;
        int ci = (int)(r % cin);
        int tap = (int)(r / cin);
        float v = 0.f;
        if (co < cout) v = transposed ? w[((int64_t)ci * cout + co) * taps + tap] : w[((int64_t)co * cin + ci) * taps + tap];
        packed[i] = v;
    }
}

__global__ void pad_copy_kernel(const float *__restrict__ src, float *__restrict__ dst, int n, int n_pad) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_pad) dst[i] = i < n ? src[i] : 0.f;
}

// All parameters of a network in one launch (miso_net_set_params: a training step changes every one of the 268 parameters, and
// 268 separate launches from Python cost ~2 ms of host time per step).  blockIdx.y = parameter (of this launch's batch),
// blockIdx.x strides over its packed elements.
struct PackDesc {
    float *dst;
    int kind, cout, cin, taps, cout_pad;  // kind: 0 conv weight, 1 transposed conv weight, 2 bias (zero padded), 3 plain copy
    long long numel, packed_elems;
};
constexpr int kPackBatch = 128;
struct PackSrcs {
    const float *src[kPackBatch];
};
__global__ void __launch_bounds__(256) pack_params_kernel(const PackDesc *__restrict__ tab, const PackSrcs srcs, int first, int count) {
    const int j = blockIdx.y;
    if (j >= count) return;
    const PackDesc d = tab[first + j];
    const float *__restrict__ w = srcs.src[j];
    if (!w) return;
    const long long total = d.kind <= 1 ? (long long)d.taps * d.cin * d.cout_pad : (d.kind == 2 ? (long long)d.cout_pad : d.numel);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        float v;
        if (d.kind <= 1) {
            const int co = (int)(i % d.cout_pad);
            const long long r = i / d.cout_pad;
            const int ci = (int)(r % d.cin), tap = (int)(r / d.cin);
            v = 0.f;
            if (co < d.cout) v = d.kind == 1 ? w[((long long)ci * d.cout + co) * d.taps + tap] : w[((long long)co * d.cin + ci) * d.taps + tap];
        } else if (d.kind == 2) {
            v = i < d.cout ? w[i] : 0.f;
        } else {
            v = w[i];
        }
        d.dst[i] = v;
    }
}

__global__ void sentinel_kernel(double *sums, int B, int ctot, int coff, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B * n) {
        int b = i / n, c = i - b * n;
        stat_set(sums + ((size_t)b * ctot + coff + c) * 2 + 1, -1.0);
    }
}

// x0 = InstanceNorm(raw bottleneck output) materialised as the TCN state [B,T,C], plus the
// per-(b,c) statistics its first TemporalBlock needs (model.py:530).
constexpr int kTcnTile = 32;
__global__ void __launch_bounds__(128) tcn_prep_kernel(const __nv_bfloat16 *__restrict__ raw, int raw_ctot, int raw_coff,
                                                       int use_lo, const double *__restrict__ raw_sums, double inv_n, float eps,
                                                       float *__restrict__ S, double *__restrict__ s_sums, int T, int C) {
    const int c = blockIdx.y * 128 + threadIdx.x;
    const int b = blockIdx.z;
    if (c >= C) return;
    const double *rs = raw_sums + ((size_t)b * raw_ctot + raw_coff + c) * 2;
    const float2 af = affine_from_sums(stat_get(rs), stat_get(rs + 1), inv_n, (double)eps);
    const int t0 = blockIdx.x * kTcnTile;
    const int t1 = min(T, t0 + kTcnTile);
    float s = 0.f, q = 0.f;
    for (int t = t0; t < t1; ++t) {
        // planes [b][hi|lo][raw_ctot/8][T][F=1][8]
        const int ca = raw_coff + c;
        const size_t e = (size_t)b * 2 * raw_ctot * T + ((size_t)(ca >> 3) * T + t) * 8 + (ca & 7);
        float r = __bfloat162float(raw[e]);
        if (use_lo) r += __bfloat162float(raw[e + (size_t)raw_ctot * T]);
        float x = fmaf(r, af.x, af.y);
        S[((size_t)b * T + t) * C + c] = x;
        s += x;
        q += x * x;
    }
    stat_add(s_sums + ((size_t)b * C + c) * 2, (double)s);
    stat_add(s_sums + ((size_t)b * C + c) * 2 + 1, (double)q);
}

// First half of DepthwiseSeparableConv fused with the block's norm/activation prologue
// (model.py:530-531 / 538-539 then 556-558): v = ELU(IN1d(u)); y = dwconv_k3_dil(v);
// p = PReLU(y); accumulates the gLN statistics of p over (C,T) per sample.
// One block = one sample x kDwFrames frames x all channels; a thread handles 4 consecutive channels of a frame
// with float4 loads of the three taps (coalesced along c) and writes 8-byte bf16 hi / lo plane quads.
constexpr int kDwFrames = 16;
// ELU(alpha = 1): the tensor-core modes store the result as bf16 hi/lo planes (16-17 mantissa bits), so exp through
// ex2.approx (absolute error ~1e-7) is exact enough there; the fp32 mode keeps expm1f
__device__ __forceinline__ float dw_elu(float x, bool fast) { return x > 0.f ? x : (fast ? __expf(x) - 1.f : expm1f(x)); }
__global__ void __launch_bounds__(256) tcn_dw_kernel(const float *__restrict__ U, const double *__restrict__ u_sums,
                                                     double inv_n, float eps, const float *__restrict__ wdw,
                                                     const float *__restrict__ alpha, float *__restrict__ P,
                                                     double *__restrict__ g_sums, int T, int C, int dil, int out_planes,
                                                     int use_lo) {
    extern __shared__ float dw_sh[];  // [C] scale, [C] shift, [3][C] taps
    pdl_trigger();
    pdl_wait();  // the state and its statistics come from the previous pointwise conv, which still reads P
    float *sc = dw_sh, *sf = dw_sh + C, *wt = dw_sh + 2 * C;
    const int b = blockIdx.y;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const double *us = u_sums + ((size_t)b * C + c) * 2;
        const float2 af = affine_from_sums(stat_get(us), stat_get(us + 1), inv_n, (double)eps);
        sc[c] = af.x;
        sf[c] = af.y;
        wt[c] = wdw[c * 3 + 0];
        wt[C + c] = wdw[c * 3 + 1];
        wt[2 * C + c] = wdw[c * 3 + 2];
    }
    __syncthreads();
    const float al = alpha[0];
    const bool fast = out_planes != 0;
    const int t0 = blockIdx.x * kDwFrames;
    const int c4n = C >> 2;
    const float *ub = U + (size_t)b * T * C;
    float s = 0.f, q = 0.f;
    for (int i = threadIdx.x; i < kDwFrames * c4n; i += blockDim.x) {
        const int tt = i / c4n, c = (i - tt * c4n) * 4;
        const int t = t0 + tt;
        if (t >= T) break;
        const float4 a4 = *reinterpret_cast<const float4 *>(sc + c), b4 = *reinterpret_cast<const float4 *>(sf + c);
        auto act = [&](int tq) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (tq >= 0 && tq < T) {
                const float4 u = __ldg(reinterpret_cast<const float4 *>(ub + (size_t)tq * C + c));
                v.x = dw_elu(fmaf(u.x, a4.x, b4.x), fast);
                v.y = dw_elu(fmaf(u.y, a4.y, b4.y), fast);
                v.z = dw_elu(fmaf(u.z, a4.z, b4.z), fast);
                v.w = dw_elu(fmaf(u.w, a4.w, b4.w), fast);
            }
            return v;
        };
        const float4 vm = act(t - dil), v0 = act(t), vp = act(t + dil);
        const float4 w0 = *reinterpret_cast<const float4 *>(wt + c), w1 = *reinterpret_cast<const float4 *>(wt + C + c),
                     w2 = *reinterpret_cast<const float4 *>(wt + 2 * C + c);
        float y[4];
        y[0] = fmaf(w0.x, vm.x, fmaf(w1.x, v0.x, w2.x * vp.x));
        y[1] = fmaf(w0.y, vm.y, fmaf(w1.y, v0.y, w2.y * vp.y));
        y[2] = fmaf(w0.z, vm.z, fmaf(w1.z, v0.z, w2.z * vp.z));
        y[3] = fmaf(w0.w, vm.w, fmaf(w1.w, v0.w, w2.w * vp.w));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            y[k] = y[k] > 0.f ? y[k] : al * y[k];
            s += y[k];
            q = fmaf(y[k], y[k], q);
        }
        if (out_planes) {
            // bf16 planes [b][hi|lo][C/8][T][8]: the A operand layout of the tensor-core pointwise conv
            __nv_bfloat16 *pp = reinterpret_cast<__nv_bfloat16 *>(P) + (size_t)b * 2 * C * T + ((size_t)(c >> 3) * T + t) * 8 + (c & 7);
            const uint32_t h0 = pack_bf16x2(y[0], y[1]), h1 = pack_bf16x2(y[2], y[3]);
            *reinterpret_cast<uint2 *>(pp) = make_uint2(h0, h1);
            if (use_lo)
                *reinterpret_cast<uint2 *>(pp + (size_t)C * T) =
                    make_uint2(pack_bf16x2(y[0] - bf16_lo(h0), y[1] - bf16_hi(h0)), pack_bf16x2(y[2] - bf16_lo(h1), y[3] - bf16_hi(h1)));
        } else {
            *reinterpret_cast<float4 *>(P + ((size_t)b * T + t) * C + c) = make_float4(y[0], y[1], y[2], y[3]);
        }
    }
    __shared__ float red[2][8];
    s = warp_sum(s);
    q = warp_sum(q);
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = s;
        red[1][threadIdx.x >> 5] = q;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ds = 0.0, dq = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
            ds += (double)red[0][w];
            dq += (double)red[1][w];
        }
        stat_add(g_sums + (size_t)b * 2, ds);
        stat_add(g_sums + (size_t)b * 2 + 1, dq);
    }
}

struct ShiftList {
    int n;
    int s[16];
};

// Writes one pixel's channel vector (<= 16 channels) into the plane layout [hi|lo][cpad/8][TF][8].
__device__ __forceinline__ void store_pixel_planes(__nv_bfloat16 *base, int cpad, int TF, int p, const float *v) {
    for (int g8 = 0; g8 < cpad; g8 += 8) {
        float h[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) h[q] = bf16_round(v[g8 + q]);
        __nv_bfloat16 *dst = base + ((size_t)(g8 >> 3) * TF + p) * 8;
        *reinterpret_cast<uint4 *>(dst) =
            make_uint4(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]), pack_bf16x2(h[4], h[5]), pack_bf16x2(h[6], h[7]));
        *reinterpret_cast<uint4 *>(dst + (size_t)cpad * TF) =
            make_uint4(pack_bf16x2(v[g8] - h[0], v[g8 + 1] - h[1]), pack_bf16x2(v[g8 + 2] - h[2], v[g8 + 3] - h[3]),
                       pack_bf16x2(v[g8 + 4] - h[4], v[g8 + 5] - h[5]), pack_bf16x2(v[g8 + 6] - h[6], v[g8 + 7] - h[7]));
    }
}

__global__ void pack_miso1_kernel(const float2 *__restrict__ mix, __nv_bfloat16 *__restrict__ x, int B, int M, int TF,
                                  ShiftList sh) {
    // one thread per (b, pixel): reads M complex values (coalesced along f), writes the 2M (padded to a
    // multiple of 8) channels re(m0..), im(m0..) of every shifted copy as 16-byte plane pixels
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * TF) return;
    int b = (int)(i / TF);
    int p = (int)(i - (int64_t)b * TF);
    const int cpad = (2 * M + 7) & ~7;
    float2 v[8];
#pragma unroll
    for (int m = 0; m < 8; ++m)
        if (m < M) v[m] = mix[((size_t)b * M + m) * TF + p];
    for (int k = 0; k < sh.n; ++k) {
        float c[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) c[j] = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (j < M) {
                int m = j + sh.s[k];
                m -= (m >= M) ? M : 0;
                float2 z = v[0];
#pragma unroll
                for (int mm = 1; mm < 8; ++mm)
                    if (mm == m) z = v[mm];
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    if (q == j) c[q] = z.x;
                    if (q == M + j) c[q] = z.y;
                }
            }
        }
        store_pixel_planes(x + ((size_t)k * B + b) * 2 * cpad * TF, cpad, TF, p, c);
    }
}

__global__ void pack_miso3_kernel(const float2 *__restrict__ mix, const float2 *__restrict__ second,
                                  const float2 *__restrict__ third, __nv_bfloat16 *__restrict__ x, int B, int M, int TF) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * TF) return;
    int b = (int)(i / TF);
    int p = (int)(i - (int64_t)b * TF);
    const int C = M + 2;
    const int cpad = (2 * C + 7) & ~7;
    float c[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) c[j] = 0.f;
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        if (m < C) {
            float2 z;
            if (m < M)
                z = mix[((size_t)b * M + m) * TF + p];
            else if (m == M)
                z = second[(size_t)b * TF + p];
            else
                z = third[(size_t)b * TF + p];
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                if (q == m) c[q] = z.x;
                if (q == C + m) c[q] = z.y;
            }
        }
    }
    store_pixel_planes(x + (size_t)b * 2 * cpad * TF, cpad, TF, p, c);
}

__global__ void unpack_complex_kernel(const float *__restrict__ y, float2 *__restrict__ out, int B, int S, int TF) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * TF) return;
    int b = (int)(i / TF);
    int p = (int)(i - (int64_t)b * TF);
    const float *src = y + (size_t)i * (2 * S);
    for (int s = 0; s < S; ++s) out[((size_t)b * S + s) * TF + p] = make_float2(src[s], src[S + s]);
}

__global__ void tap_kernel(const __nv_bfloat16 *__restrict__ buf, int ctot, int coff, int C, int use_lo,
                           const double *__restrict__ sums, double inv_n, float eps, float *__restrict__ out, int B, int TF) {
    int64_t total = (int64_t)B * C * TF;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int p = (int)(i % TF);
        int64_t r = i / TF;
        int c = (int)(r % C);
        int b = (int)(r / C);
        float2 af = make_float2(1.f, 0.f);
        if (sums) {
            const double *s = sums + ((size_t)b * ctot + coff + c) * 2;
            af = affine_from_sums(stat_get(s), stat_get(s + 1), inv_n, (double)eps);
        }
        const int ca = coff + c;
        const size_t e = (size_t)b * 2 * ctot * TF + ((size_t)(ca >> 3) * TF + p) * 8 + (ca & 7);
        float raw = __bfloat162float(buf[e]);
        if (use_lo) raw += __bfloat162float(buf[e + (size_t)ctot * TF]);
        out[i] = fmaf(raw, af.x, af.y);
    }
}

// ------------------------------------------------------------------ handle -------------
enum ParamKind { P_CONV_W, P_DECONV_W, P_PW_W, P_BIAS, P_PLAIN };

struct Param {
    std::string key;
    ParamKind kind;
    int64_t numel;
    int cout, cin, taps, cout_pad;  // conv-like
    float *d;                       // packed device storage (owned by the arena)
    size_t packed_elems;
    bool loaded;
};

struct ConvDesc {
    int w = -1, b = -1;  // param indices
    int cin = 0, cout = 0, cout_pad = 0;
};

struct TcnHalf {
    int dw, alpha, gamma, beta, pw;
};

}  // namespace
}  // namespace miso

using namespace miso;

struct miso_net {
    int in_ch, out_ch, nb, R, X, C;
    std::vector<int> en, de;  // en[0]=in_ch ... en[nb]; de[0..nb-1], de[nb]=out_ch
    std::vector<Param> params;
    std::map<std::string, int> index;
    std::vector<ConvDesc> enc_conv;                 // [nb]
    std::vector<std::vector<ConvDesc>> enc_dense;   // [nb][5] (empty if no dense block)
    std::vector<std::vector<ConvDesc>> dec_dense;   // [nb][5]
    std::vector<ConvDesc> dec_deconv;               // [nb]
    std::vector<TcnHalf> tcn;                       // [R*X*2]
    float *arena = nullptr;
    PackDesc *d_pack = nullptr;  // device table of the parameters' packing descriptors (pack_params_kernel)
    int mode = 0;
    int n_loaded = 0;
    // CUDA-graph cache of the forward launch sequence (about 200 launches, 140 tensor-map encodes):
    // keyed by everything that is baked into the kernel arguments
    struct GraphEntry {
        const void *x;
        float *y;
        void *ws;
        int B, T, F, mode, prof;
        cudaGraphExec_t exec;
        uint64_t launches;
        std::vector<miso::ProfRec> recs;  // per-launch timing events baked into the graph (prof = 1)
    };
    std::vector<GraphEntry> graphs;
    // the training step's launch sequences (forward on the training plan: ~250 launches; backward: ~900) as graphs too:
    // kind 1 = miso_net_forward_train (p = x, q = y), kind 2 = miso_net_backward (p = dL/dy, q = flat gradients).  The
    // first call with a given key runs eagerly (one-time set-up such as function attributes must not happen inside a
    // capture), the second captures, later ones replay.
    struct TrainGraph {
        int kind;
        const void *x, *p, *q;
        void *ws;
        int B, T, F, mode;
        cudaGraphExec_t exec;  // null: seen once, not captured yet
        uint64_t launches;
    };
    std::vector<TrainGraph> tgraphs;
    cudaStream_t cap_stream = nullptr;
    int use_graph = 1;
    // gradient buckets of the backward pass, in the order in which they complete (decoders top-down, TCN, encoders): an
    // event per bucket lets the caller start that bucket's all-reduce while the rest of the backward is still running
    cudaEvent_t bucket_ev[8] = {};
};

namespace miso {
namespace {

bool dense_enc(int i) { return i < 5; }   // model.py:42
bool dense_dec(int j) { return j >= 2; }  // model.py:60

int add_param(miso_net *n, const std::string &key, ParamKind kind, int64_t numel, int cout, int cin, int taps) {
    Param p;
    p.key = key;
    p.kind = kind;
    p.numel = numel;
    p.cout = cout;
    p.cin = cin;
    p.taps = taps;
    p.cout_pad = 0;
    p.d = nullptr;
    p.loaded = false;
    if (kind == P_CONV_W || kind == P_DECONV_W || kind == P_PW_W) {
        int bn = conv_fp32_tile_n(cout);
        p.cout_pad = (cout + bn - 1) / bn * bn;
        p.packed_elems = (size_t)taps * cin * p.cout_pad;
    } else if (kind == P_BIAS) {
        int bn = conv_fp32_tile_n(cout);
        p.cout_pad = (cout + bn - 1) / bn * bn;
        p.packed_elems = p.cout_pad;
    } else {
        p.packed_elems = (size_t)numel;
    }
    n->params.push_back(p);
    n->index[key] = (int)n->params.size() - 1;
    return (int)n->params.size() - 1;
}

ConvDesc add_conv(miso_net *n, const std::string &prefix, int cin, int cout, bool transposed) {
    ConvDesc d;
    d.cin = cin;
    d.cout = cout;
    d.w = add_param(n, prefix + ".weight", transposed ? P_DECONV_W : P_CONV_W, (int64_t)cin * cout * 9, cout, cin, 9);
    d.b = add_param(n, prefix + ".bias", P_BIAS, cout, cout, 0, 0);
    d.cout_pad = n->params[d.w].cout_pad;
    return d;
}

std::vector<ConvDesc> add_dense(miso_net *n, const std::string &prefix, int c, int g1, int g2) {
    std::vector<ConvDesc> v;
    for (int k = 1; k <= 5; ++k)
        v.push_back(add_conv(n, prefix + ".conv" + std::to_string(k) + ".0", c + (k - 1) * g1, k < 5 ? g1 : g2, false));
    return v;
}

// spatial size of xs[i] (encoder outputs); returns false if the bottleneck is not F == 1
bool encoder_sizes(const miso_net *n, int F, std::vector<int> &Fx) {
    Fx.assign(n->nb, 0);
    int f = F;
    for (int i = 0; i < n->nb; ++i) {
        int stride = (i == 0 || i == n->nb - 1) ? 1 : 2;
        if (f < 3) return false;
        f = (f - 3) / stride + 1;
        Fx[i] = f;
    }
    if (f != 1) return false;
    // the decoder must reproduce the encoder sizes exactly (skip concatenation, model.py:99)
    int g = 1;
    for (int j = 0; j < n->nb; ++j) {
        if (g != Fx[n->nb - 1 - j]) return false;
        int stride = (j == 0 || j == n->nb - 1) ? 1 : 2;
        g = (g - 1) * stride + 3;
    }
    return g == F;
}

struct BufDesc {
    char *p = nullptr;  // bf16 planes [B][hi|lo][ctot/8][T][F][8]
    double *sums = nullptr;
    int ctot = 0, F = 0;
    size_t lo_off = 0;  // bytes from a sample's hi plane set to its lo plane set
    float *grad = nullptr;  // training plan: gradient buffer, fp32 channels-last [B][T][F][ctot]
};

struct Plan {
    std::vector<int> Fx;
    std::vector<BufDesc> E, D, Y;
    float *S = nullptr, *U = nullptr, *P = nullptr;
    // TCN state per block: block k reads Sk[k] / Uk[k].  Inference: all alias S / U (in-place residual stream);
    // training plan: one buffer per block, kept for the backward pass.
    std::vector<float *> Sk, Uk;
    // training plan only
    bool train = false;
    char *grad_base = nullptr;  // all gradient buffers, contiguous (one memset)
    size_t grad_bytes = 0;
    float *gS = nullptr, *gU = nullptr, *tY = nullptr, *tQ = nullptr, *tDQ = nullptr, *tDN = nullptr;
    double *bred = nullptr;  // reduction scratch of the norm backward kernels
    float *wT = nullptr;     // data-gradient (transposed) weights of the layer being processed
    char *dyP = nullptr;     // dL/dy of the layer being processed as bf16 hi/lo planes (tensor-core data gradient)
    size_t dyP_bytes = 0;
    char *dgP = nullptr;     // data gradient of a DenseBlock conv as planes (row-streaming kernel), added into the fp32 buffer
    size_t dgP_bytes = 0;
    float *wgP = nullptr;    // tcgen05 weight gradient: per-CTA raw accumulators (wgrad_tc.cu)
    size_t wgP_bytes = 0;
    __nv_bfloat16 *wg_ones = nullptr;  // ... and its constant-one pixels [T * max F][8]
    size_t wg_ones_pix = 0;
    void *tcn_wimgT = nullptr; // ... and their transposes for the backward (training plans)
    void *tcn_wimg = nullptr;  // tensor-core pointwise convs: weight images and W beta / W gamma vectors (tcn.cu)
    float *tcn_wvec = nullptr;
    std::vector<double *> sS, sU, g1, g2;
    double *stats_base = nullptr;
    size_t stats_bytes = 0;
    TcScratch scratch{};
    size_t total = 0;
};

size_t plane_bytes(int B, int ctot, int T, int F) { return (size_t)B * ctot * T * F * 4; }  // hi + lo bf16

bool make_plan(const miso_net *n, int B, int T, int F, char *base, Plan &pl, bool train = false) {
    if (!encoder_sizes(n, F, pl.Fx)) return false;
    const int nb = n->nb;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        char *p = base ? base + off : nullptr;
        off += align_up(bytes, 256);
        return p;
    };
    pl.E.assign(nb, BufDesc());
    pl.D.assign(nb, BufDesc());
    pl.Y.assign(nb, BufDesc());
    // statistics first (one memset clears them all)
    size_t stats_doubles = 0;
    auto stat_take = [&](size_t doubles) {
        size_t o = stats_doubles;
        stats_doubles += (doubles + 1) & ~(size_t)1;
        return o;
    };
    std::vector<size_t> oE(nb), oD(nb), oY(nb);
    for (int i = 0; i < nb; ++i) {
        if (dense_enc(i)) {
            pl.E[i].ctot = 5 * n->en[i + 1];
            pl.E[i].F = pl.Fx[i];
            oE[i] = stat_take((size_t)B * pl.E[i].ctot * 2);
        }
    }
    for (int j = 0; j < nb; ++j) {
        pl.D[j].ctot = (dense_dec(j) ? 6 : 2) * n->de[j];
        pl.D[j].F = pl.Fx[nb - 1 - j];
        oD[j] = stat_take((size_t)B * pl.D[j].ctot * 2);
        if (dense_dec(j)) {
            pl.Y[j].ctot = 2 * n->de[j];
            pl.Y[j].F = pl.D[j].F;
            oY[j] = stat_take((size_t)B * pl.Y[j].ctot * 2);
        }
    }
    const int nblk = n->R * n->X;
    std::vector<size_t> oS(nblk), oU(nblk), o1(nblk), o2(nblk);
    for (int k = 0; k < nblk; ++k) {
        oS[k] = stat_take((size_t)B * n->C * 2);
        oU[k] = stat_take((size_t)B * n->C * 2);
        o1[k] = stat_take((size_t)B * 2);
        o2[k] = stat_take((size_t)B * 2);
    }
    pl.stats_bytes = stats_doubles * sizeof(double);
    pl.stats_base = reinterpret_cast<double *>(take(pl.stats_bytes));
    auto sp = [&](size_t o) { return pl.stats_base ? pl.stats_base + o : nullptr; };
    for (int i = 0; i < nb; ++i)
        if (dense_enc(i)) {
            pl.E[i].sums = sp(oE[i]);
            pl.E[i].p = take(plane_bytes(B, pl.E[i].ctot, T, pl.E[i].F));
            pl.E[i].lo_off = (size_t)pl.E[i].ctot * T * pl.E[i].F * 2;
        }
    for (int j = 0; j < nb; ++j) {
        pl.D[j].sums = sp(oD[j]);
        pl.D[j].p = take(plane_bytes(B, pl.D[j].ctot, T, pl.D[j].F));
        pl.D[j].lo_off = (size_t)pl.D[j].ctot * T * pl.D[j].F * 2;
        if (dense_dec(j)) {
            pl.Y[j].sums = sp(oY[j]);
            pl.Y[j].p = take(plane_bytes(B, pl.Y[j].ctot, T, pl.Y[j].F));
            pl.Y[j].lo_off = (size_t)pl.Y[j].ctot * T * pl.Y[j].F * 2;
        }
    }
    size_t tcn_bytes = (size_t)B * T * n->C * sizeof(float);
    pl.S = reinterpret_cast<float *>(take(tcn_bytes));
    pl.U = reinterpret_cast<float *>(take(tcn_bytes));
    pl.P = reinterpret_cast<float *>(take(tcn_bytes));
    pl.train = train;
    {
        const int nblk_ = n->R * n->X;
        pl.Sk.assign(nblk_, pl.S);
        pl.Uk.assign(nblk_, pl.U);
        if (train) {
            for (int k = 1; k < nblk_; ++k) pl.Sk[k] = reinterpret_cast<float *>(take(tcn_bytes));
            for (int k = 1; k < nblk_; ++k) pl.Uk[k] = reinterpret_cast<float *>(take(tcn_bytes));
            // gradient buffers: same element counts as the activation buffers, fp32 channels-last
            const size_t g0 = off;
            auto gtake = [&](const BufDesc &d) { return reinterpret_cast<float *>(take((size_t)B * d.ctot * T * d.F * sizeof(float))); };
            for (int i = 0; i < nb; ++i)
                if (dense_enc(i)) pl.E[i].grad = gtake(pl.E[i]);
            for (int j = 0; j < nb; ++j) {
                pl.D[j].grad = gtake(pl.D[j]);
                if (dense_dec(j)) pl.Y[j].grad = gtake(pl.Y[j]);
            }
            pl.grad_base = base ? base + g0 : nullptr;
            pl.grad_bytes = off - g0;
            pl.gS = reinterpret_cast<float *>(take(tcn_bytes));
            pl.gU = reinterpret_cast<float *>(take(tcn_bytes));
            pl.tY = reinterpret_cast<float *>(take(tcn_bytes));
            pl.tQ = reinterpret_cast<float *>(take(tcn_bytes));
            pl.tDQ = reinterpret_cast<float *>(take(tcn_bytes));
            pl.tDN = reinterpret_cast<float *>(take(tcn_bytes));
            int maxc = n->C;
            size_t maxw = 0;
            for (auto &p : n->params) {
                if (p.kind == P_CONV_W || p.kind == P_DECONV_W || p.kind == P_PW_W) {
                    maxc = std::max(maxc, std::max(p.cin, p.cout));
                    const int bn = conv_fp32_tile_n(p.cin);
                    maxw = std::max(maxw, (size_t)p.taps * p.cout * ((p.cin + bn - 1) / bn * bn));
                }
            }
            pl.bred = reinterpret_cast<double *>(take(((size_t)B * maxc * 2 + 2 * (size_t)B) * sizeof(double)));
            pl.wT = reinterpret_cast<float *>(take(maxw * sizeof(float)));
            // largest layer output (channels x bins): the first dense conv block's buffers bound it
            size_t maxo = 0;
            for (int i = 0; i < nb; ++i) maxo = std::max(maxo, (size_t)n->en[i + 1] * pl.Fx[i]);
            for (int j = 0; j < nb; ++j) maxo = std::max(maxo, (size_t)2 * n->de[j] * pl.D[j].F);
            for (int j = 0; j + 1 < nb; ++j) maxo = std::max(maxo, (size_t)n->de[j + 1] * pl.D[j + 1].F);
            pl.dyP_bytes = (size_t)B * maxo * T * 4;
            pl.dyP = take(pl.dyP_bytes);
            // largest DenseBlock conv input (channels x bins): the fourth / fifth conv of a block reads 4 of its 5 (6) groups
            size_t maxi = 0;
            for (int i = 0; i < nb; ++i)
                if (dense_enc(i)) maxi = std::max(maxi, (size_t)pl.E[i].ctot * pl.E[i].F);
            for (int j = 0; j < nb; ++j)
                if (dense_dec(j)) maxi = std::max(maxi, (size_t)pl.D[j].ctot * pl.D[j].F);
            pl.dgP_bytes = (size_t)B * maxi * T * 4;
            pl.dgP = take(pl.dgP_bytes);
            pl.wgP_bytes = wgrad_tc_partial_bytes(B);
            pl.wgP = reinterpret_cast<float *>(take(pl.wgP_bytes));
            pl.wg_ones_pix = (size_t)T * F;
            pl.wg_ones = reinterpret_cast<__nv_bfloat16 *>(take(pl.wg_ones_pix * 16));
        }
    }
    if (tcn_pw_eligible(n->C)) {
        size_t wi, wv;
        tcn_pw_scratch_need(n->C, 2 * nblk, &wi, &wv);
        pl.tcn_wimg = take(wi);
        pl.tcn_wvec = reinterpret_cast<float *>(take(wv));
        if (train) pl.tcn_wimgT = take(wi);  // transposed images: the pointwise convs' data gradients as tcn_pw GEMMs
    }
    pl.sS.resize(nblk);
    pl.sU.resize(nblk);
    pl.g1.resize(nblk);
    pl.g2.resize(nblk);
    for (int k = 0; k < nblk; ++k) {
        pl.sS[k] = sp(oS[k]);
        pl.sU[k] = sp(oU[k]);
        pl.g1[k] = sp(o1[k]);
        pl.g2[k] = sp(o2[k]);
    }
    pl.total = off;
    return true;
}

// tensor-core scratch (per-sample weight images, border-bias tables) goes after the activations
void plan_scratch(Plan &pl, char *base, size_t need_w, size_t need_b) {
    size_t off = pl.total;
    pl.scratch.wimg = base ? base + off : nullptr;
    pl.scratch.wimg_bytes = need_w;
    off += align_up(need_w, 256);
    pl.scratch.btab = base ? reinterpret_cast<float *>(base + off) : nullptr;
    pl.scratch.btab_bytes = need_b;
    off += align_up(need_b, 256);
    pl.total = off;
}

// where encoder i's block output xs[i] lives: the skip half of decoder (nb-1-i)'s buffer
struct ViewRef {
    const BufDesc *buf;
    int coff, c;
};
ViewRef xs_view(const miso_net *n, const Plan &pl, int i) {
    int j = n->nb - 1 - i;
    return ViewRef{&pl.D[j], n->de[j], n->en[i + 1]};
}

// A conv layer of the forward launch sequence as the backward pass needs it (recorded by the same walker that
// launches the forward, so the two can never disagree about views, strides or statistics).
struct ConvRec {
    int kind;  // 0: conv layer, 1: the TCN sits here
    ConvArgs a;
    const ConvDesc *cd;
    float *in_grad;   // gradient buffer of the input buffer (null: the network input, no data gradient)
    float *out_grad;  // gradient buffer of the output buffer (null: the network output, gradient supplied by the caller)
    bool in_grad_needed;  // false only for the first layer (pointer-independent: plans are also built without a base)
};

// One pass over the layer list.  dry = true only sizes the tensor-core scratch (no launches).
struct Walker {
    const miso_net *n;
    const Plan &pl;
    int B, T, F;
    cudaStream_t st;
    bool dry;
    size_t need_w = 0, need_b = 0;
    std::vector<ConvRec> *record = nullptr;  // non-null: list the layers instead of launching them


    ConvArgs make_args(const ConvDesc &cd, bool transposed, const BufDesc *inb, const void *in_raw, int in_ctot, int in_coff, int Fin,
                       const double *in_sums, const BufDesc *outb, float *out_raw, int out_ctot, int out_coff, int Fout, double *out_sums,
                       int stride_f, int pad_f, bool elu) const {
        ConvArgs a{};
        a.in = inb ? inb->p : in_raw;
        a.in_layout = LAYOUT_PLANES;
        a.in_lo_off = (size_t)in_ctot * T * Fin * 2;
        a.w = n->params[cd.w].d;
        a.bias = n->params[cd.b].d;
        a.out = outb ? (void *)outb->p : (void *)out_raw;
        a.out_layout = outb ? LAYOUT_PLANES : LAYOUT_CL_F32;
        a.out_lo_off = (size_t)out_ctot * T * Fout * 2;
        a.use_lo = n->mode == 2 ? 0 : 1;
        a.resid = nullptr;
        a.in_sums = in_sums;
        a.out_sums = out_sums;
        a.B = B;
        a.T = T;
        a.Fin = Fin;
        a.Fout = Fout;
        a.in_ctot = in_ctot;
        a.in_coff = in_coff;
        a.cin = cd.cin;
        a.out_ctot = out_ctot;
        a.out_coff = out_coff;
        a.cout = cd.cout;
        a.cout_pad = cd.cout_pad;
        a.KT = 3;
        a.KF = 3;
        a.stride_f = stride_f;
        a.pad_t = 1;
        a.pad_f = pad_f;
        a.transposed = transposed ? 1 : 0;
        a.norm_mode = in_sums ? NORM_IN : NORM_NONE;
        a.norm_eps = kInEps;
        a.norm_inv_n = 1.0 / ((double)T * Fin);
        a.elu = elu ? 1 : 0;
        return a;
    }

    int dispatch(const ConvArgs &a, const ConvDesc &cd, const BufDesc *inb, const BufDesc *outb) {
        if (record) {
            record->push_back(ConvRec{0, a, &cd, inb ? inb->grad : nullptr, outb ? outb->grad : nullptr, inb != nullptr});
            return MISO_OK;
        }
        const bool tc = conv_tc_eligible(a);
        if (dry) {
            if (tc) {
                size_t w, bt;
                conv_tc_scratch_need(a, 3, &w, &bt);
                need_w = std::max(need_w, w);
                need_b = std::max(need_b, bt);
            }
            return MISO_OK;
        }
        if (n->mode != 0 && tc) return launch_conv_tc(a, n->mode == 1 ? 3 : 1, pl.scratch, st);
        return launch_conv_fp32(a, st);
    }

    int conv(const ConvDesc &cd, bool transposed, const BufDesc *inb, const void *in_raw, int in_ctot, int in_coff, int Fin,
             const double *in_sums, const BufDesc *outb, float *out_raw, int out_ctot, int out_coff, int Fout, double *out_sums,
             int stride_f, int pad_f, bool elu) {
        return dispatch(make_args(cd, transposed, inb, in_raw, in_ctot, in_coff, Fin, in_sums, outb, out_raw, out_ctot, out_coff, Fout,
                                  out_sums, stride_f, pad_f, elu),
                        cd, inb, outb);
    }

    // The five convs of a DenseBlock (model.py:437-482) over the block buffer [x | y0 | y1 | y2 | y3]: conv k reads the first
    // c + (k - 1) g1 channels and writes g1 channels behind them; conv 5 writes its g2 channels to out5 at out5_coff.
    int dense(const std::vector<ConvDesc> &cds, const BufDesc &buf, int c, int g1, const BufDesc &out5, int out5_coff) {
        ConvArgs a[5];
        for (int k = 1; k <= 5; ++k) {
            if (k < 5)
                a[k - 1] = make_args(cds[k - 1], false, &buf, nullptr, buf.ctot, 0, buf.F, buf.sums, &buf, nullptr, buf.ctot, c + (k - 1) * g1, buf.F,
                                     buf.sums, 1, 1, true);
            else
                a[k - 1] = make_args(cds[k - 1], false, &buf, nullptr, buf.ctot, 0, buf.F, buf.sums, &out5, nullptr, out5.ctot, out5_coff, buf.F,
                                     out5.sums, 1, 1, true);
        }
        // block buffers are always consumed through their producer's InstanceNorm; a plan built without a base address
        // (size queries) has null statistics pointers and must still take the same decisions
        for (int k = 0; k < 5; ++k) a[k].norm_mode = NORM_IN;
        for (int k = 1; k <= 5; ++k) {
            const int rc = dispatch(a[k - 1], cds[k - 1], &buf, k < 5 ? &buf : &out5);
            if (rc) return rc;
        }
        return MISO_OK;
    }

    int run(const void *d_x, float *d_y);
};

int Walker::run(const void *d_x, float *d_y) {
    const int nb = n->nb, C = n->C;
    int rc;
    const int in_pad = (n->in_ch + 7) & ~7;
    // ---------------- encoders (model.py:40-53, 83-86) ----------------
    for (int i = 0; i < nb; ++i) {
        const BufDesc *inb = nullptr;
        const double *in_sums;
        int in_ctot, in_coff, Fin;
        if (i == 0) {
            in_sums = nullptr;
            in_ctot = in_pad;
            in_coff = 0;
            Fin = F;
        } else {
            ViewRef v = xs_view(n, pl, i - 1);
            inb = v.buf;
            in_sums = v.buf->sums;
            in_ctot = v.buf->ctot;
            in_coff = v.coff;
            Fin = v.buf->F;
        }
        const int stride = (i == 0 || i == nb - 1) ? 1 : 2;
        ViewRef xo = xs_view(n, pl, i);
        if (dense_enc(i)) {
            const BufDesc &e = pl.E[i];
            rc = conv(n->enc_conv[i], false, inb, d_x, in_ctot, in_coff, Fin, in_sums, &e, nullptr, e.ctot, 0, e.F,
                      i == 0 ? nullptr : e.sums, stride, 0, i != 0);
            if (rc) return rc;
            const int c = n->en[i + 1];
            rc = dense(n->enc_dense[i], e, c, c, *xo.buf, xo.coff);
            if (rc) return rc;
        } else {
            rc = conv(n->enc_conv[i], false, inb, d_x, in_ctot, in_coff, Fin, in_sums, xo.buf, nullptr, xo.buf->ctot, xo.coff,
                      xo.buf->F, xo.buf->sums, stride, 0, true);
            if (rc) return rc;
        }
    }

    // ---------------- TCN (model.py:486-567) ----------------
    if (record) {
        record->push_back(ConvRec{1, ConvArgs{}, nullptr, nullptr, nullptr, false});
    } else {
        const BufDesc &d0 = pl.D[0];
        const double inv_T = 1.0 / (double)T;
        const int use_lo = n->mode == 2 ? 0 : 1;
        dim3 grid(ceil_div(T, kTcnTile), ceil_div(C, 128), B);
        if (!dry) {
            tcn_prep_kernel<<<grid, 128, 0, st>>>(reinterpret_cast<const __nv_bfloat16 *>(d0.p), d0.ctot, C, use_lo, d0.sums,
                                                  inv_T, kInEps, pl.Sk[0], pl.sS[0], T, C);
            MISO_LAUNCHED("tcn_prep_kernel");
        }
        const int nblk = n->R * n->X;
        const int bn = conv_fp32_tile_n(C);
        const int cpad = (C + bn - 1) / bn * bn;
        const bool use_pw = n->mode != 0 && tcn_pw_eligible(C) && 2 * nblk <= kTcnMaxPw;
        if (!dry && use_pw) {
            TcnPwTable tab{};
            for (int i = 0; i < 2 * nblk; ++i) {
                tab.w[i] = n->params[n->tcn[i].pw].d;
                tab.gamma[i] = n->params[n->tcn[i].gamma].d;
                tab.beta[i] = n->params[n->tcn[i].beta].d;
            }
            rc = launch_tcn_wprep(tab, 2 * nblk, C, cpad, n->mode == 1 ? 3 : 1, pl.tcn_wimg, pl.tcn_wvec, st);
            if (rc) return rc;
        }
        // tensor-core modes: the whole TCN as one cluster-per-sample launch (tcn.cu, tcn_fused_kernel)
        const bool fused = !dry && use_pw && tcn_fused_eligible(C, T, 2 * nblk);
        if (fused) {
            TcnFusedArgs fa{};
            fa.nhalf = 2 * nblk;
            fa.B = B;
            fa.T = T;
            fa.C = C;
            fa.wimg = pl.tcn_wimg;
            fa.wvec = pl.tcn_wvec;
            fa.in_inv_n = inv_T;
            fa.gln_inv_n = 1.0 / ((double)C * T);
            fa.in_eps = kInEps;
            fa.gln_eps = kGlnEps;
            fa.out_ctot = d0.ctot;
            fa.use_lo = use_lo;
            fa.out_lo_off = d0.lo_off;
            for (int k = 0; k < nblk; ++k)
                for (int half = 0; half < 2; ++half) {
                    const TcnHalf &h = n->tcn[k * 2 + half];
                    TcnFusedHalf &fh = fa.h[k * 2 + half];
                    fh.u = half == 0 ? pl.Sk[k] : pl.Uk[k];
                    fh.u_sums = half == 0 ? pl.sS[k] : pl.sU[k];
                    fh.g_sums = half == 0 ? pl.g1[k] : pl.g2[k];
                    fh.dw = n->params[h.dw].d;
                    fh.alpha = n->params[h.alpha].d;
                    fh.dil = 1 << (k % n->X);
                    if (half == 0) {
                        fh.out = pl.Uk[k];
                        fh.out_sums = pl.sU[k];
                        fh.resid = nullptr;
                    } else if (k + 1 < nblk) {
                        fh.out = pl.Sk[k + 1];
                        fh.out_sums = pl.sS[k + 1];
                        fh.resid = pl.Sk[k];
                    } else {
                        fh.out = d0.p;  // TCN output = first half of decoder 0's input, consumed raw
                        fh.out_planes = 1;
                        fh.out_sums = nullptr;
                        fh.resid = pl.Sk[k];
                    }
                }
            rc = launch_tcn_fused(fa, n->mode == 1 ? 3 : 1, st);
            if (rc) return rc;
        }
        for (int k = 0; k < nblk && !fused; ++k) {
            const int dil = 1 << (k % n->X);
            for (int half = 0; half < 2; ++half) {
                const TcnHalf &h = n->tcn[k * 2 + half];
                const float *u = half == 0 ? pl.Sk[k] : pl.Uk[k];
                const double *us = half == 0 ? pl.sS[k] : pl.sU[k];
                double *gs = half == 0 ? pl.g1[k] : pl.g2[k];
                const bool tc_pw = (dry || n->mode != 0) && C % 8 == 0;
                if (!dry) {
                    MISO_CUDA(launch_pdl(tcn_dw_kernel, dim3(ceil_div(T, kDwFrames), B), dim3(256), 5 * C * sizeof(float), st, u, us, inv_T,
                                         kInEps, (const float *)n->params[h.dw].d, (const float *)n->params[h.alpha].d, pl.P, gs, T, C, dil,
                                         tc_pw ? 1 : 0, use_lo));
                    MISO_LAUNCHED("tcn_dw_kernel");
                }
                ConvArgs a{};
                a.in = pl.P;
                a.in_layout = tc_pw ? LAYOUT_PLANES : LAYOUT_CL_F32;
                a.in_lo_off = (size_t)C * T * 2;
                a.out_layout = LAYOUT_CL_F32;
                a.use_lo = use_lo;
                a.w = n->params[h.pw].d;
                a.bias = nullptr;
                a.in_sums = gs;
                a.gamma = n->params[h.gamma].d;
                a.beta = n->params[h.beta].d;
                a.B = B;
                a.T = T;
                a.Fin = 1;
                a.Fout = 1;
                a.in_ctot = C;
                a.in_coff = 0;
                a.cin = C;
                a.cout = C;
                a.cout_pad = cpad;
                a.KT = 1;
                a.KF = 1;
                a.stride_f = 1;
                a.pad_t = 0;
                a.pad_f = 0;
                a.transposed = 0;
                a.norm_mode = NORM_GLN;
                a.norm_eps = kGlnEps;
                a.norm_inv_n = 1.0 / ((double)C * T);
                a.elu = 0;
                if (half == 0) {
                    a.out = pl.Uk[k];
                    a.out_ctot = C;
                    a.out_coff = 0;
                    a.out_sums = pl.sU[k];
                    a.resid = nullptr;
                } else {
                    a.resid = pl.Sk[k];
                    a.resid_ctot = C;
                    a.resid_coff = 0;
                    if (k + 1 < nblk) {
                        a.out = pl.Sk[k + 1];  // inference: in-place residual update (each element is read once by its own writer)
                        a.out_ctot = C;
                        a.out_coff = 0;
                        a.out_sums = pl.sS[k + 1];
                    } else {
                        a.out = d0.p;  // TCN output = first half of decoder 0's input, consumed raw
                        a.out_layout = LAYOUT_PLANES;
                        a.out_lo_off = d0.lo_off;
                        a.out_ctot = d0.ctot;
                        a.out_coff = 0;
                        a.out_sums = nullptr;
                    }
                }
                if (dry) {
                    if (conv_tc_eligible(a)) {
                        size_t w, bt;
                        conv_tc_scratch_need(a, 3, &w, &bt);
                        need_w = std::max(need_w, w);
                        need_b = std::max(need_b, bt);
                    }
                    continue;
                }
                if (use_pw && (a.out_layout == LAYOUT_CL_F32 || (a.out_layout == LAYOUT_PLANES && a.out_coff == 0 && !a.out_sums))) {
                    TcnPwArgs p{};
                    p.out_planes = a.out_layout == LAYOUT_PLANES ? 1 : 0;
                    p.out_ctot = a.out_ctot;
                    p.use_lo = use_lo;
                    p.out_lo_off = a.out_lo_off;
                    p.planes = pl.P;
                    p.lo_off = a.in_lo_off;
                    p.wimg = pl.tcn_wimg;
                    p.wvec = pl.tcn_wvec;
                    p.index = k * 2 + half;
                    p.gln_sums = gs;
                    p.gln_inv_n = a.norm_inv_n;
                    p.gln_eps = a.norm_eps;
                    p.out = reinterpret_cast<float *>(a.out);
                    p.resid = a.resid;
                    p.out_sums = a.out_sums;
                    p.B = B;
                    p.T = T;
                    p.C = C;
                    rc = launch_tcn_pw(p, n->mode == 1 ? 3 : 1, st);
                } else if (tc_pw && conv_tc_eligible(a))
                    rc = launch_conv_tc(a, n->mode == 1 ? 3 : 1, pl.scratch, st);
                else
                    rc = launch_conv_fp32(a, st);  // reads either layout
                if (rc) return rc;
            }
        }
    }

    // ---------------- decoders (model.py:55-73, 97-100) ----------------
    for (int j = 0; j < nb; ++j) {
        const BufDesc &d = pl.D[j];
        const int stride = (j == 0 || j == nb - 1) ? 1 : 2;
        const bool last = j == nb - 1;
        const BufDesc *outb = last ? nullptr : &pl.D[j + 1];
        const int out_ctot = last ? n->out_ch : pl.D[j + 1].ctot;
        const int Fout = last ? F : pl.D[j + 1].F;
        double *out_sums = last ? nullptr : pl.D[j + 1].sums;
        if (dense_dec(j)) {
            const int c = 2 * n->de[j], g1 = n->de[j];
            const BufDesc &y = pl.Y[j];
            rc = dense(n->dec_dense[j], d, c, g1, y, 0);
            if (rc) return rc;
            rc = conv(n->dec_deconv[j], true, &y, nullptr, y.ctot, 0, y.F, y.sums, outb, d_y, out_ctot, 0, Fout, out_sums, stride, 0,
                      !last);
        } else {
            rc = conv(n->dec_deconv[j], true, &d, nullptr, d.ctot, 0, d.F, d.sums, outb, d_y, out_ctot, 0, Fout, out_sums, stride, 0,
                      true);
        }
        if (rc) return rc;
    }
    return MISO_OK;
}

int enqueue_forward(miso_net *net, const Plan &pl, const void *d_x, float *d_y, int B, int T, int F, cudaStream_t st);
bool capturing(cudaStream_t st) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    return cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs == cudaStreamCaptureStatusActive;
}

// Data gradient of a recorded layer as a forward conv over dL/dy (bf16 hi/lo planes in pl.dyP) with the transposed
// weights in pl.wT, accumulated into the input buffer's gradient (resid == out).  Every case maps onto a configuration
// the forward itself uses: a stride-1 pad-(1,1) conv's gradient is the same conv with the taps reversed (*flip = 1),
// a strided conv's gradient is the transposed conv of the same stride and vice versa.
ConvArgs dgrad_tc_args(const ConvArgs &f, const Plan &pl, int B, int T, float *din, int *flip) {
    const int bn = conv_fp32_tile_n(f.cin);
    ConvArgs a{};
    a.in = pl.dyP;
    a.in_layout = LAYOUT_PLANES;
    a.in_lo_off = (size_t)f.cout * T * f.Fout * 2;
    a.w = pl.wT;
    a.bias = nullptr;
    a.out = din;
    a.out_layout = LAYOUT_CL_F32;
    a.resid = din;
    a.resid_ctot = f.in_ctot;
    a.resid_coff = f.in_coff;
    a.B = B;
    a.T = T;
    a.Fin = f.Fout;
    a.Fout = f.Fin;
    a.in_ctot = f.cout;
    a.in_coff = 0;
    a.cin = f.cout;
    a.out_ctot = f.in_ctot;
    a.out_coff = f.in_coff;
    a.cout = f.cin;
    a.cout_pad = (f.cin + bn - 1) / bn * bn;
    a.KT = f.KT;
    a.KF = f.KF;
    a.stride_f = f.stride_f;
    a.pad_t = f.pad_t;
    a.pad_f = f.pad_f;
    *flip = (!f.transposed && f.stride_f == 1 && f.pad_f == 1) ? 1 : 0;
    a.transposed = (*flip || f.KT == 1) ? 0 : (f.transposed ? 0 : 1);  // a 1x1 conv's gradient is a plain 1x1 conv
    a.norm_mode = NORM_NONE;
    a.norm_eps = kInEps;
    a.norm_inv_n = 1.0;
    a.elu = 0;
    a.use_lo = 1;
    return a;
}
// forward geometry of a TCN pointwise conv (model.py:560), as far as the gradient kernels need it
ConvArgs tcn_pw_fwd_args(int C, int B, int T) {
    ConvArgs f{};
    f.in_layout = LAYOUT_CL_F32;
    f.out_layout = LAYOUT_CL_F32;
    f.use_lo = 1;
    f.B = B;
    f.T = T;
    f.Fin = 1;
    f.Fout = 1;
    f.in_ctot = C;
    f.in_coff = 0;
    f.cin = C;
    f.out_ctot = C;
    f.out_coff = 0;
    f.cout = C;
    f.KT = 1;
    f.KF = 1;
    f.stride_f = 1;
    f.pad_t = 0;
    f.pad_f = 0;
    f.transposed = 0;
    return f;
}
bool dgrad_tc_ok(const ConvArgs &f, const ConvArgs &d, const Plan &pl, int B, int T) {
    return f.cout % 8 == 0 && (size_t)B * f.cout * T * f.Fout * 4 <= pl.dyP_bytes && conv_tc_eligible(d);
}

// DenseBlock convs (stride 1, pad (1,1)): the tap-reversed gradient conv is exactly the forward's row-streaming
// configuration if it writes planes, so it runs on conv_rs in chunks of <= 64 output channels into pl.dgP (the kernel's
// N limit; the small dL/dy is re-read per chunk) and planes_accumulate_kernel adds the result into the fp32 gradient.
constexpr int kRsChunk = 64;
// DenseBlock data gradients on the row-streaming kernel in chunks of <= 64 output channels (its N limit); the kernel's
// epilogue adds into the fp32 channels-last gradient buffer itself.  With one bin per lane that read-modify-write was slower
// than a planes scratch plus planes_accumulate_kernel (52.1 against 50.1 ms per training step, 8 utterances PAPER: every
// 16-byte access of a warp touched 32 lines); turned through shared memory so that four lanes cover a bin's 64 bytes it is
// faster (43.05 against 44.07 ms).  MISO_DGRAD_RS_CL=0 selects the scratch path.
bool dgrad_rs_cl() {
    static const bool on = !(getenv("MISO_DGRAD_RS_CL") && atoi(getenv("MISO_DGRAD_RS_CL")) == 0);
    return on;
}
ConvArgs dgrad_rs_chunk(const ConvArgs &d, const Plan &pl, int T, int c0, int width) {
    ConvArgs a = d;
    a.cout = std::min(width, d.cout - c0);
    a.w = d.w ? d.w + c0 : nullptr;
    if (dgrad_rs_cl()) {
        a.out_coff = d.out_coff + c0;
        a.resid_coff = d.resid_coff + c0;
    } else {
        const int ctot8 = (d.cout + 7) & ~7;
        a.out = pl.dgP;
        a.out_layout = LAYOUT_PLANES;
        a.out_ctot = ctot8;
        a.out_coff = c0;
        a.out_lo_off = (size_t)ctot8 * T * d.Fout * 2;
        a.resid = nullptr;
    }
    return a;
}
bool dgrad_rs_ok(const ConvArgs &d, int flip, const Plan &pl, int B, int T) {
    if (!flip || d.cout % 8) return false;
    if (dgrad_rs_cl()) return conv_rs_eligible(dgrad_rs_chunk(d, pl, T, 0, d.cout), 3);
    if ((size_t)B * ((d.cout + 7) & ~7) * T * d.Fout * 4 > pl.dgP_bytes) return false;
    for (int c0 = 0; c0 < d.cout; c0 += kRsChunk)
        if (!conv_rs_eligible(dgrad_rs_chunk(d, pl, T, c0, kRsChunk), 3)) return false;
    return true;
}

bool full_plan(const miso_net *n, int B, int T, int F, char *base, Plan &pl, bool train = false) {
    if (!make_plan(n, B, T, F, base, pl, train)) return false;
    Walker w{n, pl, B, T, F, nullptr, true};
    w.run(nullptr, nullptr);
    if (train) {  // the tensor-core data-gradient convs share the per-sample operand scratch
        std::vector<ConvRec> recs;
        Walker r{n, pl, B, T, F, nullptr, false};
        r.record = &recs;
        r.run(nullptr, nullptr);
        for (auto &rc : recs) {
            if (rc.kind != 0 || !rc.in_grad_needed) continue;
            int flip;
            ConvArgs d = dgrad_tc_args(rc.a, pl, B, T, nullptr, &flip);
            if (!dgrad_tc_ok(rc.a, d, pl, B, T)) continue;
            size_t ww, bb;
            conv_tc_scratch_need(d, 3, &ww, &bb);
            w.need_w = std::max(w.need_w, ww);
            w.need_b = std::max(w.need_b, bb);
            if (dgrad_rs_ok(d, flip, pl, B, T))
                for (int c0 = 0; c0 < d.cout; c0 += kRsChunk) {
                    conv_tc_scratch_need(dgrad_rs_chunk(d, pl, T, dgrad_rs_cl() ? 0 : c0, dgrad_rs_cl() ? d.cout : kRsChunk), 3, &ww, &bb);
                    w.need_w = std::max(w.need_w, ww);
                    w.need_b = std::max(w.need_b, bb);
                }
        }
        {
            int flip;
            ConvArgs f = tcn_pw_fwd_args(n->C, B, T);
            ConvArgs d = dgrad_tc_args(f, pl, B, T, nullptr, &flip);
            if (dgrad_tc_ok(f, d, pl, B, T)) {
                size_t ww, bb;
                conv_tc_scratch_need(d, 3, &ww, &bb);
                w.need_w = std::max(w.need_w, ww);
                w.need_b = std::max(w.need_b, bb);
            }
        }
    }
    plan_scratch(pl, base, w.need_w, w.need_b);
    return true;
}

}  // namespace
}  // namespace miso

// =========================================================================== C ABI =====
extern "C" {

int miso_net_create(miso_net_t **out, int in_ch, int out_ch, int num_bottleneck, const int *en_channels,
                    const int *de_channels, int tcn_repeats, int tcn_blocks) {
    MISO_REQUIRE(out && en_channels && de_channels, "miso_net_create: null argument");
    MISO_REQUIRE(num_bottleneck >= 6 && num_bottleneck <= 12, "miso_net_create: num_bottleneck %d unsupported (6..12)",
                 num_bottleneck);
    MISO_REQUIRE(in_ch > 0 && in_ch % 4 == 0, "miso_net_create: in_ch=%d must be a positive multiple of 4", in_ch);
    MISO_REQUIRE(out_ch > 0 && out_ch % 2 == 0, "miso_net_create: out_ch=%d must be even", out_ch);
    MISO_REQUIRE(tcn_repeats > 0 && tcn_blocks > 0 && tcn_blocks <= 12, "miso_net_create: bad TCN shape");
    miso_net *n = new miso_net();
    n->in_ch = in_ch;
    n->out_ch = out_ch;
    n->nb = num_bottleneck;
    n->R = tcn_repeats;
    n->X = tcn_blocks;
    n->en.push_back(in_ch);
    for (int i = 0; i < n->nb; ++i) n->en.push_back(en_channels[i]);
    for (int i = 0; i < n->nb; ++i) n->de.push_back(de_channels[i]);
    n->de.push_back(out_ch);
    n->C = n->en[n->nb];
    for (int i = 1; i <= n->nb; ++i)
        if (n->en[i] <= 0 || n->en[i] % 4) {
            set_error("miso_net_create: encoder channels must be positive multiples of 4");
            delete n;
            return MISO_E_ARG;
        }
    for (int j = 0; j < n->nb; ++j)
        if (n->de[j] != n->en[n->nb - j]) {
            // the skip connection concatenates xs[nb-1-j] (en[nb-j] channels) with a de[j]-channel
            // tensor and the layer is declared with 2*de[j] inputs (model.py:35, 99)
            set_error("miso_net_create: de_channels[%d]=%d must equal en_channels[%d]=%d", j, n->de[j], n->nb - 1 - j,
                      n->en[n->nb - j]);
            delete n;
            return MISO_E_ARG;
        }
    const int nb = n->nb;
    n->enc_conv.resize(nb);
    n->enc_dense.resize(nb);
    n->dec_dense.resize(nb);
    n->dec_deconv.resize(nb);
    // key order follows the reference's module registration order: encoders, decoders, TCN
    for (int i = 0; i < nb; ++i) {
        std::string p = "encoders." + std::to_string(i);
        n->enc_conv[i] = add_conv(n, p + (i == 0 ? ".0.conv2d" : ".0.net.0"), n->en[i], n->en[i + 1], false);
        if (dense_enc(i)) n->enc_dense[i] = add_dense(n, p + ".1", n->en[i + 1], n->en[i + 1], n->en[i + 1]);
    }
    for (int j = 0; j < nb; ++j) {
        std::string p = "decoders." + std::to_string(j);
        int cin = 2 * n->de[j], cout = n->de[j + 1];
        if (dense_dec(j)) {
            n->dec_dense[j] = add_dense(n, p + ".0", cin, cin / 2, cin);
            n->dec_deconv[j] = add_conv(n, p + (j == nb - 1 ? ".1.deconv2d" : ".1.net.0"), cin, cout, true);
        } else {
            n->dec_deconv[j] = add_conv(n, p + ".0.net.0", cin, cout, true);
        }
    }
    const int C = n->C;
    for (int r = 0; r < n->R; ++r)
        for (int x = 0; x < n->X; ++x)
            for (int half : {2, 5}) {
                std::string p = "TCN.temporal_conv_net." + std::to_string(r) + "." + std::to_string(x) + ".net." +
                                std::to_string(half) + ".net";
                TcnHalf h;
                h.dw = add_param(n, p + ".0.weight", P_PLAIN, (int64_t)C * 3, 0, 0, 0);
                h.alpha = add_param(n, p + ".1.weight", P_PLAIN, 1, 0, 0, 0);
                h.gamma = add_param(n, p + ".2.gamma", P_PLAIN, C, 0, 0, 0);
                h.beta = add_param(n, p + ".2.beta", P_PLAIN, C, 0, 0, 0);
                h.pw = add_param(n, p + ".3.weight", P_PW_W, (int64_t)C * C, C, C, 1);
                n->tcn.push_back(h);
            }
    size_t total = 0;
    for (auto &p : n->params) total += align_up(p.packed_elems, 64);
    cudaError_t e = cudaMalloc(&n->arena, total * sizeof(float));
    if (e != cudaSuccess) {
        delete n;
        return cuda_fail(e, "cudaMalloc(packed weights)");
    }
    cudaMemset(n->arena, 0, total * sizeof(float));
    size_t off = 0;
    for (auto &p : n->params) {
        p.d = n->arena + off;
        off += align_up(p.packed_elems, 64);
    }
    *out = n;
    return MISO_OK;
}

int miso_net_destroy(miso_net_t *net) {
    if (!net) return MISO_OK;
    for (auto &g : net->graphs) cudaGraphExecDestroy(g.exec);
    for (auto &g : net->tgraphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    if (net->cap_stream) cudaStreamDestroy(net->cap_stream);
    for (auto &e : net->bucket_ev)
        if (e) cudaEventDestroy(e);
    if (net->arena) cudaFree(net->arena);
    if (net->d_pack) cudaFree(net->d_pack);
    delete net;
    return MISO_OK;
}

int miso_net_num_params(const miso_net_t *net) { return net ? (int)net->params.size() : 0; }
const char *miso_net_param_key(const miso_net_t *net, int i) {
    if (!net || i < 0 || i >= (int)net->params.size()) return nullptr;
    return net->params[i].key.c_str();
}
int64_t miso_net_param_numel(const miso_net_t *net, int i) {
    if (!net || i < 0 || i >= (int)net->params.size()) return -1;
    return net->params[i].numel;
}

int miso_net_set_param(miso_net_t *net, const char *key, const float *d_data, int64_t numel, void *stream) {
    MISO_REQUIRE(net && key && d_data, "miso_net_set_param: null argument");
    auto it = net->index.find(key);
    MISO_REQUIRE(it != net->index.end(), "miso_net_set_param: unexpected key '%s'", key);
    Param &p = net->params[it->second];
    MISO_REQUIRE(p.numel == numel, "miso_net_set_param: '%s' has %lld elements, expected %lld", key, (long long)numel,
                 (long long)p.numel);
    cudaStream_t st = as_stream(stream);
    switch (p.kind) {
        case P_CONV_W:
        case P_DECONV_W:
        case P_PW_W: {
            int blocks = (int)std::min<size_t>((p.packed_elems + 255) / 256, 4096);
            pack_conv_w_kernel<<<blocks, 256, 0, st>>>(d_data, p.d, p.cout, p.cin, p.taps, p.cout_pad,
                                                      p.kind == P_DECONV_W ? 1 : 0);
            MISO_LAUNCHED("pack_conv_w_kernel");
            break;
        }
        case P_BIAS:
            pad_copy_kernel<<<ceil_div(p.cout_pad, 128), 128, 0, st>>>(d_data, p.d, p.cout, p.cout_pad);
            MISO_LAUNCHED("pad_copy_kernel");
            break;
        case P_PLAIN:
            MISO_CUDA(cudaMemcpyAsync(p.d, d_data, (size_t)numel * sizeof(float), cudaMemcpyDeviceToDevice, st));
            break;
    }
    if (!p.loaded) {
        p.loaded = true;
        net->n_loaded++;
    }
    return MISO_OK;
}

int miso_net_set_params(miso_net_t *net, const float *const *d_data, int n, void *stream) {
    MISO_REQUIRE(net && d_data, "miso_net_set_params: null argument");
    MISO_REQUIRE(n == (int)net->params.size(), "miso_net_set_params: %d pointers for %d parameters", n, (int)net->params.size());
    cudaStream_t st = as_stream(stream);
    if (!net->d_pack) {
        std::vector<PackDesc> tab(net->params.size());
        for (size_t i = 0; i < net->params.size(); ++i) {
            const Param &p = net->params[i];
            PackDesc &d = tab[i];
            d.dst = p.d;
            d.kind = p.kind == P_DECONV_W ? 1 : ((p.kind == P_CONV_W || p.kind == P_PW_W) ? 0 : (p.kind == P_BIAS ? 2 : 3));
            d.cout = p.cout;
            d.cin = p.cin;
            d.taps = p.taps;
            d.cout_pad = p.cout_pad;
            d.numel = p.numel;
            d.packed_elems = (long long)p.packed_elems;
        }
        MISO_CUDA(cudaMalloc(&net->d_pack, tab.size() * sizeof(PackDesc)));
        MISO_CUDA(cudaMemcpy(net->d_pack, tab.data(), tab.size() * sizeof(PackDesc), cudaMemcpyHostToDevice));
    }
    for (int first = 0; first < n; first += kPackBatch) {
        const int count = std::min(kPackBatch, n - first);
        PackSrcs srcs{};
        for (int j = 0; j < count; ++j) srcs.src[j] = d_data[first + j];  // null: leave that parameter as it is
        pack_params_kernel<<<dim3(32, count), 256, 0, st>>>(net->d_pack, srcs, first, count);
        MISO_LAUNCHED("pack_params_kernel");
    }
    for (int i = 0; i < n; ++i)
        if (d_data[i] && !net->params[i].loaded) {
            net->params[i].loaded = true;
            net->n_loaded++;
        }
    return MISO_OK;
}

int miso_net_set_mode(miso_net_t *net, int mode) {
    MISO_REQUIRE(net, "miso_net_set_mode: null handle");
    MISO_REQUIRE(mode >= 0 && mode <= 2, "miso_net_set_mode: mode %d unknown (0 fp32 FMA, 1 bf16x3 tcgen05, 2 bf16 tcgen05)",
                 mode);
    net->mode = mode;
    return MISO_OK;
}

int miso_net_check_shape(const miso_net_t *net, int T, int F) {
    MISO_REQUIRE(net, "miso_net_check_shape: null handle");
    MISO_REQUIRE(T >= 1, "T=%d must be positive", T);
    std::vector<int> Fx;
    if (!encoder_sizes(net, F, Fx)) {
        // required F: 2^(nb-2)*... easiest to state by search
        int want = -1;
        for (int f = 3; f < 100000; ++f)
            if (encoder_sizes(net, f, Fx)) {
                want = f;
                break;
            }
        set_error("F=%d does not reduce to 1 at the bottleneck of a %d-block MISO network (needs F=%d)", F, net->nb, want);
        return MISO_E_ARG;
    }
    return MISO_OK;
}

size_t miso_net_workspace_bytes(const miso_net_t *net, int B, int T, int F) {
    if (!net || B <= 0 || T <= 0) return 0;
    Plan pl;
    if (!full_plan(net, B, T, F, nullptr, pl)) return 0;
    return pl.total;
}

size_t miso_net_input_bytes(const miso_net_t *net, int B, int T, int F) {
    if (!net || B <= 0 || T <= 0 || F <= 0) return 0;
    return plane_bytes(B, (net->in_ch + 7) & ~7, T, F);
}

int miso_net_forward(miso_net_t *net, const void *d_x, float *d_y, int B, int T, int F, void *d_ws, size_t ws_bytes,
                     void *stream) {
    MISO_REQUIRE(net && d_x && d_y && d_ws, "miso_net_forward: null argument");
    if (net->n_loaded != (int)net->params.size()) {
        for (auto &p : net->params)
            if (!p.loaded) {
                set_error("miso_net_forward: parameter '%s' was never set (%d of %d loaded)", p.key.c_str(), net->n_loaded,
                          (int)net->params.size());
                return MISO_E_STATE;
            }
    }
    MISO_REQUIRE(B >= 1 && B <= 65535, "miso_net_forward: batch %d out of range", B);
    int rc = miso_net_check_shape(net, T, F);
    if (rc) return rc;
    MISO_REQUIRE((reinterpret_cast<uintptr_t>(d_ws) & 255) == 0, "miso_net_forward: workspace must be 256-byte aligned");
    MISO_REQUIRE((reinterpret_cast<uintptr_t>(d_x) & 127) == 0, "miso_net_forward: input planes must be 128-byte aligned");
    Plan pl;
    full_plan(net, B, T, F, reinterpret_cast<char *>(d_ws), pl);
    if (pl.total > ws_bytes) {
        set_error("miso_net_forward: workspace %zu < required %zu bytes", ws_bytes, pl.total);
        return MISO_E_WORKSPACE;
    }
    cudaStream_t st = as_stream(stream);
    // Replay a captured graph when nothing baked into the kernel arguments changed.  With per-launch
    // profiling on (miso_prof_enable) the graph additionally carries event-record nodes around every conv.
    if (net->use_graph) {
        const int prof = prof_enabled() ? 1 : 0;
        for (auto &g : net->graphs)
            if (g.x == d_x && g.y == d_y && g.ws == d_ws && g.B == B && g.T == T && g.F == F && g.mode == net->mode &&
                g.prof == prof) {
                MISO_CUDA(cudaGraphLaunch(g.exec, st));
                g_launch_count.fetch_add(g.launches, std::memory_order_relaxed);
                if (prof) prof_replayed(g.recs);
                return MISO_OK;
            }
        rc = conv_tc_init();
        if (rc) return rc;
        if (!net->cap_stream) MISO_CUDA(cudaStreamCreateWithFlags(&net->cap_stream, cudaStreamNonBlocking));
        const uint64_t l0 = g_launch_count.load();
        std::vector<ProfRec> recs;
        MISO_CUDA(cudaStreamBeginCapture(net->cap_stream, cudaStreamCaptureModeThreadLocal));
        prof_capture(&recs);
        rc = enqueue_forward(net, pl, d_x, d_y, B, T, F, net->cap_stream);
        prof_capture(nullptr);
        cudaGraph_t graph = nullptr;
        cudaError_t ce = cudaStreamEndCapture(net->cap_stream, &graph);
        if (rc) {
            if (graph) cudaGraphDestroy(graph);
            return rc;
        }
        if (ce != cudaSuccess) return cuda_fail(ce, "cudaStreamEndCapture");
        miso_net::GraphEntry e{d_x, d_y, d_ws, B, T, F, net->mode, prof, nullptr, g_launch_count.load() - l0, recs};
        ce = cudaGraphInstantiate(&e.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ce != cudaSuccess) return cuda_fail(ce, "cudaGraphInstantiate");
        if (net->graphs.size() >= 8) {
            cudaGraphExecDestroy(net->graphs.front().exec);
            net->graphs.erase(net->graphs.begin());
        }
        net->graphs.push_back(e);
        MISO_CUDA(cudaGraphLaunch(e.exec, st));
        if (prof) prof_replayed(e.recs);
        return MISO_OK;
    }
    return enqueue_forward(net, pl, d_x, d_y, B, T, F, st);
}

int miso_debug_tc_trace(long long *d_buf, int cin, int fin) {
    // both kernels log into the same buffer; MISO_TRACE_KERNEL=rs|tc restricts the trace to one of them (a strided conv
    // and a DenseBlock conv can share (cin, Fin), and the later launch would overwrite the earlier one's log)
    const char *which = getenv("MISO_TRACE_KERNEL");
    const bool tc = !which || which[0] == 't', rs = !which || which[0] == 'r';
    tcn_fused_set_trace(cin == -7 ? d_buf : nullptr);  // (cin = -7: the fused TCN launch, [half][16] stamps)
    conv_tc_set_trace(tc ? d_buf : nullptr, cin, fin);
    conv_rs_set_trace(rs ? d_buf : nullptr, cin, fin);
    return MISO_OK;
}

int miso_net_set_graph(miso_net_t *net, int on) {
    MISO_REQUIRE(net, "miso_net_set_graph: null handle");
    net->use_graph = on ? 1 : 0;
    return MISO_OK;
}

}  // extern "C"

namespace miso {
namespace {
int enqueue_forward(miso_net *net, const Plan &pl, const void *d_x, float *d_y, int B, int T, int F, cudaStream_t st) {
    const int C = net->C;

    MISO_CUDA(cudaMemsetAsync(pl.stats_base, 0, pl.stats_bytes, st));
    {
        // raw (never normalised) channels: enc0's first conv output (model.py:401-406) and the TCN output
        const BufDesc &e0 = pl.E[0];
        sentinel_kernel<<<ceil_div(B * net->en[1], 128), 128, 0, st>>>(e0.sums, B, e0.ctot, 0, net->en[1]);
        MISO_LAUNCHED("sentinel_kernel");
        sentinel_kernel<<<ceil_div(B * C, 128), 128, 0, st>>>(pl.D[0].sums, B, pl.D[0].ctot, 0, C);
        MISO_LAUNCHED("sentinel_kernel");
    }
    Walker w{net, pl, B, T, F, st, false};
    return w.run(d_x, d_y);
}
}  // namespace
}  // namespace miso

extern "C" {

int64_t miso_net_tap(miso_net_t *net, const char *name, float *d_out, int64_t capacity, int B, int T, int F, void *d_ws,
                     void *stream) {
    MISO_REQUIRE(net && name && d_out && d_ws, "miso_net_tap: null argument");
    Plan pl;
    if (!full_plan(net, B, T, F, reinterpret_cast<char *>(d_ws), pl)) {
        set_error("miso_net_tap: bad shape");
        return MISO_E_ARG;
    }
    const miso_net *n = net;
    std::string s(name);
    const BufDesc *buf = nullptr;
    int coff = 0, c = 0;
    if (s.rfind("enc", 0) == 0) {
        int i = atoi(s.c_str() + 3);
        MISO_REQUIRE(i >= 0 && i < n->nb, "miso_net_tap: no such tap '%s'", name);
        ViewRef v = xs_view(n, pl, i);
        buf = v.buf;
        coff = v.coff;
        c = v.c;
    } else if (s == "tcn") {
        buf = &pl.D[0];
        coff = 0;
        c = n->C;
    } else if (s.rfind("dec", 0) == 0) {
        int j = atoi(s.c_str() + 3);
        MISO_REQUIRE(j >= 0 && j < n->nb - 1, "miso_net_tap: no such tap '%s' (the last decoder is the network output)",
                     name);
        buf = &pl.D[j + 1];
        coff = 0;
        c = n->de[j + 1];
    } else {
        set_error("miso_net_tap: no such tap '%s'", name);
        return MISO_E_ARG;
    }
    int64_t total = (int64_t)B * c * T * buf->F;
    MISO_REQUIRE(total <= capacity, "miso_net_tap: output capacity %lld < %lld", (long long)capacity, (long long)total);
    int blocks = (int)std::min<int64_t>((total + 255) / 256, 8192);
    tap_kernel<<<blocks, 256, 0, as_stream(stream)>>>(reinterpret_cast<const __nv_bfloat16 *>(buf->p), buf->ctot, coff, c,
                                                      net->mode == 2 ? 0 : 1, buf->sums, 1.0 / ((double)T * buf->F), kInEps,
                                                      d_out, B, T * buf->F);
    MISO_LAUNCHED("tap_kernel");
    return total;
}

int miso_pack_miso1(const void *d_mix, void *d_x, int B, int M, int T, int F, const int *shifts, int n_shift,
                    void *stream) {
    MISO_REQUIRE(d_mix && d_x && shifts, "miso_pack_miso1: null argument");
    MISO_REQUIRE(M >= 1 && M <= 8, "miso_pack_miso1: M=%d unsupported (1..8)", M);
    MISO_REQUIRE(n_shift >= 1 && n_shift <= 16, "miso_pack_miso1: n_shift=%d unsupported (1..16)", n_shift);
    ShiftList sh;
    sh.n = n_shift;
    for (int k = 0; k < n_shift; ++k) sh.s[k] = ((shifts[k] % M) + M) % M;
    int64_t n = (int64_t)B * T * F;
    pack_miso1_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const float2 *>(d_mix), reinterpret_cast<__nv_bfloat16 *>(d_x), B, M, T * F, sh);
    MISO_LAUNCHED("pack_miso1_kernel");
    return MISO_OK;
}

int miso_pack_miso3(const void *d_mix, const void *d_second, const void *d_third, void *d_x, int B, int M, int T, int F,
                    void *stream) {
    MISO_REQUIRE(d_mix && d_second && d_third && d_x, "miso_pack_miso3: null argument");
    MISO_REQUIRE(M >= 1 && M <= 6, "miso_pack_miso3: M=%d unsupported (1..6)", M);
    int64_t n = (int64_t)B * T * F;
    pack_miso3_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const float2 *>(d_mix), reinterpret_cast<const float2 *>(d_second),
        reinterpret_cast<const float2 *>(d_third), reinterpret_cast<__nv_bfloat16 *>(d_x), B, M, T * F);
    MISO_LAUNCHED("pack_miso3_kernel");
    return MISO_OK;
}

int miso_unpack_complex(const float *d_y, void *d_out, int B, int S, int T, int F, void *stream) {
    MISO_REQUIRE(d_y && d_out && S >= 1, "miso_unpack_complex: bad argument");
    int64_t n = (int64_t)B * T * F;
    unpack_complex_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(d_y, reinterpret_cast<float2 *>(d_out),
                                                                                  B, S, T * F);
    MISO_LAUNCHED("unpack_complex_kernel");
    return MISO_OK;
}

}  // extern "C"

// =========================================================================== training =====
// Forward that keeps what the backward pass needs (per-block TCN state) and the backward pass itself
// (reference trainer.py:159-212: model(mix) -> loss_uPIT -> loss.backward()).
namespace miso {
namespace {

int param_grad_offsets(const miso_net *n, std::vector<int64_t> &off) {
    off.assign(n->params.size() + 1, 0);
    for (size_t i = 0; i < n->params.size(); ++i) off[i + 1] = off[i] + n->params[i].numel;
    return MISO_OK;
}

// MISO_BWD_PROF=1: CUDA-event timing of the backward's phases (eager launches; printed to stderr after every backward)
struct BwdProf {
    bool on = getenv("MISO_BWD_PROF") && atoi(getenv("MISO_BWD_PROF")) != 0;
    std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> recs;
    cudaEvent_t open = nullptr;
    void begin(cudaStream_t st) {
        if (!on) return;
        cudaEventCreate(&open);
        cudaEventRecord(open, st);
    }
    void end(cudaStream_t st, int phase) {
        if (!on) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, st);
        recs.push_back({phase, {open, e}});
    }
    void report() {
        if (!on) return;
        static const char *names[] = {"in_bwd (IN/ELU backward)", "wgrad", "dgrad", "tcn", "other"};
        double tot[5] = {0, 0, 0, 0, 0};
        for (auto &r : recs) {
            cudaEventSynchronize(r.second.second);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, r.second.first, r.second.second);
            tot[r.first] += ms;
            cudaEventDestroy(r.second.first);
            cudaEventDestroy(r.second.second);
        }
        recs.clear();
        fprintf(stderr, "backward phases (ms):");
        for (int i = 0; i < 5; ++i) fprintf(stderr, " %s %.2f |", names[i], tot[i]);
        fprintf(stderr, "\n");
    }
};
BwdProf g_bwd_prof;

struct Backward {
    const miso_net *n;
    const Plan &pl;
    int B, T, F;
    cudaStream_t st;
    float *grads;
    std::vector<int64_t> goff;

    float *g(int param) const { return grads + goff[param]; }

    // data gradient of one layer through the forward FMA kernel with transposed weights: in' = dL/dy (channels-last),
    // out' = the input buffer's gradient, accumulated in place (resid == out)
    bool use_tc(const ConvArgs &f) const {
        if (n->mode == 0) return false;  // fp32 mode: FMA kernels throughout
        int flip;
        ConvArgs d = dgrad_tc_args(f, pl, B, T, nullptr, &flip);
        return dgrad_tc_ok(f, d, pl, B, T);
    }

    int dgrad(const ConvArgs &f, const float *w_packed, int cout_pad_fwd, const float *dy, float *din, bool tc = false) {
        const int bn = conv_fp32_tile_n(f.cin);
        const int cin_pad = (f.cin + bn - 1) / bn * bn;
        if (tc) {
            int flip;
            ConvArgs d = dgrad_tc_args(f, pl, B, T, din, &flip);
            int rc = launch_dgrad_pack(w_packed, pl.wT, f.KT * f.KF, f.cin, f.cout, cout_pad_fwd, cin_pad, flip, st);
            if (rc) return rc;
            static const bool no_rs = getenv("MISO_DGRAD_RS") && atoi(getenv("MISO_DGRAD_RS")) == 0;  // debugging: general kernel only
            if (!no_rs && dgrad_rs_ok(d, flip, pl, B, T)) {
                if (dgrad_rs_cl()) return launch_conv_tc(dgrad_rs_chunk(d, pl, T, 0, d.cout), 3, pl.scratch, st);  // all chunks in one launch
                for (int c0 = 0; c0 < d.cout; c0 += kRsChunk) {
                    rc = launch_conv_tc(dgrad_rs_chunk(d, pl, T, c0, kRsChunk), 3, pl.scratch, st);
                    if (rc) return rc;
                }
                return launch_planes_accumulate(reinterpret_cast<const __nv_bfloat16 *>(pl.dgP), (d.cout + 7) & ~7, din, d.out_ctot,
                                                d.out_coff, d.cout, B, T * d.Fout, st);
            }
            return launch_conv_tc(d, 3, pl.scratch, st);
        }
        int rc = launch_dgrad_pack(w_packed, pl.wT, f.KT * f.KF, f.cin, f.cout, cout_pad_fwd, cin_pad, 0, st);
        if (rc) return rc;
        ConvArgs a{};
        a.in = dy;
        a.in_layout = LAYOUT_CL_F32;
        a.w = pl.wT;
        a.bias = nullptr;
        a.out = din;
        a.out_layout = LAYOUT_CL_F32;
        a.resid = din;
        a.resid_ctot = f.in_ctot;
        a.resid_coff = f.in_coff;
        a.in_sums = nullptr;
        a.out_sums = nullptr;
        a.B = B;
        a.T = T;
        a.Fin = f.Fout;
        a.Fout = f.Fin;
        a.in_ctot = f.out_ctot;
        a.in_coff = f.out_coff;
        a.cin = f.cout;
        a.out_ctot = f.in_ctot;
        a.out_coff = f.in_coff;
        a.cout = f.cin;
        a.cout_pad = cin_pad;
        a.KT = f.KT;
        a.KF = f.KF;
        a.stride_f = f.stride_f;
        a.pad_t = f.pad_t;
        a.pad_f = f.pad_f;
        a.transposed = f.transposed ? 0 : 1;
        a.norm_mode = NORM_NONE;
        a.elu = 0;
        a.use_lo = 1;
        return launch_conv_fp32(a, st);
    }

    int conv_layer(const ConvRec &r, float *d_gy) {
        ConvArgs f = r.a;
        const int cout_real = f.cout;
        float *og = r.out_grad ? r.out_grad : d_gy;
        if (!r.out_grad) {
            // the network output: the caller's gradient is padded to a multiple of 8 channels (miso_grad_pack); the pad
            // channels carry zero gradient and meet zero-padded packed weights in the data gradient
            f.cout = (f.cout + 7) & ~7;  // a whole 8-channel plane group: the tensor-core gradient paths take the layer too
            f.out_ctot = f.cout;
        }
        int rc;
        InBwdArgs ib{};
        ib.e = reinterpret_cast<const __nv_bfloat16 *>(f.out);
        ib.g = og;
        ib.sums = f.out_sums;
        ib.red = pl.bred;
        ib.dbias = g(r.cd->b);
        ib.B = B;
        ib.npix = T * f.Fout;
        ib.ctot = f.out_ctot;
        ib.coff = f.out_coff;
        ib.c = f.cout;
        ib.c_real = cout_real;
        ib.use_lo = f.use_lo;
        ib.inv_n = 1.0 / ((double)T * f.Fout);
        ib.eps = kInEps;
        ib.plain = f.elu ? 0 : 1;
        const bool tc = r.in_grad != nullptr && use_tc(f);
        // dy as bf16 hi/lo planes: the input of the tensor-core data gradient AND of the tcgen05 weight gradient (which
        // also serves layers without a data gradient, i.e. the first conv)
        const bool planes_fit = n->mode != 0 && f.cout % 8 == 0 && (size_t)B * f.cout * T * f.Fout * 4 <= pl.dyP_bytes;
        ib.dyp = (tc || planes_fit) ? reinterpret_cast<__nv_bfloat16 *>(pl.dyP) : nullptr;
        if (!ib.plain && (!f.out_sums || f.out_layout != LAYOUT_PLANES)) {
            set_error("backward: normalised layer without statistics");
            return MISO_E_STATE;
        }
        g_bwd_prof.begin(st);
        rc = launch_in_bwd(ib, st);
        g_bwd_prof.end(st, 0);
        if (rc) return rc;
        WgradArgs w{};
        w.x = f.in;
        w.x_layout = f.in_layout;
        w.use_lo = f.use_lo;
        w.x_sums = f.in_sums;
        w.inv_n = f.norm_inv_n;
        w.eps = f.norm_eps;
        w.dy = og;
        w.dw = g(r.cd->w);
        w.B = B;
        w.T = T;
        w.Fin = f.Fin;
        w.Fout = f.Fout;
        w.x_ctot = f.in_ctot;
        w.x_coff = f.in_coff;
        w.cin = f.cin;
        w.dy_ctot = f.out_ctot;
        w.dy_coff = f.out_coff;
        w.cout = f.cout;
        w.cout_real = cout_real;
        w.KT = f.KT;
        w.KF = f.KF;
        w.stride_f = f.stride_f;
        w.pad_t = f.pad_t;
        w.pad_f = f.pad_f;
        w.transposed = f.transposed;
        if (n->mode != 0 && ib.dyp) {  // tensor-core modes: the DenseBlock convs take the tcgen05 GEMM over the raw planes
            w.dyp = ib.dyp;
            w.partial = pl.wgP;
            w.partial_bytes = pl.wgP_bytes;
            w.ones = pl.wg_ones;
        }
        g_bwd_prof.begin(st);
        rc = launch_wgrad(w, st);
        g_bwd_prof.end(st, 1);
        if (rc) return rc;
        g_bwd_prof.begin(st);
        if (r.in_grad) rc = dgrad(f, n->params[r.cd->w].d, r.cd->cout_pad, og, r.in_grad, tc);
        g_bwd_prof.end(st, 2);
        return rc;
    }

    // one half of a TemporalBlock: upstream `dout` (gradient of the pointwise conv's output) -> gradient of the half's
    // input written / accumulated into `din`
    int tcn_half(int k, int half, const float *dout, float *din, int accumulate) {
        const int C = n->C;
        const TcnHalf &h = n->tcn[k * 2 + half];
        TcnBwdArgs a{};
        a.u = half == 0 ? pl.Sk[k] : pl.Uk[k];
        a.u_sums = half == 0 ? pl.sS[k] : pl.sU[k];
        a.inv_T = 1.0 / (double)T;
        a.in_eps = kInEps;
        a.g_sums = half == 0 ? pl.g1[k] : pl.g2[k];
        a.gln_inv_n = 1.0 / ((double)C * T);
        a.gln_eps = kGlnEps;
        a.wdw = n->params[h.dw].d;
        a.alpha = n->params[h.alpha].d;
        a.gamma = n->params[h.gamma].d;
        a.beta = n->params[h.beta].d;
        a.B = B;
        a.T = T;
        a.C = C;
        a.dil = 1 << (k % n->X);
        int rc = launch_tcn_recompute(a, pl.tY, pl.tQ, st);
        if (rc) return rc;
        // pointwise conv: weight gradient from q, data gradient into tDQ
        ConvArgs f = tcn_pw_fwd_args(C, B, T);
        f.in = pl.tQ;
        WgradArgs w{};
        w.x = pl.tQ;
        w.x_layout = LAYOUT_CL_F32;
        w.use_lo = 1;
        w.x_sums = nullptr;
        w.inv_n = 1.0;
        w.eps = 0.f;
        w.dy = dout;
        w.dw = g(h.pw);
        w.B = B;
        w.T = T;
        w.Fin = 1;
        w.Fout = 1;
        w.x_ctot = C;
        w.x_coff = 0;
        w.cin = C;
        w.dy_ctot = C;
        w.dy_coff = 0;
        w.cout = C;
        w.cout_real = C;
        w.KT = 1;
        w.KF = 1;
        w.stride_f = 1;
        w.pad_t = 0;
        w.pad_f = 0;
        w.transposed = 0;
        rc = launch_wgrad(w, st);
        if (rc) return rc;
        const bool tc = use_tc(f);
        if (tc) {
            rc = launch_cl_to_planes(dout, reinterpret_cast<__nv_bfloat16 *>(pl.dyP), B, T, C, st);
            if (rc) return rc;
        }
        if (tc && pw_dgrad_gemm) {
            // dq = dy W^T as the forward's pointwise GEMM kernel over the transposed weight image (built once per backward)
            TcnPwArgs p{};
            p.planes = pl.dyP;
            p.lo_off = (size_t)C * T * 2;
            p.wimg = pl.tcn_wimgT;
            p.wvec = pl.tcn_wvec;
            p.index = k * 2 + half;
            p.plain = 1;
            p.out = pl.tDQ;
            p.B = B;
            p.T = T;
            p.C = C;
            rc = launch_tcn_pw(p, 3, st);
        } else {
            MISO_CUDA(cudaMemsetAsync(pl.tDQ, 0, (size_t)B * T * C * sizeof(float), st));
            rc = dgrad(f, n->params[h.pw].d, n->params[h.pw].cout_pad, dout, pl.tDQ, tc);
        }
        if (rc) return rc;
        rc = launch_gln_bwd(a, pl.tDQ, pl.tY, pl.bred, g(h.gamma), g(h.beta), g(h.alpha), st);
        if (rc) return rc;
        return launch_dw_bwd(a, pl.tDQ, pl.tDN, pl.bred, g(h.dw), din, accumulate, st);
    }

    bool pw_dgrad_gemm = false;  // the pointwise convs' data gradients run on tcn_pw_kernel (transposed images in pl.tcn_wimgT)

    int tcn() {
        const int C = n->C;
        const BufDesc &d0 = pl.D[0];
        const int64_t rows = (int64_t)B * T;
        {
            static const bool off = getenv("MISO_TCN_DGRAD_PW") && atoi(getenv("MISO_TCN_DGRAD_PW")) == 0;
            const int nblk = n->R * n->X;
            pw_dgrad_gemm = !off && n->mode == 1 && pl.tcn_wimgT && tcn_pw_eligible(C) && 2 * nblk <= kTcnMaxPw;
            if (pw_dgrad_gemm) {
                const int bn = conv_fp32_tile_n(C);
                const int cpad = (C + bn - 1) / bn * bn;
                TcnPwTable tab{};
                for (int i = 0; i < 2 * nblk; ++i) {
                    tab.w[i] = n->params[n->tcn[i].pw].d;
                    tab.gamma[i] = n->params[n->tcn[i].gamma].d;
                    tab.beta[i] = n->params[n->tcn[i].beta].d;
                }
                int rc0 = launch_tcn_wprep(tab, 2 * nblk, C, cpad, 3, pl.tcn_wimgT, nullptr, st, 1);
                if (rc0) return rc0;
            }
        }
        // upstream: the TCN output is consumed raw as channels [0, C) of decoder 0's buffer (model.py:97-99)
        int rc = launch_copy_channels(d0.grad, d0.ctot, 0, pl.gS, C, 0, C, rows, 0, st);
        if (rc) return rc;
        for (int k = n->R * n->X - 1; k >= 0; --k) {
            rc = tcn_half(k, 1, pl.gS, pl.gU, 0);  // y = half1(U_k) + S_k
            if (rc) return rc;
            rc = tcn_half(k, 0, pl.gU, pl.gS, 1);  // dS_k = dS_{k+1} + d half0
            if (rc) return rc;
        }
        // S_0 = InstanceNorm(last encoder's output) = the normalised view of channels [C, 2C) of decoder 0's buffer
        return launch_copy_channels(pl.gS, C, 0, d0.grad, d0.ctot, C, C, rows, 1, st);
    }
};

}  // namespace
}  // namespace miso

extern "C" {

}  // extern "C"

namespace miso {
namespace {
// Bucket k covers the parameters [first[k], first[k + 1]) of one completion group; groups in completion order:
// decoders nb/2..nb-1, decoders 0..nb/2-1, TCN, encoders nb/2..nb-1, encoders 0..nb/2-1 (the backward walks the layers
// from the output to the input, model.py:97-106 reversed).  Parameters are in key order (encoders, decoders, TCN), so each
// group is one contiguous range of the flat gradient buffer.
constexpr int kGradBuckets = 5;
struct BucketTable {
    int lo[kGradBuckets], hi[kGradBuckets];  // parameter index ranges, completion order
};
BucketTable grad_buckets(const miso_net *n) {
    auto first_with = [&](const std::string &prefix) {
        for (size_t i = 0; i < n->params.size(); ++i)
            if (n->params[i].key.compare(0, prefix.size(), prefix) == 0) return (int)i;
        return (int)n->params.size();
    };
    const int h = n->nb / 2;
    const int e0 = 0, e1 = first_with("encoders." + std::to_string(h) + "."), d0 = first_with("decoders.0."),
              d1 = first_with("decoders." + std::to_string(h) + "."), t0 = first_with("TCN."), end = (int)n->params.size();
    BucketTable b;
    b.lo[0] = d1, b.hi[0] = t0;   // upper decoders (run first)
    b.lo[1] = d0, b.hi[1] = d1;   // lower decoders
    b.lo[2] = t0, b.hi[2] = end;  // TCN
    b.lo[3] = e1, b.hi[3] = d0;   // upper encoders
    b.lo[4] = e0, b.hi[4] = e1;   // lower encoders (run last)
    return b;
}
int bucket_of_param(const BucketTable &b, int p) {
    for (int k = 0; k < kGradBuckets; ++k)
        if (p >= b.lo[k] && p < b.hi[k]) return k;
    return kGradBuckets - 1;
}
}  // namespace
}  // namespace miso

extern "C" {

int miso_net_grad_buckets(const miso_net_t *net, int64_t *begin, int64_t *end, int capacity) {
    MISO_REQUIRE(net && begin && end, "miso_net_grad_buckets: null argument");
    MISO_REQUIRE(capacity >= kGradBuckets, "miso_net_grad_buckets: capacity %d < %d", capacity, kGradBuckets);
    std::vector<int64_t> off;
    param_grad_offsets(net, off);
    const BucketTable b = grad_buckets(net);
    for (int k = 0; k < kGradBuckets; ++k) {
        begin[k] = off[b.lo[k]];
        end[k] = off[b.hi[k]];
    }
    return kGradBuckets;
}

int miso_net_wait_grad_bucket(miso_net_t *net, int bucket, void *stream) {
    MISO_REQUIRE(net && bucket >= 0 && bucket < kGradBuckets, "miso_net_wait_grad_bucket: bad bucket %d", bucket);
    MISO_REQUIRE(net->bucket_ev[bucket], "miso_net_wait_grad_bucket: no backward pass has run on this handle");
    MISO_CUDA(cudaStreamWaitEvent(as_stream(stream), net->bucket_ev[bucket], 0));
    return MISO_OK;
}

int64_t miso_net_grad_numel(const miso_net_t *net) {
    if (!net) return -1;
    int64_t t = 0;
    for (auto &p : net->params) t += p.numel;
    return t;
}

size_t miso_net_train_workspace_bytes(const miso_net_t *net, int B, int T, int F) {
    if (!net || B <= 0 || T <= 0) return 0;
    Plan pl;
    if (!full_plan(net, B, T, F, nullptr, pl, true)) return 0;
    return pl.total;
}

}  // extern "C"

namespace miso {
namespace {
// Runs `enqueue(stream)` eagerly the first time a key is seen, captures it into a CUDA graph the second time and replays
// the graph afterwards (miso_net::TrainGraph).
template <class Fn>
int train_graphed(miso_net *net, int kind, const void *x, const void *p, const void *q, void *ws, int B, int T, int F, cudaStream_t st,
                  Fn enqueue) {
    static const bool off = getenv("MISO_TRAIN_GRAPH") && atoi(getenv("MISO_TRAIN_GRAPH")) == 0;
    if (!net->use_graph || off || prof_enabled()) return enqueue(st);
    miso_net::TrainGraph *hit = nullptr;
    for (auto &g : net->tgraphs)
        if (g.kind == kind && g.x == x && g.p == p && g.q == q && g.ws == ws && g.B == B && g.T == T && g.F == F && g.mode == net->mode) {
            hit = &g;
            break;
        }
    if (hit && hit->exec) {
        MISO_CUDA(cudaGraphLaunch(hit->exec, st));
        g_launch_count.fetch_add(hit->launches, std::memory_order_relaxed);
        return MISO_OK;
    }
    if (!hit) {
        if (net->tgraphs.size() >= 8) {
            if (net->tgraphs.front().exec) cudaGraphExecDestroy(net->tgraphs.front().exec);
            net->tgraphs.erase(net->tgraphs.begin());
        }
        net->tgraphs.push_back(miso_net::TrainGraph{kind, x, p, q, ws, B, T, F, net->mode, nullptr, 0});
        return enqueue(st);
    }
    if (!net->cap_stream) MISO_CUDA(cudaStreamCreateWithFlags(&net->cap_stream, cudaStreamNonBlocking));
    const uint64_t l0 = g_launch_count.load();
    MISO_CUDA(cudaStreamBeginCapture(net->cap_stream, cudaStreamCaptureModeThreadLocal));
    int rc = enqueue(net->cap_stream);
    cudaGraph_t graph = nullptr;
    cudaError_t ce = cudaStreamEndCapture(net->cap_stream, &graph);
    if (rc) {
        if (graph) cudaGraphDestroy(graph);
        return rc;
    }
    if (ce != cudaSuccess) return cuda_fail(ce, "cudaStreamEndCapture(training)");
    ce = cudaGraphInstantiate(&hit->exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) {
        hit->exec = nullptr;
        return cuda_fail(ce, "cudaGraphInstantiate(training)");
    }
    hit->launches = g_launch_count.load() - l0;
    MISO_CUDA(cudaGraphLaunch(hit->exec, st));
    return MISO_OK;
}
}  // namespace
}  // namespace miso

extern "C" {

int miso_net_forward_train(miso_net_t *net, const void *d_x, float *d_y, int B, int T, int F, void *d_ws, size_t ws_bytes,
                           void *stream) {
    MISO_REQUIRE(net && d_x && d_y && d_ws, "miso_net_forward_train: null argument");
    MISO_REQUIRE(net->n_loaded == (int)net->params.size(), "miso_net_forward_train: %d of %d parameters loaded", net->n_loaded,
                 (int)net->params.size());
    MISO_REQUIRE(B >= 1 && B <= 65535, "miso_net_forward_train: batch %d out of range", B);
    int rc = miso_net_check_shape(net, T, F);
    if (rc) return rc;
    MISO_REQUIRE((reinterpret_cast<uintptr_t>(d_ws) & 255) == 0, "miso_net_forward_train: workspace must be 256-byte aligned");
    MISO_REQUIRE((reinterpret_cast<uintptr_t>(d_x) & 127) == 0, "miso_net_forward_train: input planes must be 128-byte aligned");
    Plan pl;
    full_plan(net, B, T, F, reinterpret_cast<char *>(d_ws), pl, true);
    if (pl.total > ws_bytes) {
        set_error("miso_net_forward_train: workspace %zu < required %zu bytes", ws_bytes, pl.total);
        return MISO_E_WORKSPACE;
    }
    rc = conv_tc_init();
    if (rc) return rc;
    return train_graphed(net, 1, d_x, d_x, d_y, d_ws, B, T, F, as_stream(stream),
                         [&](cudaStream_t s) { return enqueue_forward(net, pl, d_x, d_y, B, T, F, s); });
}

int miso_net_backward(miso_net_t *net, const void *d_x, float *d_gy, int B, int T, int F, void *d_ws, size_t ws_bytes,
                      float *d_grads, void *stream) {
    MISO_REQUIRE(net && d_x && d_gy && d_ws && d_grads, "miso_net_backward: null argument");
    int rc = miso_net_check_shape(net, T, F);
    if (rc) return rc;
    Plan pl;
    full_plan(net, B, T, F, reinterpret_cast<char *>(d_ws), pl, true);
    if (pl.total > ws_bytes) {
        set_error("miso_net_backward: workspace %zu < required %zu bytes", ws_bytes, pl.total);
        return MISO_E_WORKSPACE;
    }
    std::vector<ConvRec> recs;
    Walker w{net, pl, B, T, F, nullptr, false};
    w.record = &recs;
    rc = w.run(d_x, nullptr);
    if (rc) return rc;
    const BucketTable bt = grad_buckets(net);
    for (int k = 0; k < kGradBuckets; ++k)
        if (!net->bucket_ev[k]) MISO_CUDA(cudaEventCreateWithFlags(&net->bucket_ev[k], cudaEventDisableTiming));
    return train_graphed(net, 2, d_x, d_gy, d_grads, d_ws, B, T, F, as_stream(stream), [&](cudaStream_t st) -> int {
    Backward bw{net, pl, B, T, F, st, d_grads, {}};
    param_grad_offsets(net, bw.goff);
    MISO_CUDA(cudaMemsetAsync(d_grads, 0, (size_t)bw.goff.back() * sizeof(float), st));
    MISO_CUDA(cudaMemsetAsync(pl.grad_base, 0, pl.grad_bytes, st));
    rc = wgrad_tc_fill_ones(pl.wg_ones, pl.wg_ones_pix, st);
    if (rc) return rc;
    auto bucket_of = [&](const ConvRec &r) { return r.kind == 1 ? 2 : bucket_of_param(bt, r.cd->w); };
    for (int i = (int)recs.size() - 1; i >= 0; --i) {
        if (recs[i].kind == 1) g_bwd_prof.begin(st);
        rc = recs[i].kind == 1 ? bw.tcn() : bw.conv_layer(recs[i], d_gy);
        if (recs[i].kind == 1) g_bwd_prof.end(st, 3);
        if (rc) return rc;
        // the last layer of a completion group: every gradient of the bucket has been written
        // (an EXTERNAL record: inside a captured graph it becomes an event-record node that fires on every replay)
        if (i == 0 || bucket_of(recs[i - 1]) != bucket_of(recs[i]))
            MISO_CUDA(cudaEventRecordWithFlags(net->bucket_ev[bucket_of(recs[i])], st, capturing(st) ? cudaEventRecordExternal : cudaEventRecordDefault));
    }
    g_bwd_prof.report();
    return MISO_OK;
    });
}

}  // extern "C"
