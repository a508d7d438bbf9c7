// Pairwise spectrogram distances, permutation decisions and the L1 losses
// (A1/A2/L1/L2 of SURVEY.md section 8(a); reference tester.py:1043-1065, 889-915,
// criterion.py:8-63, 121-141).
//
// Two-stage reductions: every CTA writes fp64 partial sums, a second small kernel adds
// them in a fixed order, so the argmin decisions are independent of scheduling.
#include <algorithm>

#include "common.cuh"

namespace miso {
namespace {

constexpr int kMaxS = 4;
constexpr int kMaxPerm = 24;
constexpr int kPairThreads = 256;

struct PermTable {
    int n;  // number of permutations
    signed char p[kMaxPerm][kMaxS];
};

PermTable make_perms(int S) {
    PermTable t;
    int idx[kMaxS] = {0, 1, 2, 3};
    t.n = 0;
    do {  // lexicographic order == itertools.permutations(range(S)) (criterion.py:49)
        for (int i = 0; i < kMaxS; ++i) t.p[t.n][i] = (signed char)(i < S ? idx[i] : 0);
        t.n++;
    } while (std::next_permutation(idx, idx + S));
    return t;
}

int pair_chunks(int B, int64_t n) {
    int64_t want = (n + kPairThreads * 8 - 1) / (kPairThreads * 8);
    int64_t cap = std::max<int64_t>(1, (4 * 148 + B - 1) / B);
    return (int)std::max<int64_t>(1, std::min(want, cap));
}

__device__ __forceinline__ double block_sum_256(double v, double *sh) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x == 0)
        for (int i = 0; i < kPairThreads / 32; ++i) r += sh[i];
    return r;  // valid on thread 0
}

template <int S, int MODE>
__global__ void __launch_bounds__(kPairThreads) pair_partial_kernel(const float2 *__restrict__ a, int64_t a_sb, int64_t a_ss,
                                                                   const float2 *__restrict__ bq, int64_t b_sb, int64_t b_ss,
                                                                   int64_t n, double *__restrict__ partial) {
    __shared__ double sh[kPairThreads / 32];
    const int b = blockIdx.y;
    const int nch = gridDim.x;
    float acc[S][S];
    double dacc[S][S];
#pragma unroll
    for (int i = 0; i < S; ++i)
#pragma unroll
        for (int j = 0; j < S; ++j) {
            acc[i][j] = 0.f;
            dacc[i][j] = 0.0;
        }
    int cnt = 0;
    for (int64_t e = (int64_t)blockIdx.x * kPairThreads + threadIdx.x; e < n; e += (int64_t)nch * kPairThreads) {
        float2 av[S], bv[S];
        float am[S], bm[S];
#pragma unroll
        for (int i = 0; i < S; ++i) {
            av[i] = a[b * a_sb + i * a_ss + e];
            bv[i] = bq[b * b_sb + i * b_ss + e];
            if (MODE == 0)
                am[i] = sqrtf(av[i].x * av[i].x + av[i].y * av[i].y);
            else
                am[i] = sqrtf(av[i].x * av[i].x + av[i].y * av[i].y + 1e-8f);  // criterion.py:30 (EPS inside the sqrt)
            bm[i] = sqrtf(bv[i].x * bv[i].x + bv[i].y * bv[i].y);
        }
#pragma unroll
        for (int i = 0; i < S; ++i)
#pragma unroll
            for (int j = 0; j < S; ++j) {
                float d = fabsf(am[i] - bm[j]);
                if (MODE == 1) d += fabsf(av[i].x - bv[j].x) + fabsf(av[i].y - bv[j].y);
                acc[i][j] += d;
            }
        if (++cnt == 16) {  // flush the short fp32 runs into fp64
            cnt = 0;
#pragma unroll
            for (int i = 0; i < S; ++i)
#pragma unroll
                for (int j = 0; j < S; ++j) {
                    dacc[i][j] += (double)acc[i][j];
                    acc[i][j] = 0.f;
                }
        }
    }
#pragma unroll
    for (int i = 0; i < S; ++i)
#pragma unroll
        for (int j = 0; j < S; ++j) {
            double r = block_sum_256(dacc[i][j] + (double)acc[i][j], sh);
            if (threadIdx.x == 0) partial[(((size_t)b * nch + blockIdx.x) * S + i) * S + j] = r;
        }
}

__global__ void pair_decide_kernel(const double *__restrict__ partial, int nch, int B, int S, PermTable perms,
                                   float *__restrict__ pair_out, int64_t *__restrict__ idx_out, float *__restrict__ loss_out,
                                   double *__restrict__ minscore) {
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        double pr[kMaxS][kMaxS];
        for (int i = 0; i < S; ++i)
            for (int j = 0; j < S; ++j) {
                double s = 0.0;
                for (int c = 0; c < nch; ++c) s += partial[(((size_t)b * nch + c) * S + i) * S + j];
                pr[i][j] = s;
                if (pair_out) pair_out[((size_t)b * S + i) * S + j] = (float)s;
            }
        int best = 0;
        double bests = 0.0;
        for (int p = 0; p < perms.n; ++p) {
            double s = 0.0;
            for (int i = 0; i < S; ++i) s += pr[i][(int)perms.p[p][i]];
            if (p == 0 || s < bests) {  // strict: the first minimum wins, like torch.argmin
                bests = s;
                best = p;
            }
        }
        if (idx_out) idx_out[b] = best;
        minscore[b] = bests;
    }
    __syncthreads();
    if (threadIdx.x == 0 && loss_out) {
        double s = 0.0;
        for (int b = 0; b < B; ++b) s += minscore[b];
        loss_out[0] = (float)(s / (double)B);  // criterion.py:61-63
    }
}

__global__ void perm_gather_kernel(const float2 *__restrict__ src, int64_t src_sb, int64_t src_ss, float2 *__restrict__ dst,
                                   int64_t dst_sb, int64_t dst_ss, const int64_t *__restrict__ idx, int64_t n,
                                   PermTable perms) {
    const int s = blockIdx.y, b = blockIdx.z;
    int p = (int)idx[b];
    p = p < 0 ? 0 : (p >= perms.n ? perms.n - 1 : p);
    const int from = perms.p[p][s];
    const float2 *sp = src + b * src_sb + from * src_ss;
    float2 *dp = dst + b * dst_sb + s * dst_ss;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) dp[e] = sp[e];
}

__global__ void __launch_bounds__(kPairThreads) enhance_partial_kernel(const float2 *__restrict__ est,
                                                                      const float2 *__restrict__ ref, int64_t n,
                                                                      double *__restrict__ partial) {
    __shared__ double sh[kPairThreads / 32];
    double dacc = 0.0;
    float acc = 0.f;
    int cnt = 0;
    for (int64_t e = (int64_t)blockIdx.x * kPairThreads + threadIdx.x; e < n; e += (int64_t)gridDim.x * kPairThreads) {
        float2 a = est[e], r = ref[e];
        float am = sqrtf(a.x * a.x + a.y * a.y + 1e-8f);  // criterion.py:131
        float rm = sqrtf(r.x * r.x + r.y * r.y);
        acc += fabsf(a.x - r.x) + fabsf(a.y - r.y) + fabsf(am - rm);
        if (++cnt == 16) {
            cnt = 0;
            dacc += (double)acc;
            acc = 0.f;
        }
    }
    double r = block_sum_256(dacc + (double)acc, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = r;
}

__global__ void enhance_final_kernel(const double *__restrict__ partial, int nch, int B, float *__restrict__ loss) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double s = 0.0;
        for (int c = 0; c < nch; ++c) s += partial[c];
        loss[0] = (float)(s / (double)B);
    }
}

}  // namespace
}  // namespace miso

using namespace miso;

extern "C" {

size_t miso_pair_workspace_bytes(int B, int S, int T, int F) {
    if (B < 1 || S < 1 || S > kMaxS) return 0;
    int nch = pair_chunks(B, (int64_t)T * F);
    return align_up((size_t)B * nch * S * S * sizeof(double), 256) + align_up((size_t)B * sizeof(double), 256);
}

int miso_pair_fwd(const void *d_a, int64_t a_sb, int64_t a_ss, const void *d_b, int64_t b_sb, int64_t b_ss, int B, int S,
                  int T, int F, int mode, float *d_pair, int64_t *d_perm_idx, float *d_loss, void *d_ws, size_t ws_bytes,
                  void *stream) {
    MISO_REQUIRE(d_a && d_b && d_ws, "miso_pair_fwd: null argument");
    MISO_REQUIRE(S >= 1 && S <= kMaxS, "miso_pair_fwd: S=%d unsupported (1..%d)", S, kMaxS);
    MISO_REQUIRE(B >= 1 && B <= 65535 && T >= 1 && F >= 1, "miso_pair_fwd: bad shape");
    MISO_REQUIRE(mode == 0 || mode == 1, "miso_pair_fwd: bad mode %d", mode);
    const int64_t n = (int64_t)T * F;
    const int nch = pair_chunks(B, n);
    const size_t need = miso_pair_workspace_bytes(B, S, T, F);
    if (need > ws_bytes) {
        set_error("miso_pair_fwd: workspace %zu < required %zu bytes", ws_bytes, need);
        return MISO_E_WORKSPACE;
    }
    double *partial = reinterpret_cast<double *>(d_ws);
    double *minscore =
        reinterpret_cast<double *>(reinterpret_cast<char *>(d_ws) + align_up((size_t)B * nch * S * S * sizeof(double), 256));
    cudaStream_t st = as_stream(stream);
    const float2 *a = reinterpret_cast<const float2 *>(d_a);
    const float2 *b = reinterpret_cast<const float2 *>(d_b);
    dim3 grid(nch, B);
#define MISO_PAIR_CASE(s)                                                                                       \
    case s:                                                                                                     \
        if (mode == 0)                                                                                          \
            pair_partial_kernel<s, 0><<<grid, kPairThreads, 0, st>>>(a, a_sb, a_ss, b, b_sb, b_ss, n, partial); \
        else                                                                                                    \
            pair_partial_kernel<s, 1><<<grid, kPairThreads, 0, st>>>(a, a_sb, a_ss, b, b_sb, b_ss, n, partial); \
        break
    switch (S) {
        MISO_PAIR_CASE(1);
        MISO_PAIR_CASE(2);
        MISO_PAIR_CASE(3);
        MISO_PAIR_CASE(4);
    }
#undef MISO_PAIR_CASE
    MISO_LAUNCHED("pair_partial_kernel");
    pair_decide_kernel<<<1, 256, 0, st>>>(partial, nch, B, S, make_perms(S), d_pair, d_perm_idx, d_loss, minscore);
    MISO_LAUNCHED("pair_decide_kernel");
    return MISO_OK;
}

int miso_perm_gather(const void *d_src, int64_t src_sb, int64_t src_ss, void *d_dst, int64_t dst_sb, int64_t dst_ss,
                     const int64_t *d_perm_idx, int B, int S, int T, int F, void *stream) {
    MISO_REQUIRE(d_src && d_dst && d_perm_idx, "miso_perm_gather: null argument");
    MISO_REQUIRE(S >= 1 && S <= kMaxS, "miso_perm_gather: S=%d unsupported", S);
    MISO_REQUIRE(B >= 1 && B <= 65535, "miso_perm_gather: bad batch");
    const int64_t n = (int64_t)T * F;
    dim3 grid((unsigned)std::min<int64_t>((n + 255) / 256, 64), S, B);
    perm_gather_kernel<<<grid, 256, 0, as_stream(stream)>>>(reinterpret_cast<const float2 *>(d_src), src_sb, src_ss,
                                                           reinterpret_cast<float2 *>(d_dst), dst_sb, dst_ss, d_perm_idx, n,
                                                           make_perms(S));
    MISO_LAUNCHED("perm_gather_kernel");
    return MISO_OK;
}

int miso_loss_enhance_fwd(const void *d_est, const void *d_ref, int B, int64_t n_per_batch, float *d_loss, void *d_ws,
                          size_t ws_bytes, void *stream) {
    MISO_REQUIRE(d_est && d_ref && d_loss && d_ws, "miso_loss_enhance_fwd: null argument");
    MISO_REQUIRE(B >= 1 && n_per_batch >= 1, "miso_loss_enhance_fwd: bad shape");
    const int64_t n = (int64_t)B * n_per_batch;
    const int nch = (int)std::max<int64_t>(1, std::min<int64_t>((n + kPairThreads * 8 - 1) / (kPairThreads * 8), 1024));
    if ((size_t)nch * sizeof(double) > ws_bytes) {
        set_error("miso_loss_enhance_fwd: workspace %zu < required %zu bytes", ws_bytes, (size_t)nch * sizeof(double));
        return MISO_E_WORKSPACE;
    }
    double *partial = reinterpret_cast<double *>(d_ws);
    cudaStream_t st = as_stream(stream);
    enhance_partial_kernel<<<nch, kPairThreads, 0, st>>>(reinterpret_cast<const float2 *>(d_est),
                                                        reinterpret_cast<const float2 *>(d_ref), n, partial);
    MISO_LAUNCHED("enhance_partial_kernel");
    enhance_final_kernel<<<1, 32, 0, st>>>(partial, nch, B, d_loss);
    MISO_LAUNCHED("enhance_final_kernel");
    return MISO_OK;
}

}  // extern "C"
