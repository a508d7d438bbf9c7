// Per-frequency MVDR beamformer (M1..M7 of SURVEY.md section 8(a); reference
// tester.py:1071-1167, 1211-1228 and its two verbatim twins).
//
// Four kernels, all HBM-bound or trivially small:
//   scm_kernel    streams source + mixture once (coalesced along f, the contiguous axis of
//                 [B,M,T,F]) and accumulates both spatial covariance matrices per (b,s,f);
//                 T is split across CTAs and the partial sums are combined in a fixed order.
//   eig_kernel    Hermitian Jacobi eigen-decomposition per (b,s,f) -> principal eigenvector
//                 (np.linalg.eigh + argmax, tester.py:1107-1115) -> reference-mic and
//                 sqrt(M/||d||) normalisation (tester.py:1119-1123, reproduced as written).
//   solve_kernel  phase correction as a prefix product of unit phasors (equivalent to the
//                 sequential recurrence of tester.py:1163-1166), diagonal loading,
//                 pivoted 6x6 complex solve and w = u / (d^H u) (tester.py:1211-1225).
//   apply_kernel  y[t] = sum_m conj(w[m]) x[m,t] for all sources in one pass over the mixture.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace miso {
namespace {

template <int M>
struct Tri {
    static constexpr int N = M * (M + 1) / 2;  // complex entries of the upper triangle
    static constexpr int NV = 4 * N;           // floats per (b,s,f): re/im x {source, noise}
};

constexpr int kFx = 32, kTy = 8;

// L2 eviction-priority hints: the mixture is read by the covariance pass and again by the filter-and-sum pass of the
// same utterance chunk (evict_last keeps it resident in the 126 MB L2 in between), the sources are streamed once
// (evict_first keeps them from pushing the mixture out).
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ float2 ld_hint(const float2 *p, uint64_t policy) {
    float2 v;
    asm volatile("ld.global.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;" : "=f"(v.x), "=f"(v.y) : "l"(p), "l"(policy));
    return v;
}

// Both spatial covariances (source, noise = mixture - source) of up to two sources in ONE pass over the mixture: the
// block's eight time lanes are split between the sources (ty = sl * lanes + tl), the lanes of both sources walk the
// frames in lockstep, so the second reader of a mixture element hits L1 / L2 and DRAM sees the mixture once.
template <int M>
__global__ void __launch_bounds__(kFx *kTy) scm_kernel(const float2 *__restrict__ src, int64_t src_ss,
                                                        const float2 *__restrict__ mix, int64_t sb, int64_t sm, int64_t st,
                                                        int64_t sf, float *__restrict__ partial, int S, int B, int T, int F,
                                                        int tsplit, int b0) {
    constexpr int NV = Tri<M>::NV;
    __shared__ float red[2][NV][kFx];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int f = blockIdx.x * kFx + tx;
    const int split = blockIdx.y;
    const int b = blockIdx.z + b0;  // utterance chunk [b0, b0 + gridDim.z)
    const int tc = (T + tsplit - 1) / tsplit;
    const int t0 = split * tc, t1 = min(T, t0 + tc);
    const bool fok = f < F;
    const float2 *mix_b = mix + b * sb + (int64_t)f * sf;
    const uint64_t pol_mix = l2_policy_evict_last(), pol_src = l2_policy_evict_first();

    for (int s0 = 0; s0 < S; s0 += 2) {
        const int ns = min(2, S - s0);           // sources of this pass
        const int lanes = kTy / ns;              // time lanes per source
        const int sl = ty / lanes, tl = ty - sl * lanes;
        const int s = s0 + sl;
        float acc[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) acc[i] = 0.f;
        if (fok) {
            const float2 *src_b = src + s * src_ss + b * sb + (int64_t)f * sf;
            for (int t = t0 + tl; t < t1; t += lanes) {
                float2 x[M], y[M];
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    const float2 mx = ld_hint(mix_b + m * sm + t * st, pol_mix);
                    const float2 sx = ld_hint(src_b + m * sm + t * st, pol_src);
                    x[m] = sx;
                    y[m] = make_float2(mx.x - sx.x, mx.y - sx.y);  // noise = mix - source (tester.py:1095)
                }
                int k = 0;
#pragma unroll
                for (int i = 0; i < M; ++i)
#pragma unroll
                    for (int j = i; j < M; ++j) {
                        // Phi[i][j] += x_i conj(x_j)
                        acc[4 * k + 0] = fmaf(x[i].x, x[j].x, fmaf(x[i].y, x[j].y, acc[4 * k + 0]));
                        acc[4 * k + 1] = fmaf(x[i].y, x[j].x, fmaf(-x[i].x, x[j].y, acc[4 * k + 1]));
                        acc[4 * k + 2] = fmaf(y[i].x, y[j].x, fmaf(y[i].y, y[j].y, acc[4 * k + 2]));
                        acc[4 * k + 3] = fmaf(y[i].y, y[j].x, fmaf(-y[i].x, y[j].y, acc[4 * k + 3]));
                        ++k;
                    }
            }
        }
        // ordered reduction over the time lanes of each source
        for (int r = 0; r < lanes; ++r) {
            if (tl == r) {
#pragma unroll
                for (int i = 0; i < NV; ++i) red[sl][i][tx] = (r == 0 ? 0.f : red[sl][i][tx]) + acc[i];
            }
            __syncthreads();
        }
        if (fok) {
            for (int q = 0; q < ns; ++q) {
                float *dst = partial + (((size_t)((s0 + q) * B + b) * tsplit + split) * NV) * F + f;  // problem index = s*B + b
                for (int i = ty; i < NV; i += kTy) dst[(size_t)i * F] = red[q][i][tx];
            }
        }
        __syncthreads();
    }
}

// sum the T-split partials in a fixed order; which = 0 source, 1 noise.  Returns the full
// Hermitian matrix scaled by 1/T (the 0.5*(R+R^H) of tester.py:1092,1100 is implicit: only
// the upper triangle is accumulated and the diagonal is real by construction).
template <int M>
__device__ void load_scm(const float *__restrict__ partial, int bs, int f, int F, int tsplit, int which, double inv_t,
                         double (*Ar)[M], double (*Ai)[M]) {
    constexpr int NV = Tri<M>::NV;
    int k = 0;
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
        for (int j = i; j < M; ++j) {
            double re = 0.0, im = 0.0;
            for (int sp = 0; sp < tsplit; ++sp) {
                const float *p = partial + (((size_t)bs * tsplit + sp) * NV) * F + f;
                re += (double)p[(size_t)(4 * k + 2 * which) * F];
                im += (double)p[(size_t)(4 * k + 2 * which + 1) * F];
            }
            re *= inv_t;
            im *= inv_t;
            if (i == j) im = 0.0;
            Ar[i][j] = re;
            Ai[i][j] = im;
            Ar[j][i] = re;
            Ai[j][i] = -im;
            ++k;
        }
}

// any microphone count: one thread per problem, (p,q)-indexed sweep (dynamically indexed arrays)
template <int M>
__global__ void __launch_bounds__(64) eig_generic_kernel(const float *__restrict__ partial, double2 *__restrict__ steer, int nprob,
                                                 int F, int T, int tsplit, int B, int b0, int Bc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nprob) return;
    const int f = i % F, bl = i / F;
    const int bs = (bl / Bc) * B + b0 + bl % Bc;  // problems of the utterance chunk [b0, b0 + Bc): (s, b) -> s * B + b
    double Ar[M][M], Ai[M][M], Vr[M][M], Vi[M][M];
    load_scm<M>(partial, bs, f, F, tsplit, 0, 1.0 / (double)T, Ar, Ai);
    for (int p = 0; p < M; ++p)
        for (int q = 0; q < M; ++q) {
            Vr[p][q] = p == q ? 1.0 : 0.0;
            Vi[p][q] = 0.0;
        }
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0.0, dg = 0.0;
        for (int p = 0; p < M; ++p) {
            dg += Ar[p][p] * Ar[p][p];
            for (int q = p + 1; q < M; ++q) off += Ar[p][q] * Ar[p][q] + Ai[p][q] * Ai[p][q];
        }
        if (off <= 1e-30 * dg || off == 0.0) break;
        // unroll 1: nvcc 12.9 -O3 miscompiles the fully unrolled rotation sweep (verified on B200:
        // wrong eigenvectors at -O3, correct at -O0 and with rolled loops; tools/eig_test.cu)
#pragma unroll 1
        for (int p = 0; p < M - 1; ++p)
#pragma unroll 1
            for (int q = p + 1; q < M; ++q) {
                const double xr = Ar[p][q], xi = Ai[p][q];
                const double ax = sqrt(xr * xr + xi * xi);
                if (ax == 0.0) continue;
                const double tau = (Ar[q][q] - Ar[p][p]) / (2.0 * ax);
                const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                const double c = 1.0 / sqrt(1.0 + t * t), s = t * c;
                const double er = xr / ax, ei = -xi / ax;  // e = conj(x)/|x|
                const double ser = s * er, sei = s * ei, cer = c * er, cei = c * ei;
                // columns: X[:,p] = c X[:,p] - s e X[:,q];  X[:,q] = s X[:,p] + c e X[:,q]
                for (int r = 0; r < M; ++r) {
                    double pr = Ar[r][p], pi = Ai[r][p], qr = Ar[r][q], qi = Ai[r][q];
                    Ar[r][p] = c * pr - (ser * qr - sei * qi);
                    Ai[r][p] = c * pi - (ser * qi + sei * qr);
                    Ar[r][q] = s * pr + (cer * qr - cei * qi);
                    Ai[r][q] = s * pi + (cer * qi + cei * qr);
                    pr = Vr[r][p], pi = Vi[r][p], qr = Vr[r][q], qi = Vi[r][q];
                    Vr[r][p] = c * pr - (ser * qr - sei * qi);
                    Vi[r][p] = c * pi - (ser * qi + sei * qr);
                    Vr[r][q] = s * pr + (cer * qr - cei * qi);
                    Vi[r][q] = s * pi + (cer * qi + cei * qr);
                }
                // rows: A[p,:] = c A[p,:] - s conj(e) A[q,:];  A[q,:] = s A[p,:] + c conj(e) A[q,:]
                for (int r = 0; r < M; ++r) {
                    double pr = Ar[p][r], pi = Ai[p][r], qr = Ar[q][r], qi = Ai[q][r];
                    Ar[p][r] = c * pr - (ser * qr + sei * qi);
                    Ai[p][r] = c * pi - (ser * qi - sei * qr);
                    Ar[q][r] = s * pr + (cer * qr + cei * qi);
                    Ai[q][r] = s * pi + (cer * qi - cei * qr);
                }
                Ar[p][q] = Ai[p][q] = Ar[q][p] = Ai[q][p] = 0.0;
                Ai[p][p] = Ai[q][q] = 0.0;
            }
    }
    int kmax = 0;
    for (int p = 1; p < M; ++p)
        if (Ar[p][p] > Ar[kmax][kmax]) kmax = p;
    // d = v / v[0];  d *= sqrt(M / ||d||_2)     (tester.py:1119-1123)
    const double v0r = Vr[0][kmax], v0i = Vi[0][kmax];
    const double den = v0r * v0r + v0i * v0i;
    double dr[M], di[M], nrm = 0.0;
    for (int m = 0; m < M; ++m) {
        const double vr = Vr[m][kmax], vi = Vi[m][kmax];
        dr[m] = (vr * v0r + vi * v0i) / den;
        di[m] = (vi * v0r - vr * v0i) / den;
        nrm += dr[m] * dr[m] + di[m] * di[m];
    }
    const double g = sqrt((double)M / sqrt(nrm));
    for (int m = 0; m < M; ++m) steer[((size_t)bs * F + f) * M + m] = make_double2(dr[m] * g, di[m] * g);
}

// Principal eigenvector by cyclic Jacobi in ROUND-ROBIN order: a round rotates the three disjoint index pairs
// (0,1), (2,3), (4,5) at once (their 2x2 rotations commute), then a fixed permutation of the index positions brings
// three new pairs into those places; five rounds visit all 15 pairs and return the positions to their original order.
// The rotated pairs are therefore compile-time constants, so the Hermitian matrix (upper triangle) and the eigenvector
// rows live in registers (the (p,q)-indexed classic sweep needs dynamically indexed arrays = local memory, 6x slower),
// and the three rotations of a round give the fp64 pipe independent work.  Two adjacent lanes share a problem: both
// rotate the matrix, each accumulates three of the six eigenvector rows (V <- V J acts on rows independently).
template <int M>
__global__ void __launch_bounds__(64) eig6_kernel(const float *__restrict__ partial, double2 *__restrict__ steer, int nprob,
                                                  int F, int T, int tsplit, int B, int b0, int Bc) {
    const int gi = blockIdx.x * blockDim.x + threadIdx.x;
    const int h = gi & 1;
    const bool live = (gi >> 1) < nprob;
    const int i = live ? (gi >> 1) : nprob - 1;
    const int f = i % F, bl = i / F;
    const int bs = (bl / Bc) * B + b0 + bl % Bc;  // problems of the utterance chunk [b0, b0 + Bc): (s, b) -> s * B + b
    double Ar[M][M], Ai[M][M];  // only [i][j], i <= j is maintained after the load
    load_scm<M>(partial, bs, f, F, tsplit, 0, 1.0 / (double)T, Ar, Ai);
    double Vr[3][M], Vi[3][M];  // rows 3h .. 3h+2 of V
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int q = 0; q < M; ++q) {
            Vr[r][q] = (3 * h + r == q) ? 1.0 : 0.0;
            Vi[r][q] = 0.0;
        }
    constexpr int perm[6] = {0, 3, 1, 5, 2, 4};  // new position i holds the old position perm[i]
#pragma unroll 1
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0.0, dg = 0.0;
#pragma unroll
        for (int p = 0; p < M; ++p) {
            dg += Ar[p][p] * Ar[p][p];
#pragma unroll
            for (int q = p + 1; q < M; ++q) off += Ar[p][q] * Ar[p][q] + Ai[p][q] * Ai[p][q];
        }
        // the principal eigenvector feeds a complex64 beamformer (and the reference's np.linalg.eigh runs in single
        // precision, tester.py:1107): an off-diagonal mass of 1e-18 of the diagonal's bounds the eigenvector error near
        // 1e-9, far below both -- two sweeps fewer than the 1e-30 of the generic kernel (Jacobi converges quadratically)
        if (off <= 1e-18 * dg || off == 0.0) break;
#pragma unroll 1
        for (int rnd = 0; rnd < 5; ++rnd) {
            double c[3], s[3], er[3], ei[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const int p = 2 * a, q = 2 * a + 1;
                const double xr = Ar[p][q], xi = Ai[p][q];
                const double ax = sqrt(xr * xr + xi * xi);
                if (ax == 0.0) {
                    c[a] = 1.0;
                    s[a] = 0.0;
                    er[a] = 1.0;
                    ei[a] = 0.0;
                } else {
                    const double tau = (Ar[q][q] - Ar[p][p]) / (2.0 * ax);
                    const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                    c[a] = 1.0 / sqrt(1.0 + t * t);
                    s[a] = t * c[a];
                    er[a] = xr / ax;  // e = conj(x) / |x|
                    ei[a] = -xi / ax;
                    Ar[p][p] -= t * ax;
                    Ar[q][q] += t * ax;
                    Ar[p][q] = 0.0;
                    Ai[p][q] = 0.0;
                }
            }
            // off-diagonal 2x2 blocks: B <- Ja^H (B Jb), with J = [[c, s], [-s e, c e]]
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = a + 1; b < 3; ++b) {
                    const double sber = s[b] * er[b], sbei = s[b] * ei[b], cber = c[b] * er[b], cbei = c[b] * ei[b];
                    double tr[2][2], ti[2][2];
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        const double pr = Ar[2 * a + r][2 * b], pi = Ai[2 * a + r][2 * b];
                        const double qr = Ar[2 * a + r][2 * b + 1], qi = Ai[2 * a + r][2 * b + 1];
                        tr[r][0] = c[b] * pr - (sber * qr - sbei * qi);
                        ti[r][0] = c[b] * pi - (sber * qi + sbei * qr);
                        tr[r][1] = s[b] * pr + (cber * qr - cbei * qi);
                        ti[r][1] = s[b] * pi + (cber * qi + cbei * qr);
                    }
                    const double saer = s[a] * er[a], saei = s[a] * ei[a], caer = c[a] * er[a], caei = c[a] * ei[a];
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const double pr = tr[0][k], pi = ti[0][k], qr = tr[1][k], qi = ti[1][k];
                        // row p = c p - s conj(e) q ; row q = s p + c conj(e) q
                        Ar[2 * a][2 * b + k] = c[a] * pr - (saer * qr + saei * qi);
                        Ai[2 * a][2 * b + k] = c[a] * pi - (saer * qi - saei * qr);
                        Ar[2 * a + 1][2 * b + k] = s[a] * pr + (caer * qr + caei * qi);
                        Ai[2 * a + 1][2 * b + k] = s[a] * pi + (caer * qi - caei * qr);
                    }
                }
            // V <- V J on this lane's rows
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const int p = 2 * a, q = 2 * a + 1;
                const double ser = s[a] * er[a], sei = s[a] * ei[a], cer = c[a] * er[a], cei = c[a] * ei[a];
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const double pr = Vr[r][p], pi = Vi[r][p], qr = Vr[r][q], qi = Vi[r][q];
                    Vr[r][p] = c[a] * pr - (ser * qr - sei * qi);
                    Vi[r][p] = c[a] * pi - (ser * qi + sei * qr);
                    Vr[r][q] = s[a] * pr + (cer * qr - cei * qi);
                    Vi[r][q] = s[a] * pi + (cer * qi + cei * qr);
                }
            }
            // move the index positions: new [i][j] = old [perm i][perm j] (conjugated when that entry is below the diagonal)
            double Nr[M][M], Ni[M][M];
#pragma unroll
            for (int p = 0; p < M; ++p)
#pragma unroll
                for (int q = p; q < M; ++q) {
                    const int pp = perm[p], qq = perm[q];
                    if (pp <= qq) {
                        Nr[p][q] = Ar[pp][qq];
                        Ni[p][q] = Ai[pp][qq];
                    } else {
                        Nr[p][q] = Ar[qq][pp];
                        Ni[p][q] = -Ai[qq][pp];
                    }
                }
#pragma unroll
            for (int p = 0; p < M; ++p)
#pragma unroll
                for (int q = p; q < M; ++q) {
                    Ar[p][q] = Nr[p][q];
                    Ai[p][q] = p == q ? 0.0 : Ni[p][q];
                }
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                double wr[M], wi[M];
#pragma unroll
                for (int q = 0; q < M; ++q) {
                    wr[q] = Vr[r][perm[q]];
                    wi[q] = Vi[r][perm[q]];
                }
#pragma unroll
                for (int q = 0; q < M; ++q) {
                    Vr[r][q] = wr[q];
                    Vi[r][q] = wi[q];
                }
            }
        }
    }
    int kmax = 0;
    double best = Ar[0][0];
#pragma unroll
    for (int p = 1; p < M; ++p)
        if (Ar[p][p] > best) {
            best = Ar[p][p];
            kmax = p;
        }
    double vr[3], vi[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        vr[r] = Vr[r][0];
        vi[r] = Vi[r][0];
#pragma unroll
        for (int q = 1; q < M; ++q)
            if (q == kmax) {
                vr[r] = Vr[r][q];
                vi[r] = Vi[r][q];
            }
    }
    // d = v / v[0];  d *= sqrt(M / ||d||_2)     (tester.py:1119-1123); v[0] lives in the even lane
    const int src_lane = (threadIdx.x & 31) & ~1;
    const double v0r = __shfl_sync(0xffffffffu, vr[0], src_lane), v0i = __shfl_sync(0xffffffffu, vi[0], src_lane);
    const double den = v0r * v0r + v0i * v0i;
    double dr[3], di[3], nrm = 0.0;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        dr[r] = (vr[r] * v0r + vi[r] * v0i) / den;
        di[r] = (vi[r] * v0r - vr[r] * v0i) / den;
        nrm += dr[r] * dr[r] + di[r] * di[r];
    }
    const double other = __shfl_xor_sync(0xffffffffu, nrm, 1);
    nrm = h == 0 ? nrm + other : other + nrm;  // same summation order in both lanes
    const double g = sqrt((double)M / sqrt(nrm));
    if (live) {
#pragma unroll
        for (int r = 0; r < 3; ++r) steer[((size_t)bs * F + f) * M + 3 * h + r] = make_double2(dr[r] * g, di[r] * g);
    }
}

template <int M>
__global__ void __launch_bounds__(256) solve_kernel(const float *__restrict__ partial, const double2 *__restrict__ steer,
                                                    float2 *__restrict__ wout, int F, int T, int tsplit, double epsi, int B, int b0,
                                                    int Bc) {
    extern __shared__ double2 ph[];  // [F] phasors, then their prefix products
    const int bs = ((int)blockIdx.x / Bc) * B + b0 + (int)blockIdx.x % Bc;
    const double2 *d = steer + (size_t)bs * F * M;
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
        double2 p = make_double2(1.0, 0.0);
        if (f > 0) {
            double cr = 0.0, ci = 0.0;
#pragma unroll
            for (int m = 0; m < M; ++m) {
                double2 a = d[(size_t)f * M + m], bq = d[(size_t)(f - 1) * M + m];
                cr += a.x * bq.x + a.y * bq.y;  // a * conj(b)
                ci += a.y * bq.x - a.x * bq.y;
            }
            double n = sqrt(cr * cr + ci * ci);
            if (n > 0.0) p = make_double2(cr / n, -ci / n);  // exp(-j angle(c))
        }
        ph[f] = p;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double2 acc = ph[0];
        for (int f = 1; f < F; ++f) {
            double2 p = ph[f];
            acc = make_double2(acc.x * p.x - acc.y * p.y, acc.x * p.y + acc.y * p.x);
            ph[f] = acc;
        }
    }
    __syncthreads();
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
        double Ar[M][M], Ai[M][M];
        load_scm<M>(partial, bs, f, F, tsplit, 1, 1.0 / (double)T, Ar, Ai);
        double dr[M], di[M], ur[M], ui[M];
        const double2 P = ph[f];
#pragma unroll
        for (int m = 0; m < M; ++m) {
            double2 a = d[(size_t)f * M + m];
            dr[m] = a.x * P.x - a.y * P.y;
            di[m] = a.x * P.y + a.y * P.x;
            ur[m] = dr[m];
            ui[m] = di[m];
            Ar[m][m] += epsi;  // tester.py:1221
        }
        // Phi_n + epsi I is Hermitian positive definite, so Gaussian elimination needs no pivoting (the reference's
        // LAPACK gesv pivots, tester.py:1222; both are exact to rounding, and this one runs in fp64) and preserves the
        // Hermitian structure of the trailing block: only the upper triangle is touched, every index is a compile-time
        // constant, and the system stays in registers.
#pragma unroll
        for (int k = 0; k < M; ++k) {
            const double dk = Ar[k][k];
            const double inv = 1.0 / (fabs(dk) > 1e-300 ? dk : 1e-300);
#pragma unroll
            for (int r = k + 1; r < M; ++r) {
                // l = A[r][k] / A[k][k] = conj(A[k][r]) / A[k][k]
                const double lr = Ar[k][r] * inv, li = -Ai[k][r] * inv;
#pragma unroll
                for (int c = r; c < M; ++c) {
                    const double kr = Ar[k][c], ki = Ai[k][c];
                    Ar[r][c] -= lr * kr - li * ki;
                    Ai[r][c] -= lr * ki + li * kr;
                }
                Ai[r][r] = 0.0;
                const double uk = ur[k], uik = ui[k];
                ur[r] -= lr * uk - li * uik;
                ui[r] -= lr * uik + li * uk;
            }
        }
#pragma unroll
        for (int k = M - 1; k >= 0; --k) {
            double sr = ur[k], si = ui[k];
#pragma unroll
            for (int c = k + 1; c < M; ++c) {
                sr -= Ar[k][c] * ur[c] - Ai[k][c] * ui[c];
                si -= Ar[k][c] * ui[c] + Ai[k][c] * ur[c];
            }
            const double dk = Ar[k][k];
            const double inv = 1.0 / (fabs(dk) > 1e-300 ? dk : 1e-300);
            ur[k] = sr * inv;
            ui[k] = si * inv;
        }
        // w = u / (d^H u)
        double nr = 0.0, ni = 0.0;
#pragma unroll
        for (int m = 0; m < M; ++m) {
            nr += dr[m] * ur[m] + di[m] * ui[m];
            ni += dr[m] * ui[m] - di[m] * ur[m];
        }
        const double nd = nr * nr + ni * ni;
#pragma unroll
        for (int m = 0; m < M; ++m) {
            const double wr = (ur[m] * nr + ui[m] * ni) / nd;
            const double wi = (ui[m] * nr - ur[m] * ni) / nd;
            wout[((size_t)bs * F + f) * M + m] = make_float2((float)wr, (float)wi);
        }
    }
}

constexpr int kMaxS = 4;
constexpr int kApplyTile = 64;

template <int M>
__global__ void __launch_bounds__(kFx *kTy) apply_kernel(const float2 *__restrict__ mix, int64_t sb, int64_t sm, int64_t st,
                                                          int64_t sf, const float2 *__restrict__ w, float2 *__restrict__ out,
                                                          int S, int B, int T, int F, int b0) {
    const int f = blockIdx.x * kFx + threadIdx.x;
    const int b = blockIdx.z + b0;
    if (f >= F) return;
    float2 wv[kMaxS][M];
#pragma unroll
    for (int s = 0; s < kMaxS; ++s)
        if (s < S) {
#pragma unroll
            for (int m = 0; m < M; ++m) wv[s][m] = w[((size_t)(s * B + b) * F + f) * M + m];
        }
    const int t0 = blockIdx.y * kApplyTile;
    const int t1 = min(T, t0 + kApplyTile);
    const float2 *mix_b = mix + b * sb + (int64_t)f * sf;
    const uint64_t pol = l2_policy_evict_first();  // last use of the mixture
    for (int t = t0 + threadIdx.y; t < t1; t += kTy) {
        float2 x[M];
#pragma unroll
        for (int m = 0; m < M; ++m) x[m] = ld_hint(mix_b + m * sm + t * st, pol);
#pragma unroll
        for (int s = 0; s < kMaxS; ++s)
            if (s < S) {
                float yr = 0.f, yi = 0.f;
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    // conj(w) * x
                    yr = fmaf(wv[s][m].x, x[m].x, fmaf(wv[s][m].y, x[m].y, yr));
                    yi = fmaf(wv[s][m].x, x[m].y, fmaf(-wv[s][m].y, x[m].x, yi));
                }
                out[((size_t)(s * B + b) * T + t) * F + f] = make_float2(yr, yi);
            }
    }
}

int pick_tsplit(int B, int F) {
    int ctas = ceil_div(F, kFx) * B;
    int ts = ceil_div(2 * 148, ctas);
    return ts < 1 ? 1 : (ts > 8 ? 8 : ts);
}

struct MvdrWs {
    float *partial;
    double2 *steer;
    float2 *w;
    size_t total;
};

template <int M>
MvdrWs carve(char *base, int S, int B, int T, int F, int tsplit) {
    MvdrWs ws;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        char *p = base ? base + off : nullptr;
        off += align_up(bytes, 256);
        return p;
    };
    ws.partial = reinterpret_cast<float *>(take((size_t)B * S * tsplit * Tri<M>::NV * F * sizeof(float)));
    ws.steer = reinterpret_cast<double2 *>(take((size_t)B * S * F * M * sizeof(double2)));
    ws.w = reinterpret_cast<float2 *>(take((size_t)B * S * F * M * sizeof(float2)));
    ws.total = off;
    return ws;
}

template <int M>
int run(const void *d_src, int64_t src_ss, const void *d_mix, int64_t sb, int64_t sm, int64_t st, int64_t sf, void *d_out,
        void *d_weights, int S, int B, int T, int F, float epsi, void *d_ws, size_t ws_bytes, cudaStream_t stream) {
    const int tsplit = pick_tsplit(B, F);
    MvdrWs ws = carve<M>(reinterpret_cast<char *>(d_ws), S, B, T, F, tsplit);
    if (ws.total > ws_bytes) {
        set_error("miso_mvdr_fwd: workspace %zu < required %zu bytes", ws_bytes, ws.total);
        return MISO_E_WORKSPACE;
    }
    float2 *w = d_weights ? reinterpret_cast<float2 *>(d_weights) : ws.w;
    const float2 *src = reinterpret_cast<const float2 *>(d_src);
    const float2 *mix = reinterpret_cast<const float2 *>(d_mix);
    dim3 blk(kFx, kTy);
    // Optional utterance chunks (MISO_MVDR_L2_MB) sized so that a chunk's mixture, read by the covariance pass and again by
    // filter-and-sum, could stay in L2 between the two passes.  Measured on B200 (profiles/r2_mvdr_*.json): the eigenvector
    // and solve kernels are latency bound (75 / 46 us whatever the problem count), so every extra chunk costs more than
    // the 197 MB of DRAM reads it saves -- 1.08 ms with 40 MB chunks against 0.36 ms unchunked.  Default: one chunk.
    // tsplit and the problem indexing are those of the whole batch, so the result does not depend on the chunking.
    const size_t mix_bytes = (size_t)M * T * F * sizeof(float2);
    static const size_t l2_budget = getenv("MISO_MVDR_L2_MB") ? (size_t)atoi(getenv("MISO_MVDR_L2_MB")) << 20 : ~(size_t)0;
    const int Bc_max = (int)std::max<size_t>(1, std::min<size_t>((size_t)B, l2_budget / std::max<size_t>(mix_bytes, 1)));
    for (int b0 = 0; b0 < B; b0 += Bc_max) {
        const int Bc = std::min(Bc_max, B - b0);
        prof_begin(stream);
        scm_kernel<M><<<dim3(ceil_div(F, kFx), tsplit, Bc), blk, 0, stream>>>(src, src_ss, mix, sb, sm, st, sf, ws.partial, S, B, T, F, tsplit,
                                                                           b0);
        MISO_LAUNCHED("scm_kernel");
        // problems are ordered (s, b, f): bs = s*B + b, matching the [S,B,...] outputs
        const int nprob = Bc * S * F;
        if constexpr (M == 6)
            eig6_kernel<M><<<ceil_div(2 * nprob, 64), 64, 0, stream>>>(ws.partial, ws.steer, nprob, F, T, tsplit, B, b0, Bc);
        else
            eig_generic_kernel<M><<<ceil_div(nprob, 64), 64, 0, stream>>>(ws.partial, ws.steer, nprob, F, T, tsplit, B, b0, Bc);
        MISO_LAUNCHED("eig_kernel");
        solve_kernel<M><<<Bc * S, 256, (size_t)F * sizeof(double2), stream>>>(ws.partial, ws.steer, w, F, T, tsplit, (double)epsi, B, b0, Bc);
        MISO_LAUNCHED("solve_kernel");
        apply_kernel<M><<<dim3(ceil_div(F, kFx), ceil_div(T, kApplyTile), Bc), blk, 0, stream>>>(
            mix, sb, sm, st, sf, w, reinterpret_cast<float2 *>(d_out), S, B, T, F, b0);
        MISO_LAUNCHED("apply_kernel");
        // algorithmic bytes: mixture + S sources in, S outputs out (160 T F per utterance for S = 2, M = 6)
        prof_end(stream, 0.0, (double)Bc * T * F * sizeof(float2) * ((double)M * (1 + S) + S), MISO_PROF_MVDR);
    }
    return MISO_OK;
}

// Staged form: the covariance sums of one recording may be spread over ranks (utterance-level MVDR of a chunked
// recording, tester.py:425-449): every rank computes its frames' partial sums (same layout as the fused path:
// [S*B][split][NV][F]), the partial sums of all ranks are concatenated along the split axis, and the eigenvector /
// solve kernels add them in that fixed order in fp64.
template <int M>
int run_scm(const void *d_src, int64_t src_ss, const void *d_mix, int64_t sb, int64_t sm, int64_t st, int64_t sf, float *d_partial,
            int S, int B, int T, int F, cudaStream_t stream) {
    const int tsplit = pick_tsplit(B, F);
    scm_kernel<M><<<dim3(ceil_div(F, kFx), tsplit, B), dim3(kFx, kTy), 0, stream>>>(
        reinterpret_cast<const float2 *>(d_src), src_ss, reinterpret_cast<const float2 *>(d_mix), sb, sm, st, sf, d_partial, S, B, T, F, tsplit, 0);
    MISO_LAUNCHED("scm_kernel");
    return MISO_OK;
}

template <int M>
int run_weights(const float *d_partial, int nsplit, int T_total, void *d_weights, int S, int B, int F, float epsi, void *d_ws,
                size_t ws_bytes, cudaStream_t stream) {
    const size_t need = (size_t)B * S * F * M * sizeof(double2);
    if (ws_bytes < need) {
        set_error("miso_mvdr_weights: workspace %zu < required %zu bytes", ws_bytes, need);
        return MISO_E_WORKSPACE;
    }
    double2 *steer = reinterpret_cast<double2 *>(d_ws);
    const int nprob = B * S * F;
    if constexpr (M == 6)
        eig6_kernel<M><<<ceil_div(2 * nprob, 64), 64, 0, stream>>>(d_partial, steer, nprob, F, T_total, nsplit, B, 0, B);
    else
        eig_generic_kernel<M><<<ceil_div(nprob, 64), 64, 0, stream>>>(d_partial, steer, nprob, F, T_total, nsplit, B, 0, B);
    MISO_LAUNCHED("eig_kernel");
    solve_kernel<M><<<B * S, 256, (size_t)F * sizeof(double2), stream>>>(d_partial, steer, reinterpret_cast<float2 *>(d_weights), F, T_total,
                                                                      nsplit, (double)epsi, B, 0, B);
    MISO_LAUNCHED("solve_kernel");
    return MISO_OK;
}

template <int M>
int run_apply(const void *d_mix, int64_t sb, int64_t sm, int64_t st, int64_t sf, const void *d_weights, void *d_out, int S, int B, int T,
              int F, cudaStream_t stream) {
    apply_kernel<M><<<dim3(ceil_div(F, kFx), ceil_div(T, kApplyTile), B), dim3(kFx, kTy), 0, stream>>>(
        reinterpret_cast<const float2 *>(d_mix), sb, sm, st, sf, reinterpret_cast<const float2 *>(d_weights), reinterpret_cast<float2 *>(d_out), S,
        B, T, F, 0);
    MISO_LAUNCHED("apply_kernel");
    return MISO_OK;
}

template <int M>
size_t ws_bytes_for(int S, int B, int T, int F) {
    return carve<M>(nullptr, S, B, T, F, pick_tsplit(B, F)).total;
}

}  // namespace
}  // namespace miso

using namespace miso;

extern "C" {

size_t miso_mvdr_workspace_bytes(int S, int B, int M, int T, int F) {
    if (S < 1 || B < 1 || T < 1 || F < 1) return 0;
    switch (M) {
        case 2: return ws_bytes_for<2>(S, B, T, F);
        case 3: return ws_bytes_for<3>(S, B, T, F);
        case 4: return ws_bytes_for<4>(S, B, T, F);
        case 5: return ws_bytes_for<5>(S, B, T, F);
        case 6: return ws_bytes_for<6>(S, B, T, F);
        case 7: return ws_bytes_for<7>(S, B, T, F);
        case 8: return ws_bytes_for<8>(S, B, T, F);
    }
    return 0;
}

int miso_mvdr_fwd(const void *d_src, int64_t src_ss, const void *d_mix, int64_t sb, int64_t sm, int64_t st, int64_t sf,
                  void *d_out, void *d_weights, int S, int B, int M, int T, int F, float epsi, void *d_ws, size_t ws_bytes,
                  void *stream) {
    MISO_REQUIRE(d_src && d_mix && d_out && d_ws, "miso_mvdr_fwd: null argument");
    MISO_REQUIRE(S >= 1 && S <= kMaxS, "miso_mvdr_fwd: S=%d unsupported (1..%d)", S, kMaxS);
    MISO_REQUIRE(M >= 2 && M <= 8, "miso_mvdr_fwd: M=%d unsupported (2..8)", M);
    MISO_REQUIRE(B >= 1 && B <= 65535 && T >= 1 && F >= 1 && F <= 2048, "miso_mvdr_fwd: bad shape B=%d T=%d F=%d", B, T, F);
    cudaStream_t s = as_stream(stream);
#define MISO_MVDR_CASE(m) \
    case m: return run<m>(d_src, src_ss, d_mix, sb, sm, st, sf, d_out, d_weights, S, B, T, F, epsi, d_ws, ws_bytes, s)
    switch (M) {
        MISO_MVDR_CASE(2);
        MISO_MVDR_CASE(3);
        MISO_MVDR_CASE(4);
        MISO_MVDR_CASE(5);
        MISO_MVDR_CASE(6);
        MISO_MVDR_CASE(7);
        MISO_MVDR_CASE(8);
    }
#undef MISO_MVDR_CASE
    return MISO_E_ARG;
}

int miso_mvdr_tsplit(int B, int F) { return (B < 1 || F < 1) ? -1 : pick_tsplit(B, F); }

size_t miso_mvdr_partial_bytes(int S, int B, int M, int F) {
    if (S < 1 || B < 1 || F < 1 || M < 2 || M > 8) return 0;
    return (size_t)S * B * pick_tsplit(B, F) * (size_t)(2 * M * (M + 1)) * F * sizeof(float);
}

#define MISO_MVDR_SWITCH(call)                 \
    switch (M) {                               \
        case 2: return call(2);                \
        case 3: return call(3);                \
        case 4: return call(4);                \
        case 5: return call(5);                \
        case 6: return call(6);                \
        case 7: return call(7);                \
        case 8: return call(8);                \
    }                                          \
    return MISO_E_ARG

int miso_mvdr_scm(const void *d_src, int64_t src_ss, const void *d_mix, int64_t sb, int64_t sm, int64_t st, int64_t sf, void *d_partial,
                  int S, int B, int M, int T, int F, void *stream) {
    MISO_REQUIRE(d_src && d_mix && d_partial, "miso_mvdr_scm: null argument");
    MISO_REQUIRE(S >= 1 && S <= kMaxS && M >= 2 && M <= 8, "miso_mvdr_scm: S=%d M=%d unsupported", S, M);
    MISO_REQUIRE(B >= 1 && B <= 65535 && T >= 1 && F >= 1 && F <= 2048, "miso_mvdr_scm: bad shape B=%d T=%d F=%d", B, T, F);
#define MISO_CALL(m) run_scm<m>(d_src, src_ss, d_mix, sb, sm, st, sf, reinterpret_cast<float *>(d_partial), S, B, T, F, as_stream(stream))
    MISO_MVDR_SWITCH(MISO_CALL);
#undef MISO_CALL
}

int miso_mvdr_weights(const void *d_partial, int nsplit, int T_total, void *d_weights, int S, int B, int M, int F, float epsi, void *d_ws,
                      size_t ws_bytes, void *stream) {
    MISO_REQUIRE(d_partial && d_weights && d_ws, "miso_mvdr_weights: null argument");
    MISO_REQUIRE(S >= 1 && S <= kMaxS && M >= 2 && M <= 8, "miso_mvdr_weights: S=%d M=%d unsupported", S, M);
    MISO_REQUIRE(B >= 1 && nsplit >= 1 && T_total >= 1 && F >= 1 && F <= 2048, "miso_mvdr_weights: bad shape");
#define MISO_CALL(m) run_weights<m>(reinterpret_cast<const float *>(d_partial), nsplit, T_total, d_weights, S, B, F, epsi, d_ws, ws_bytes, as_stream(stream))
    MISO_MVDR_SWITCH(MISO_CALL);
#undef MISO_CALL
}

int miso_mvdr_apply(const void *d_mix, int64_t sb, int64_t sm, int64_t st, int64_t sf, const void *d_weights, void *d_out, int S, int B,
                    int M, int T, int F, void *stream) {
    MISO_REQUIRE(d_mix && d_weights && d_out, "miso_mvdr_apply: null argument");
    MISO_REQUIRE(S >= 1 && S <= kMaxS && M >= 2 && M <= 8, "miso_mvdr_apply: S=%d M=%d unsupported", S, M);
    MISO_REQUIRE(B >= 1 && B <= 65535 && T >= 1 && F >= 1, "miso_mvdr_apply: bad shape");
#define MISO_CALL(m) run_apply<m>(d_mix, sb, sm, st, sf, d_weights, d_out, S, B, T, F, as_stream(stream))
    MISO_MVDR_SWITCH(MISO_CALL);
#undef MISO_CALL
}

}  // extern "C"
