// Multi-channel STFT front end and its inverse (S1 of SURVEY.md section 8(a), ISTFT of 8(f) rank 2; reference
// dataloader/data.py:49-66,77-79 = tester.py:992-1012): zero-padded framing, periodic
// hann window, unnormalised real FFT.  One warp transforms one (b, mic, frame) with a
// shared-memory radix-2 FFT; twiddles and the window are computed once per CTA in fp64.
#include <algorithm>

#include "common.cuh"

namespace miso {
namespace {

constexpr int kWarps = 4;

template <int N>
__global__ void __launch_bounds__(kWarps * 32) stft_kernel(const float *__restrict__ x, int64_t sb, int64_t sn, int64_t sm,
                                                           float2 *__restrict__ out, int B, int Ns, int M, int T, int hop) {
    constexpr int LOG2N = (N == 256) ? 8 : 9;
    __shared__ float2 tw[N / 2];
    __shared__ float win[N];
    __shared__ float2 buf[kWarps][N];
    for (int i = threadIdx.x; i < N / 2; i += blockDim.x) {
        double s, c;
        sincospi(2.0 * (double)i / (double)N, &s, &c);
        tw[i] = make_float2((float)c, (float)(-s));
    }
    for (int i = threadIdx.x; i < N; i += blockDim.x) win[i] = (float)(0.5 - 0.5 * cospi(2.0 * (double)i / (double)N));
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float2 *bf = buf[warp];
    // the fp64 twiddle / window tables above cost more than a frame's transform: a CTA keeps them for many frames
    for (int64_t g = (int64_t)blockIdx.x * kWarps + warp; g < (int64_t)B * M * T; g += (int64_t)gridDim.x * kWarps) {
    const int t = (int)(g % T);
    const int m = (int)((g / T) % M);
    const int b = (int)(g / ((int64_t)T * M));
    const float *xb = x + b * sb + m * sm;
    __syncwarp();
    for (int i = lane; i < N; i += 32) {
        int p = t * hop + i - N / 2;
        float v = (p >= 0 && p < Ns) ? xb[(int64_t)p * sn] * win[i] : 0.f;
        int r = (int)(__brev((unsigned)i) >> (32 - LOG2N));
        bf[r] = make_float2(v, 0.f);
    }
    __syncwarp();
#pragma unroll
    for (int s = 0; s < LOG2N; ++s) {
        const int half = 1 << s;
        for (int j = lane; j < N / 2; j += 32) {
            int pos = j & (half - 1);
            int i0 = ((j >> s) << (s + 1)) + pos;
            int i1 = i0 + half;
            float2 w = tw[pos << (LOG2N - 1 - s)];
            float2 a = bf[i0], c = bf[i1];
            float2 wc = make_float2(w.x * c.x - w.y * c.y, w.x * c.y + w.y * c.x);
            bf[i0] = make_float2(a.x + wc.x, a.y + wc.y);
            bf[i1] = make_float2(a.x - wc.x, a.y - wc.y);
        }
        __syncwarp();
    }
    float2 *o = out + (size_t)g * (N / 2 + 1);
    for (int k = lane; k <= N / 2; k += 32) o[k] = bf[k];
    }
}

// Inverse: one warp rebuilds one (signal, frame): Hermitian extension of the one-sided spectrum, the same radix-2
// network with conjugated twiddles, real part / N, times the synthesis window -> frames[g][N].
template <int N>
__global__ void __launch_bounds__(kWarps * 32) istft_frames_kernel(const float2 *__restrict__ spec, int64_t ss, int64_t st, int64_t sf,
                                                                   float *__restrict__ frames, int S, int T) {
    constexpr int LOG2N = (N == 256) ? 8 : 9;
    __shared__ float2 tw[N / 2];
    __shared__ float win[N];
    __shared__ float2 buf[kWarps][N];
    for (int i = threadIdx.x; i < N / 2; i += blockDim.x) {
        double s, c;
        sincospi(2.0 * (double)i / (double)N, &s, &c);
        tw[i] = make_float2((float)c, (float)s);
    }
    for (int i = threadIdx.x; i < N; i += blockDim.x) win[i] = (float)(0.5 - 0.5 * cospi(2.0 * (double)i / (double)N));
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float2 *bf = buf[warp];
    for (int64_t g = (int64_t)blockIdx.x * kWarps + warp; g < (int64_t)S * T; g += (int64_t)gridDim.x * kWarps) {
    const int t = (int)(g % T);
    const int sidx = (int)(g / T);
    const float2 *x = spec + sidx * ss + t * st;
    __syncwarp();
    for (int k = lane; k <= N / 2; k += 32) {
        const float2 v = x[k * sf];
        bf[(int)(__brev((unsigned)k) >> (32 - LOG2N))] = v;
        if (k > 0 && k < N / 2) bf[(int)(__brev((unsigned)(N - k)) >> (32 - LOG2N))] = make_float2(v.x, -v.y);
    }
    __syncwarp();
#pragma unroll
    for (int s = 0; s < LOG2N; ++s) {
        const int half = 1 << s;
        for (int j = lane; j < N / 2; j += 32) {
            int pos = j & (half - 1);
            int i0 = ((j >> s) << (s + 1)) + pos;
            int i1 = i0 + half;
            float2 w = tw[pos << (LOG2N - 1 - s)];
            float2 a = bf[i0], c = bf[i1];
            float2 wc = make_float2(w.x * c.x - w.y * c.y, w.x * c.y + w.y * c.x);
            bf[i0] = make_float2(a.x + wc.x, a.y + wc.y);
            bf[i1] = make_float2(a.x - wc.x, a.y - wc.y);
        }
        __syncwarp();
    }
    float *o = frames + (size_t)g * N;
    for (int i = lane; i < N; i += 32) o[i] = bf[i].x * (1.f / (float)N) * win[i];
    }
}

// Overlap-add as a gather in a fixed order (deterministic), divided by the summed squared window, with the
// nperseg/2 boundary samples trimmed (scipy.signal.istft defaults; tester.py:979-990).
template <int N>
__global__ void __launch_bounds__(256) istft_ola_kernel(const float *__restrict__ frames, float *__restrict__ out, int S, int T, int hop,
                                                        int n_out) {
    __shared__ float w2[N];
    for (int k = threadIdx.x; k < N; k += blockDim.x) {
        const float w = (float)(0.5 - 0.5 * cospi(2.0 * (double)k / (double)N));
        w2[k] = w * w;
    }
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (int64_t)S * n_out; i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i % n_out);
    const int sidx = (int)(i / n_out);
    const int p = n + N / 2;  // position in the untrimmed signal
    int t_hi = p / hop;
    if (t_hi > T - 1) t_hi = T - 1;
    int t_lo = (p - N + hop) / hop;  // first frame with t*hop + N > p
    if (p - N + 1 <= 0) t_lo = 0;
    float acc = 0.f, norm = 0.f;
    for (int t = t_lo; t <= t_hi; ++t) {
        const int k = p - t * hop;
        if (k < 0 || k >= N) continue;
        acc += frames[((size_t)sidx * T + t) * N + k];
        norm += w2[k];
    }
    out[i] = acc / (norm > 1e-10f ? norm : 1.f);
    }
}

}  // namespace
}  // namespace miso

using namespace miso;

// Output sample format of the reference's wav writer (tester.py:155-157, 444-446, 950-952): wave * MaxINT16 in double,
// then numpy's astype(np.int16) = truncation toward zero (saturated here; numpy's overflow is undefined).
__global__ void wave_to_int16_kernel(const float *__restrict__ x, int16_t *__restrict__ out, int64_t n, double scale) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double v = trunc((double)x[i] * scale);
        v = fmin(fmax(v, -32768.0), 32767.0);
        out[i] = (int16_t)v;
    }
}

extern "C" {

int miso_stft_num_frames(int n_samples, int nperseg, int hop) {
    // scipy.signal.stft defaults: boundary='zeros' pads nperseg/2 both sides, padded=True
    // zero-extends to a whole number of hops (dataloader/data.py:58).
    if (n_samples < 0 || nperseg <= 0 || hop <= 0 || hop > nperseg) return -1;
    int64_t total = (int64_t)n_samples + 2 * (nperseg / 2);
    int64_t rem = (total - nperseg) % hop;
    int64_t nadd = ((hop - rem) % hop) % nperseg;
    total += nadd;
    return (int)((total - (nperseg - hop)) / hop);
}

int miso_stft_fwd(const float *d_x, int64_t sb, int64_t sn, int64_t sm, void *d_out, int B, int N, int M, int nperseg,
                  int hop, void *stream) {
    MISO_REQUIRE(d_x && d_out, "miso_stft_fwd: null argument");
    MISO_REQUIRE(nperseg == 256 || nperseg == 512, "miso_stft_fwd: nperseg=%d unsupported (256 or 512)", nperseg);
    MISO_REQUIRE(hop > 0 && hop <= nperseg, "miso_stft_fwd: bad hop %d", hop);
    MISO_REQUIRE(B >= 1 && M >= 1 && N >= 1, "miso_stft_fwd: bad shape");
    const int T = miso_stft_num_frames(N, nperseg, hop);
    const int64_t frames = (int64_t)B * M * T;
    const unsigned blocks = (unsigned)std::min<int64_t>((frames + kWarps - 1) / kWarps, 148 * 16);
    cudaStream_t st = as_stream(stream);
    if (nperseg == 256)
        stft_kernel<256><<<blocks, kWarps * 32, 0, st>>>(d_x, sb, sn, sm, reinterpret_cast<float2 *>(d_out), B, N, M, T, hop);
    else
        stft_kernel<512><<<blocks, kWarps * 32, 0, st>>>(d_x, sb, sn, sm, reinterpret_cast<float2 *>(d_out), B, N, M, T, hop);
    MISO_LAUNCHED("stft_kernel");
    return MISO_OK;
}

int miso_istft_num_samples(int T, int nperseg, int hop) {
    if (T < 1 || nperseg <= 0 || hop <= 0 || hop > nperseg) return -1;
    return (T - 1) * hop;  // nperseg + (T-1) hop minus the two trimmed half windows
}

size_t miso_istft_workspace_bytes(int S, int T, int nperseg) { return (size_t)S * T * nperseg * sizeof(float); }

int miso_istft_fwd(const void *d_spec, int64_t ss, int64_t st, int64_t sf, float *d_out, int S, int T, int nperseg, int hop,
                   void *d_ws, size_t ws_bytes, void *stream) {
    MISO_REQUIRE(d_spec && d_out && d_ws, "miso_istft_fwd: null argument");
    MISO_REQUIRE(nperseg == 256 || nperseg == 512, "miso_istft_fwd: nperseg=%d unsupported (256 or 512)", nperseg);
    MISO_REQUIRE(hop > 0 && hop <= nperseg, "miso_istft_fwd: bad hop %d", hop);
    MISO_REQUIRE(S >= 1 && T >= 1, "miso_istft_fwd: bad shape");
    if (ws_bytes < miso_istft_workspace_bytes(S, T, nperseg)) {
        set_error("miso_istft_fwd: workspace too small (%zu < %zu bytes)", ws_bytes, miso_istft_workspace_bytes(S, T, nperseg));
        return MISO_E_WORKSPACE;
    }
    const int n_out = miso_istft_num_samples(T, nperseg, hop);
    cudaStream_t stq = as_stream(stream);
    float *frames = reinterpret_cast<float *>(d_ws);
    const int64_t nfr = (int64_t)S * T;
    const unsigned blocks = (unsigned)std::min<int64_t>((nfr + kWarps - 1) / kWarps, 148 * 16);
    const unsigned oblocks = (unsigned)std::min<int64_t>(((int64_t)S * n_out + 255) / 256, 148 * 32);
    if (nperseg == 256) {
        istft_frames_kernel<256><<<blocks, kWarps * 32, 0, stq>>>(reinterpret_cast<const float2 *>(d_spec), ss, st, sf, frames, S, T);
        MISO_LAUNCHED("istft_frames_kernel");
        if (n_out > 0) istft_ola_kernel<256><<<oblocks, 256, 0, stq>>>(frames, d_out, S, T, hop, n_out);
    } else {
        istft_frames_kernel<512><<<blocks, kWarps * 32, 0, stq>>>(reinterpret_cast<const float2 *>(d_spec), ss, st, sf, frames, S, T);
        MISO_LAUNCHED("istft_frames_kernel");
        if (n_out > 0) istft_ola_kernel<512><<<oblocks, 256, 0, stq>>>(frames, d_out, S, T, hop, n_out);
    }
    MISO_LAUNCHED("istft_ola_kernel");
    return MISO_OK;
}

int miso_wave_to_int16(const float *d_x, int16_t *d_out, int64_t n, float scale, void *stream) {
    MISO_REQUIRE(d_x && d_out && n >= 0, "miso_wave_to_int16: bad argument");
    if (n == 0) return MISO_OK;
    const int blocks = (int)std::min<int64_t>((n + 255) / 256, 148 * 16);
    wave_to_int16_kernel<<<blocks, 256, 0, as_stream(stream)>>>(d_x, d_out, n, (double)scale);
    MISO_LAUNCHED("wave_to_int16_kernel");
    return MISO_OK;
}

}  // extern "C"
