// Multi-channel STFT front end (S1 of SURVEY.md section 8(a); reference
// dataloader/data.py:49-66,77-79 = tester.py:992-1012): zero-padded framing, periodic
// hann window, unnormalised real FFT.  One warp transforms one (b, mic, frame) with a
// shared-memory radix-2 FFT; twiddles and the window are computed once per CTA in fp64.
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
    const int64_t g = (int64_t)blockIdx.x * kWarps + warp;
    if (g >= (int64_t)B * M * T) return;
    const int t = (int)(g % T);
    const int m = (int)((g / T) % M);
    const int b = (int)(g / ((int64_t)T * M));
    float2 *bf = buf[warp];
    const float *xb = x + b * sb + m * sm;
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

}  // namespace
}  // namespace miso

using namespace miso;

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
    const unsigned blocks = (unsigned)((frames + kWarps - 1) / kWarps);
    cudaStream_t st = as_stream(stream);
    if (nperseg == 256)
        stft_kernel<256><<<blocks, kWarps * 32, 0, st>>>(d_x, sb, sn, sm, reinterpret_cast<float2 *>(d_out), B, N, M, T, hop);
    else
        stft_kernel<512><<<blocks, kWarps * 32, 0, st>>>(d_x, sb, sn, sm, reinterpret_cast<float2 *>(d_out), B, N, M, T, hop);
    MISO_LAUNCHED("stft_kernel");
    return MISO_OK;
}

}  // extern "C"
