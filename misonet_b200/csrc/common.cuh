// Shared helpers for the misonet_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>
#include <vector>

#include "../../include/misonet_b200.h"

namespace miso {

// ---- error reporting (thread-local message, C ABI returns an int) ---------------
void set_error(const char *fmt, ...);
extern std::atomic<uint64_t> g_launch_count;

inline int cuda_fail(cudaError_t e, const char *what) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return MISO_E_CUDA;
}

#define MISO_CUDA(call)                                                      \
    do {                                                                     \
        cudaError_t _e = (call);                                             \
        if (_e != cudaSuccess) return miso::cuda_fail(_e, #call);            \
    } while (0)

// call after every kernel launch
#define MISO_LAUNCHED(name)                                                  \
    do {                                                                     \
        miso::g_launch_count.fetch_add(1, std::memory_order_relaxed);        \
        cudaError_t _e = cudaGetLastError();                                 \
        if (_e != cudaSuccess) return miso::cuda_fail(_e, name);             \
    } while (0)

#define MISO_REQUIRE(cond, ...)                                              \
    do {                                                                     \
        if (!(cond)) {                                                       \
            miso::set_error(__VA_ARGS__);                                    \
            return MISO_E_ARG;                                               \
        }                                                                    \
    } while (0)

// optional per-launch event timing of the dominant kernel family (runtime.cu)
struct ProfRec {
    cudaEvent_t a, b;
    double flops, bytes;
    double exec_flops;  // tensor-pipe work actually issued (3 MMAs per product in bf16x3, N / M / K padding included)
    int family;
    bool persistent;
};
bool prof_enabled();
void prof_capture(std::vector<ProfRec> *sink);
void prof_replayed(const std::vector<ProfRec> &recs);
void prof_begin(cudaStream_t st);
void prof_end(cudaStream_t st, double flops, double bytes, int family, double exec_flops = 0.0);

// Programmatic dependent launch: a kernel launched through launch_pdl may start while its predecessor in the stream is
// still running; it must execute pdl_wait() before touching anything the predecessor reads or writes (everything before
// that point -- barrier / tensor-memory set-up -- overlaps the predecessor's tail).  pdl_trigger() lets the successor's
// launch begin.  Off by default (it measured 2.6 % slower on the bench workload: the early-resident dependents compete with
// the predecessor's tail); MISO_PDL=1 turns the attribute on (without it the device-side instructions are no-ops).
bool pdl_enabled();
int pdl_level();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_if(bool on, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = on ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
    return launch_pdl_if(pdl_level() == 1, kernel, grid, block, smem, st, static_cast<Args &&>(args)...);
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---- device helpers -----------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float elu1(float x) { return x > 0.f ? x : expm1f(x); }

// Statistics convention: every activation channel owns two 64-bit accumulators (sum, sum of squares) in
// FIXED POINT (value * 2^20 as int64): integer atomics make the totals independent of the order in which
// CTAs arrive, so a forward is bitwise reproducible (fp64 atomics are not: a last-bit difference in a sum
// occasionally flips the bf16 rounding of an activation downstream).  The buffers are declared double*
// for historical reasons and only touched through stat_add / stat_get / stat_set.  A negative
// sum-of-squares marks a channel that is consumed raw (no normalisation in the reference at that point).
// Range: |sum| < 2^43 ~ 8.8e12, i.e. a channel RMS up to ~8e3 over a 500 x 257 plane -- unit-scale spectra (the
// reference divides the STFT by the window sum, dataloader/data.py:77) sit nine orders of magnitude below.  A partial
// sum beyond that is clamped instead of wrapping, so an out-of-range input gives a saturated (finite, wrong-scale)
// normalisation rather than garbage with a flipped sign.
constexpr double kStatScale = 1048576.0;
constexpr double kStatClamp = 4.0e18;  // < 2^62: one clamped addend cannot wrap the accumulator
__device__ __forceinline__ void stat_add(double *p, double v) {
    const double s = fmin(fmax(v * kStatScale, -kStatClamp), kStatClamp);
    atomicAdd(reinterpret_cast<unsigned long long *>(p), (unsigned long long)__double2ll_rn(s));
}
__device__ __forceinline__ double stat_get(const double *p) {
    return (double)(*reinterpret_cast<const long long *>(p)) * (1.0 / kStatScale);
}
__device__ __forceinline__ void stat_set(double *p, double v) {
    *reinterpret_cast<long long *>(p) = __double2ll_rn(v * kStatScale);
}
__device__ __forceinline__ float2 affine_from_sums(double s, double q, double inv_n, double eps) {
    if (q < 0.0) return make_float2(1.f, 0.f);
    double mean = s * inv_n;
    double var = q * inv_n - mean * mean;
    if (var < 0.0) var = 0.0;
    double r = rsqrt(var + eps);
    return make_float2((float)r, (float)(-mean * r));
}

}  // namespace miso
