// Argument block shared by the (de)convolution kernels of the MISO conv stack.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace miso {

enum NormMode { NORM_NONE = 0, NORM_IN = 1, NORM_GLN = 2 };

// Activation layouts in HBM.
//   LAYOUT_CL_F32 : fp32 channels-last [B, T, F, Ctot]  (TCN state, final network output)
//   LAYOUT_PLANES : bf16 "NC/8HWC8" planes [B, Ctot/8, T, F, 8] stored twice: a hi plane set
//                   (the value rounded to bf16) and a lo plane set (the bf16-rounded remainder),
//                   lo = hi base + lo_off bytes.  hi + lo carries 16-17 mantissa bits.  A
//                   [T, F] tile of 8 channels is a dense box of 16-byte pixels, which is what
//                   both TMA and the no-swizzle K-major UMMA operand layout want.
enum ActLayout { LAYOUT_CL_F32 = 0, LAYOUT_PLANES = 1 };

// One 2-D (de)convolution over an activation view.
//   logical input  : channels [in_coff, in_coff + cin) of a buffer with in_ctot channels
//   logical output : channels [out_coff, out_coff + cout) of a buffer with out_ctot channels
// The consumer applies the producer's normalisation:
//   NORM_IN  : per (b, channel) instance norm from fp64 (sum, sumsq) accumulators
//              (model.py:411-414 order is conv -> ELU -> InstanceNorm, so the stored
//              tensor is the raw ELU output and its statistics);
//   NORM_GLN : per-sample global layer norm with gamma/beta (model.py:609-632).
// Zero padding is applied AFTER the normalisation, as in the reference.
struct ConvArgs {
    const void *in;
    const float *w;      // packed fp32 [KT*KF][cin][cout_pad] (gather form for transposed convs)
    const float *bias;   // [cout_pad] or null
    void *out;
    const float *resid;  // optional fp32 channels-last residual view added before the store (model.py:549)
    const double *in_sums;  // NORM_IN: [B][in_ctot][2]   NORM_GLN: [B][2]
    const float *gamma;     // NORM_GLN: [cin]
    const float *beta;
    double *out_sums;       // [B][out_ctot][2] or null
    int B, T, Fin, Fout;
    int in_ctot, in_coff, cin;
    int out_ctot, out_coff, cout, cout_pad;
    int resid_ctot, resid_coff;
    int KT, KF, stride_f, pad_t, pad_f;
    int transposed;  // ConvTranspose2d gather form (model.py:418-433)
    int norm_mode;
    float norm_eps;
    double norm_inv_n;  // 1 / (elements per statistic)
    int elu;
    int in_layout, out_layout;
    size_t in_lo_off, out_lo_off;  // bytes from the hi plane set to the lo plane set
    int use_lo;                    // 0: bf16 throughput mode (hi planes only are read and written)
};

inline size_t plane_set_bytes(int B, int ctot, int T, int F) { return (size_t)B * ctot * T * F * 2; }

int launch_conv_fp32(const ConvArgs &a, cudaStream_t stream);

// tensor-core path (conv_tc.cu): split = 1 (bf16) or 3 (bf16x3, parity-grade)
struct TcScratch {
    void *wimg;      // per-sample weight images
    float *btab;     // per-sample bias tables
    size_t wimg_bytes, btab_bytes;
};
int conv_tc_init();
void conv_tc_set_trace(long long *d_buf, int cin, int fin);
bool conv_tc_eligible(const ConvArgs &a);
void conv_tc_scratch_need(const ConvArgs &a, int split, size_t *wimg_bytes, size_t *btab_bytes);
int launch_conv_tc(const ConvArgs &a, int split, const TcScratch &scratch, cudaStream_t stream);

// row-streaming variant with the frame taps merged into N (conv_rs.cu): stride-1 pad-(1,1) 3x3 convs at F+1 = 128 / 256
int conv_rs_init();
void conv_rs_set_trace(long long *d_buf, int cin, int fin);
bool conv_rs_eligible(const ConvArgs &a, int split);
void conv_rs_scratch_need(const ConvArgs &a, int split, size_t *wimg_bytes, size_t *btab_bytes);
int launch_conv_rs(const ConvArgs &a, int split, const TcScratch &scratch, cudaStream_t stream);

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&v);
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

}  // namespace miso
