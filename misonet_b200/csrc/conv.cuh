// Argument block shared by the (de)convolution kernels of the MISO conv stack.
#pragma once
#include "common.cuh"

namespace miso {

enum NormMode { NORM_NONE = 0, NORM_IN = 1, NORM_GLN = 2 };

// One 2-D (de)convolution over a channels-last activation view.
//   logical input  : in[b, t, f, in_coff + ci],  ci < cin, pixel pitch in_ctot
//   logical output : out[b, t, f, out_coff + co], co < cout, pixel pitch out_ctot
// The consumer applies the producer's normalisation while loading:
//   NORM_IN  : per (b, channel) instance norm from fp64 (sum, sumsq) accumulators
//              (model.py:411-414 order is conv -> ELU -> InstanceNorm, so the stored
//              tensor is the raw ELU output and its statistics);
//   NORM_GLN : per-sample global layer norm with gamma/beta (model.py:609-632).
// Zero padding is applied AFTER the normalisation, as in the reference.
struct ConvArgs {
    const float *in;
    const float *w;      // packed [KT*KF][cin][cout_pad]
    const float *bias;   // [cout_pad] or null
    float *out;
    const float *resid;  // optional residual view added before the store (model.py:549)
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
    // tcgen05 path only (conv_tc.cu)
    const void *w_tc;  // bf16 weight image in shared-memory order, or null
    int cout_pad16;    // cout rounded up to the UMMA N granularity (16 at M = 128)
    int tc_G;          // 128-row M tiles per CTA (set by the launcher)
};

int launch_conv_fp32(const ConvArgs &a, cudaStream_t stream);

// tensor-core path: split = 1 (bf16) or 3 (bf16x3, parity-grade)
bool conv_tc_eligible(const ConvArgs &a);
int launch_conv_tc(const ConvArgs &a, int split, cudaStream_t stream);
size_t conv_tc_weight_elems(int cin, int cout_pad16, int nsp);
int pack_conv_tc_weights(const float *d_w, void *d_img, int cout, int cin, int cout_pad16, int nsp, int transposed,
                         cudaStream_t stream);

}  // namespace miso
