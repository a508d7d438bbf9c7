// Backward (training) kernels of the MISO conv stack and TCN (backward.cu); SURVEY.md section 8(f) rank 1,
// reference trainer.py:159-212 (loss.backward() through model.py:76-111) and criterion.py:8-63.
#pragma once
#include "conv.cuh"

namespace miso {

// InstanceNorm2d + ELU backward of one conv layer's output view (model.py:411-414 order conv -> ELU -> IN):
// the gradient buffer holds dL/dz (z = the normalised tensor its consumers read) and is turned in place into
// dL/dy (y = the conv output before the ELU); the bias gradient is accumulated on the way.
struct InBwdArgs {
    const __nv_bfloat16 *e;  // forward planes of the buffer [B][hi|lo][ctot/8][npix][8] (raw ELU outputs)
    float *g;                // gradient buffer, fp32 channels-last [B][npix][ctot]
    const double *sums;      // forward statistics [B][ctot][2]
    double *red;             // [B][c][2] fp64 scratch: sum dz, sum dz * z
    float *dbias;            // [c] (torch layout), accumulated
    int B, npix, ctot, coff, c;
    int c_real;  // channels that own a bias (c_real < c only for a network output padded to a multiple of 4 channels)
    int use_lo;
    double inv_n;
    float eps;
    int plain;  // layer without ELU / IN (model.py:401-406, 418-423): dy = dz, only the bias gradient is taken
    // optional copy of dy as bf16 hi/lo planes [B][hi|lo][c/8][npix][8] (c % 8 == 0): the input of the tensor-core
    // data-gradient conv
    __nv_bfloat16 *dyp;
};
int launch_in_bwd(const InBwdArgs &a, cudaStream_t st);

// weight gradient dW[tap][ci][co] = sum_{b, p} xhat[b, ci, src(p, tap)] * dy[b, co, p]  (p over the layer's OUTPUT grid;
// src = the forward kernel's gather rule, conv.cuh), written with atomics straight into the torch layout.
struct WgradArgs {
    const void *x;  // forward input of the layer: planes or fp32 channels-last
    int x_layout, use_lo;
    const double *x_sums;  // NORM_IN statistics of x or null
    double inv_n;
    float eps;
    const float *dy;  // fp32 channels-last [B][T*Fout][dy_ctot]
    float *dw;        // Conv2d [cout][cin][taps] / ConvTranspose2d [cin][cout][taps]
    int B, T, Fin, Fout;
    int x_ctot, x_coff, cin;
    int dy_ctot, dy_coff, cout;
    int cout_real;  // output channels of the parameter tensor (cout_real < cout only for a padded network output)
    int KT, KF, stride_f, pad_t, pad_f, transposed;
    // tcgen05 path (wgrad_tc.cu; the stride-1 pad-(1,1) 3x3 convs): dy as bf16 hi/lo planes [B][hi|lo][cout/8][T*Fout][8]
    // (in_bwd_apply writes them for the data-gradient conv), a partial-accumulator scratch, and the constant-one pixels
    __nv_bfloat16 *dyp = nullptr;
    float *partial = nullptr;
    size_t partial_bytes = 0;
    __nv_bfloat16 *ones = nullptr;
};
int launch_wgrad(const WgradArgs &a, cudaStream_t st);
bool wgrad_tc_eligible(const WgradArgs &a);
size_t wgrad_tc_partial_bytes(int B);
int wgrad_tc_fill_ones(__nv_bfloat16 *ones, size_t npix, cudaStream_t st);
int launch_wgrad_tc(const WgradArgs &a, cudaStream_t st);

// packed forward weights [taps][cin][cout_pad] -> data-gradient weights [taps][cout][cin_pad]
// flip = 1 reverses the tap order (a stride-1 transposed conv as a plain conv)
int launch_dgrad_pack(const float *src, float *dst, int taps, int cin, int cout, int cout_pad, int cin_pad, int flip, cudaStream_t st);

// One half of a TemporalBlock (model.py:530-531 / 538-539 + DepthwiseSeparableConv model.py:553-567)
struct TcnBwdArgs {
    const float *u;        // input of the half, fp32 [B][T][C]
    const double *u_sums;  // its InstanceNorm1d statistics [B][C][2]
    double inv_T;
    float in_eps;
    const double *g_sums;  // forward gLN statistics of the PReLU output [B][2]
    double gln_inv_n;
    float gln_eps;
    const float *wdw, *alpha, *gamma, *beta;
    int B, T, C, dil;
};
int launch_tcn_recompute(const TcnBwdArgs &a, float *Y, float *Q, cudaStream_t st);
int launch_gln_bwd(const TcnBwdArgs &a, float *DQ, const float *Y, double *gred, float *dgamma, float *dbeta, float *dalpha,
                   cudaStream_t st);
int launch_dw_bwd(const TcnBwdArgs &a, const float *DY, float *DN, double *ired, float *dwdw, float *out, int accumulate,
                  cudaStream_t st);
// fp32 channels-last [B][npix][C] -> bf16 hi/lo planes [B][hi|lo][C/8][npix][8] (C % 8 == 0)
int launch_cl_to_planes(const float *src, __nv_bfloat16 *dst, int B, int npix, int C, cudaStream_t st);
// dst[b][p][dcoff + c] += (hi + lo)[b][c][p] for bf16 hi/lo planes src [B][hi|lo][sctot/8][npix][8], c < C (C % 4 == 0)
int launch_planes_accumulate(const __nv_bfloat16 *src, int sctot, float *dst, int dctot, int dcoff, int C, int B, int npix,
                             cudaStream_t st);
// dst[b][p][dcoff + c] (+)= src[b][p][scoff + c]
int launch_copy_channels(const float *src, int sctot, int scoff, float *dst, int dctot, int dcoff, int C, int64_t rows,
                         int accumulate, cudaStream_t st);

}  // namespace miso
