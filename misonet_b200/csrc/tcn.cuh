// Tensor-core pointwise convolutions of the TCN bottleneck (tcn.cu).
#pragma once
#include "common.cuh"

namespace miso {

constexpr int kTcnMaxPw = 64;  // pointwise convs per network (2 * R * X)

struct TcnPwTable {  // device pointers of every pointwise conv's parameters (model.py:556-561)
    const float *w[kTcnMaxPw];      // packed fp32 [cin][cout_pad]
    const float *gamma[kTcnMaxPw];  // gLN gamma / beta of the conv's input (model.py:609-632)
    const float *beta[kTcnMaxPw];
};

struct TcnPwArgs {
    const void *planes;  // A operand: bf16 planes [B][hi|lo][C/8][T][8] written by tcn_dw_kernel
    size_t lo_off;
    const void *wimg;    // images of all convs (tcn_wprep_kernel)
    const float *wvec;   // [nconv][2][C]: W beta, W gamma
    int index;           // which conv
    const double *gln_sums;
    double gln_inv_n;
    float gln_eps;
    float *out;
    const float *resid;
    double *out_sums;
    int B, T, C;
    // out_planes: `out` is a bf16 hi/lo plane buffer [B][hi|lo][out_ctot/8][T][8] (the TCN's last half writes the decoder's input)
    int out_planes, out_ctot, use_lo;
    size_t out_lo_off;
    int plain;  // no gLN epilogue (rstd 1, no bias, wvec unused): the data gradient dq = dy W^T over transposed images
};

// The whole TCN as ONE launch (tcn_fused_kernel, tcn.cu): a thread-block cluster per sample (one CTA per 128 frames) walks the
// 2 R X half-blocks; cluster barriers stand where the per-sample statistics (InstanceNorm1d, gLN) used to need a kernel boundary.
constexpr int kTcnMaxHalf = 32;
struct TcnFusedHalf {
    const float *u;         // input state, fp32 channels-last [B][T][C]
    const double *u_sums;   // its InstanceNorm1d statistics [B][C][2]
    double *g_sums;         // gLN statistics of the PReLU output [B][2] (accumulated by this half)
    void *out;              // fp32 channels-last [B][T][C], or (out_planes) bf16 hi/lo planes [B][hi|lo][out_ctot/8][T][8]
    double *out_sums;       // InstanceNorm1d statistics of the output [B][C][2], or null
    const float *resid;     // residual, layout of u, or null (may alias out: each element is read by its own writer)
    const float *dw;        // depthwise taps [C][3]
    const float *alpha;     // PReLU slope
    int dil, out_planes;
};
struct TcnFusedArgs {
    int nhalf, B, T, C;
    const void *wimg;       // images of all pointwise convs (tcn_wprep_kernel)
    const float *wvec;      // [nhalf][2][C]: W beta, W gamma
    double in_inv_n, gln_inv_n;
    float in_eps, gln_eps;
    int out_ctot, use_lo;   // geometry of a planes output
    size_t out_lo_off;
    long long *trace;       // debug: clock64 stamps of CTA 0, [half][16] (tools/tcn_trace.py), or null
    TcnFusedHalf h[kTcnMaxHalf];
};
bool tcn_fused_eligible(int C, int T, int nhalf);
int launch_tcn_fused(const TcnFusedArgs &a, int split, cudaStream_t stream);
void tcn_fused_set_trace(long long *d_buf);

bool tcn_pw_eligible(int C);
void tcn_pw_scratch_need(int C, int nconv, size_t *wimg_bytes, size_t *wvec_bytes);
int tcn_pw_init();
int launch_tcn_wprep(const TcnPwTable &tab, int nconv, int C, int cpad, int split, void *wimg, float *wvec, cudaStream_t stream,
                     int transposed = 0);
int launch_tcn_pw(const TcnPwArgs &p, int split, cudaStream_t stream);

}  // namespace miso
