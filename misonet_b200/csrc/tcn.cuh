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
};

bool tcn_pw_eligible(int C);
void tcn_pw_scratch_need(int C, int nconv, size_t *wimg_bytes, size_t *wvec_bytes);
int tcn_pw_init();
int launch_tcn_wprep(const TcnPwTable &tab, int nconv, int C, int cpad, int split, void *wimg, float *wvec, cudaStream_t stream);
int launch_tcn_pw(const TcnPwArgs &p, int split, cudaStream_t stream);

}  // namespace miso
