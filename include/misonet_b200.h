/*
 * misonet_b200 -- C ABI of the B200-native MISO-BF-MISO hot path.
 *
 * The reference (yuhogun0908/MISOnet @ 79b3190) is pure Python and has no FFI; its
 * "interface" for this path is a set of Python callables.  Every entry point below
 * names the reference callable it stands in for (file:line into the reference repo).
 * The Python host side (misonet_b200/*.py) mirrors those callables and binds these
 * symbols with ctypes; INTEGRATION.md shows the binding a maintainer would add.
 *
 * Conventions
 *   - every pointer named d_* is a DEVICE pointer, everything else is host memory;
 *   - all tensors are dense fp32 / complex64 (float2 = re,im) unless strides are given;
 *     strides are in ELEMENTS of the pointed-to type;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *   - the caller owns inputs, outputs and workspaces; the library owns only the
 *     opaque net handle (packed weights);
 *   - calls are asynchronous on `stream`, never synchronise, never throw, never block
 *     (cf. the reference's NaN guard that drops into pdb, model.py:109-110);
 *   - return value: 0 = ok, negative = error (MISO_E_*); miso_last_error() gives a
 *     thread-local message for the last failing call.
 *   - one process per GPU; calls are re-entrant across streams except that a net
 *     handle's forward must not run concurrently with itself on the same workspace.
 */
#ifndef MISONET_B200_H
#define MISONET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MISO_OK 0
#define MISO_E_ARG (-1)       /* bad argument / unsupported shape */
#define MISO_E_CUDA (-2)      /* CUDA runtime error (launch, alloc) */
#define MISO_E_STATE (-3)     /* handle not ready (missing parameters) */
#define MISO_E_WORKSPACE (-4) /* workspace too small */
#define MISO_E_ARCH (-5)      /* device is not sm_100 */

#define MISO_ABI_VERSION 1

/* ---- library ---------------------------------------------------------------- */
int miso_abi_version(void);
const char *miso_last_error(void);
/* 0 when the current device is compute capability 10.x, MISO_E_ARCH otherwise. */
int miso_check_device(void);
/* number of kernels this library has launched in this process (all streams). */
uint64_t miso_launch_count(void);

/* ---- measurement hooks (bench.py's roofline block) ------------------------------
 * When enabled, every (de)convolution launch of the conv stack -- the dominant kernel
 * family -- is bracketed by cudaEvents on its own stream.  miso_prof_collect waits for
 * the recorded events, returns the summed device time, the ALGORITHMIC flops and bytes
 * of those launches (2*MAC; each conv reads its logical input
 * once and writes its output once) and their count, and clears the record. */
#define MISO_PROF_ALL (-1)
#define MISO_PROF_CONV_FP32 0 /* fp32 FMA implicit-GEMM (de)conv / pointwise kernels */
#define MISO_PROF_CONV_TC 1   /* tcgen05 implicit-GEMM conv kernels */
#define MISO_PROF_TCN 2       /* tcgen05 pointwise convs of the TCN */
#define MISO_PROF_MVDR 3      /* MVDR kernels */
#define MISO_PROF_PREP 5      /* per-sample operand preparation of the tensor-core convs (weight images, border-bias sums) */
#define MISO_PROF_CONV_RS 4   /* row-streaming tcgen05 conv (frame taps merged into N): the DenseBlock convs */
int miso_prof_enable(int on);
/* collects (and clears) the records of one family, or of all with MISO_PROF_ALL */
int miso_prof_collect(int family, double *total_ms, double *total_flops, double *total_bytes, uint64_t *launches);
/* the same plus the tensor-pipe work the family's kernels actually ISSUED (bf16x3: three MMAs per product; N, M and K
 * padding of the MMA tiles included): executed / peak is how busy the tensor pipe is, algorithmic / peak is the roofline
 * fraction */
int miso_prof_collect2(int family, double *total_ms, double *total_flops, double *total_bytes, double *total_exec_flops,
                       uint64_t *launches);
/* per-launch view of the queued records, in launch order: fills up to `capacity` entries of each
 * non-NULL array and returns the count; the records stay queued for miso_prof_collect. */
int miso_prof_dump(double *ms, double *flops, int *family, int capacity);

/* ---- S1: STFT front end -------------------------------------------------------
 * replaces AudioDataset.STFT + "/scale" + permute, dataloader/data.py:49-66,77-79
 * (= tester.py:992-1012): zero-pad nperseg/2 both sides, frames of nperseg with hop
 * (nperseg - noverlap), periodic hann, UNNORMALISED rFFT.
 *   d_x   : fp32, sample (b, n, m) at d_x[b*sb + n*sn + m*sm]
 *   d_out : complex64 [B, M, T, nperseg/2+1], T = miso_stft_num_frames(N, nperseg, hop)
 * nperseg must be 256 or 512. */
int miso_stft_num_frames(int n_samples, int nperseg, int hop);
int miso_stft_fwd(const float *d_x, int64_t sb, int64_t sn, int64_t sm, void *d_out, int B, int N, int M,
                  int nperseg, int hop, void *stream);

/* ---- ISTFT back end ----------------------------
 * replaces Tester_*.ISTFT applied to ``spec * scale`` (tester.py:186-198, 545-556, 979-990; call sites
 * tester.py:949-957): scipy.signal.istft(window='hann', nperseg, noverlap) with its defaults
 * (one-sided input, boundary trimmed by nperseg/2 both sides, division by the summed squared window)
 * of the spectrogram times scale = 1/sum(window), i.e. the exact inverse of miso_stft_fwd.
 *   d_spec : complex64, bin (s, t, f) of signal s at d_spec[s*ss + t*st + f*sf]  (element strides)
 *   d_out  : fp32 [S, miso_istft_num_samples(T, nperseg, hop)]
 *   d_ws   : miso_istft_workspace_bytes(S, T, nperseg) bytes (the windowed frames before overlap-add) */
int miso_istft_num_samples(int T, int nperseg, int hop);
size_t miso_istft_workspace_bytes(int S, int T, int nperseg);
int miso_istft_fwd(const void *d_spec, int64_t ss, int64_t st, int64_t sf, float *d_out, int S, int T, int nperseg,
                   int hop, void *d_ws, size_t ws_bytes, void *stream);

/* ---- N1/N2: MISO_1 / MISO_3 network body --------------------------------------
 * replaces model.MISO_1 / model.MISO_3 (model.py:8-111, 282-395) and the layers they
 * are built from (model.py:401-632).  The handle is created from the constructor
 * arguments (channel lists WITHOUT the in/out channels that model.py:16-17 inserts)
 * and filled with the reference's own state_dict tensors, key by key. */
typedef struct miso_net miso_net_t;

int miso_net_create(miso_net_t **out, int in_ch, int out_ch, int num_bottleneck, const int *en_channels,
                    const int *de_channels, int tcn_repeats, int tcn_blocks);
int miso_net_destroy(miso_net_t *net);
/* number of state_dict entries and the i-th key / element count (key order = the
 * reference's state_dict order). */
int miso_net_num_params(const miso_net_t *net);
const char *miso_net_param_key(const miso_net_t *net, int i);
int64_t miso_net_param_numel(const miso_net_t *net, int i);
/* (re)pack one parameter from a DEVICE fp32 tensor in the reference's layout
 * (Conv2d [Cout,Cin,3,3], ConvTranspose2d [Cin,Cout,3,3], Conv1d [C,1,3]/[C,C,1], ...). */
int miso_net_set_param(miso_net_t *net, const char *key, const float *d_data, int64_t numel, void *stream);
/* the same for ALL parameters in one or two launches: d_data[i] = device fp32 tensor of parameter i (key order of
 * miso_net_param_key), or NULL to leave that parameter untouched; n = miso_net_num_params().  A training step changes
 * every parameter (trainer.py:212 optimizer.step()): one call instead of 268. */
int miso_net_set_params(miso_net_t *net, const float *const *d_data, int n, void *stream);
/* compute path of the 3x3 (de)convs: 0 = fp32 FMA kernels everywhere; 1 = tcgen05 bf16x3 split
 * (a_hi*w_hi + a_lo*w_hi + a_hi*w_lo, fp32 accumulate: parity-grade, 3 MMAs per product); 2 = tcgen05
 * bf16 operands and hi-only activation planes (throughput mode, ~1e-2 relative error).
 * In modes 1 / 2 the TCN's pointwise convs run on tcgen05 too (tcn_pw_kernel, same split); its depthwise convs,
 * norms and the residual stream stay fp32. */
int miso_net_set_mode(miso_net_t *net, int mode);
/* F must reduce to exactly 1 at the bottleneck (129 for 7 blocks, 257 for 8); returns
 * MISO_E_ARG with a clear message otherwise (the reference raises an opaque conv error). */
int miso_net_check_shape(const miso_net_t *net, int T, int F);
size_t miso_net_workspace_bytes(const miso_net_t *net, int B, int T, int F);
/* Network input layout ("planes"): bf16 [B][2: hi, lo][in_pad/8][T][F][8] with in_pad = in_ch rounded
 * up to 8; value = hi + lo (16-17 mantissa bits), padding channels are zero.  Written by
 * miso_pack_miso1 / miso_pack_miso3; miso_net_input_bytes gives the buffer size (128-byte aligned base).
 * All internal activations use the same layout: a [T, F] tile of 8 channels is a dense box of
 * 16-byte pixels, which TMA moves straight into the UMMA operand layout. */
size_t miso_net_input_bytes(const miso_net_t *net, int B, int T, int F);
/* d_x : input planes (above); d_y : fp32 channels-last [B, T, F, out_ch]. */
int miso_net_forward(miso_net_t *net, const void *d_x, float *d_y, int B, int T, int F, void *d_ws,
                     size_t ws_bytes, void *stream);
/* The forward's ~200 launches are captured into a CUDA graph the first time a given (d_x, d_y, d_ws,
 * B, T, F, mode) is seen and replayed afterwards (weights are read through the handle, so repacking
 * them needs no re-capture).  on = 0 disables this (every call enqueues the launches one by one);
 * per-launch profiling (miso_prof_enable) always uses the eager path. */
int miso_net_set_graph(miso_net_t *net, int on);
/* debugging: subsequent tensor-core conv launches whose (cin, Fin) match write a clock64 event log of
 * their CTA 0 into d_buf (3 x 4096 int64: producer / MMA issuer / epilogue regions of (tag, clock) pairs;
 * see tools/tc_trace.py).  d_buf = NULL turns it off.  Use with miso_net_set_graph(net, 0). */
int miso_debug_tc_trace(long long *d_buf, int cin, int fin);
/* debugging / parity taps: copy an internal activation of the LAST forward on this
 * workspace into a dense NCHW fp32 tensor (normalised as the reference sees it).
 * name: "enc<i>", "tcn", "dec<i>".  Returns the element count or a negative error. */
int64_t miso_net_tap(miso_net_t *net, const char *name, float *d_out, int64_t capacity, int B, int T, int F,
                     void *d_ws, void *stream);

/* Output sample format of the reference's wav writer (tester.py:155-157, 444-446, 950-952): (int16) trunc(x * scale)
 * with the product in double, scale = 32767 (np.iinfo(np.int16).max), saturated. */
int miso_wave_to_int16(const float *d_x, int16_t *d_out, int64_t n, float scale, void *stream);

/* ---- training (reference trainer.py:159-212: estimate = model(mix); loss = loss_uPIT(...); loss.backward()) ----
 * miso_net_forward_train is miso_net_forward on the TRAINING workspace plan: every TemporalBlock keeps its input and
 * mid state (the inference plan updates the residual stream in place), no CUDA graph.  miso_net_backward then consumes
 * the activations and statistics the forward left in the SAME workspace:
 *   d_gy    : dL/d(output), fp32 channels-last [B, T, F, out_ch rounded up to a multiple of 8] (the layout of d_y with
 *             zero padding channels; written by miso_grad_pack); clobbered
 *   d_grads : gradient of every parameter, flat fp32 in miso_net_param_key order, each in its torch layout
 *             (miso_net_grad_numel elements in total); overwritten.
 * Both run in the conv mode of miso_net_set_mode.  Mode 0: all backward arithmetic is fp32 FMA.  Modes 1 / 2: the data
 * gradients run on the forward's tcgen05 conv kernels (dL/dy as bf16 hi/lo planes x transposed weights) and the weight
 * gradients on tensor cores with the same bf16 hi/lo split and fp32 accumulation; the InstanceNorm2d / ELU / gLN /
 * PReLU / depthwise / InstanceNorm1d backward kernels are fp32 elementwise / reduction kernels in every mode. */
int64_t miso_net_grad_numel(const miso_net_t *net);
/* Gradient buckets for a data-parallel step (trainer.py:205-212 is the step; the reference itself is single-GPU): the flat
 * gradient buffer as contiguous element ranges [begin[k], end[k]) in the ORDER IN WHICH miso_net_backward COMPLETES THEM
 * (upper decoders, lower decoders, TCN, upper encoders, lower encoders).  miso_net_backward records an event after the last
 * kernel of every bucket; miso_net_wait_grad_bucket makes `stream` wait for bucket k of the most recent backward, so a
 * communication stream can all-reduce bucket k while the rest of the backward is still running.  Returns the bucket count. */
int miso_net_grad_buckets(const miso_net_t *net, int64_t *begin, int64_t *end, int capacity);
int miso_net_wait_grad_bucket(miso_net_t *net, int bucket, void *stream);
size_t miso_net_train_workspace_bytes(const miso_net_t *net, int B, int T, int F);
int miso_net_forward_train(miso_net_t *net, const void *d_x, float *d_y, int B, int T, int F, void *d_ws,
                           size_t ws_bytes, void *stream);
int miso_net_backward(miso_net_t *net, const void *d_x, float *d_gy, int B, int T, int F, void *d_ws, size_t ws_bytes,
                      float *d_grads, void *stream);
/* complex64 gradient [B, S, T, F] (PyTorch convention dL/dre + i dL/dim) -> d_gy layout [B, T, F, round_up(2S, 8)] */
int miso_grad_pack(const void *d_grad, float *d_gy, int B, int S, int T, int F, void *stream);
/* gradient of loss_uPIT (criterion.py:8-63) w.r.t. the estimate for the winning permutation d_perm_idx[b]
 * (int64, from miso_pair_fwd mode 1), scaled by the upstream scalar *d_gout: complex64 [B, S, T, F] dense. */
/* gradient of loss_Enhance (criterion.py:121-141) w.r.t. the estimate, scaled by *d_gout: complex64, same shape */
int miso_loss_enhance_bwd(const void *d_est, const void *d_ref, int B, int64_t n_per_batch, const float *d_gout,
                          void *d_grad, void *stream);
int miso_upit_bwd(const void *d_est, int64_t e_sb, int64_t e_ss, const void *d_ref, int64_t r_sb, int64_t r_ss,
                  const int64_t *d_perm_idx, int B, int S, int T, int F, const float *d_gout, void *d_grad,
                  void *stream);

/* input / output layout adapters (model.py:76-80, 109-111; 358-366):
 * pack_miso1 : mixture complex64 [B,M,T,F] -> input planes of n_shift*B samples with the mic axis
 *              circularly rolled by -shift[k] (torch.roll(mix,-q,dims=1), tester.py:1034,1049);
 *              batch index of (k, b) is k*B + b;  channels = re(m0..), im(m0..).
 * pack_miso3 : (mix [B,M,T,F], second [B,1,T,F], third [B,1,T,F]) -> input planes, 2(M+2) channels
 *              in the positional order every caller uses (tester.py:1242).
 * unpack     : channels-last [B,T,F,2S] -> complex64 [B,S,T,F]. */
int miso_pack_miso1(const void *d_mix, void *d_x, int B, int M, int T, int F, const int *shifts, int n_shift,
                    void *stream);
int miso_pack_miso3(const void *d_mix, const void *d_second, const void *d_third, void *d_x, int B, int M, int T,
                    int F, void *stream);
int miso_unpack_complex(const float *d_y, void *d_out, int B, int S, int T, int F, void *stream);

/* ---- A1/A2/L1/L2: pairwise distances, permutation decisions, losses ------------
 * miso_pair_fwd computes, for a = [B,S,T,F] and b = [B,S,T,F] complex64 (batch/speaker
 * strides in elements, [T,F] planes dense):
 *   mode 0 (tester.py:1043-1054, 903-905):  pair[b,i,j] = sum | |a_i| - |b_j| |
 *   mode 1 (criterion.py:36-47):            pair[b,i,j] = sum |re a_i - re b_j| + |im a_i - im b_j|
 *                                                          + | sqrt(|a_i|^2 + 1e-8) - |b_j| |
 * then scores[b,p] = sum_i pair[b,i,perm_p[i]] over itertools.permutations order,
 * d_perm_idx[b] = argmin_p (int64, first minimum), d_loss[0] = mean_b min_p scores
 * (criterion.py:56-63).  Sums are accumulated in fp64 in a fixed order, so decisions
 * do not depend on the launch configuration.  S <= 4.
 * outputs d_pair (fp32 [B,S,S]), d_perm_idx, d_loss may each be NULL. */
size_t miso_pair_workspace_bytes(int B, int S, int T, int F);
int miso_pair_fwd(const void *d_a, int64_t a_sb, int64_t a_ss, const void *d_b, int64_t b_sb, int64_t b_ss, int B,
                  int S, int T, int F, int mode, float *d_pair, int64_t *d_perm_idx, float *d_loss, void *d_ws,
                  size_t ws_bytes, void *stream);
/* out[s][b] = src[b][perm_{idx[b]}[s]]  ([T,F] planes; tester.py:1061-1065, 907-910). */
int miso_perm_gather(const void *d_src, int64_t src_sb, int64_t src_ss, void *d_dst, int64_t dst_sb, int64_t dst_ss,
                     const int64_t *d_perm_idx, int B, int S, int T, int F, void *stream);
/* criterion.py:121-141: d_loss[0] = (sum|dre| + sum|dim| + sum|sqrt(|a|^2+1e-8) - |b||) / B
 * over n_per_batch complex elements per batch row. */
int miso_loss_enhance_fwd(const void *d_est, const void *d_ref, int B, int64_t n_per_batch, float *d_loss, void *d_ws,
                          size_t ws_bytes, void *stream);

/* ---- M1..M7: MVDR beamformer --------------------------------------------------
 * replaces Tester_*.Apply_Beamforming and its helpers (tester.py:1071-1167,1211-1228)
 * for S sources that share one mixture:
 *   Phi_s = 1/T sum_t s s^H, Phi_n = 1/T sum_t (x-s)(x-s)^H (Hermitian-symmetrised),
 *   v = principal eigenvector of Phi_s, d = v/v[0], d *= sqrt(M/||d||_2)  (as written),
 *   sequential-over-f phase correction, (Phi_n + epsi I) u = d, w = u / (d^H u),
 *   y[t] = sum_m conj(w[m]) x[m,t].
 *   d_src : complex64, element (s,b,m,t,f) at s*src_ss + b*sb + m*sm + t*st + f*sf
 *   d_mix : complex64, element (b,m,t,f)   at            b*sb + m*sm + t*st + f*sf
 *   d_out : complex64 dense [S, B, T, F]   (tester.py:1134 returns [B,T,F] per source)
 *   d_weights (optional, may be NULL): complex64 [S, B, F, M] beamformer weights.
 * 2 <= M <= 8.  Streaming work is fp32/complex64 like the reference; the 6x6 eigen and
 * linear solves run in fp64. */
size_t miso_mvdr_workspace_bytes(int S, int B, int M, int T, int F);
int miso_mvdr_fwd(const void *d_src, int64_t src_ss, const void *d_mix, int64_t sb, int64_t sm, int64_t st, int64_t sf,
                  void *d_out, void *d_weights, int S, int B, int M, int T, int F, float epsi, void *d_ws,
                  size_t ws_bytes, void *stream);

/* Staged form of the same stage, for covariance statistics that span ranks: the reference's utterance-level
 * beamforming of a chunked recording (Tester_Beamforming.inference with utterance_flag, tester.py:425-449) takes the
 * spatial covariances over ALL frames of the recording.  Every rank computes the partial sums of its frames
 * (miso_mvdr_scm: float [S*B][miso_mvdr_tsplit(B,F)][2 M (M+1)][F]), the ranks' partial sums are concatenated along the
 * split axis (an all-gather of miso_mvdr_partial_bytes each), miso_mvdr_weights adds them in that fixed order in fp64 and
 * returns the beamformers ([S,B,F,M] complex64; d_ws: S*B*F*M*16 bytes), and miso_mvdr_apply filters the local frames.
 * With one rank the three calls equal miso_mvdr_fwd bit for bit. */
int miso_mvdr_tsplit(int B, int F);
size_t miso_mvdr_partial_bytes(int S, int B, int M, int F);
int miso_mvdr_scm(const void *d_src, int64_t src_ss, const void *d_mix, int64_t sb, int64_t sm, int64_t st, int64_t sf,
                  void *d_partial, int S, int B, int M, int T, int F, void *stream);
int miso_mvdr_weights(const void *d_partial, int nsplit, int T_total, void *d_weights, int S, int B, int M, int F,
                      float epsi, void *d_ws, size_t ws_bytes, void *stream);
int miso_mvdr_apply(const void *d_mix, int64_t sb, int64_t sm, int64_t st, int64_t sf, const void *d_weights, void *d_out,
                    int S, int B, int M, int T, int F, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MISONET_B200_H */
