"""numpy restatement of the non-network stages of the MISO-BF-MISO hot path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Every function cites
the reference file:line it follows (paths are relative to the reference repo,
yuhogun0908/MISOnet @ 79b3190).  Pinned against the reference's own outputs
through ``tests/golden/*.npz`` (``tests/test_oracle_golden.py``).

All functions keep the reference's arithmetic precision: inputs that arrive as
complex64 stay complex64 (the reference's MVDR runs in single precision,
SURVEY.md section 8(a) M1), unless ``dtype`` says otherwise.
"""
from itertools import permutations

import numpy as np


# --------------------------------------------------------------------------
# S1: STFT front end
# --------------------------------------------------------------------------
def hann_periodic(nperseg):
    """scipy.signal.get_window('hann', n) (fftbins=True -> periodic hann).
    dataloader/data.py:37."""
    n = np.arange(nperseg, dtype=np.float64)
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * n / nperseg)


def stft(time_sig, nperseg=256, noverlap=192):
    """dataloader/data.py:49-66 followed by ``/scale`` and the [M,F,T]->[M,T,F]
    permute of data.py:77-79.

    scipy.signal.stft(window='hann', nperseg, noverlap) with its defaults
    (boundary='zeros': nperseg//2 zeros each side; padded=True: zero-extend so
    that an integer number of hops fits; scaling='spectrum': divide by sum(w))
    and then ``/scale`` with scale = sqrt(1/sum(w)^2) = 1/sum(w) (data.py:37-38)
    i.e. the *unnormalised* windowed rFFT.

    time_sig : float [N, M] (samples, mics)
    returns  : complex64 [M, T, F]
    """
    x = np.asarray(time_sig)
    assert x.ndim == 2 and x.shape[0] > x.shape[1]
    n, m = x.shape
    hop = nperseg - noverlap
    w = hann_periodic(nperseg)
    half = nperseg // 2
    xp = np.concatenate([np.zeros((half, m)), x.astype(np.float64), np.zeros((half, m))], axis=0)
    # padded=True: extend so that (len - nperseg) % hop == 0
    nadd = (-(xp.shape[0] - nperseg) % hop) % nperseg
    if nadd:
        xp = np.concatenate([xp, np.zeros((nadd, m))], axis=0)
    n_frames = (xp.shape[0] - noverlap) // hop
    idx = np.arange(nperseg)[None, :] + hop * np.arange(n_frames)[:, None]
    frames = xp[idx, :]                       # [T, nperseg, M]
    frames = frames * w[None, :, None]
    spec = np.fft.rfft(frames, axis=1)        # [T, F, M]
    # scipy returns complex64 for float32 input; the reference then divides by a
    # python float, staying complex64.
    return np.ascontiguousarray(np.transpose(spec, (2, 0, 1))).astype(np.complex64)


def istft(spec_tf, nperseg=256, noverlap=192):
    """Tester_*.ISTFT applied to ``spec * scale`` (tester.py:979-990, call sites 949-957): scipy.signal.istft with its
    defaults -- per frame irfft(Z * sum(w)) (scaling='spectrum'; times scale = 1/sum(w) that is irfft of the
    unnormalised spectrum), synthesis window, overlap-add, division by the overlap-added squared window where it
    exceeds 1e-10, and nperseg//2 samples trimmed at both ends (boundary=True).

    spec_tf : complex [T, F] (one signal; the reference passes the [F, T] transpose)
    returns : float32 [(T - 1) * hop]"""
    z = np.asarray(spec_tf)
    t_frames, f_bins = z.shape
    assert f_bins == nperseg // 2 + 1
    hop = nperseg - noverlap
    w = hann_periodic(nperseg)
    xs = np.fft.irfft(z.astype(np.complex128), n=nperseg, axis=1) * w[None, :]
    total = nperseg + (t_frames - 1) * hop
    x = np.zeros(total)
    norm = np.zeros(total)
    for t in range(t_frames):
        x[t * hop:t * hop + nperseg] += xs[t]
        norm[t * hop:t * hop + nperseg] += w * w
    half = nperseg // 2
    x, norm = x[half:total - half], norm[half:total - half]
    return (x / np.where(norm > 1e-10, norm, 1.0)).astype(np.float32)


def stft_num_frames(n_samples, nperseg=256, noverlap=192):
    hop = nperseg - noverlap
    total = n_samples + 2 * (nperseg // 2)
    total += (-(total - nperseg) % hop) % nperseg
    return (total - noverlap) // hop


# --------------------------------------------------------------------------
# A1/A2/L1: permutation tables and decisions
# --------------------------------------------------------------------------
def perm_table(num_spks):
    """list(itertools.permutations(range(S))) -- criterion.py:49, tester.py:1055."""
    return np.array(list(permutations(range(num_spks))), dtype=np.int64)


def best_perm(pair, num_spks):
    """einsum('bij,pij->bp', pair, one_hot(perms)) then argmin over p
    (criterion.py:56-58; tester.py:1058-1059).  pair: [B,S,S].
    Returns (argmin index int64 [B], scores [B,P])."""
    perms = perm_table(num_spks)
    rows = np.arange(num_spks)
    scores = np.stack([pair[:, rows, p].sum(axis=1, dtype=pair.dtype) for p in perms], axis=1)
    return np.argmin(scores, axis=1).astype(np.int64), scores


def align_distance(ref_est, other_est):
    """tester.py:1043-1054 (MISO1_Inference) and tester.py:903-905 (clean
    alignment): D[b,i,j] = sum_{t,f} | |ref_est[b,i]| - |other_est[b,j]| |.

    ref_est, other_est : complex64 [B,S,T,F];  returns float32 [B,S,S].
    """
    a = np.abs(np.sqrt(ref_est.real.astype(np.float32) ** 2 + ref_est.imag.astype(np.float32) ** 2))
    b = np.abs(other_est).astype(np.float32)
    d = np.abs(a[:, :, None] - b[:, None, :])
    return d.sum(axis=(3, 4), dtype=np.float32)


def miso1_align(ref_est, shift_est):
    """Permutation decision of MISO1_Inference (tester.py:1053-1065).
    Returns (perm index [B], gather index [B,S]) with
    out[spk][b] = shift_est[b, gather[b, spk]]."""
    s = ref_est.shape[1]
    idx, _ = best_perm(align_distance(ref_est, shift_est), s)
    return idx, perm_table(s)[idx]


def clean_align(clean, est):
    """tester.py:889-915: D[b,i,j] = sum | |est_j| - |clean_i| |, argmin over perms;
    e_clean_MISO1[spk][b] = est[perm[spk]][b].
    clean, est : complex64 [B,S,T,F]."""
    s = clean.shape[1]
    mag_est = np.abs(np.sqrt(est.real.astype(np.float32) ** 2 + est.imag.astype(np.float32) ** 2))
    mag_clean = np.abs(clean).astype(np.float32)
    d = np.abs(mag_est[:, None, :] - mag_clean[:, :, None]).sum(axis=(3, 4), dtype=np.float32)
    idx, _ = best_perm(d, s)
    return idx, perm_table(s)[idx]


def loss_upit(estimate, ref, eps=1e-8):
    """criterion.py:8-63.  estimate, ref : complex64 [B,S,T,F].
    P[b,i,j] = sum|re_i - re_j| + sum|im_i - im_j| + sum| sqrt(re_i^2+im_i^2+eps) - |ref_j| |
    (i = estimate speaker, j = reference speaker).
    Returns (loss float32 scalar, argmin int64 [B], pair [B,S,S])."""
    er = estimate.real.astype(np.float32)[:, :, None]
    ei = estimate.imag.astype(np.float32)[:, :, None]
    rr = ref.real.astype(np.float32)[:, None]
    ri = ref.imag.astype(np.float32)[:, None]
    l1_re = np.abs(er - rr).sum(axis=(3, 4), dtype=np.float32)
    l1_im = np.abs(ei - ri).sum(axis=(3, 4), dtype=np.float32)
    emag = np.abs(np.sqrt(er * er + ei * ei + np.float32(eps)))
    rmag = np.abs(ref).astype(np.float32)[:, None]
    l1_mag = np.abs(emag - rmag).sum(axis=(3, 4), dtype=np.float32)
    pair = l1_re + l1_im + l1_mag
    idx, scores = best_perm(pair, estimate.shape[1])
    loss = scores.min(axis=1).mean(dtype=np.float32)
    return np.float32(loss), idx, pair


def loss_enhance(estimate, ref, eps=1e-8):
    """criterion.py:121-141. estimate, ref: complex64 [B,Ch,T,F] -> float32."""
    er, ei = estimate.real.astype(np.float32), estimate.imag.astype(np.float32)
    l1_re = np.abs(er - ref.real.astype(np.float32)).sum(dtype=np.float32)
    l1_im = np.abs(ei - ref.imag.astype(np.float32)).sum(dtype=np.float32)
    emag = np.abs(np.sqrt(er * er + ei * ei + np.float32(eps)))
    l1_mag = np.abs(emag - np.abs(ref).astype(np.float32)).sum(dtype=np.float32)
    return np.float32((l1_re + l1_im + l1_mag) / np.float32(estimate.shape[0]))


# --------------------------------------------------------------------------
# M1..M7: MVDR
# --------------------------------------------------------------------------
def spatial_covariance(obs):
    """tester.py:1138-1152 + the Hermitian symmetrisation of tester.py:1092,1100.
    obs: complex [B,F,C,T] -> [B,F,C,C] = 0.5*(R + R^H), R = (1/T) sum_t x x^H."""
    t = obs.shape[-1]
    r = np.einsum('...dt,...et->...de', obs, obs.conj())
    r = r / np.asarray(t, dtype=r.real.dtype)
    return 0.5 * (r + np.conj(np.swapaxes(r, -1, -2)))


def principal_eigvec(scm):
    """tester.py:1107-1115: batched eigh, column of the largest eigenvalue."""
    shape = scm.shape
    vals, vecs = np.linalg.eigh(scm.reshape((-1,) + shape[-2:]))
    k = np.argmax(vals, axis=-1)
    v = vecs[np.arange(vecs.shape[0]), :, k]
    return v.reshape(shape[:-1])


def steering_normalise(v):
    """tester.py:1119-1123: d = v / v[0]; d *= sqrt(M / ||d||_2)  (norm, not norm^2 --
    reproduced as written)."""
    m = v.shape[-1]
    d = v / v[..., :1]
    nrm = np.linalg.norm(d, axis=-1, keepdims=True)
    return d * np.sqrt(m / nrm)


def phase_correction(w):
    """tester.py:1154-1167: sequential over f,
    w[f] *= exp(-1j*angle(sum_m w[f,m] * conj(w[f-1,m]))) using the corrected w[f-1]."""
    w = w.copy()
    for f in range(1, w.shape[1]):
        c = np.sum(w[:, f, :] * w[:, f - 1, :].conj(), axis=-1, keepdims=True)
        w[:, f, :] = w[:, f, :] * np.exp(-1j * np.angle(c))
    return w


def mvdr_weights(steering, noise_scm, epsi=1e-6):
    """tester.py:1211-1225: (Phi_n + epsi*I) u = d ; w = u / (d^H u)."""
    m = steering.shape[-1]
    a = noise_scm + epsi * np.eye(m)
    u = np.linalg.solve(a, steering[..., None])[..., 0]
    denom = np.einsum('...d,...d->...', steering.conj(), u)
    return u / denom[..., None]


def apply_beamforming(source_stft, mix_stft, epsi=1e-6, return_parts=False):
    """tester.py:1071-1136 (Apply_Beamforming).

    source_stft, mix_stft : complex [B,F,C,T]
    returns               : complex [B,T,F]   (tester.py:1134 permute)

    Dtype follows numpy promotion in the reference: complex64 inputs give
    complex64 SCMs and eigenvectors; ``steering / steering[...,0]`` stays
    complex64; PhaseCorrection multiplies by a complex128 phasor in place (so
    stays complex64); ``R_noise += delta`` is in place (complex64);
    ``solve``/``einsum`` stay complex64.
    """
    src = np.asarray(source_stft)
    mix = np.asarray(mix_stft)
    m = src.shape[2]
    scm_s = spatial_covariance(src)
    scm_n = spatial_covariance(mix - src)
    v = principal_eigvec(scm_s)
    d = steering_normalise(v).astype(v.dtype)
    d = phase_correction(d)
    a = scm_n + (epsi * np.eye(m)).astype(scm_n.dtype)
    u = np.linalg.solve(a, d[..., None])[..., 0]
    denom = np.einsum('...d,...d->...', d.conj(), u)
    w = u / denom[..., None]
    y = np.einsum('...a,...at->...t', w.conj(), mix)           # [B,F,T]
    out = np.ascontiguousarray(np.transpose(y, (0, 2, 1)))
    if return_parts:
        return out, dict(scm_s=scm_s, scm_n=scm_n, eigvec=v, steering=d, weights=w)
    return out
