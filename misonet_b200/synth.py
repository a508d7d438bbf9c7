"""Seeded synthetic inputs (host side, numpy only): "SMS-WSJ-shaped" multi-mic
mixtures, STFT-domain random spectrograms and MVDR test cases.

There is no network access for datasets, so bench.py and the tests use these.
The recipe follows SURVEY.md section 8(d): two AR-filtered noise "speakers" with a
syllabic envelope, each convolved with M random exponentially decaying room
impulse responses, summed with sensor noise, peak-normalised to 0.1 (the shipped
sample/Clean wavs peak at 0.08-0.11).
"""
import numpy as np


def random_spec(seed, shape, scale=1.0):
    """complex64 spectrogram-like tensor [..., T, F]: smooth magnitude envelope times
    random phase, plus a noise floor, O(1) magnitudes like an unnormalised STFT of a
    0.1-peak signal."""
    rng = np.random.default_rng(seed)
    t, f = shape[-2], shape[-1]
    env_t = 0.3 + np.abs(np.sin(np.linspace(0, 3.0, t)[:, None] * (1 + rng.random(shape[:-2] + (1, 1)) * 3)))
    env_f = np.exp(-np.linspace(0, 2.5, f))[None, :] * (0.5 + rng.random(shape[:-2] + (1, f)))
    mag = env_t * env_f
    z = (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)) * mag
    z += 0.01 * (rng.standard_normal(shape) + 1j * rng.standard_normal(shape))
    return (scale * z).astype(np.complex64)


def mvdr_case(seed, b, f, m, t, snr_db=15.0):
    """Rank-1-structured source images + diffuse noise, [B,F,M,T] complex64, so the
    source SCM has a large eigen-gap (iid Gaussians make the
    principal eigenvector ill-conditioned and are unsuitable for parity)."""
    rng = np.random.default_rng(seed)
    steer = rng.standard_normal((b, f, m, 1)) + 1j * rng.standard_normal((b, f, m, 1))
    steer /= np.abs(steer[:, :, :1])
    s = rng.standard_normal((b, f, 1, t)) + 1j * rng.standard_normal((b, f, 1, t))
    src = steer * s
    src += 0.03 * (rng.standard_normal((b, f, m, t)) + 1j * rng.standard_normal((b, f, m, t)))
    steer2 = rng.standard_normal((b, f, m, 1)) + 1j * rng.standard_normal((b, f, m, 1))
    interf = steer2 * (rng.standard_normal((b, f, 1, t)) + 1j * rng.standard_normal((b, f, 1, t)))
    g = 10.0 ** (-snr_db / 20.0)
    noise = g * (rng.standard_normal((b, f, m, t)) + 1j * rng.standard_normal((b, f, m, t)))
    mix = src + 0.7 * interf + noise
    return src.astype(np.complex64), mix.astype(np.complex64)


def _ar_speaker(rng, n, fs):
    """white noise through a random stable 12-pole AR filter, gated by a 4 Hz envelope."""
    from scipy.signal import lfilter
    poles = []
    for _ in range(6):
        r = rng.uniform(0.85, 0.97)
        th = rng.uniform(0.05, 0.9) * np.pi
        poles += [r * np.exp(1j * th), r * np.exp(-1j * th)]
    a = np.real(np.poly(poles))
    x = lfilter([1.0], a, rng.standard_normal(n))
    tt = np.arange(n) / fs
    env = 0.5 * (1 - np.cos(2 * np.pi * 4.0 * tt + rng.uniform(0, 2 * np.pi)))
    gate = (np.sin(2 * np.pi * 0.4 * tt + rng.uniform(0, 2 * np.pi)) > -0.3).astype(np.float64)
    return x * env * gate


def make_utterance(utt_idx, n_samples=32000, num_mics=6, num_spks=2, fs=8000):
    """Returns (mix float32 [N,M], sources float32 [S,N,M]) -- source images at every mic."""
    from scipy.signal import fftconvolve
    rng = np.random.default_rng(1234 + utt_idx)
    ang = 2 * np.pi * np.arange(num_mics) / num_mics
    mic_xy = 0.05 * np.stack([np.cos(ang), np.sin(ang)], axis=1)          # 10 cm circular array
    images = np.zeros((num_spks, n_samples, num_mics))
    for s in range(num_spks):
        dry = _ar_speaker(rng, n_samples, fs)
        doa = rng.uniform(0, 2 * np.pi)
        t60 = rng.uniform(0.2, 0.5)
        rir_len = int(0.3 * fs)
        decay = np.exp(-6.9 * np.arange(rir_len) / (t60 * fs))
        for m in range(num_mics):
            delay = 20 + int(round(fs * (mic_xy[m] @ np.array([np.cos(doa), np.sin(doa)])) / 343.0 * 4))
            h = 0.3 * rng.standard_normal(rir_len) * decay
            h[:delay] = 0.0
            h[delay] = 1.0
            images[s, :, m] = fftconvolve(dry, h)[:n_samples]
    mix = images.sum(axis=0)
    snr = rng.uniform(20, 30)
    mix = mix + rng.standard_normal(mix.shape) * mix.std() * 10 ** (-snr / 20)
    g = 0.1 / np.abs(mix).max()
    return (mix * g).astype(np.float32), (images * g).astype(np.float32)
