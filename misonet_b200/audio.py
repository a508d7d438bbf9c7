"""STFT front end (mirror of AudioDataset.STFT + "/scale", dataloader/data.py:49-66,77-79;
tester.py:992-1012).  libs/audio.py of the reference is a dead stub."""
import torch

from . import _lib


def stft_num_frames(n_samples, nperseg=256, noverlap=192):
    return int(_lib.load().miso_stft_num_frames(int(n_samples), int(nperseg), int(nperseg - noverlap)))


def stft(time_sig, nperseg=256, noverlap=192):
    """time_sig: float CUDA tensor [N, M] (the reference's per-utterance layout, data.py:49-51)
    or [B, N, M]  ->  complex64 [M, T, F] / [B, M, T, F] with F = nperseg/2 + 1.

    Equals ``scipy.signal.stft(x, window='hann', nperseg, noverlap)[2] / scale`` with
    ``scale = 1/sum(window)`` permuted to [M, T, F]: the unnormalised windowed rFFT."""
    _lib.require_cuda(time_sig, "time_sig")
    _lib.check_device(time_sig.device)
    squeeze = time_sig.dim() == 2
    x = time_sig.unsqueeze(0) if squeeze else time_sig
    if x.dim() != 3:
        raise ValueError("time_sig must be [N, M] or [B, N, M]")
    if x.dtype != torch.float32:
        x = x.float()
    B, N, M = x.shape
    hop = nperseg - noverlap
    T = stft_num_frames(N, nperseg, noverlap)
    out = torch.empty(B, M, T, nperseg // 2 + 1, dtype=torch.complex64, device=x.device)
    sb, sn, sm = x.stride()
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().miso_stft_fwd(_lib.ptr(x), sb, sn, sm, _lib.ptr(out), B, N, M, nperseg, hop, _lib.stream_ptr()),
                   "miso_stft_fwd")
    return out[0] if squeeze else out
