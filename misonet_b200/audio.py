"""STFT front end (mirror of AudioDataset.STFT + "/scale", dataloader/data.py:49-66,77-79;
tester.py:992-1012) and ISTFT back end (mirror of Tester_*.ISTFT of ``spec * scale``, tester.py:979-990, 949-957).  libs/audio.py of the reference is a dead stub."""
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


_ws = {}


def istft_num_samples(n_frames, nperseg=256, noverlap=192):
    return int(_lib.load().miso_istft_num_samples(int(n_frames), int(nperseg), int(nperseg - noverlap)))


def istft(spec, nperseg=256, noverlap=192):
    """spec: complex CUDA tensor [..., T, F] (the layout the networks produce; the reference permutes to [F, T] and
    multiplies by ``scale`` first, tester.py:949)  ->  float32 [..., (T - 1) * hop].

    Equals ``scipy.signal.istft(spec.T * scale, window='hann', nperseg, noverlap)[1]`` with ``scale = 1/sum(window)``:
    the inverse of :func:`stft` (scipy defaults: half-window boundary trimmed, squared-window normalisation)."""
    _lib.require_cuda(spec, "spec")
    _lib.check_device(spec.device)
    if spec.dim() < 2 or spec.shape[-1] != nperseg // 2 + 1:
        raise ValueError(f"spec must be [..., T, {nperseg // 2 + 1}], got {tuple(spec.shape)}")
    x = spec.to(torch.complex64).contiguous()
    lead = x.shape[:-2]
    T, F = x.shape[-2], x.shape[-1]
    S = 1
    for d in lead:
        S *= int(d)
    hop = nperseg - noverlap
    n_out = istft_num_samples(T, nperseg, noverlap)
    out = torch.empty(S, n_out, dtype=torch.float32, device=x.device)
    lib = _lib.load()
    need = int(lib.miso_istft_workspace_bytes(S, T, nperseg))
    ws = _ws.get(x.device)
    if ws is None or ws.numel() < need:
        ws = torch.empty(max(need, 1 << 20), dtype=torch.uint8, device=x.device)
        _ws[x.device] = ws
    with torch.cuda.device(x.device):
        _lib.check(lib.miso_istft_fwd(_lib.ptr(x), T * F, F, 1, _lib.ptr(out), S, T, nperseg, hop, _lib.ptr(ws), ws.numel(),
                                      _lib.stream_ptr()), "miso_istft_fwd")
    return out.reshape(*lead, n_out)
