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


MAX_INT16 = 32767     # np.iinfo(np.int16).max (tester.py:36,280)


def to_int16(wave):
    """float CUDA waveform -> int16, the reference's output sample format (``wave * MaxINT16`` then ``astype(np.int16)``,
    tester.py:155-157, 444-446, 950-952)."""
    _lib.require_cuda(wave, "wave")
    _lib.check_device(wave.device)
    x = wave.to(torch.float32).contiguous()
    out = torch.empty(x.shape, dtype=torch.int16, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().miso_wave_to_int16(_lib.ptr(x), _lib.ptr(out), x.numel(), float(MAX_INT16), _lib.stream_ptr()),
                   "miso_wave_to_int16")
    return out


def write_wav(path, pcm, fs, subtype="PCM_24"):
    """Write int16 samples [N] or [N, channels] (a CPU tensor / array; the reference passes ``x.T``) as a RIFF/WAVE file
    the way ``sf.write(path, int16_data, fs, 'PCM_24')`` does (tester.py:447, 971-972): libsndfile widens int16 to 24 bits
    by a left shift of 8.  ``subtype="PCM_16"`` writes the samples as they are.  File I/O only -- no arithmetic."""
    import struct
    import numpy as np
    a = np.asarray(pcm.cpu() if hasattr(pcm, "cpu") else pcm)
    if a.dtype != np.int16:
        raise TypeError("write_wav expects int16 samples (audio.to_int16)")
    if a.ndim == 1:
        a = a[:, None]
    n, ch = a.shape
    if subtype == "PCM_24":
        width = 3
        v = a.astype(np.int32) << 8
        raw = np.empty((n, ch, 3), dtype=np.uint8)
        raw[..., 0] = v & 0xFF
        raw[..., 1] = (v >> 8) & 0xFF
        raw[..., 2] = (v >> 16) & 0xFF
        data = raw.tobytes()
    elif subtype == "PCM_16":
        width = 2
        data = a.astype("<i2").tobytes()
    else:
        raise ValueError("subtype must be PCM_24 or PCM_16")
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVE")
        f.write(b"fmt " + struct.pack("<IHHIIHH", 16, 1, ch, int(fs), int(fs) * ch * width, ch * width, 8 * width))
        f.write(b"data" + struct.pack("<I", len(data)))
        f.write(data)


def read_wav(path):
    """Read a PCM (16 / 24 / 32 bit) or float32 RIFF/WAVE file as ``sf.read`` does (data.py:516, tester.py: the samples as
    floats in [-1, 1), [N, channels]); returns (float32 numpy array, fs)."""
    import struct
    import numpy as np
    raw = open(path, "rb").read()
    if raw[:4] != b"RIFF" or raw[8:12] != b"WAVE":
        raise ValueError(f"{path}: not a RIFF/WAVE file")
    pos, fmt, data = 12, None, None
    while pos + 8 <= len(raw):
        cid, size = raw[pos:pos + 4], struct.unpack("<I", raw[pos + 4:pos + 8])[0]
        body = raw[pos + 8:pos + 8 + size]
        if cid == b"fmt ":
            fmt = struct.unpack("<HHIIHH", body[:16])
        elif cid == b"data":
            data = body
        pos += 8 + size + (size & 1)
    if fmt is None or data is None:
        raise ValueError(f"{path}: missing fmt / data chunk")
    tag, ch, fs, _, _, bits = fmt
    if tag == 3 and bits == 32:
        x = np.frombuffer(data, dtype="<f4").astype(np.float32)
    elif tag in (1, 0xFFFE) and bits == 16:
        x = np.frombuffer(data, dtype="<i2").astype(np.float32) / 32768.0
    elif tag in (1, 0xFFFE) and bits == 24:
        b = np.frombuffer(data, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        v = np.where(v >= 1 << 23, v - (1 << 24), v)
        x = v.astype(np.float32) / float(1 << 23)
    elif tag in (1, 0xFFFE) and bits == 32:
        x = np.frombuffer(data, dtype="<i4").astype(np.float32) / float(1 << 31)
    else:
        raise ValueError(f"{path}: unsupported wav format tag {tag}, {bits} bits")
    return x.reshape(-1, ch), fs
