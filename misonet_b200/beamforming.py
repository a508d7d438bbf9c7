"""MVDR beamformer (mirror of Tester_*.Apply_Beamforming, tester.py:1071-1136, and of its
helpers get_spatial_covariance_matrix / PhaseCorrection / get_mvdr_beamformer /
apply_beamformer, tester.py:1138-1167, 1211-1228)."""
import numpy as np
import torch

from . import _lib

_ws_cache = {}


def _workspace(nbytes, dev):
    ws = _ws_cache.get(dev)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=dev)
        _ws_cache[dev] = ws
    return ws


def mvdr(sources, mix, epsi=1e-6, return_weights=False):
    """Device-native form: all sources of one mixture in one pass.

    sources : complex64 CUDA [S, B, M, T, F] (or a list of S tensors [B, M, T, F])
    mix     : complex64 CUDA [B, M, T, F]
    returns : complex64 [S, B, T, F]  (and the beamformer weights [S, B, F, M])"""
    if isinstance(sources, (list, tuple)):
        sources = torch.stack([s.to(torch.complex64) for s in sources], dim=0)
    _lib.require_cuda(sources, "sources")
    _lib.require_cuda(mix, "mix")
    _lib.check_device(mix.device)
    src = sources.to(torch.complex64).contiguous()
    mx = mix.to(torch.complex64).contiguous()
    S, B, M, T, F = src.shape
    if tuple(mx.shape) != (B, M, T, F):
        raise ValueError(f"mix shape {tuple(mx.shape)} does not match sources {tuple(src.shape)}")
    lib = _lib.load()
    out = torch.empty(S, B, T, F, dtype=torch.complex64, device=mx.device)
    w = torch.empty(S, B, F, M, dtype=torch.complex64, device=mx.device)
    nbytes = lib.miso_mvdr_workspace_bytes(S, B, M, T, F)
    ws = _workspace(nbytes, mx.device)
    sb, sm, st, sf = mx.stride()
    with torch.cuda.device(mx.device):
        _lib.check(lib.miso_mvdr_fwd(_lib.ptr(src), src.stride(0), _lib.ptr(mx), sb, sm, st, sf, _lib.ptr(out), _lib.ptr(w),
                                     S, B, M, T, F, float(epsi), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()), "miso_mvdr_fwd")
    return (out, w) if return_weights else out


def mvdr_utterance(sources, mix, epsi=1e-6, group=None):
    """Utterance-level MVDR of a recording whose frames are spread over the ranks of ``group`` (the reference's
    ``utterance_flag`` mode, tester.py:425-449: spatial covariances over ALL frames, then every frame filtered with
    the same beamformer).  Every rank passes ITS frames: sources complex64 CUDA [S, B, M, T_local, F], mix
    [B, M, T_local, F]; returns its frames of the output [S, B, T_local, F] and the beamformers [S, B, F, M].

    Collective: one all-gather of the partial covariance sums (S*B*tsplit*84*F floats per rank for 6 mics) and one
    all-reduce of the frame count.  With one rank the result equals :func:`mvdr` bit for bit."""
    import torch.distributed as dist
    if isinstance(sources, (list, tuple)):
        sources = torch.stack([s.to(torch.complex64) for s in sources], dim=0)
    _lib.require_cuda(sources, "sources")
    _lib.require_cuda(mix, "mix")
    _lib.check_device(mix.device)
    src = sources.to(torch.complex64).contiguous()
    mx = mix.to(torch.complex64).contiguous()
    S, B, M, T, F = src.shape
    if tuple(mx.shape) != (B, M, T, F):
        raise ValueError(f"mix shape {tuple(mx.shape)} does not match sources {tuple(src.shape)}")
    lib = _lib.load()
    dev = mx.device
    tsplit = int(lib.miso_mvdr_tsplit(B, F))
    nv = 2 * M * (M + 1)
    partial = torch.empty(S * B, tsplit, nv, F, dtype=torch.float32, device=dev)
    sb, sm, st, sf = mx.stride()
    on = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    with torch.cuda.device(dev):
        stream = _lib.stream_ptr()
        _lib.check(lib.miso_mvdr_scm(_lib.ptr(src), src.stride(0), _lib.ptr(mx), sb, sm, st, sf, _lib.ptr(partial), S, B, M, T, F,
                                     stream), "miso_mvdr_scm")
        t_total = torch.tensor([T], dtype=torch.int64, device=dev)
        if on:
            world = dist.get_world_size(group)
            parts = [torch.empty_like(partial) for _ in range(world)]
            dist.all_gather(parts, partial, group=group)
            partial = torch.stack(parts, dim=1).reshape(S * B, world * tsplit, nv, F).contiguous()   # rank-major split axis
            dist.all_reduce(t_total, group=group)
        w = torch.empty(S, B, F, M, dtype=torch.complex64, device=dev)
        ws = _workspace(S * B * F * M * 16, dev)
        _lib.check(lib.miso_mvdr_weights(_lib.ptr(partial), partial.shape[1], int(t_total.item()), _lib.ptr(w), S, B, M, F,
                                         float(epsi), _lib.ptr(ws), ws.numel(), stream), "miso_mvdr_weights")
        out = torch.empty(S, B, T, F, dtype=torch.complex64, device=dev)
        _lib.check(lib.miso_mvdr_apply(_lib.ptr(mx), sb, sm, st, sf, _lib.ptr(w), _lib.ptr(out), S, B, M, T, F, stream),
                   "miso_mvdr_apply")
    return out, w


def Apply_Beamforming(source_stft, mix_stft, epsi=1e-6, device=None):
    """Drop-in form of tester.py:1071-1136.

    source_stft, mix_stft : complex [B, F, Ch, T], numpy arrays or torch tensors (CPU or CUDA)
    returns               : torch complex64 [B, T, F] on the device of the inputs (CPU inputs
                            give a CPU result, as the reference's callers expect, tester.py:924-931)."""
    was_cpu = True
    if isinstance(source_stft, np.ndarray):
        source_stft = torch.from_numpy(np.ascontiguousarray(source_stft))
    else:
        was_cpu = not source_stft.is_cuda
    if isinstance(mix_stft, np.ndarray):
        mix_stft = torch.from_numpy(np.ascontiguousarray(mix_stft))
    if device is None:
        device = source_stft.device if source_stft.is_cuda else torch.device("cuda", torch.cuda.current_device())
    # [B,F,Ch,T] -> [B,Ch,T,F]: the layout the kernels stream (F contiguous)
    src = source_stft.to(device=device, dtype=torch.complex64).permute(0, 2, 3, 1).contiguous()
    mix = mix_stft.to(device=device, dtype=torch.complex64).permute(0, 2, 3, 1).contiguous()
    out = mvdr(src.unsqueeze(0), mix, epsi)[0]
    return out.cpu() if was_cpu else out
