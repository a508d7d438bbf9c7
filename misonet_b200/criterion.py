"""Loss functions and permutation decisions (mirror of criterion.py:8-63 loss_uPIT and
criterion.py:121-141 loss_Enhance, plus the |.|-distance alignment used by
tester.py:1043-1065 and tester.py:889-915)."""
from itertools import permutations

import torch

from . import _lib

_ws_cache = {}


def _workspace(nbytes, dev):
    ws = _ws_cache.get(dev)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 16), dtype=torch.uint8, device=dev)
        _ws_cache[dev] = ws
    return ws


def perm_table(num_spks, device=None):
    """list(itertools.permutations(range(S))) as int64 [P, S] (criterion.py:49)."""
    return torch.tensor(list(permutations(range(num_spks))), dtype=torch.long, device=device)


def _planes(t):
    """[B,S,T,F] complex64 with dense [T,F] planes; returns (tensor, batch stride, speaker stride)."""
    t = t.to(torch.complex64)
    if t.stride(-1) != 1 or t.stride(-2) != t.shape[-1]:
        t = t.contiguous()
    return t, t.stride(0), t.stride(1)


def pair_decide(a, b, mode, want_loss=False):
    """a, b: complex CUDA [B,S,T,F].  mode 0: sum||a_i|-|b_j|| ; mode 1: the uPIT L1 triple.
    Returns (pair float32 [B,S,S], perm index int64 [B], loss float32 [] or None)."""
    _lib.require_cuda(a, "a")
    _lib.require_cuda(b, "b")
    _lib.check_device(a.device)
    a, a_sb, a_ss = _planes(a)
    b, b_sb, b_ss = _planes(b)
    B, S, T, F = a.shape
    if tuple(b.shape) != (B, S, T, F):
        raise ValueError(f"shape mismatch {tuple(a.shape)} vs {tuple(b.shape)}")
    lib = _lib.load()
    dev = a.device
    pair = torch.empty(B, S, S, dtype=torch.float32, device=dev)
    idx = torch.empty(B, dtype=torch.long, device=dev)
    loss = torch.empty((), dtype=torch.float32, device=dev) if want_loss else None
    nbytes = lib.miso_pair_workspace_bytes(B, S, T, F)
    ws = _workspace(nbytes, dev)
    with torch.cuda.device(dev):
        _lib.check(lib.miso_pair_fwd(_lib.ptr(a), a_sb, a_ss, _lib.ptr(b), b_sb, b_ss, B, S, T, F, int(mode), _lib.ptr(pair),
                                     _lib.ptr(idx), _lib.ptr(loss) if want_loss else None, _lib.ptr(ws), ws.numel(),
                                     _lib.stream_ptr()), "miso_pair_fwd")
    return pair, idx, loss


def perm_gather(src, idx, out=None, out_view=None):
    """out[s][b] = src[b][perm_{idx[b]}[s]] for src complex64 [B,S,T,F].
    ``out_view`` (optional) is a [B,S,T,F]-shaped strided view to write into."""
    src, s_sb, s_ss = _planes(src)
    B, S, T, F = src.shape
    if out_view is None:
        out_view = torch.empty_like(src)
    if out_view.stride(-1) != 1 or out_view.stride(-2) != F or out_view.dtype != torch.complex64:
        raise ValueError("out_view must have dense complex64 [T,F] planes")
    with torch.cuda.device(src.device):
        _lib.check(_lib.load().miso_perm_gather(_lib.ptr(src), s_sb, s_ss, _lib.ptr(out_view), out_view.stride(0),
                                                out_view.stride(1), _lib.ptr(idx), B, S, T, F, _lib.stream_ptr()),
                   "miso_perm_gather")
    return out_view


class _UpitFunction(torch.autograd.Function):
    """loss_uPIT with its gradient w.r.t. the estimate (miso_upit_bwd, include/misonet_b200.h)."""

    @staticmethod
    def forward(ctx, est, ref):
        _, idx, loss = pair_decide(est, ref, 1, want_loss=True)
        ctx.save_for_backward(est, ref, idx)
        ctx.mark_non_differentiable(idx)
        return loss, idx

    @staticmethod
    def backward(ctx, gloss, _gidx):
        est, ref, idx = ctx.saved_tensors
        est, e_sb, e_ss = _planes(est.detach())
        ref, r_sb, r_ss = _planes(ref)
        B, S, T, F = est.shape
        grad = torch.empty(B, S, T, F, dtype=torch.complex64, device=est.device)
        g = gloss.detach().to(torch.float32).reshape(1).contiguous()
        with torch.cuda.device(est.device):
            _lib.check(_lib.load().miso_upit_bwd(_lib.ptr(est), e_sb, e_ss, _lib.ptr(ref), r_sb, r_ss, _lib.ptr(idx), B, S, T, F,
                                                 _lib.ptr(g), _lib.ptr(grad), _lib.stream_ptr()), "miso_upit_bwd")
        return grad, None


def loss_uPIT(num_spks, estimate_clean, ref_clean, return_perm=False):
    """criterion.py:8-63.  estimate_clean: complex [B,Spks,T,F]; ref_clean: list[Spks] of complex
    [B,T,F] (or a stacked [B,Spks,T,F] tensor).  Returns the scalar loss (float32 CUDA tensor);
    with return_perm=True also the argmin permutation index int64 [B]."""
    ref = torch.stack(list(ref_clean), dim=1) if isinstance(ref_clean, (list, tuple)) else ref_clean
    if estimate_clean.shape[1] != num_spks:
        raise ValueError("estimate does not have num_spks speakers")
    ref = ref.to(estimate_clean.device)
    if torch.is_grad_enabled() and estimate_clean.requires_grad:
        loss, idx = _UpitFunction.apply(estimate_clean, ref)      # training: trainer.py:170-172, loss.backward()
    else:
        _, idx, loss = pair_decide(estimate_clean, ref, 1, want_loss=True)
    return (loss, idx) if return_perm else loss


class _EnhanceFunction(torch.autograd.Function):
    """loss_Enhance with its gradient w.r.t. the estimate (miso_loss_enhance_bwd)."""

    @staticmethod
    def forward(ctx, est, rf):
        B = est.shape[0]
        loss = torch.empty((), dtype=torch.float32, device=est.device)
        ws = _workspace(1024 * 8, est.device)
        with torch.cuda.device(est.device):
            _lib.check(_lib.load().miso_loss_enhance_fwd(_lib.ptr(est), _lib.ptr(rf), B, est.numel() // B, _lib.ptr(loss), _lib.ptr(ws),
                                                         ws.numel(), _lib.stream_ptr()), "miso_loss_enhance_fwd")
        ctx.save_for_backward(est, rf)
        return loss

    @staticmethod
    def backward(ctx, gloss):
        est, rf = ctx.saved_tensors
        B = est.shape[0]
        grad = torch.empty_like(est)
        g = gloss.detach().to(torch.float32).reshape(1).contiguous()
        with torch.cuda.device(est.device):
            _lib.check(_lib.load().miso_loss_enhance_bwd(_lib.ptr(est), _lib.ptr(rf), B, est.numel() // B, _lib.ptr(g), _lib.ptr(grad),
                                                         _lib.stream_ptr()), "miso_loss_enhance_bwd")
        return grad, None


def loss_Enhance(estimate, ref):
    """criterion.py:121-141.  estimate, ref: complex [B,Ch,T,F] -> float32 scalar."""
    _lib.require_cuda(estimate, "estimate")
    _lib.check_device(estimate.device)
    est = estimate.to(torch.complex64).contiguous()
    rf = ref.to(device=est.device, dtype=torch.complex64).contiguous()
    if est.shape != rf.shape:
        raise ValueError("shape mismatch")
    if torch.is_grad_enabled() and estimate.requires_grad:
        return _EnhanceFunction.apply(est, rf)               # training: trainer.py:398-443
    B = est.shape[0]
    n = est.numel() // B
    loss = torch.empty((), dtype=torch.float32, device=est.device)
    ws = _workspace(1024 * 8, est.device)
    with torch.cuda.device(est.device):
        _lib.check(_lib.load().miso_loss_enhance_fwd(_lib.ptr(est), _lib.ptr(rf), B, n, _lib.ptr(loss), _lib.ptr(ws),
                                                     ws.numel(), _lib.stream_ptr()), "miso_loss_enhance_fwd")
    return loss
