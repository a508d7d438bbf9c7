"""MISO1 inference over all circular microphone shifts and the speaker alignments
(mirror of Tester_*.MISO1_Inference, tester.py:1014-1068, and of the clean-reference
alignment of tester.py:889-915).

Per-utterance-correct batch semantics: the reference writes the last batch row into all
rows (tester.py:1065) and only ever runs with batch_size 1 (config/NN_BSS.yml:108-111);
at B = 1 both agree."""
import numpy as np
import torch

from . import criterion


def miso1_inference(model_sep, mix_stft, ref_ch=0, return_perm=False, stacked=False):
    """mix_stft: complex CUDA [B, Mic, T, F] -> list[Spk] of complex64 [B, Mic, T, F]
    (views of one [Spk, B, Mic, T, F] tensor; ``stacked=True`` returns that tensor).

    All M shifted forwards run as one batch of M*B samples; the permutation decision and
    the scatter into the output stay on the device (the reference does M implicit D2H
    copies, tester.py:1027,1065)."""
    B, M, T, F = mix_stft.shape
    S = model_sep.num_spks
    order = [int(q) for q in np.roll(np.arange(M), -ref_ch)]       # tester.py:1029-1030
    with torch.no_grad():
        est = model_sep.forward_shifts(mix_stft, order)             # [M*B, S, T, F]
    est = est.view(M, B, S, T, F)
    out = torch.empty(S, B, M, T, F, dtype=torch.complex64, device=est.device)
    ref = est[0]
    out[:, :, ref_ch] = ref.transpose(0, 1)                         # tester.py:1037-1038
    perm_idx = torch.zeros(M, B, dtype=torch.long, device=est.device)
    for k in range(1, M):
        q = order[k]
        _, idx, _ = criterion.pair_decide(ref, est[k], 0)           # tester.py:1043-1059
        perm_idx[q] = idx
        # out[s][b, q] = est[k][b, perm[s]]  -> a [B,S,T,F]-shaped view of `out`
        criterion.perm_gather(est[k], idx, out_view=out[:, :, q].transpose(0, 1))
    res = out if stacked else [out[s] for s in range(S)]
    return (res, perm_idx) if return_perm else res


def align_to_clean(clean_ref, miso1_stft, ref_ch=0, return_perm=False):
    """tester.py:889-915: reorder the per-speaker MISO1 outputs so that they match the clean
    references at the reference microphone.

    clean_ref  : complex CUDA [B, Spk, T, F] (clean sources at ref_ch)
    miso1_stft : [Spk, B, Mic, T, F] tensor or list[Spk] of [B, Mic, T, F]
    returns    : complex64 [Spk, B, Mic, T, F] (and the permutation index int64 [B])"""
    if isinstance(miso1_stft, (list, tuple)):
        miso1_stft = torch.stack(list(miso1_stft), dim=0)
    S, B, M, T, F = miso1_stft.shape
    est_ref = miso1_stft[:, :, ref_ch].transpose(0, 1)             # [B,S,T,F] view
    # D[b,i,j] = sum | |est_j| - |clean_i| |   (i = clean speaker, j = estimate)
    _, idx, _ = criterion.pair_decide(clean_ref.to(miso1_stft.device), est_ref, 0)
    out = torch.empty_like(miso1_stft)
    for m in range(M):
        criterion.perm_gather(miso1_stft[:, :, m].transpose(0, 1), idx, out_view=out[:, :, m].transpose(0, 1))
    return (out, idx) if return_perm else out
