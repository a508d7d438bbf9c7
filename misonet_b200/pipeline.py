"""The MISO-BF-MISO chunk pipeline on one GPU (what Tester_Enhance.inference does per 4 s
chunk, tester.py:857-939): STFT -> MISO1 over all mic shifts -> speaker alignment -> MVDR per
speaker -> MISO3 per speaker.  Utterances are independent (every normalisation is
per-sample, MVDR is per utterance and frequency), so multi-GPU use is plain utterance
sharding: see ``shard_range`` and bench.py."""
import torch

from . import audio, beamforming, separation


def shard_range(n_items, rank, world_size):
    """Contiguous block partition of ``n_items`` utterances: [lo, hi) for ``rank``."""
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class MisoBfMiso:
    def __init__(self, model_sep, model_enh, num_spks=2, ref_ch=0, nperseg=256, noverlap=192, epsi=1e-6):
        self.model_sep, self.model_enh = model_sep, model_enh
        self.num_spks, self.ref_ch = num_spks, ref_ch
        self.nperseg, self.noverlap, self.epsi = nperseg, noverlap, epsi

    @torch.no_grad()
    def from_stft(self, mix_stft, clean_ref=None):
        """mix_stft: complex CUDA [B, Mic, T, F]; clean_ref (optional, evaluation only):
        complex [B, Spk, T, F] clean sources at the reference mic (tester.py:889-915).
        Returns dict(miso1 [S,B,M,T,F], beamformed [S,B,T,F], enhanced [B,S,T,F])."""
        miso1 = separation.miso1_inference(self.model_sep, mix_stft, self.ref_ch, stacked=True)
        if clean_ref is not None:
            miso1 = separation.align_to_clean(clean_ref, miso1, self.ref_ch)
        bf = beamforming.mvdr(miso1, mix_stft, self.epsi)                          # [S,B,T,F]
        enhanced = []
        for s in range(self.num_spks):                                            # tester.py:935-939
            e = self.model_enh(mix_stft, bf[s].unsqueeze(1), miso1[s][:, self.ref_ch].unsqueeze(1))
            enhanced.append(e[:, 0])
        return dict(miso1=miso1, beamformed=bf, enhanced=torch.stack(enhanced, dim=1))

    @torch.no_grad()
    def __call__(self, mix_time, clean_ref=None):
        """mix_time: float CUDA [B, N, Mic] time-domain chunks."""
        mix_stft = audio.stft(mix_time, self.nperseg, self.noverlap)
        out = self.from_stft(mix_stft, clean_ref)
        out["mix_stft"] = mix_stft
        return out
