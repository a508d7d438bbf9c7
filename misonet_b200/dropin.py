"""Method-level drop-ins for the reference's tester/dataset classes.

The reference has no plugin interface: MVDR, the shifted MISO1 inference and the STFT are
methods of ``Tester_*`` / ``AudioDataset`` (tester.py:992-1244, dataloader/data.py:49-66).
``B200HotPath`` overrides exactly those methods, with the reference's names, argument
meaning and return conventions, so that

    class Tester(B200HotPath, tester.Tester_Enhance): pass

runs the reference's own ``inference`` loop on the CUDA library."""
import numpy as np
import torch

from . import audio, beamforming, separation


class B200HotPath:
    # attributes the reference classes already provide: model_sep, model, num_spks, ref_ch,
    # nperseg, noverlap, scale, device

    def _b200_device(self):
        dev = getattr(self, "device", None)
        if isinstance(dev, int):
            return torch.device("cuda", dev)
        if dev is None or (isinstance(dev, str) and not dev.startswith("cuda")):
            return torch.device("cuda", torch.cuda.current_device())
        return torch.device(dev)

    def MISO1_Inference(self, mix_stft, ref_ch=0):
        """tester.py:1014-1068: [B,Mic,T,F] -> list[Spk] of complex64 [B,Mic,T,F] on the CPU
        (the reference allocates its outputs with torch.empty on the CPU, tester.py:1027)."""
        out = separation.miso1_inference(self.model_sep, mix_stft.to(self._b200_device()), ref_ch)
        return [o.cpu() for o in out]

    def Apply_Beamforming(self, source_stft, mix_stft, epsi=1e-6):
        """tester.py:1071-1136: numpy/torch [B,F,Ch,T] x2 -> torch complex64 [B,T,F]."""
        return beamforming.Apply_Beamforming(source_stft, mix_stft, epsi, device=self._b200_device())

    def MISO3_inference(self, mix_stft, bf_stft, MISO1_stft):
        """tester.py:1231-1244."""
        with torch.no_grad():
            return self.model(mix_stft, bf_stft, MISO1_stft)

    def STFT(self, time_sig):
        """tester.py:992-1012: [T,Nch] -> torch complex64 [Nch,F,T] with scipy's 'spectrum'
        scaling (the caller divides by self.scale afterwards, dataloader/data.py:77)."""
        x = torch.as_tensor(np.asarray(time_sig), dtype=torch.float32)
        if x.shape[1] > x.shape[0]:
            x = x.T
        spec = audio.stft(x.to(self._b200_device()), self.nperseg, self.noverlap)     # [M,T,F], unnormalised
        return (spec * float(self.scale)).permute(0, 2, 1).cpu()

    def ISTFT(self, FT_sig):
        """tester.py:979-990: [F,T] (or [C,F,T]) spectrogram already multiplied by self.scale -> numpy [T_samples]
        (or [C,T_samples]), i.e. scipy.signal.istft with the tester's window / nperseg / noverlap."""
        z = torch.as_tensor(np.asarray(FT_sig)) if not torch.is_tensor(FT_sig) else FT_sig
        z = z.to(self._b200_device()).to(torch.complex64) / float(self.scale)      # back to the unnormalised spectrum
        return audio.istft(z.transpose(-1, -2), self.nperseg, self.noverlap).cpu().numpy()
