"""misonet_b200: B200-native MISO-BF-MISO hot path (STFT, MISO_1/MISO_3 forward, speaker
alignment / uPIT decisions, MVDR) behind the call signatures of yuhogun0908/MISOnet.

Importing the package does not load the CUDA library; the first call does, and raises if
it is missing (there is no CPU or stock-PyTorch fallback)."""
__all__ = ["MISO_1", "MISO_3", "Apply_Beamforming", "mvdr", "miso1_inference", "align_to_clean", "loss_uPIT",
           "loss_Enhance", "stft", "istft", "MisoBfMiso", "B200HotPath", "separate_recording"]


def __getattr__(name):
    if name in ("MISO_1", "MISO_3"):
        from . import model
        return getattr(model, name)
    if name in ("Apply_Beamforming", "mvdr"):
        from . import beamforming
        return getattr(beamforming, name)
    if name in ("miso1_inference", "align_to_clean"):
        from . import separation
        return getattr(separation, name)
    if name in ("loss_uPIT", "loss_Enhance"):
        from . import criterion
        return getattr(criterion, name)
    if name == "stft":
        from . import audio
        return audio.stft
    if name == "istft":
        from . import audio
        return audio.istft
    if name == "separate_recording":
        from . import continuous
        return continuous.separate_recording
    if name == "MisoBfMiso":
        from . import pipeline
        return pipeline.MisoBfMiso
    if name == "B200HotPath":
        from . import dropin
        return dropin.B200HotPath
    raise AttributeError(name)
