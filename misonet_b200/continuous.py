"""Long recordings: the reference's chunked inference (AudioDataset_Test.__getitem__, dataloader/data.py:524-597,
and the per-chunk loop of Tester_Enhance.inference, tester.py:857-974) on top of the chunk pipeline.

A recording is cut into independent, non-overlapping chunks of ``chunk_size`` samples (4 s), the last one
zero-padded by ``gap`` samples; every chunk runs STFT -> MISO1 x mics -> MVDR -> MISO3 -> ISTFT on its own (its own
normalisation statistics and spatial covariances); the waveforms are concatenated and the padding is cut off the
last chunk (tester.py:961-969).  Chunks are independent, so they are block-partitioned over the ranks
(BASELINE configs[4]: 8000 frames = 16 chunks of 500) and the only collective is the gather of the waveforms."""
import torch
import torch.distributed as dist

from . import audio
from .pipeline import shard_range

MAX_INT16 = audio.MAX_INT16  # tester.py: self.MaxInt16


def chunk_signal(wav, chunk_size):
    """wav: [N, M] -> (chunks [C, chunk_size, M], gap).  dataloader/data.py:538-597: full chunks, then the remainder
    zero-padded by ``gap``.  (For N an exact multiple of chunk_size the reference appends one all-zero chunk whose
    output is trimmed away entirely, and for N == chunk_size it leaves ``gap`` unbound; both cases are C = N /
    chunk_size chunks and gap = 0 here.)"""
    n, m = wav.shape
    if n <= 0:
        raise ValueError("empty recording")
    c = (n + chunk_size - 1) // chunk_size
    gap = c * chunk_size - n
    if gap:
        wav = torch.cat([wav, wav.new_zeros(gap, m)], dim=0)
    return wav.reshape(c, chunk_size, m), gap


def gather_chunks(local, n_total, world=None):
    """local: [c_local, ...] results of this rank's chunks (block partition over ``world`` ranks, default: the process
    group's size) -> [n_total, ...] on every rank."""
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1) or world == 1:
        return local
    if world is None:
        world = dist.get_world_size()
    if world != dist.get_world_size():
        raise ValueError(f"chunks were partitioned over {world} ranks but the process group has {dist.get_world_size()}")
    cap = (n_total + world - 1) // world
    buf = local.new_zeros((cap,) + tuple(local.shape[1:]))
    buf[: local.shape[0]] = local
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    out = []
    for r, p in enumerate(parts):
        lo, hi = shard_range(n_total, r, world)
        out.append(p[: hi - lo])
    return torch.cat(out, dim=0)


@torch.no_grad()
def separate_recording(pipe, wav, chunk_size=32000, rank=0, world=1, to_int16=False, batch=8, gather=True):
    """pipe: pipeline.MisoBfMiso; wav: float CUDA [N, Mic].  Returns the enhanced sources [Spk, N] (float32, or int16
    scaled by 32767 as the reference writes them, tester.py:950-951) -- of this rank's chunks only when
    ``gather=False`` (then [c_local, Spk, chunk_size])."""
    chunks, gap = chunk_signal(wav, chunk_size)
    n_chunks = chunks.shape[0]
    lo, hi = shard_range(n_chunks, rank, world)
    outs = []
    for i in range(lo, hi, batch):
        res = pipe(chunks[i:min(hi, i + batch)])                         # enhanced: [b, Spk, T, F]
        w = audio.istft(res["enhanced"], pipe.nperseg, pipe.noverlap)   # [b, Spk, (T-1) hop]
        if w.shape[-1] != chunk_size:                                    # tester.py:953 assert
            raise RuntimeError(f"ISTFT length {w.shape[-1]} does not match the chunk size {chunk_size}")
        outs.append(w)
    spk = pipe.num_spks
    local = torch.cat(outs, dim=0) if outs else wav.new_zeros((0, spk, chunk_size))
    if not gather:
        return local
    full = gather_chunks(local, n_chunks, world)                          # [C, Spk, chunk]
    sig = full.permute(1, 0, 2).reshape(spk, n_chunks * chunk_size)
    sig = sig[:, : n_chunks * chunk_size - gap]                           # tester.py:961-963
    if to_int16:
        sig = audio.to_int16(sig.contiguous())                           # wave_to_int16_kernel: x * MaxINT16 in double, truncation (tester.py:155-157)
    return sig


@torch.no_grad()
def beamform_recording(model_sep, wav, chunk_size=32000, ref_ch=0, nperseg=256, noverlap=192, epsi=1e-6, clean=None,
                       rank=0, world=1, batch=4):
    """Utterance-wise beamforming of a long recording (Tester_Beamforming.inference with ``utterance_flag``,
    tester.py:340-449): per chunk MISO1 over all mic shifts (and, for evaluation, alignment to the clean references),
    ISTFT of every speaker image at every mic and concatenation along time (tester.py:395-423); then ONE STFT of the
    whole recording (tester.py:426-434) and an MVDR whose spatial covariances span all frames (tester.py:442).

    wav: float CUDA [N, Mic]; clean (optional): float CUDA [Spk, N] clean sources at ``ref_ch``.
    Returns dict(beamformed=[Spk, T_all, F] complex64, wav=[Spk, N'] float32 with N' = (T_all - 1) * hop,
    miso1_wav=[Spk, Mic, N] the concatenated MISO1 images).  Chunks are block-partitioned over the ranks; every rank
    then holds the whole recording (one waveform gather) and the final MVDR runs replicated."""
    from . import beamforming, separation
    n, n_mic = wav.shape
    chunks, gap = chunk_signal(wav, chunk_size)
    cchunks = None
    if clean is not None:
        cchunks = torch.stack([chunk_signal(c[:, None], chunk_size)[0][:, :, 0] for c in clean], dim=1)   # [C, Spk, chunk]
    n_chunks = chunks.shape[0]
    lo, hi = shard_range(n_chunks, rank, world)
    outs = []
    for i in range(lo, hi, batch):
        j = min(hi, i + batch)
        mix_stft = audio.stft(chunks[i:j], nperseg, noverlap)                                  # [b, Mic, T, F]
        miso1 = separation.miso1_inference(model_sep, mix_stft, ref_ch, stacked=True)          # [Spk, b, Mic, T, F]
        if cchunks is not None:
            cref = audio.stft(cchunks[i:j].permute(0, 2, 1).contiguous(), nperseg, noverlap)   # [b, Spk, T, F]
            miso1 = separation.align_to_clean(cref, miso1, ref_ch)
        outs.append(audio.istft(miso1, nperseg, noverlap).permute(1, 0, 2, 3))                 # [b, Spk, Mic, chunk]
    spk = model_sep.num_spks
    local = torch.cat(outs, dim=0) if outs else wav.new_zeros((0, spk, n_mic, chunk_size))
    full = gather_chunks(local.contiguous(), n_chunks, world)                                  # [C, Spk, Mic, chunk]
    img = full.permute(1, 2, 0, 3).reshape(spk, n_mic, n_chunks * chunk_size)[:, :, :n]        # gap trimmed (tester.py:405-417)
    src_stft = audio.stft(img.permute(0, 2, 1).contiguous(), nperseg, noverlap)                # [Spk, Mic, T_all, F]
    mix_all = audio.stft(wav[None], nperseg, noverlap)                                         # [1, Mic, T_all, F]
    bf = beamforming.mvdr(src_stft[:, None], mix_all, epsi)[:, 0]                              # [Spk, T_all, F]
    return dict(beamformed=bf, wav=audio.istft(bf, nperseg, noverlap), miso1_wav=img)
