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

MAX_INT16 = 32767  # tester.py: self.MaxInt16


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
        sig = (sig * MAX_INT16).to(torch.int16)                          # numpy astype truncates toward zero, like .to()
    return sig
