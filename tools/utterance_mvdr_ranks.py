"""Utterance-level MVDR across ranks (torchrun): every rank holds a slice of the frames of one recording; the result must
match the one-rank computation to rounding (the partial covariance sums are added in a different order)."""
import json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from misonet_b200 import beamforming, distributed as D, synth
from misonet_b200.pipeline import shard_range
rank, world, local = D.init_from_env()
torch.cuda.set_device(local)
T, F = 8016, 257
src_np, mix_np = synth.mvdr_case(3, 1, F, 6, T)
src = torch.from_numpy(src_np).permute(0, 2, 3, 1).contiguous().cuda()[None]   # [S=1,B,M,T,F]
mix = torch.from_numpy(mix_np).permute(0, 2, 3, 1).contiguous().cuda()
full = beamforming.mvdr(src, mix)
lo, hi = shard_range(T, rank, world)
out, w = beamforming.mvdr_utterance(src[:, :, :, lo:hi].contiguous(), mix[:, :, lo:hi].contiguous())
err = float((out - full[:, :, lo:hi]).abs().max() / full.abs().max())
err = D.max_over_ranks(err, "cuda")
if rank == 0:
    print(json.dumps({"check": "utterance-level MVDR, frames sharded over ranks, vs one rank", "n_gpus": world, "frames": T, "bins": F,
                      "max_abs_err_over_peak": err, "ok": err < 1e-5}))
