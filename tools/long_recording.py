"""BASELINE configs[4]-shaped run: one long 6-channel recording (16 chunks of ~500 frames, 257 bins) through the chunked
MISO-BF-MISO path (continuous.separate_recording), chunks block-partitioned over the ranks (torchrun), waveforms gathered.
Prints one JSON line on rank 0 (device-timed, max over ranks) and compares the N-rank result with the 1-rank result
(bit for bit when the per-call batches have the same composition; to rounding otherwise -- the tensor-core path's
statistics partial sums depend on the tiling, hence on the batch size)."""
import json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from misonet_b200 import continuous, distributed as D
from misonet_b200.model import MISO_1, MISO_3
from misonet_b200.pipeline import MisoBfMiso

rank, world, local = D.init_from_env()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
en, de = bench.LAYOUTS["PAPER"]
m1 = MISO_1(2, 6, len(en), list(en), list(de), "IN"); m1.load_state_dict(bench.make_state_dict_np(m1, 0))
m3 = MISO_3(1, 6, len(en), list(en), list(de), "IN"); m3.load_state_dict(bench.make_state_dict_np(m3, 1))
m1 = m1.to(dev).eval(); m3 = m3.to(dev).eval()
m1.conv_mode = m3.conv_mode = "bf16x3"
pipe = MisoBfMiso(m1, m3, nperseg=512, noverlap=384)
chunk, n_chunks = 64000, 16                      # 4 s at 16 kHz: 501 frames of 257 bins per chunk
rng = np.random.default_rng(77)
wav = torch.from_numpy((0.05 * rng.standard_normal((chunk * n_chunks - 12345, 6))).astype(np.float32)).to(dev)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
out = continuous.separate_recording(pipe, wav, chunk, rank, world, batch=4)      # warm-up (graph capture)
torch.cuda.synchronize(); D.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    out = continuous.separate_recording(pipe, wav, chunk, rank, world, batch=4)
e1.record()
torch.cuda.synchronize(); D.barrier()
ms = D.max_over_ranks(e0.elapsed_time(e1) / steps, dev)
same, diff = None, None
if world > 1:   # every rank also runs the whole recording alone
    alone = continuous.separate_recording(pipe, wav, chunk, 0, 1, batch=4)
    same = bool(torch.equal(alone, out))
    diff = float((alone - out).norm() / alone.norm())
if rank == 0:
    frames = n_chunks * 501
    print(json.dumps({"workload": "long recording, 6 ch x 257 bins x %d frames (16 chunks), chunk pipeline MISO1x6 -> MVDRx2 -> MISO3x2 -> ISTFT" % frames,
                      "n_gpus": world, "ms_per_recording": ms, "frames_per_s": frames / ms * 1e3, "samples": int(wav.shape[0]),
                      "output_shape": list(out.shape), "bit_identical_to_one_rank": same, "rel_diff_to_one_rank": diff,
                      "collective": "one all-gather of the chunk waveforms (continuous.gather_chunks); no data-path collective before it",
                      "timing": "CUDA events around the timed region on every rank, max over ranks", "conv_mode": "bf16x3"}))
