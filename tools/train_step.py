"""Training-step timing (BASELINE configs[3]: MISO1 training step with the uPIT loss, utterances sharded over the GPUs).
One step = forward (training plan) + loss_uPIT + backward + gradient all-reduce + clip + Adam, on synthetic spectrograms
resident in HBM.  usage: [torchrun ...] python tools/train_step.py [--layout PAPER] [--batch 8] [--steps 3] [--warmup 2]
Prints one JSON line (rank 0): frames/s of the whole job (max over ranks), and the phase split of rank 0."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from misonet_b200 import criterion, distributed, synth  # noqa: E402
from misonet_b200.model import MISO_1  # noqa: E402

LAYOUTS = {"REF": ([24, 32, 32, 32, 32, 64, 128], [128, 64, 32, 32, 32, 32, 24], 129, 501),
           "PAPER": ([24, 32, 32, 32, 32, 64, 128, 384], [384, 128, 64, 32, 32, 32, 32, 24], 257, 500)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layout", default="PAPER", choices=sorted(LAYOUTS))
    ap.add_argument("--batch", type=int, default=8, help="utterances per GPU")
    ap.add_argument("--frames", type=int, default=0)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--conv-mode", default="bf16x3")
    args = ap.parse_args()
    rank, world, local = distributed.init_from_env()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    en, de, F, T = LAYOUTS[args.layout]
    T = args.frames or T
    B = args.batch
    torch.manual_seed(0)
    model = MISO_1(2, 6, len(en), list(en), list(de), "IN").to(dev).train()
    model.conv_mode = args.conv_mode
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    mix = torch.from_numpy(synth.random_spec(100 + rank, (B, 6, T, F))).to(dev)
    refs = [torch.from_numpy(synth.random_spec(200 + 10 * rank + s, (B, T, F))).to(dev) for s in range(2)]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    phases = [0.0, 0.0, 0.0, 0.0]
    total_ms = 0.0
    loss = None
    for it in range(args.warmup + args.steps):
        if it == args.warmup:
            distributed.barrier()
            torch.cuda.synchronize()
        opt.zero_grad(set_to_none=True)
        ev[0].record()
        est = model(mix)
        loss = criterion.loss_uPIT(2, est, refs)
        ev[1].record()
        loss.backward()
        ev[2].record()
        distributed.allreduce_gradients(model.parameters(), n_local=B)
        ev[3].record()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 5.0)
        opt.step()
        ev[4].record()
        torch.cuda.synchronize()
        if it >= args.warmup:
            for k in range(4):
                phases[k] += ev[k].elapsed_time(ev[k + 1])
            total_ms += ev[0].elapsed_time(ev[4])
    distributed.barrier()
    ms = distributed.max_over_ranks(total_ms / args.steps, device=dev)
    if rank == 0:
        print(json.dumps({
            "metric": "frames/sec MISO1 training step (fwd + loss_uPIT + bwd + grad all-reduce + Adam)", "value": world * B * T / (ms * 1e-3),
            "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "config": {"workload": f"MISO1 training step, {B} utterances per GPU x 6ch x {F}bin x {T}fr (BASELINE configs[3] shape)",
                       "layout": args.layout, "conv_mode": args.conv_mode, "global_batch": world * B},
            "phases_ms_rank0": {"forward+loss": phases[0] / args.steps, "backward": phases[1] / args.steps,
                                "grad_allreduce": phases[2] / args.steps, "clip+adam": phases[3] / args.steps},
            "loss": float(loss), "train_workspace_gb": model._ws_train.numel() / 2 ** 30,
            "data": "synthetic (seeded random spectrograms, default PyTorch initialisation)"}))


if __name__ == "__main__":
    main()
