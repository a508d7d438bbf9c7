"""2-GPU check of the data-parallel training step: rank r back-propagates its own shard with model.data_parallel = True;
the all-reduced gradients must equal the utterance-weighted mean of the per-shard gradients (which every rank also
computes locally, shard by shard, with data_parallel off).  Shards are deliberately unequal (2 and 1 utterances).
usage: torchrun --nproc-per-node 2 tools/ddp_check.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from misonet_b200 import criterion, distributed, synth  # noqa: E402
from misonet_b200.model import MISO_1  # noqa: E402

rank, world, local = distributed.init_from_env()
assert world == 2, "run under torchrun with 2 ranks"
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
en, de = [24, 32, 32, 32, 32, 64, 128], [128, 64, 32, 32, 32, 32, 24]
torch.manual_seed(0)
model = MISO_1(2, 6, 7, list(en), list(de), "IN").to(dev).train()
model.conv_mode = "bf16x3"
T, F = 32, 129
shards = [(0, 2), (2, 3)]


def grads_of(lo, hi, ddp):
    model.zero_grad(set_to_none=True)
    model.data_parallel = ddp
    mix = torch.from_numpy(synth.random_spec(7, (3, 6, T, F)))[lo:hi].to(dev)
    refs = torch.from_numpy(synth.random_spec(8, (3, 2, T, F)))[lo:hi].to(dev)
    loss = criterion.loss_uPIT(2, model(mix), [refs[:, 0], refs[:, 1]])
    loss.backward()
    return torch.cat([p.grad.flatten() for p in model.parameters()]).clone()


parts = [grads_of(lo, hi, False) for lo, hi in shards]
expected = (2.0 * parts[0] + 1.0 * parts[1]) / 3.0
got = grads_of(*shards[rank], True)
err = float((got - expected).norm() / expected.norm())
other = [torch.empty_like(got) for _ in range(2)]
torch.distributed.all_gather(other, got)
same = float((other[0] - other[1]).abs().max())
print(f"rank {rank}: all-reduced gradient vs weighted mean of shard gradients: rel err {err:.2e}; max |rank0 - rank1| = {same:.1e}")
assert err < 1e-4 and same == 0.0
distributed.barrier()
if rank == 0:
    print("DDP_OK")
torch.distributed.destroy_process_group()
