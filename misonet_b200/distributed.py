"""Multi-GPU plumbing: one process per GPU, utterances sharded across ranks, and the only
data-path collectives the forward path needs: a SUM all-reduce of
[loss_sum, count] and an all-gather of the int64 permutation indices.  The reference has no
distributed code at all (single ``model.cuda(gpu_num)``, run.py:68); this is torch.distributed
(NCCL over NVLink on the GPU box, gloo in the CPU tests) used as plumbing."""
import os

import torch
import torch.distributed as dist

from .pipeline import shard_range  # noqa: F401  (re-export)


def init_from_env(backend=None):
    """Reads RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (torchrun).  Returns (rank, world, local_rank).
    With WORLD_SIZE unset or 1 nothing is initialised."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def _on():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def reduce_metrics(loss_sum, count, device="cpu"):
    """Global mean of a per-utterance metric: all-reduce(SUM) of [loss_sum, count] in fp64."""
    v = torch.tensor([float(loss_sum), float(count)], dtype=torch.float64, device=device)
    if _on():
        dist.all_reduce(v, op=dist.ReduceOp.SUM)
    return (v[0] / v[1].clamp(min=1)).item(), int(v[1].item())


def max_over_ranks(value, device="cpu"):
    v = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if _on():
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
    return v.item()


def gather_perm_indices(idx, n_total):
    """idx: int64 [n_local] decisions of this rank's shard (block partition by shard_range)
    -> int64 [n_total] on every rank, in global utterance order."""
    if not _on():
        return idx.clone()
    world = dist.get_world_size()
    cap = (n_total + world - 1) // world
    buf = torch.full((cap,), -1, dtype=torch.long, device=idx.device)
    buf[: idx.numel()] = idx
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    out = []
    for r, p in enumerate(parts):
        lo, hi = shard_range(n_total, r, world)
        out.append(p[: hi - lo])
    return torch.cat(out)


def barrier():
    if _on():
        dist.barrier()


def allreduce_gradients(params, n_local=None):
    """Data-parallel gradient reduction of a training step (one all-reduce of the 2.59 M /
    8.58 M fp32 gradients per step).  Every rank holds the gradient of the MEAN loss over its own ``n_local``
    utterances (criterion.py:59 averages over the batch); the result on every rank is the gradient of the mean over
    all utterances of the step: sum_r n_r * g_r / sum_r n_r (a plain average for equal shards).  The gradients are
    packed into ONE flat bucket (the backward pass produces them as one flat buffer anyway, so this is a single
    latency-bound collective on NVLink / NVSwitch) and written back in place.  Returns the global utterance count."""
    params = [p for p in params if p.grad is not None]
    if not _on():
        return n_local
    flat = torch.cat([p.grad.reshape(-1) for p in params])
    w = 1.0 if n_local is None else float(n_local)
    cnt = torch.tensor([w], dtype=torch.float32, device=flat.device)
    if n_local is not None:
        flat.mul_(w)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    total = float(cnt.item()) if n_local is not None else float(dist.get_world_size())
    flat.div_(total)
    off = 0
    for p in params:
        n = p.grad.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n
    return int(total) if n_local is not None else None
