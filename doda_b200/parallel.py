"""Data-parallel plumbing for the one thing that shards on this path: whole scenes.
The sparse conv itself never communicates; the only collective is the gradient mean after backward
(reference: DistributedDataParallel at tool/train.py:361, DistributedSampler at dataset/__init__.py:62-69)."""
import torch
import torch.distributed as dist


def scene_ids(rank, world, scenes_per_rank):
    """Seeds of the synthetic scenes rank `rank` owns: 1000*rank + i; disjoint across ranks."""
    return [1000 * rank + i for i in range(scenes_per_rank)]


def allreduce_grads(params, world):
    """Mean of the gradients over ranks through ONE flat buffer (one NCCL launch instead of one per tensor)."""
    if world <= 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat)
    flat.div_(world)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


def max_over_ranks(value, device):
    """device-timed durations are reported as the max over ranks"""
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
