"""Data-parallel plumbing for the one thing that shards on this path: whole scenes.
The sparse conv itself never communicates; the only collective is the gradient mean after backward
(reference: DistributedDataParallel at tool/train.py:361, DistributedSampler at dataset/__init__.py:62-69)."""
import torch
import torch.distributed as dist


def scene_ids(rank, world, scenes_per_rank):
    """Seeds of the synthetic scenes rank `rank` owns: 1000*rank + i; disjoint across ranks."""
    return [1000 * rank + i for i in range(scenes_per_rank)]


def _avg_all_reduce(t, world):
    if dist.get_backend() == "nccl":
        dist.all_reduce(t, op=dist.ReduceOp.AVG)
    else:  # gloo has no AVG
        dist.all_reduce(t)
        t.div_(world)


def allreduce_grads(params, world):
    """Mean of the gradients over ranks, after backward, in as few collectives as there are gradient STORAGES.

    The engine hands out every conv weight gradient of a step as a slice of one zero-filled arena block
    (`ops._ZeroArena`), in the same order on every rank, so all conv gradients (30 MB at m=16) are reduced IN PLACE
    with one NCCL call over the span they cover -- no flatten / copy-back and no per-parameter hook on the host
    (DistributedDataParallel's hooks cost ~2 ms per step here, tool/train.py:361).  The remaining small gradients
    (BN, linear) go through one flat buffer."""
    if world <= 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    spans, rest = {}, []
    for g in grads:
        st = g.untyped_storage()
        if g.is_contiguous() and st.nbytes() >= (1 << 20) and g.numel() * g.element_size() < st.nbytes():
            lo = g.storage_offset()
            e = spans.setdefault((st.data_ptr(), g.dtype), [g, lo, lo + g.numel()])
            e[1], e[2] = min(e[1], lo), max(e[2], lo + g.numel())
        else:
            rest.append(g)
    for (_, dtype), (g, lo, hi) in spans.items():  # insertion order = backward order: identical on every rank
        flat = torch.empty(0, dtype=dtype, device=g.device).set_(g.untyped_storage(), lo, (hi - lo,))
        _avg_all_reduce(flat, world)
    if rest:
        flat = torch.cat([g.reshape(-1) for g in rest])
        _avg_all_reduce(flat, world)
        torch._foreach_copy_(rest, [v.view_as(g) for v, g in zip(flat.split([g.numel() for g in rest]), rest)])


def max_over_ranks(value, device):
    """device-timed durations are reported as the max over ranks"""
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
