"""Data-parallel plumbing for the one thing that shards on this path: whole scenes.
The sparse conv itself never communicates; the only collective is the gradient mean after backward
(reference: DistributedDataParallel at tool/train.py:361, DistributedSampler at dataset/__init__.py:62-69)."""
import torch
import torch.distributed as dist


def scene_ids(rank, world, scenes_per_rank):
    """Seeds of the synthetic scenes rank `rank` owns: 1000*rank + i; disjoint across ranks."""
    return [1000 * rank + i for i in range(scenes_per_rank)]


def _avg_all_reduce(t, world, async_op=False):
    """in-place mean over ranks; async_op -> returns a callable that makes the current stream wait for it"""
    if dist.get_backend() == "nccl":
        work = dist.all_reduce(t, op=dist.ReduceOp.AVG, async_op=async_op)
        return work.wait if async_op else None
    work = dist.all_reduce(t, async_op=async_op)  # gloo has no AVG
    if not async_op:
        t.div_(world)
        return None

    def done(work=work, t=t):
        work.wait()
        t.div_(world)
    return done


def allreduce_grads(params, world, async_op=False):
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
            e = spans.setdefault((st.data_ptr(), g.dtype), [g, lo, lo + g.numel(), 0, []])
            e[1], e[2] = min(e[1], lo), max(e[2], lo + g.numel())
            e[3] += (g.numel() + 63) // 64 * 64  # the arena hands out 64-float aligned slices
            e[4].append(g)
        else:
            rest.append(g)
    # a span is reduced in place only if these gradients (almost) fill it: anything else living between them in the
    # arena block (another model's gradients, a stale slice) must not be averaged along
    for key in [k for k, e in spans.items() if e[3] < 0.98 * (e[2] - e[1])]:
        rest.extend(spans.pop(key)[4])
    # every rank must bring the same span lengths (same sequence of arena slices since the block was taken): checked
    # with one tiny all-gather whenever the signature changes, instead of hanging or corrupting inside NCCL
    sig = tuple(e[2] - e[1] for e in spans.values()) + (len(rest), sum(g.numel() for g in rest))
    if sig not in _checked_signatures:
        mine = torch.tensor(list(sig) + [0] * (32 - len(sig)) if len(sig) <= 32 else [hash(sig) % (1 << 40)] * 32,
                            dtype=torch.int64, device=grads[0].device)
        every = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(every, mine)
        if any(not torch.equal(t, mine) for t in every):
            raise RuntimeError("allreduce_grads: ranks disagree on the gradient layout %s -- a rank-asymmetric backward "
                               "(skipped step, different need_dw); use torch DDP or reduce per parameter" % (sig,))
        _checked_signatures.add(sig)
    waits = []
    for (_, dtype), (g, lo, hi, _, _) in spans.items():  # insertion order = backward order: identical on every rank
        flat = torch.empty(0, dtype=dtype, device=g.device).set_(g.untyped_storage(), lo, (hi - lo,))
        w = _avg_all_reduce(flat, world, async_op)
        if w is not None:
            waits.append(w)
    if rest:
        flat = torch.cat([g.reshape(-1) for g in rest])
        w = _avg_all_reduce(flat, world, async_op)

        def copy_back(w=w, flat=flat, rest=rest):
            if w is not None:
                w()
            torch._foreach_copy_(rest, [v.view_as(g) for v, g in zip(flat.split([g.numel() for g in rest]), rest)])
        if async_op:
            waits.append(copy_back)
        else:
            copy_back()
    return waits if async_op else None


_checked_signatures = set()


class OverlappedGradReducer(object):
    """Gradient mean over ranks that STARTS during backward (the role of DistributedDataParallel's bucket hooks,
    tool/train.py:361, without its per-parameter host cost).

    DODA's U-Net holds 97 % of its parameters below level 1 (`unet.u`), and backward leaves that sub-network with the
    level-1 down conv, two residual blocks and the input conv (~2 ms of GPU work) still to run.  `attach(module)`
    arranges that, the moment the gradient of `module`'s input features is produced -- i.e. its whole backward is
    done -- every gradient accumulated so far is reduced with an ASYNC NCCL call (one in-place call over the dW arena
    span they fill, one flat call for the small BN / linear ones) on the process group's own stream, overlapping the
    rest of backward.  `finish()` after `loss.backward()` reduces what came later and makes the current stream wait.

        reducer = OverlappedGradReducer(params, world); reducer.attach(model.unet.u)
        loss.backward(); reducer.finish()
    """

    def __init__(self, params, world):
        self.params, self.world = list(params), world
        self._done = set()
        self._work = []

    def attach(self, module):
        def pre(mod, inputs):
            x = inputs[0]
            feats = getattr(x, "features", x)
            if torch.is_tensor(feats) and feats.requires_grad and self.world > 1:
                feats.register_hook(self._on_grad)
        return module.register_forward_pre_hook(pre)

    def _on_grad(self, grad):
        self._reduce([p for p in self.params if p.grad is not None and id(p) not in self._done], async_op=True)
        return None

    def _reduce(self, ps, async_op):
        if not ps or self.world <= 1:
            return
        for p in ps:
            self._done.add(id(p))
        self._work.extend(allreduce_grads(ps, self.world, async_op=async_op) or [])

    def finish(self):
        self._reduce([p for p in self.params if p.grad is not None and id(p) not in self._done], async_op=True)
        for w in self._work:
            w()  # current stream waits for the collective (and the small gradients are copied back)
        self._work, self._done = [], set()


def broadcast_parameters(module, src=0):
    """what DistributedDataParallel does at construction (tool/train.py:361): every rank starts from rank `src`'s
    parameters and buffers.  Call once before training when `allreduce_grads` replaces DDP."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() <= 1:
        return
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t, src)


def max_over_ranks(value, device):
    """device-timed durations are reported as the max over ranks"""
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
