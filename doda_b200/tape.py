"""One autograd node for a whole U-Net sub-tree (`UBlock`: model/unet_block.py:42-100).

Why: one training step of DODA's U-Net is ~270 autograd nodes, and the host pays ~25 us of Python / autograd-engine
work per node and direction -- 11.5 ms per step whatever the scene size, while the GPU needs 11-12 ms at 2 x 150 k
voxels (the step is host-bound, and N ranks share the box's host cores).  The deep levels, where a
kernel is 8-30 us, are where the host falls behind.  This module runs a `UBlock` -- its residual / VGG blocks, the
BN-ReLU-down conv, the child level, the BN-ReLU-inverse conv, the concat and the tail blocks -- as ONE
`torch.autograd.Function`: the forward calls the same per-layer C entry points (`ops._layer_fwd`) in a plain Python
recursion with no autograd bookkeeping and returns a closure that replays the chain backwards (`ops._layer_bwd`), so
the per-layer host cost drops to the C call plus a few allocations.  Same kernels, same arithmetic, same order of
operations as the module-by-module path (`tests/test_parity_gpu.py::test_taped_ublock_*`).

It applies to a sub-tree whose modules are exactly the shapes DODA builds (duck-typed on the attribute names of
`model/unet_block.py`: `blocks`, `conv`, `u`, `deconv`, `blocks_tail`, `conv_branch`, `i_branch`, `conv_layers`), in
training mode, fp32, on CUDA; anything else falls back to the modules' own forward.  Retained graphs / double backward
are not supported through a taped sub-tree (the closure frees its activations after one backward).
"""
import torch
from torch import nn
from torch.autograd import Function

from . import ops as _ops
from .spconv.modules import SparseSequential, is_sparse_conv, _is_bn_like

enabled = True          # module-level switch (A/B tests); `B200SP_TAPE=0` in the environment turns it off too
runs = 0                # taped sub-tree executions so far (tests assert the path was taken)
capture = None          # (masks dict, {id(bn module): key}) while oracle/gates.py records ReLU gates, else None

import os as _os
if _os.environ.get("B200SP_TAPE", "1") == "0":
    enabled = False


def _triplet_of(seq):
    """[BN-like, ReLU, sparse conv] -> (bn, conv) or None"""
    mods = list(seq._modules.values())
    if len(mods) == 3 and _is_bn_like(mods[0]) and isinstance(mods[1], nn.ReLU) and is_sparse_conv(mods[2]):
        return mods[0], mods[2]
    return None


def _block_plan(blk):
    """-> ("res", [(bn, conv), (bn, conv)], skip conv or None) | ("vgg", [(bn, conv)], None) | None"""
    cb = getattr(blk, "conv_branch", None)
    ib = getattr(blk, "i_branch", None)
    if isinstance(cb, SparseSequential) and isinstance(ib, SparseSequential):
        mods = list(cb._modules.values())
        if len(mods) != 6:
            return None
        trips = []
        for j in (0, 3):
            if not (_is_bn_like(mods[j]) and isinstance(mods[j + 1], nn.ReLU) and is_sparse_conv(mods[j + 2])):
                return None
            trips.append((mods[j], mods[j + 2]))
        skip = list(ib._modules.values())
        if len(skip) != 1:
            return None
        if isinstance(skip[0], nn.Identity):
            return "res", trips, None
        if is_sparse_conv(skip[0]) and skip[0].conv1x1 and skip[0].bias is None:
            return "res", trips, skip[0]
        return None
    cl = getattr(blk, "conv_layers", None)
    if isinstance(cl, SparseSequential):
        t = _triplet_of(cl)
        return ("vgg", [t], None) if t is not None else None
    return None


def plan(ub):
    """the executable description of a UBlock sub-tree, or None when it is not the shape this module knows"""
    blocks = getattr(ub, "blocks", None)
    if not isinstance(blocks, SparseSequential):
        return None
    p = {"blocks": [], "tail": None}
    for blk in blocks._modules.values():
        bp = _block_plan(blk)
        if bp is None:
            return None
        p["blocks"].append(bp)
    child = getattr(ub, "u", None)
    if child is not None:
        conv, deconv, tail = getattr(ub, "conv", None), getattr(ub, "deconv", None), getattr(ub, "blocks_tail", None)
        if not (isinstance(conv, SparseSequential) and isinstance(deconv, SparseSequential) and isinstance(tail, SparseSequential)):
            return None
        p["down"], p["up"] = _triplet_of(conv), _triplet_of(deconv)
        if p["down"] is None or p["up"] is None or p["down"][1].bias is not None or p["up"][1].bias is not None:
            return None
        if not p["up"][1].inverse:
            return None
        p["child"] = plan(child)
        if p["child"] is None:
            return None
        p["tail"] = []
        for blk in tail._modules.values():
            bp = _block_plan(blk)
            if bp is None:
                return None
            p["tail"].append(bp)
    for kind, trips, skip in p["blocks"] + (p["tail"] or []):
        for bn, conv in trips:
            if conv.bias is not None:
                return None
    return p


def _plan_modules(p):
    for kind, trips, skip in p["blocks"] + (p["tail"] or []):
        for bn, conv in trips:
            yield bn
            yield conv
        if skip is not None:
            yield skip
    if p["tail"] is not None:
        for bn, conv in (p["down"], p["up"]):
            yield bn
            yield conv
        for m in _plan_modules(p["child"]):
            yield m


def cached_plan(ub):
    """plan(ub), cached on the module and re-derived when any module of the sub-tree has been replaced since
    (convert_dsnorm swaps the BatchNorm modules of a built model, model/dsnorm.py:90-110)"""
    c = ub.__dict__.get("_b200sp_tape_plan")
    if c is not None:
        p, links = c
        ok = True
        for parent, name, mod in links:
            if parent._modules.get(name) is not mod:
                ok = False
                break
        if ok:
            return p
    p = plan(ub)
    links = []
    for parent in ub.modules():
        for name, mod in parent._modules.items():
            links.append((parent, name, mod))
    if p is not None:
        mods = list(_plan_modules(p))
        p["_bns"] = [m for m in mods if not is_sparse_conv(m)]
        p["_convs"] = [m for m in mods if is_sparse_conv(m)]
        params, seen = [], set()
        for m in mods:
            for prm in m.parameters(recurse=False):
                if id(prm) not in seen:
                    seen.add(id(prm))
                    params.append(prm)
        p["_params"] = params
    ub.__dict__["_b200sp_tape_plan"] = (p, links)
    return p


def usable(ub, p, x):
    """taped execution only where the per-layer executor would take every layer: training-mode batch statistics, fp32
    CUDA features, contiguous weights, no profiling pass"""
    if not (enabled and _ops.layer_exec and _ops._prof is None and p is not None and torch.is_grad_enabled()):
        return False
    f = x.features
    if not (f.is_cuda and f.dtype == torch.float32 and f.dim() == 2 and x.indices.shape[0] > 0):
        return False
    for m in p["_bns"]:
        if not m.training:
            return False
    for m in p["_convs"]:
        if not m.weight.is_contiguous():
            return False
    return True


# ---------------------------------------------------------------------------------------------------
# forward primitives: each returns (output features, backward closure).  A closure takes the gradient of its output
# (contiguous fp32) and the dict {id(parameter): gradient} it adds its parameter gradients to; it returns the gradient
# of its input and drops its saved activations.
# ---------------------------------------------------------------------------------------------------
def _triplet(bn, conv, x, t, res=None):
    """res: [rows, Cout] added to the conv output in its epilogue; the closure's `add`: added to the returned input
    gradient inside the BatchNorm backward (both replace a separate elementwise kernel, same fp32 adds)"""
    kind, rb, outids, oshape = conv._resolve(t)
    rm, rv, nbt, momentum = _ops.bn_batch_stats_args(bn)
    prep = _ops.prepared_weights(conv)
    W = conv.weight
    out, y, stats = _ops._layer_fwd(kind, x, W, rb, prep, (bn.weight, bn.bias, rm, rv, nbt, momentum, bn.eps), res=res)
    if capture is not None:
        key = capture[1].get(id(bn))
        if key is not None:
            capture[0][key] = (y > 0).cpu()
    t.indices, t.spatial_shape = outids, oshape
    saved = [x, y, stats]

    def bw(g, grads, add=None):
        x_, y_, stats_ = saved
        saved[:] = (None, None, None)
        dx, dW, dwb = _ops._layer_bwd(kind, x_, y_, W, g, rb, prep, True, W.requires_grad, bn=(bn.weight, bn.bias), stats=stats_,
                                      dx_add=add)
        if dW is not None:
            grads[id(W)] = dW
        if bn.weight is not None and bn.weight.requires_grad:
            grads[id(bn.weight)] = dwb[0]
        if bn.bias is not None and bn.bias.requires_grad:
            grads[id(bn.bias)] = dwb[1]
        return dx

    return out, bw


def _plain_conv(conv, x, t):
    kind, rb, outids, oshape = conv._resolve(t)
    prep = _ops.prepared_weights(conv)
    W = conv.weight
    out = _ops._layer_fwd(kind, x, W, rb, prep)[0]
    saved = [x]

    def bw(g, grads):
        x_ = saved[0]
        saved[0] = None
        din, dW, _ = _ops._layer_bwd(kind, None, x_, W, g, rb, prep, True, W.requires_grad)
        if dW is not None:
            grads[id(W)] = dW
        return din

    return out, bw


def _block(bp, x, t):
    kind, trips, skip = bp
    if kind == "vgg":
        return _triplet(trips[0][0], trips[0][1], x, t)
    h1, bw1 = _triplet(trips[0][0], trips[0][1], x, t)
    if skip is None:
        # out.features += identity.features (model/unet_block.py:37): folded into the second conv's epilogue, and the
        # identity's gradient into the first BatchNorm's backward
        h2, bw2 = _triplet(trips[1][0], trips[1][1], h1, t, res=x)

        def bw(g, grads):
            return bw1(bw2(g, grads), grads, add=g)
    else:
        s, bws = _plain_conv(skip, x, t)
        h2, bw2 = _triplet(trips[1][0], trips[1][1], h1, t, res=s)

        def bw(g, grads):
            ds = bws(g, grads)
            return bw1(bw2(g, grads), grads, add=ds)
    return h2, bw


def _ublock(p, x, t):
    bws = []
    h = x
    for bp in p["blocks"]:
        h, b = _block(bp, h, t)
        bws.append(b)
    if p["tail"] is None:
        def bw(g, grads):
            for b in reversed(bws):
                g = b(g, grads)
            return g
        return h, bw
    c0 = h.shape[1]
    fine_idx, fine_shape = t.indices, t.spatial_shape
    d, bw_down = _triplet(p["down"][0], p["down"][1], h, t)
    u, bw_child = _ublock(p["child"], d, t)
    up, bw_up = _triplet(p["up"][0], p["up"][1], u, t)
    assert t.indices.shape[0] == fine_idx.shape[0] and list(t.spatial_shape) == list(fine_shape)
    h = torch.cat((h, up), dim=1)  # model/unet_block.py:95
    tails = []
    for bp in p["tail"]:
        h, b = _block(bp, h, t)
        tails.append(b)

    def bw(g, grads):
        for b in reversed(tails):
            g = b(g, grads)
        g_skip = g[:, :c0].contiguous()
        g_up = g[:, c0:].contiguous()
        g = bw_down(bw_child(bw_up(g_up, grads), grads), grads, add=g_skip)  # + the skip connection's gradient
        for b in reversed(bws):
            g = b(g, grads)
        return g

    return h, bw


class UBlockTapeFunction(Function):
    """forward(features, plan, metadata tensor, *parameters) -> features of the sub-tree's output"""

    @staticmethod
    def forward(ctx, feats, p, t, *params):
        feats = _ops._f32c(feats)
        out, bw = _ublock(p, feats, t)
        ctx.bw = bw
        ctx.params = params
        return out

    @staticmethod
    def backward(ctx, g):
        bw, ctx.bw = ctx.bw, None
        if bw is None:
            raise RuntimeError("doda_b200.tape: a taped U-Net sub-tree supports ONE backward pass (its activations are "
                               "released as the gradient passes; retain_graph / double backward need B200SP_TAPE=0)")
        grads = {}
        try:
            dx = bw(_ops._f32c(g), grads)
        except BaseException:
            _ops._drop_pending_join()
            raise
        need = ctx.needs_input_grad
        return (dx if need[0] else None, None, None) + tuple(
            grads.get(id(prm)) if need[3 + i] else None for i, prm in enumerate(ctx.params))


def run(ub, p, x):
    """x: SparseConvTensor entering the UBlock `ub` (plan p) -> SparseConvTensor leaving it"""
    from .spconv import SparseConvTensor
    global runs
    runs += 1
    t = SparseConvTensor(None, x.indices, x.spatial_shape, x.batch_size)
    t.indice_dict = x.indice_dict
    t.grid = x.grid
    out = UBlockTapeFunction.apply(x.features, p, t, *p["_params"])
    res = SparseConvTensor(out, t.indices, t.spatial_shape, x.batch_size)
    res.indice_dict = x.indice_dict
    res.grid = x.grid
    return res


def attach(model):
    """Opt a model built from the REFERENCE's own classes (model/unet_block.py UBlock / ResidualBlock / VGGBlock on the
    engine's spconv surface) into taped execution: every sub-module whose structure `plan` recognises gets an
    instance-level `forward` that runs the sub-tree as one autograd node when it qualifies and the class's own forward
    otherwise.  The engine's mirror (doda_b200/unet.py) does this by itself.  Returns the number of sub-trees wrapped
    (nested ones included: an outer taped run never reaches the inner wrappers)."""
    n = 0
    for m in model.modules():
        if "forward" in m.__dict__ or plan(m) is None:
            continue
        cls_forward = m.forward

        def fwd(x, _m=m, _orig=cls_forward):
            if enabled and _m.training and getattr(_m, "tape", True):
                p = cached_plan(_m)
                if usable(_m, p, x):
                    return run(_m, p, x)
            return _orig(x)

        m.forward = fwd
        n += 1
    return n
