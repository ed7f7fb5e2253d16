"""Metric epilogue of the training loop on the device.

`intersectionAndUnionGPU` keeps the name, arguments and return value of the reference's
`util/common_utils.py:233-247`, which clones both tensors, masks, and runs three `torch.histc` on CPU copies
(three device->host syncs per iteration, `tool/train.py:113-118`).  Here: one pass on the device
(`b200sp_intersection_union`, `csrc/loss.cu`), float counts like histc returns, no host sync -- the caller decides
when to read them (`update_meter` all-reduces them first under DDP)."""
import torch

from . import ops as _ops
from ._lib import lib, check


def intersectionAndUnionGPU(output, target, K, ignore_index=255):
    """-> (area_intersection, area_union, area_target), float32 CUDA tensors of K per-class counts.
    `output` (predictions) and `target` (labels): integer tensors of the same shape, any of N / N x L / N x H x W."""
    assert output.dim() in (1, 2, 3)
    assert output.shape == target.shape
    if not (output.is_cuda and target.is_cuda):
        raise RuntimeError("doda_b200.metrics needs CUDA tensors (no CPU fallback)")
    o = output.reshape(-1)
    t = target.reshape(-1)
    o = o.contiguous() if o.dtype == torch.int64 else o.long()
    t = t.contiguous() if t.dtype == torch.int64 else t.long()
    K = int(K)
    out = torch.empty((3, K), dtype=torch.float32, device=o.device)
    ws = _ops._workspace(12 * K, o.device, "iou")
    check(lib.b200sp_intersection_union(o.data_ptr(), t.data_ptr(), o.numel(), K, int(ignore_index), out.data_ptr(),
                                        ws.data_ptr(), ws.numel(), _ops._stream()), "intersection_union")
    return out[0], out[1], out[2]


def intersection_and_union_ref(output, target, K, ignore_index=255):
    """the reference formula on CPU tensors (util/common_utils.py:233-247), for tests: NOT used by the product path"""
    output = output.reshape(-1).clone().cpu()
    target = target.reshape(-1).clone().cpu()
    output[target == ignore_index] = ignore_index
    intersection = output[output == target]
    ai = torch.histc(intersection.float(), bins=K, min=0, max=K - 1)
    ao = torch.histc(output.float(), bins=K, min=0, max=K - 1)
    at = torch.histc(target.float(), bins=K, min=0, max=K - 1)
    return ai, ao + at - ai, at


def packed_epilogue(loss, preds, labels, n_classes, ignore_label=255, dist_train=False):
    """The per-iteration metric epilogue of the reference's training loop (tool/train.py:107-118 +
    util/common_utils.py:250-256) in ONE device pass, ONE collective and ONE device->host read.

    The reference all-reduces the loss, the point count and the three per-class histograms separately (5 tiny
    collectives) and reads them back with `.item()` / three `.cpu()` calls (6+ host syncs per iteration).  Here the
    histograms come from `b200sp_intersection_union`, everything is packed into one float64 vector
    [loss * n, n, intersection[K], union[K], target[K]], summed over ranks by one all-reduce, and copied to the host
    once.  float64 keeps the counts exact (the reference's float32 `histc` counts are exact below 2^24 as well).
    -> (mean loss over all ranks' points, n_total, intersection, union, target) with numpy float32 histograms like the
    reference's `update_meter` returns."""
    K = int(n_classes)
    inter, union, target = intersectionAndUnionGPU(preds, labels, K, ignore_label)
    n = preds.shape[0]
    pack = torch.empty(2 + 3 * K, dtype=torch.float64, device=preds.device)
    pack[0] = loss.detach().double() * n
    pack[1] = n
    pack[2:2 + K] = inter
    pack[2 + K:2 + 2 * K] = union
    pack[2 + 2 * K:] = target
    if dist_train:
        import torch.distributed as dist
        dist.all_reduce(pack)
    host = pack.cpu().numpy()  # the one host sync of the iteration
    n_tot = int(round(host[1]))
    f32 = host[2:].astype("float32")
    return float(host[0] / max(n_tot, 1)), n_tot, f32[:K], f32[K:2 * K], f32[2 * K:]


def update_meter(intersection_meter, union_meter, target_meter, preds, labels, n_classes, ignore_label, dist_train):
    """Drop-in for the reference's `update_meter` (util/common_utils.py:250-256): same arguments, same return tuple
    (meters, running accuracy, this iteration's intersection / union / target as numpy arrays); the three histograms
    cross the ranks in one packed all-reduce and reach the host in one copy."""
    K = int(n_classes)
    inter, union, target = intersectionAndUnionGPU(preds, labels, K, ignore_label)
    pack = torch.cat((inter, union, target))
    if dist_train:
        import torch.distributed as dist
        dist.all_reduce(pack)
    host = pack.cpu().numpy()
    intersection, union, target = host[:K].copy(), host[K:2 * K].copy(), host[2 * K:].copy()
    intersection_meter.update(intersection), union_meter.update(union), target_meter.update(target)
    accuracy = sum(intersection_meter.val) / (sum(target_meter.val) + 1e-10)
    return intersection_meter, union_meter, target_meter, accuracy, intersection, union, target
