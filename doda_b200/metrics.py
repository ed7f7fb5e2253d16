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
