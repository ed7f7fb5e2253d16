"""spconv.modules — SparseModule marker and SparseSequential (spconv v1.2 semantics).

SparseSequential additionally recognises the `BatchNorm -> ReLU` pairs DODA puts in front of every conv
(model/unet_block.py:24-28,46-47,68-69,76-77; model/unet.py:43-44) and runs them as ONE fused sm_100a kernel
sequence (statistics + normalise + ReLU) instead of cuDNN BN + a separate ReLU pass.  Semantics are unchanged:
the result is written back into `input.features` of the SAME SparseConvTensor object.
"""
from collections import OrderedDict

import torch
from torch import nn

from .. import ops as _ops
from .tensor import SparseConvTensor


class SparseModule(nn.Module):
    """Marker base: modules deriving from it receive the SparseConvTensor itself."""
    pass


fuse_conv = True  # run [BatchNorm, ReLU, sparse conv] triplets as one autograd node
fuse_bn = True    # run BatchNorm-like modules through the engine's kernels (False: call the module itself; A/B tests)


def is_sparse_conv(module):
    from .conv import SparseConvolution
    return isinstance(module, SparseConvolution)


_bn_like_types = {}


def _is_bn_like(m):
    t = type(m)
    v = _bn_like_types.get(t)
    if v is None:
        v = _bn_like_types[t] = _is_bn_like_slow(m)
    return v


def _is_bn_like_slow(m):
    if isinstance(m, nn.SyncBatchNorm):
        return False  # needs cross-rank statistics: keep torch's implementation
    if isinstance(m, nn.modules.batchnorm._BatchNorm):
        return True
    # DODA's DSNorm (model/dsnorm.py) derives from nn.Module, not _BatchNorm
    return all(hasattr(m, a) for a in ("eps", "momentum", "track_running_stats", "weight", "bias",
                                       "running_mean_source", "running_var_source", "domain_label"))


class SparseSequential(SparseModule):
    def __init__(self, *args, **kwargs):
        super().__init__()
        if len(args) == 1 and isinstance(args[0], OrderedDict):
            for key, module in args[0].items():
                self.add_module(key, module)
        else:
            for idx, module in enumerate(args):
                self.add_module(str(idx), module)
        for name, module in kwargs.items():
            if name in self._modules:
                raise ValueError("name exists.")
            self.add_module(name, module)
        self._sparity_dict = {}

    def __getitem__(self, idx):
        if not (-len(self) <= idx < len(self)):
            raise IndexError("index {} is out of range".format(idx))
        if idx < 0:
            idx += len(self)
        it = iter(self._modules.values())
        for _ in range(idx):
            next(it)
        return next(it)

    def __len__(self):
        return len(self._modules)

    @property
    def sparity_dict(self):
        return self._sparity_dict

    def add(self, module, name=None):
        if name is None:
            name = str(len(self._modules))
            if name in self._modules:
                raise KeyError("name exists")
        self.add_module(name, module)

    def forward(self, input):
        mods = list(self._modules.values())
        i, n = 0, len(mods)
        while i < n:
            module = mods[i]
            if isinstance(module, SparseModule):
                input = module(input)
                i += 1
                continue
            if isinstance(input, SparseConvTensor):
                if input.indices.shape[0] != 0:
                    feats = input.features
                    if fuse_bn and _is_bn_like(module) and feats.is_cuda and feats.dim() == 2:
                        fuse_relu = i + 1 < n and isinstance(mods[i + 1], nn.ReLU)
                        if fuse_conv and fuse_relu and i + 2 < n and is_sparse_conv(mods[i + 2]) \
                                and feats.dtype == torch.float32:
                            stats_args = _ops.bn_batch_stats_args(module)
                            if stats_args is not None:  # training-mode statistics: the whole triplet in one node
                                input = mods[i + 2].forward_after_bn_relu(input, module, stats_args)
                                i += 3
                                continue
                        input.features = _ops.batch_norm_relu(feats, module, relu=fuse_relu)
                        i += 2 if fuse_relu else 1
                        continue
                    input.features = module(feats)
            else:
                input = module(input)
            i += 1
        return input
