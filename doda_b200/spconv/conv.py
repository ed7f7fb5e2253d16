"""spconv.conv — SparseConvolution and the three 3-D variants DODA uses (SubMConv3d k=3/k=1, SparseConv3d k=2 s=2,
SparseInverseConv3d k=2; model/unet.py:36, model/unet_block.py:20,26,29,48,70,78).  Same constructor signature,
parameter names and shapes as spconv v1.2 (`weight` [k,k,k,Cin,Cout], optional `bias` [Cout]) so reference
checkpoints load (util/model_utils.py:42-94)."""
import math

import numpy as np
import torch
from torch import nn
from torch.nn import init
from torch.nn.parameter import Parameter

from .. import ops as _ops
from .modules import SparseModule
from .tensor import SparseConvTensor


def _ntuple(v, n):
    if isinstance(v, (list, tuple)):
        assert len(v) == n
        return [int(x) for x in v]
    return [int(v)] * n


class SparseConvolution(SparseModule):
    def __init__(self, ndim, in_channels, out_channels, kernel_size=3, stride=1, padding=0, dilation=1, groups=1,
                 bias=True, subm=False, output_padding=0, transposed=False, inverse=False, indice_key=None,
                 fused_bn=False, use_hash=False, algo=None):
        super().__init__()
        assert groups == 1, "groups != 1 is not supported (spconv v1.2 asserts the same)"
        assert ndim == 3, "only the 3-D operators DODA uses are built"
        self.ndim = ndim
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = _ntuple(kernel_size, ndim)
        self.conv1x1 = int(np.prod(self.kernel_size)) == 1
        self.stride = _ntuple(stride, ndim)
        self.padding = _ntuple(padding, ndim)
        self.dilation = _ntuple(dilation, ndim)
        self.transposed = transposed
        self.inverse = inverse
        self.output_padding = _ntuple(output_padding, ndim)
        self.groups = groups
        self.subm = subm
        self.indice_key = indice_key
        self.fused_bn = fused_bn
        self.use_hash = use_hash  # accepted for API parity; the engine always hashes
        self.algo = algo
        self.weight = Parameter(torch.Tensor(*self.kernel_size, in_channels, out_channels))
        if bias:
            self.bias = Parameter(torch.Tensor(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()
        _ops.register_conv_module(self)

    def reset_parameters(self):
        init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            fan_in, _ = init._calculate_fan_in_and_fan_out(self.weight)
            bound = 1 / math.sqrt(fan_in)
            init.uniform_(self.bias, -bound, bound)

    def extra_repr(self):
        return "{in_channels}, {out_channels}, kernel_size={kernel_size}, stride={stride}, subm={subm}, " \
               "inverse={inverse}, indice_key={indice_key}".format(**self.__dict__)

    def _resolve(self, input):
        """-> (kind, rulebook, output indices, output spatial shape) for this conv on `input`"""
        if self.conv1x1:
            return "dense", None, input.indices, input.spatial_shape
        rb = input.find_indice_pair(self.indice_key)
        if self.inverse:
            assert rb is not None and self.indice_key is not None, "inverse conv needs the rulebook of its key"
            assert rb.kind == "conv" and rb.K == int(np.prod(self.kernel_size)), "kernel size mismatch with the key"
            return "inverse", rb, rb.indices, rb.spatial_shape
        if rb is None:
            rb = _ops.build_rulebook(input.indices, input.batch_size, input.spatial_shape, self.kernel_size, self.stride,
                                     self.padding, self.dilation, subm=self.subm)
            if self.indice_key is not None:
                input.indice_dict[self.indice_key] = rb
        return ("subm" if self.subm else "conv"), rb, rb.outids, rb.out_spatial_shape

    def _wrap(self, input, out_features, outids, out_spatial_shape):
        if self.bias is not None:
            out_features = out_features + self.bias
        out_tensor = SparseConvTensor(out_features, outids, out_spatial_shape, input.batch_size)
        out_tensor.indice_dict = input.indice_dict
        out_tensor.grid = input.grid
        return out_tensor

    def forward(self, input):
        assert isinstance(input, SparseConvTensor)
        kind, rb, outids, oshape = self._resolve(input)
        prep = _ops.prepared_weights(self)
        if kind == "dense":
            out_features = _ops.DenseConvFunction.apply(input.features, self.weight, prep)
        else:
            out_features = _ops._CONV_FN[kind].apply(input.features, self.weight, rb, prep)
        return self._wrap(input, out_features, outids, oshape)

    def forward_after_bn_relu(self, input, bn, stats_args):
        """[BatchNorm (batch statistics), ReLU, this conv] on `input` as one autograd node (SparseSequential calls
        this when it sees the triplet).  `input.features` is rebound to the BN+ReLU activation, as the unfused
        sequence would leave it."""
        kind, rb, outids, oshape = self._resolve(input)
        rm, rv, nbt, momentum = stats_args
        out_features, act = _ops.BNReLUConvFunction.apply(input.features, bn.weight, bn.bias, self.weight, rm, rv, nbt,
                                                          momentum, bn.eps, rb, _ops.prepared_weights(self), kind)
        input.features = act
        return self._wrap(input, out_features, outids, oshape)


class SparseConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None, use_hash=False, algo=None):
        super().__init__(3, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias,
                         indice_key=indice_key, use_hash=use_hash, algo=algo)


class SubMConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None, use_hash=False, algo=None):
        super().__init__(3, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias, True,
                         indice_key=indice_key, use_hash=use_hash, algo=algo)


class SparseInverseConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, indice_key, bias=True, algo=None):
        super().__init__(3, in_channels, out_channels, kernel_size, bias=bias, inverse=True, indice_key=indice_key,
                         algo=algo)
