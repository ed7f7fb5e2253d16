"""spconv.functional — autograd entry points with spconv v1.2 names."""
from ..ops import (SubMConvFunction, SparseConvFunction, SparseInverseConvFunction, DenseConvFunction,
                   indice_conv as _raw_indice_conv, indice_conv_backward as _raw_indice_conv_backward)
from torch.autograd import Function


class _RawIndiceConv(Function):
    """indice_conv / indice_subm_conv / indice_inverse_conv on raw (pairs, pairnum) tensors."""

    @staticmethod
    def forward(ctx, features, filters, indice_pairs, indice_pair_num, num_activate_out, inverse, subm):
        ctx.save_for_backward(indice_pairs, indice_pair_num, features, filters)
        ctx.inverse, ctx.subm = inverse, subm
        return _raw_indice_conv(features, filters, indice_pairs, indice_pair_num, num_activate_out, inverse, subm)

    @staticmethod
    def backward(ctx, grad_output):
        indice_pairs, indice_pair_num, features, filters = ctx.saved_tensors
        din, dW = _raw_indice_conv_backward(features, filters, grad_output, indice_pairs, indice_pair_num,
                                            ctx.inverse, ctx.subm)
        return din, dW, None, None, None, None, None


def indice_conv(features, filters, indice_pairs, indice_pair_num, num_activate_out):
    return _RawIndiceConv.apply(features, filters, indice_pairs, indice_pair_num, num_activate_out, False, False)


def indice_subm_conv(features, filters, indice_pairs, indice_pair_num, num_activate_out):
    return _RawIndiceConv.apply(features, filters, indice_pairs, indice_pair_num, num_activate_out, False, True)


def indice_inverse_conv(features, filters, indice_pairs, indice_pair_num, num_activate_out):
    return _RawIndiceConv.apply(features, filters, indice_pairs, indice_pair_num, num_activate_out, True, False)
