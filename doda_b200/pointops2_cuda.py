"""Module with the function surface of the reference's `pointops2_cuda` extension
(lib/pointops2/src/pointops_api.cpp:13-23), backed by libb200sparse.so.  `compat/pointops2_cuda.py` aliases it so
lib/pointops2/functions/pointops2.py:7-30 imports it instead of JIT-compiling the reference sources."""
import torch

from ._lib import lib, check


def _s():
    from .ops import _stream
    return _stream()


def _cuda(*ts):
    for t in ts:
        if not t.is_cuda or not t.is_contiguous():
            raise RuntimeError("pointops2_cuda: expected contiguous CUDA tensors (no CPU fallback)")


def knnquery_cuda(m, nsample, xyz, new_xyz, offset, new_offset, idx, dist2):
    _cuda(xyz, new_xyz, offset, new_offset, idx, dist2)
    check(lib.b200sp_knnquery(m, nsample, xyz.data_ptr(), new_xyz.data_ptr(), offset.data_ptr(),
                              new_offset.data_ptr(), idx.data_ptr(), dist2.data_ptr(), _s()), "knnquery")


def furthestsampling_cuda(b, n_max, xyz, offset, new_offset, tmp, idx):
    _cuda(xyz, offset, new_offset, tmp, idx)
    check(lib.b200sp_furthestsampling(b, int(n_max), 3, xyz.data_ptr(), offset.data_ptr(), new_offset.data_ptr(),
                                      tmp.data_ptr(), idx.data_ptr(), _s()), "furthestsampling")


def furthestsampling_dim_cuda(b, n_max, dim, xyz, offset, new_offset, tmp, idx):
    _cuda(xyz, offset, new_offset, tmp, idx)
    check(lib.b200sp_furthestsampling(b, int(n_max), int(dim), xyz.data_ptr(), offset.data_ptr(),
                                      new_offset.data_ptr(), tmp.data_ptr(), idx.data_ptr(), _s()),
          "furthestsampling_dim")


def grouping_forward_cuda(m, nsample, c, input, idx, output):
    _cuda(input, idx, output)
    check(lib.b200sp_grouping_fwd(m, nsample, c, input.data_ptr(), idx.data_ptr(), output.data_ptr(), _s()),
          "grouping_fwd")


def grouping_backward_cuda(m, nsample, c, grad_output, idx, grad_input):
    _cuda(grad_output, idx, grad_input)
    check(lib.b200sp_grouping_bwd(m, nsample, c, grad_output.data_ptr(), idx.data_ptr(), grad_input.data_ptr(),
                                  _s()), "grouping_bwd")


def interpolation_forward_cuda(n, c, k, input, idx, weight, output):
    _cuda(input, idx, weight, output)
    check(lib.b200sp_interpolation_fwd(n, c, k, input.data_ptr(), idx.data_ptr(), weight.data_ptr(),
                                       output.data_ptr(), _s()), "interpolation_fwd")


def interpolation_backward_cuda(n, c, k, grad_output, idx, weight, grad_input):
    _cuda(grad_output, idx, weight, grad_input)
    check(lib.b200sp_interpolation_bwd(n, c, k, grad_output.data_ptr(), idx.data_ptr(), weight.data_ptr(),
                                       grad_input.data_ptr(), _s()), "interpolation_bwd")


def subtraction_forward_cuda(n, nsample, c, input1, input2, idx, output):
    _cuda(input1, input2, idx, output)
    check(lib.b200sp_subtraction_fwd(n, nsample, c, input1.data_ptr(), input2.data_ptr(), idx.data_ptr(),
                                     output.data_ptr(), _s()), "subtraction_fwd")


def subtraction_backward_cuda(n, nsample, c, idx, grad_output, grad_input1, grad_input2):
    _cuda(idx, grad_output, grad_input1, grad_input2)
    check(lib.b200sp_subtraction_bwd(n, nsample, c, idx.data_ptr(), grad_output.data_ptr(), grad_input1.data_ptr(),
                                     grad_input2.data_ptr(), _s()), "subtraction_bwd")


def aggregation_forward_cuda(n, nsample, c, w_c, input, position, weight, idx, output):
    _cuda(input, position, weight, idx, output)
    check(lib.b200sp_aggregation_fwd(n, nsample, c, w_c, input.data_ptr(), position.data_ptr(), weight.data_ptr(),
                                     idx.data_ptr(), output.data_ptr(), _s()), "aggregation_fwd")


def aggregation_backward_cuda(n, nsample, c, w_c, input, position, weight, idx, grad_output, grad_input,
                              grad_position, grad_weight):
    _cuda(input, position, weight, idx, grad_output, grad_input, grad_position, grad_weight)
    check(lib.b200sp_aggregation_bwd(n, nsample, c, w_c, input.data_ptr(), position.data_ptr(), weight.data_ptr(),
                                     idx.data_ptr(), grad_output.data_ptr(), grad_input.data_ptr(),
                                     grad_position.data_ptr(), grad_weight.data_ptr(), _s()), "aggregation_bwd")
