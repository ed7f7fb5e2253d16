"""spconv.ops — spconv v1.2 names over the sm_100a engine."""
from ..ops import get_indice_pairs, indice_conv, indice_conv_backward, build_rulebook, conv_out_shape  # noqa: F401


def get_conv_output_size(input_size, kernel_size, stride, padding, dilation):
    return conv_out_shape(input_size, kernel_size, stride, padding, dilation)
