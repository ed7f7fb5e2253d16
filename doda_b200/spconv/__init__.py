"""Drop-in for the `spconv` v1.2 package surface DODA imports (model/unet.py:3, model/unet_block.py:3-4)."""
from . import modules, conv, ops, functional  # noqa: F401
from .tensor import SparseConvTensor
from .modules import SparseModule, SparseSequential
from .conv import SparseConvolution, SparseConv3d, SubMConv3d, SparseInverseConv3d

__version__ = "1.2.1+b200"
__all__ = ["SparseConvTensor", "SparseModule", "SparseSequential", "SparseConvolution", "SparseConv3d",
           "SubMConv3d", "SparseInverseConv3d", "modules", "conv", "ops", "functional"]
