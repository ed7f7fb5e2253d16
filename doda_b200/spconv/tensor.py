"""spconv.SparseConvTensor (spconv v1.2 container)."""
import numpy as np
import torch


class SparseConvTensor(object):
    def __init__(self, features, indices, spatial_shape, batch_size, grid=None):
        """
        features: [num_points, num_features] float
        indices:  [num_points, ndim + 1] int32, batch index first, axes in the order of spatial_shape
        """
        self.features = features
        self.indices = indices
        if self.indices.dtype != torch.int32:
            self.indices.int()  # spconv v1.2 does the same no-op; callers pass .int() (model/unet.py:94)
        self.spatial_shape = spatial_shape
        self.batch_size = batch_size
        self.indice_dict = {}
        self.grid = grid

    @property
    def spatial_size(self):
        return int(np.prod(self.spatial_shape))

    def find_indice_pair(self, key):
        if key is None:
            return None
        if key in self.indice_dict:
            return self.indice_dict[key]
        return None

    def dense(self, channels_first=True):
        ndim = len(self.spatial_shape)
        out_shape = [self.batch_size] + [int(s) for s in self.spatial_shape] + [self.features.shape[1]]
        res = torch.zeros(out_shape, dtype=self.features.dtype, device=self.features.device)
        idx = self.indices.long()
        res[tuple(idx[:, i] for i in range(ndim + 1))] = self.features
        if not channels_first:
            return res
        perm = [0, ndim + 1] + list(range(1, ndim + 1))
        return res.permute(*perm).contiguous()

    @property
    def sparity(self):
        return self.indices.shape[0] / self.spatial_size / self.batch_size
