"""Host-side mirror of DODA's sparse U-Net (model/unet.py:15-69, model/unet_block.py:10-100) built on the
engine's spconv surface.  It exists so that bench.py / smoke() / the GPU tests can run the hot path on a box
where /root/reference is absent; module names, parameter names and shapes are identical to the reference
(tests/golden/unet_state_dict.json pins that), so a reference checkpoint loads into it and vice versa.
The reference's own model files run unchanged on the same surface.
"""
import functools
from collections import OrderedDict

import torch
from torch import nn

from . import spconv
from . import ops as _ops
from . import pointgroup_ops
from . import tape as _tape


class ResidualBlock(spconv.SparseModule):
    """pre-activation residual unit: BN-ReLU-SubM3, BN-ReLU-SubM3, plus identity or a 1x1 SubM on the skip."""

    def __init__(self, in_channels, out_channels, norm_fn, indice_key=None):
        super().__init__()
        skip = nn.Identity() if in_channels == out_channels else \
            spconv.SubMConv3d(in_channels, out_channels, kernel_size=1, bias=False)
        self.i_branch = spconv.SparseSequential(skip)
        self.conv_branch = spconv.SparseSequential(
            norm_fn(in_channels), nn.ReLU(),
            spconv.SubMConv3d(in_channels, out_channels, kernel_size=3, padding=1, bias=False, indice_key=indice_key),
            norm_fn(out_channels), nn.ReLU(),
            spconv.SubMConv3d(out_channels, out_channels, kernel_size=3, padding=1, bias=False, indice_key=indice_key))

    def forward(self, x):
        skip_in = spconv.SparseConvTensor(x.features, x.indices, x.spatial_shape, x.batch_size)
        out = self.conv_branch(x)
        out.features += self.i_branch(skip_in).features
        return out


class VGGBlock(spconv.SparseModule):
    def __init__(self, in_channels, out_channels, norm_fn, indice_key=None):
        super().__init__()
        self.conv_layers = spconv.SparseSequential(
            norm_fn(in_channels), nn.ReLU(),
            spconv.SubMConv3d(in_channels, out_channels, kernel_size=3, padding=1, bias=False, indice_key=indice_key))

    def forward(self, x):
        return self.conv_layers(x)


class UBlock(nn.Module):
    """One U-Net level: blocks -> (BN-ReLU-down k2s2 -> child level -> BN-ReLU-inverse k2) -> concat -> tail."""

    def __init__(self, nPlanes, norm_fn, block_reps, block, indice_key_id=1):
        super().__init__()
        self.nPlanes = nPlanes
        c0 = nPlanes[0]
        sub_key = "subm%d" % indice_key_id
        self.blocks = spconv.SparseSequential(OrderedDict(
            ("block%d" % i, block(c0, c0, norm_fn, indice_key=sub_key)) for i in range(block_reps)))
        if len(nPlanes) > 1:
            down_key = "spconv%d" % indice_key_id
            self.conv = spconv.SparseSequential(
                norm_fn(c0), nn.ReLU(),
                spconv.SparseConv3d(c0, nPlanes[1], kernel_size=2, stride=2, bias=False, indice_key=down_key))
            self.u = UBlock(nPlanes[1:], norm_fn, block_reps, block, indice_key_id=indice_key_id + 1)
            self.deconv = spconv.SparseSequential(
                norm_fn(nPlanes[1]), nn.ReLU(),
                spconv.SparseInverseConv3d(nPlanes[1], c0, kernel_size=2, bias=False, indice_key=down_key))
            self.blocks_tail = spconv.SparseSequential(OrderedDict(
                ("block%d" % i, block(c0 * (2 - i), c0, norm_fn, indice_key=sub_key)) for i in range(block_reps)))

    tape = True  # run this sub-tree as ONE autograd node when it qualifies (doda_b200/tape.py); False: module by module

    def forward(self, x):
        if self.tape and _tape.enabled and self.training:
            p = _tape.cached_plan(self)
            if _tape.usable(self, p, x):
                return _tape.run(self, p, x)
        out = self.blocks(x)
        skip = spconv.SparseConvTensor(out.features, out.indices, out.spatial_shape, out.batch_size)
        if len(self.nPlanes) > 1:
            dec = self.deconv(self.u(self.conv(out)))
            out.features = torch.cat((skip.features, dec.features), dim=1)
            out = self.blocks_tail(out)
        return out


class SparseConvNet(nn.Module):
    def __init__(self, in_channel=3, mid_channel=16, n_classes=11, block_reps=2, block_residual=True):
        super().__init__()
        norm_fn = functools.partial(nn.BatchNorm1d, eps=1e-4, momentum=0.1)
        block = ResidualBlock if block_residual else VGGBlock
        m = mid_channel
        self.input_conv = spconv.SparseSequential(
            spconv.SubMConv3d(in_channel, m, kernel_size=3, padding=1, bias=False, indice_key="subm1"))
        self.unet = UBlock([m * i for i in range(1, 8)], norm_fn, block_reps, block, indice_key_id=1)
        self.output_layer = spconv.SparseSequential(norm_fn(m), nn.ReLU())
        self.linear = nn.Linear(m, n_classes)
        for mod in self.modules():
            if isinstance(mod, nn.modules.batchnorm._BatchNorm):
                mod.weight.data.fill_(1.0)
                mod.bias.data.fill_(0.0)

    def forward(self, input, input_map, return_mid_feat=False, v2p_map=None):
        out = self.output_layer(self.unet(self.input_conv(input)))
        # voxel -> points (model/unet.py:62); with the voxelizer's v2p map the backward is an atomics-free segmented sum
        point_feats = _ops.devoxelize(out.features, input_map, v2p_map)
        scores = self.linear(point_feats)
        return (point_feats, scores) if return_mid_feat else scores


def model_step(model, batch, voxel_mode=4, criterion=None, device="cuda", coords_pending=False):
    """One forward of the reference's `model_fn` (model/unet.py:72-99,154-198) on a collated batch dict:
    H2D copies, voxelize the point features, build the SparseConvTensor, run the net, cross-entropy.

    The voxel coordinates go through `ops.stage_coords` (copy + int cast on the engine's index stream) so that the
    rulebooks of this step do not queue behind the previous step's backward; device-resident `voxel_locs` must be
    complete, or pass coords_pending=True."""
    staged = batch.get("_staged_event")
    if staged is not None:  # ops.stage_batch started the copies on the index stream
        torch.cuda.current_stream().wait_event(staged)
    voxel_coords = _ops.stage_coords(batch["voxel_locs"], device, pending=coords_pending)
    p2v_map = batch["p2v_map"].to(device, non_blocking=True)
    v2p_map = batch["v2p_map"].to(device, non_blocking=True)
    feats = batch["feats"].to(device, non_blocking=True)
    labels = batch["labels"].to(device, non_blocking=True)
    batch_size = batch["offsets"].size(0) - 1
    voxel_feats = pointgroup_ops.voxelization(feats, v2p_map, voxel_mode)
    x = spconv.SparseConvTensor(voxel_feats, voxel_coords, batch["spatial_shape"], batch_size)
    scores = model(x, p2v_map, v2p_map=v2p_map) if isinstance(model, SparseConvNet) else model(x, p2v_map)
    if criterion is None:  # the engine's one-pass cross-entropy (same result as nn.CrossEntropyLoss(ignore_index=255))
        loss = _ops.cross_entropy(scores, labels, ignore_index=255)
    else:
        loss = criterion(scores, labels)
    return loss, scores
