"""Functional CPU restatement of DODA's SparseConvNet forward (model/unet.py:58-69 with the blocks of
model/unet_block.py:32-38 ResidualBlock, 87-100 UBlock) on top of the oracle rulebook/conv.  Parameters come from a
state_dict with the reference's key names; gradients come from torch autograd, so one call gives the forward AND
backward baseline (`loss.backward()`)."""
import numpy as np
import torch
import torch.nn.functional as F

from .rulebook import get_indice_pairs_ref
from .conv import indice_conv_ref


class _Sp(object):
    def __init__(self, feats, idx, shape, bs, rbs):
        self.f, self.idx, self.shape, self.bs, self.rbs = feats, idx, shape, bs, rbs


# Optional {BatchNorm key: bool mask [rows, C]} that PINS the ReLU gates (tests only): with the gates taken from the
# implementation under test, the net is a smooth function of its inputs and whole-net gradients can be compared at a
# fixed tight bar -- without it, two fp32-accurate evaluations disagree on the sign of a ~1e-6 fraction of the
# pre-activations and every flipped gate moves the gradients by a whole row's contribution.
_RELU_MASKS = None


def _bn_relu(sd, pre, x, training, eps=1e-4, momentum=0.1):
    # nn.BatchNorm1d(eps=1e-4, momentum=0.1) + nn.ReLU (model/unet.py:28,43-44)
    rm, rv = sd.get(pre + ".running_mean"), sd.get(pre + ".running_var")
    y = F.batch_norm(x, None if training else rm, None if training else rv, sd[pre + ".weight"], sd[pre + ".bias"],
                     training, momentum, eps)
    if _RELU_MASKS is not None and pre in _RELU_MASKS:
        return y * _RELU_MASKS[pre].to(y.dtype)
    return F.relu(y)


def _subm3(sd, name, t, key):
    if key not in t.rbs:
        t.rbs[key] = get_indice_pairs_ref(t.idx, t.bs, t.shape, 3, 1, 1, 1, subm=True)
    _, pairs, pairnum, _ = t.rbs[key]
    return indice_conv_ref(t.f, sd[name], pairs, pairnum, t.f.shape[0], subm=True)


def _residual(sd, pre, t, key, training):
    # model/unet_block.py:32-38
    x = t.f
    w_skip = sd.get(pre + ".i_branch.0.weight")
    skip = x if w_skip is None else x @ w_skip.reshape(w_skip.shape[-2], w_skip.shape[-1])  # 1x1 SubM == GEMM
    h = _bn_relu(sd, pre + ".conv_branch.0", x, training)
    h = _subm3(sd, pre + ".conv_branch.2.weight", _Sp(h, t.idx, t.shape, t.bs, t.rbs), key)
    h = _bn_relu(sd, pre + ".conv_branch.3", h, training)
    h = _subm3(sd, pre + ".conv_branch.5.weight", _Sp(h, t.idx, t.shape, t.bs, t.rbs), key)
    return _Sp(h + skip, t.idx, t.shape, t.bs, t.rbs)


def _vgg(sd, pre, t, key, training):
    # model/unet_block.py:41-52 (block_residual: False): BN-ReLU-SubM3, no skip
    h = _bn_relu(sd, pre + ".conv_layers.0", t.f, training)
    h = _subm3(sd, pre + ".conv_layers.2.weight", _Sp(h, t.idx, t.shape, t.bs, t.rbs), key)
    return _Sp(h, t.idx, t.shape, t.bs, t.rbs)


def _block(sd, pre, t, key, training):
    # the block type is read off the state_dict keys (ResidualBlock has a conv_branch, VGGBlock has conv_layers)
    if (pre + ".conv_layers.2.weight") in sd:
        return _vgg(sd, pre, t, key, training)
    return _residual(sd, pre, t, key, training)


def _ublock(sd, pre, t, level, nlevels, reps, training):
    # model/unet_block.py:87-100
    key = "subm%d" % level
    for i in range(reps):
        t = _block(sd, "%s.blocks.block%d" % (pre, i), t, key, training)
    if level < nlevels:
        skip = t.f
        h = _bn_relu(sd, pre + ".conv.0", t.f, training)
        outids, pairs, pairnum, oshape = get_indice_pairs_ref(t.idx, t.bs, t.shape, 2, 2, 0, 1, subm=False)
        d = indice_conv_ref(h, sd[pre + ".conv.2.weight"], pairs, pairnum, outids.shape[0])
        child = _ublock(sd, pre + ".u", _Sp(d, outids, oshape, t.bs, t.rbs), level + 1, nlevels, reps, training)
        h = _bn_relu(sd, pre + ".deconv.0", child.f, training)
        up = indice_conv_ref(h, sd[pre + ".deconv.2.weight"], pairs, pairnum, t.f.shape[0], inverse=True)
        t = _Sp(torch.cat((skip, up), dim=1), t.idx, t.shape, t.bs, t.rbs)
        for i in range(reps):
            t = _block(sd, "%s.blocks_tail.block%d" % (pre, i), t, key, training)
    return t


def unet_forward_ref(sd, voxel_feats, voxel_coords, spatial_shape, batch_size, p2v, training=True, nlevels=7,
                     reps=2):
    """-> per-point scores [N, n_classes]  (model/unet.py:58-69)"""
    idx = np.asarray(voxel_coords, dtype=np.int64)
    t = _Sp(voxel_feats, idx, [int(s) for s in spatial_shape], batch_size, {})
    t = _Sp(_subm3(sd, "input_conv.0.weight", t, "subm1"), idx, t.shape, batch_size, t.rbs)
    t = _ublock(sd, "unet", t, 1, nlevels, reps, training)
    h = _bn_relu(sd, "output_layer.0", t.f, training)
    pts = h[torch.as_tensor(np.asarray(p2v), dtype=torch.int64)]
    return pts @ sd["linear.weight"].t() + sd["linear.bias"]


def voxelize_mean_ref(feats, v2p):
    """mode-4 voxelization of point features (voxelize.cu:10-23): mean over the voxel's points."""
    v2p = torch.as_tensor(np.asarray(v2p), dtype=torch.int64)
    cnt = v2p[:, 0].clamp(min=1).to(feats.dtype)
    out = torch.zeros((v2p.shape[0], feats.shape[1]), dtype=feats.dtype)
    for j in range(1, v2p.shape[1]):
        m = v2p[:, 0] >= j
        out[m] += feats[v2p[m, j]]
    return out / cnt[:, None]


def model_step_ref(sd, batch, training=True, relu_masks=None):
    """model_fn forward (model/unet.py:72-99,154-198): voxelize features, net, CE(ignore 255). -> (loss, scores)
    relu_masks: see _RELU_MASKS"""
    global _RELU_MASKS
    vf = voxelize_mean_ref(batch["feats"], batch["v2p_map"])
    _RELU_MASKS = relu_masks
    try:
        scores = unet_forward_ref(sd, vf, batch["voxel_locs"].numpy(), batch["spatial_shape"],
                                  batch["offsets"].shape[0] - 1, batch["p2v_map"].numpy(), training)
    finally:
        _RELU_MASKS = None
    loss = F.cross_entropy(scores, batch["labels"], ignore_index=255)
    return loss, scores


def encoder_decoder_ref(weights_down, weights_up, bn_down, bn_up, x, coords, spatial_shape, batch_size, training=True,
                        relu_masks=None):
    """BASELINE configs[3]: the stride-2 SparseConv3d / SparseInverseConv3d encoder-decoder of UBlock
    (model/unet_block.py:67-79) without the SubM blocks.  weights_down[l] / weights_up[l]: [2,2,2,Cin,Cout] filters of
    level l's down conv / inverse conv; bn_down[l] / bn_up[l]: (weight, bias) of the BatchNorm in front of each (None =
    no BN/ReLU).  -> features on the input's own active set [M, C0].
    relu_masks: optional list of 0/1 masks, one per BN+ReLU in execution order, that replace the ReLU gates (the
    pinned-gate comparison of oracle/gates.py)."""
    masks = list(relu_masks) if relu_masks is not None else None

    def act(y):
        if masks is None:
            return F.relu(y)
        return y * masks.pop(0).to(y.dtype)

    idx = np.asarray(coords, dtype=np.int64)
    shape = [int(s) for s in spatial_shape]
    stack, f = [], x
    for l, W in enumerate(weights_down):
        if bn_down is not None:
            f = act(F.batch_norm(f, None, None, bn_down[l][0], bn_down[l][1], True, 0.1, 1e-4))
        outids, pairs, pairnum, oshape = get_indice_pairs_ref(idx, batch_size, shape, 2, 2, 0, 1, subm=False)
        stack.append((pairs, pairnum, f.shape[0]))
        f = indice_conv_ref(f, W, pairs, pairnum, outids.shape[0])
        idx, shape = outids, oshape
    for l in reversed(range(len(weights_up))):
        pairs, pairnum, n_fine = stack[l]
        if bn_up is not None:
            f = act(F.batch_norm(f, None, None, bn_up[l][0], bn_up[l][1], True, 0.1, 1e-4))
        f = indice_conv_ref(f, weights_up[l], pairs, pairnum, n_fine, inverse=True)
    return f
