import numpy as np
import torch


def rel_err(a, b):
    """max|a-b| / max(max|b|, 1e-6)"""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-6))


def random_coords(seed, n, batch, shape):
    """n distinct voxels per batch element inside `shape`, in random (unsorted) order -> int32 [batch*n, 4]"""
    rng = np.random.RandomState(seed)
    out = []
    vol = int(np.prod(shape))
    for b in range(batch):
        flat = rng.choice(vol, size=min(n, vol), replace=False)
        c = np.stack(np.unravel_index(flat, shape), axis=1)
        out.append(np.concatenate([np.full((c.shape[0], 1), b), c], axis=1))
    return np.concatenate(out, 0).astype(np.int32)


def surface_coords(seed, target, batch):
    from doda_b200 import scenes
    out = []
    for b in range(batch):
        v = scenes.scene_with_voxels(seed + b, target)
        v = v[np.random.RandomState(seed + b).permutation(v.shape[0])]
        out.append(np.concatenate([np.full((v.shape[0], 1), b), v], axis=1))
    c = np.concatenate(out, 0).astype(np.int32)
    shape = (c[:, 1:].max(0) + 1).tolist()
    return c, shape
