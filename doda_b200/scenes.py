"""Seeded synthetic scenes with ScanNet / S3DIS-like surface statistics and the collated batch
schema of the reference (dataset/dataset.py:121-187): no dataset or network needed."""
import numpy as np
import torch


def surface_scene(seed, W=300, D=225, H=135, nbox=14, dropout=0.2):
    """Integer voxel coords [M,3] of a room: floor, 4 partial-height walls, `nbox` boxes (top + 4 sides)."""
    rng = np.random.RandomState(seed)
    occ = np.zeros((W, D, H), dtype=bool)
    occ[:, :, 0] = True
    for (sl, h) in ((np.s_[0, :], 0.9), (np.s_[W - 1, :], 0.8), (np.s_[:, 0], 0.7), (np.s_[:, D - 1], 0.85)):
        hh = int(H * h)
        occ[sl + (slice(0, hh),)] = True
    for _ in range(nbox):
        w = rng.randint(min(15, W // 4), max(min(70, W // 3), min(15, W // 4) + 1))
        d = rng.randint(min(15, D // 4), max(min(70, D // 3), min(15, D // 4) + 1))
        h = rng.randint(min(15, H // 4), max(int(H * 0.6), min(15, H // 4) + 1))
        x0, y0 = rng.randint(1, W - w - 1), rng.randint(1, D - d - 1)
        occ[x0:x0 + w, y0:y0 + d, h] = True
        occ[x0, y0:y0 + d, 1:h] = True
        occ[x0 + w - 1, y0:y0 + d, 1:h] = True
        occ[x0:x0 + w, y0, 1:h] = True
        occ[x0:x0 + w, y0 + d - 1, 1:h] = True
    coords = np.argwhere(occ)
    keep = rng.rand(coords.shape[0]) >= dropout
    return coords[keep].astype(np.int64)


def scene_with_voxels(seed, target_voxels, aspect=(300, 225, 135), nbox=14):
    """A room scaled so that it holds a bit more than `target_voxels` active voxels, thinned at random to exactly
    that count (so BASELINE.json's "150k voxels" means 150 000 rows)."""
    base = 200000.0  # voxels of the 300x225x135 reference room
    s = (1.15 * target_voxels / base) ** 0.5
    while True:
        W, D = max(int(aspect[0] * s), 40), max(int(aspect[1] * s), 40)
        H = max(int(aspect[2] * min(s, 1.0) ** 0.5), 40)
        vox = surface_scene(seed, W, D, H, nbox=max(int(nbox * s * s), 3), dropout=0.15)
        if vox.shape[0] >= target_voxels:
            break
        s *= 1.1
    keep = np.sort(np.random.RandomState(seed + 777).permutation(vox.shape[0])[:target_voxels])
    return vox[keep]


def uniform_scene(seed, n_voxels, occupancy):
    """n_voxels distinct voxels drawn uniformly from a cube sized for the requested occupancy."""
    rng = np.random.RandomState(seed)
    side = int(np.ceil((n_voxels / occupancy) ** (1.0 / 3.0)))
    flat = rng.choice(side ** 3, size=n_voxels, replace=False)
    return np.stack(np.unravel_index(flat, (side, side, side)), axis=1).astype(np.int64)


def collate(scenes, seed=0, n_classes=11, dup_max=1, ignore_frac=0.05, full_scale_min=128, mode=4, voxelize=None):
    """scenes: list of int64 [M_i,3] voxel coords -> the reference's batch dict (CPU tensors).
    Points = voxels repeated 1..dup_max times (exercises maxActive > 1), shuffled like the augmentor does.
    `voxelize(locs, batch_size, mode)`: the CPU voxelizer to collate with; default = the engine's
    `pointgroup_ops.voxelization_idx` (imported on first use, so that generating scenes loads no native library)."""
    rng = np.random.RandomState(seed + 12345)
    locs, feats, labels, offsets = [], [], [], [0]
    for b, vox in enumerate(scenes):
        rep = rng.randint(1, dup_max + 1, size=vox.shape[0]) if dup_max > 1 else np.ones(vox.shape[0], dtype=np.int64)
        pts = np.repeat(vox, rep, axis=0)
        pts = pts[rng.permutation(pts.shape[0])]
        xyz = pts.astype(np.float32) + rng.rand(*pts.shape).astype(np.float32)
        f = xyz - xyz.mean(0, keepdims=True)
        f = f / max(float(np.abs(f).max()), 1.0)
        lab = rng.randint(0, n_classes, size=pts.shape[0]).astype(np.int64)
        lab[rng.rand(pts.shape[0]) < ignore_frac] = 255
        locs.append(np.concatenate([np.full((pts.shape[0], 1), b, dtype=np.int64), pts], axis=1))
        feats.append(f.astype(np.float32))
        labels.append(lab)
        offsets.append(offsets[-1] + pts.shape[0])
    locs = torch.from_numpy(np.concatenate(locs, 0))
    feats = torch.from_numpy(np.concatenate(feats, 0))
    labels = torch.from_numpy(np.concatenate(labels, 0))
    spatial_shape = np.clip((locs.max(0)[0][1:] + 1).numpy(), full_scale_min, None)  # dataset/dataset.py:176
    if voxelize is None:
        from . import pointgroup_ops
        voxelize = pointgroup_ops.voxelization_idx
    voxel_locs, p2v_map, v2p_map = voxelize(locs, len(scenes), mode)
    return {"locs": locs, "voxel_locs": voxel_locs, "p2v_map": p2v_map, "v2p_map": v2p_map,
            "locs_float": locs[:, 1:].float(), "feats": feats, "labels": labels,
            "offsets": torch.tensor(offsets, dtype=torch.int32), "spatial_shape": spatial_shape,
            "id": list(range(len(scenes)))}
