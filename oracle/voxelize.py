"""numpy restatement of the reference's CPU hash voxelizer `voxelize_idx` for the sum / mean modes the data path uses
(lib/pointgroup_ops/src/voxelize/voxelize.cpp:62-155: first-touch voxel ids in scan order, per voxel the ascending list
of its points, output_map rows = [count, pt0, pt1, ..., -1 pad]) -- TEST INFRASTRUCTURE, and the collate step of
bench.py's reference arm (which must not load the product library).  Pinned against tests/golden/voxelize_idx.npz,
which was produced by the reference's own compiled code (tests/golden/make_golden.py)."""
import numpy as np
import torch


def voxelize_idx_ref(coords, batch_size, mode=4):
    """coords int64 [N, 4] = (batch, x, y, z) -> (output_coords int64 [M,4], input_map int32 [N], output_map int32
    [M, 1 + maxActive]) exactly as `PG_OP.voxelize_idx(..., mode in {3, 4})` returns them."""
    if int(mode) not in (3, 4):
        raise NotImplementedError("oracle voxelizer: sum / mean modes only (voxelize.cpp:143-153)")
    c = np.ascontiguousarray(coords.numpy() if isinstance(coords, torch.Tensor) else coords).astype(np.int64)
    N = c.shape[0]
    if N == 0:
        return (torch.zeros((0, 4), dtype=torch.int64), torch.zeros(0, dtype=torch.int32),
                torch.zeros((0, 1), dtype=torch.int32))
    assert c.min() >= 0 and c[:, 1:].max() < (1 << 20) and c[:, 0].max() < (1 << 3)
    key = ((c[:, 0] << 60) | (c[:, 1] << 40) | (c[:, 2] << 20) | c[:, 3])
    uniq, first, inv, counts = np.unique(key, return_index=True, return_inverse=True, return_counts=True)
    order = np.argsort(first, kind="stable")          # voxels in first-touch order (voxelize.cpp:78-85)
    rank = np.empty_like(order)
    rank[order] = np.arange(order.shape[0])
    p2v = rank[inv]
    M = order.shape[0]
    cnt = counts[order]
    A = int(cnt.max())
    omap = np.full((M, 1 + A), -1, dtype=np.int32)
    omap[:, 0] = cnt
    by_vox = np.argsort(p2v, kind="stable")            # points grouped by voxel, ascending point index inside
    start = np.concatenate([[0], np.cumsum(cnt)[:-1]])
    slot = np.arange(N) - np.repeat(start, cnt)
    omap[p2v[by_vox], 1 + slot] = by_vox
    return (torch.from_numpy(c[first[order]].copy()), torch.from_numpy(p2v.astype(np.int32)),
            torch.from_numpy(omap))


def init_state_dict(shapes, seed=0):
    """a SparseConvNet state_dict with the reference's key names / shapes (tests/golden/unet_state_dict.json, made
    from the reference's own model file): conv weights U(+-1/sqrt(fan_in)) with fan_in = k*k*Cin*Cout... as spconv
    v1.2 initialises them, BatchNorm weight 1 / bias 0 (model/unet.py:51-56), Linear like nn.Linear."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k in sorted(shapes):
        shp = list(shapes[k])
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.zeros((), dtype=torch.long)
        elif k.endswith("running_mean"):
            sd[k] = torch.zeros(shp)
        elif k.endswith("running_var"):
            sd[k] = torch.ones(shp)
        elif len(shp) == 5:   # sparse conv [k,k,k,Cin,Cout]: kaiming_uniform_(a=sqrt(5)) on that shape
            fan_in = shp[1] * shp[2] * shp[3] * shp[4]
            sd[k] = (torch.rand(shp, generator=g) * 2 - 1) / fan_in ** 0.5
        elif k == "linear.weight":
            sd[k] = (torch.rand(shp, generator=g) * 2 - 1) / shp[1] ** 0.5
        elif k == "linear.bias":
            sd[k] = (torch.rand(shp, generator=g) * 2 - 1) / 16 ** 0.5
        elif k.endswith(".weight"):
            sd[k] = torch.ones(shp)
        else:
            sd[k] = torch.zeros(shp)
    return sd
