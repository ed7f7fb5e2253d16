"""CPU restatement (numpy only, no scipy) of the augmentation hot spots -- TEST INFRASTRUCTURE, never on the product
path.  Follows dataset/augmentor/augmentor_utils.py: elastic 61-80, crop 449-472.  Pinned against the reference's own
functions (staged unmodified under oracle/_ref/src by oracle/stage_ref.py) in tests/test_oracle_cpu.py.
"""
import math

import numpy as np


def _box3(n, axis):
    """scipy.ndimage.convolve(n, ones(3)/3 along axis, mode='constant', cval=0): double accumulation, float32 out
    (augmentor_utils.py:62-64, 68-73)"""
    w = np.float64(np.float32(1.0) / np.float32(3.0))
    a = n.astype(np.float64)
    up = np.zeros_like(a)
    dn = np.zeros_like(a)
    sl = [slice(None)] * 3
    s_hi, s_lo = list(sl), list(sl)
    s_hi[axis], s_lo[axis] = slice(1, None), slice(None, -1)
    up[tuple(s_lo)] = a[tuple(s_hi)]   # neighbour at +1
    dn[tuple(s_hi)] = a[tuple(s_lo)]   # neighbour at -1
    return ((w * up + w * a) + w * dn).astype(np.float32)


def elastic_ref(x, gran, mag, noise=None):
    """augmentor_utils.py:61-80; `noise` replaces the three np.random.randn grids when given"""
    x = np.asarray(x)
    bb = np.abs(x).max(0).astype(np.int32) // gran + 3
    if noise is None:
        noise = [np.random.randn(bb[0], bb[1], bb[2]).astype("float32") for _ in range(3)]
    noise = [np.asarray(n, dtype=np.float32) for n in noise]
    for axis in (0, 1, 2, 0, 1, 2):
        noise = [_box3(n, axis) for n in noise]
    ax = [np.linspace(-(b - 1) * gran, (b - 1) * gran, b) for b in bb]
    xd = x.astype(np.float64)
    idx, t, inside = [], [], np.ones(x.shape[0], dtype=bool)
    for d in range(3):
        # RegularGridInterpolator._find_indices: searchsorted - 1 clipped to [0, n - 2]; linear weight
        i = np.clip(np.searchsorted(ax[d], xd[:, d]) - 1, 0, len(ax[d]) - 2)
        idx.append(i)
        t.append((xd[:, d] - ax[d][i]) / (ax[d][i + 1] - ax[d][i]))
        inside &= (xd[:, d] >= ax[d][0]) & (xd[:, d] <= ax[d][-1])
    out = np.zeros((x.shape[0], 3))
    for c in range(3):
        v = np.zeros(x.shape[0])
        for e0 in (0, 1):
            for e1 in (0, 1):
                for e2 in (0, 1):
                    wgt = (t[0] if e0 else 1 - t[0]) * (t[1] if e1 else 1 - t[1]) * (t[2] if e2 else 1 - t[2])
                    v += wgt * noise[c][idx[0] + e0, idx[1] + e1, idx[2] + e2].astype(np.float64)
        out[:, c] = np.where(inside, v, 0.0)
    return x + out * mag


def crop_ref(xyz, full_scale, point_range, max_npoint):
    """augmentor_utils.py:449-472 (consumes np.random.rand(3) once per loop iteration, like the reference)"""
    xyz = np.asarray(xyz)
    xyz_offset = xyz.copy()
    valid = xyz_offset.min(1) >= 0
    assert valid.sum() == xyz.shape[0]
    full = np.array([full_scale[1]] * 3, dtype=np.float64)
    room_range = xyz.max(0) - xyz.min(0)
    curr_scale = room_range[0] * room_range[1] * room_range[2]
    if curr_scale > point_range:
        crop_scale = math.sqrt(point_range / curr_scale)
        full = np.minimum(full, np.array([crop_scale * room_range[0], crop_scale * room_range[1], room_range[2]]))
        valid = (xyz_offset < full).sum(1) == 3
    while valid.sum() > max_npoint:
        offset = np.clip(full - room_range + 0.001, None, 0) * np.random.rand(3)
        xyz_offset = xyz + offset
        valid = valid & (xyz_offset.min(1) >= 0) & ((xyz_offset < full).sum(1) == 3)
        full[:2] -= 32
    return xyz_offset, valid
