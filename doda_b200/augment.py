"""Augmentation hot spots on the device behind the reference's own function names.

`elastic`, `crop` and `scene_aug` take the arguments of dataset/augmentor/augmentor_utils.py (elastic 61-80, scene_aug
85-104, crop 449-472) and consume numpy's global random stream exactly as the reference does (same calls, same order),
so a seeded pipeline produces the same scenes whichever implementation runs; the arrays live on the GPU and the
per-point / per-cell work runs in the kernels of csrc/augment.cu.  There is no CPU path: without the CUDA library the
import of `_lib` raises.
"""
import math

import numpy as np
import torch

from . import ops as _ops
from ._lib import lib, check


def _dev_points(x, device=None):
    """numpy or tensor [N,3] -> contiguous device tensor (float32 stays float32, everything else becomes float64)"""
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    if device is None:
        device = x.device if x.is_cuda else torch.device("cuda")
    x = x.to(device)
    if x.dtype not in (torch.float32, torch.float64):
        x = x.double()
    return x.contiguous()


def elastic(x, gran, mag, noise=None):
    """augmentor_utils.py:61-80.  x: [N,3] numpy or tensor; returns a float64 CUDA tensor [N,3].
    noise: optional list of three float32 arrays of shape bb (what np.random.randn would have produced)."""
    xd = _dev_points(x)
    N = xd.shape[0]
    # bb = np.abs(x).max(0).astype(np.int32) // gran + 3   (augmentor_utils.py:66)
    amax = xd.abs().amax(0).cpu().numpy() if N else np.zeros(3)
    bb = (amax.astype(np.int32) // gran + 3)
    bb = [int(b) for b in bb]
    if noise is None:
        noise = [np.random.randn(bb[0], bb[1], bb[2]).astype("float32") for _ in range(3)]
    grids = torch.from_numpy(np.stack([np.asarray(n, dtype=np.float32).reshape(bb) for n in noise])).to(xd.device).contiguous()
    scratch = torch.empty_like(grids)
    st = _ops._stream()
    check(lib.b200sp_elastic_blur(grids.data_ptr(), scratch.data_ptr(), bb[0], bb[1], bb[2], st), "elastic_blur")
    out = torch.empty(N, 3, dtype=torch.float64, device=xd.device)
    check(lib.b200sp_elastic_apply(xd.data_ptr(), 1 if xd.dtype == torch.float64 else 0, N, grids.data_ptr(), bb[0], bb[1], bb[2],
                                   float(gran), float(mag), out.data_ptr(), st), "elastic_apply")
    return out


def affine(xyz, m):
    """xyz @ m for a 3x3 matrix m (float64 result), the last line of scene_aug (augmentor_utils.py:103)"""
    xd = _dev_points(xyz)
    m9 = np.ascontiguousarray(np.asarray(m, dtype=np.float64).reshape(9))
    out = torch.empty(xd.shape[0], 3, dtype=torch.float64, device=xd.device)
    check(lib.b200sp_affine3(xd.data_ptr(), 1 if xd.dtype == torch.float64 else 0, xd.shape[0], m9.ctypes.data, out.data_ptr(),
                             _ops._stream()), "affine3")
    return out


def _enabled(key):
    # check_key, augmentor_utils.py:13-24
    if key is None:
        return False
    if isinstance(key, bool):
        return key
    if isinstance(key, dict):
        return key.get("enabled", True)
    return True


def _passes(key):
    # check_p, augmentor_utils.py:27-28
    return (not isinstance(key, dict)) or ("p" not in key) or (np.random.rand() < key["p"])


def scene_matrix(aug):
    """the 3x3 matrix scene_aug builds (augmentor_utils.py:87-102), drawing from np.random in the reference's order"""
    m = np.eye(3)
    if _enabled(aug.jitter):
        m += np.random.randn(3, 3) * 0.1
    if _enabled(aug.flip) and _passes(aug.flip):
        m[0][0] *= -1
    if _enabled(aug.rotation) and _passes(aug.rotation):
        th = [(np.random.rand() * 2 * math.pi - math.pi) * aug.rotation.value[i] for i in range(3)]
        cx, sx, cy, sy, cz, sz = math.cos(th[0]), math.sin(th[0]), math.cos(th[1]), math.sin(th[1]), math.cos(th[2]), math.sin(th[2])
        Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
        Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
        Rz = np.array([[cz, sz, 0], [-sz, cz, 0], [0, 0, 1]])
        m = np.matmul(m, Rx.dot(Ry).dot(Rz))
    return m


def scene_aug(aug, xyz):
    """augmentor_utils.py:85-104"""
    assert xyz.ndim == 2
    return affine(xyz, scene_matrix(aug))


def crop(xyz, full_scale, point_range, max_npoint):
    """augmentor_utils.py:449-472.  Returns (xyz_offset float64 CUDA [N,3], valid_idxs bool CUDA [N]).  One 4-byte
    device->host read per iteration of the point-count loop (the loop's exit test is the count)."""
    xd = _dev_points(xyz).double()
    N = xd.shape[0]
    dev = xd.device
    lo, hi = (xd.amin(0).cpu().numpy(), xd.amax(0).cpu().numpy()) if N else (np.zeros(3), np.zeros(3))
    assert N == 0 or lo.min() >= 0, "crop: negative coordinates (the reference asserts the same)"
    full = np.array([full_scale[1]] * 3, dtype=np.float64)
    room_range = hi - lo
    curr_scale = room_range[0] * room_range[1] * room_range[2]
    valid = torch.ones(N, dtype=torch.uint8, device=dev)
    count = torch.empty(1, dtype=torch.int32, device=dev)
    xyz_offset = xd.clone()
    st = _ops._stream()
    zero = np.zeros(3, dtype=np.float64)
    n_valid = N

    def mask(offset, full_, write):
        o = np.ascontiguousarray(offset, dtype=np.float64)
        f = np.ascontiguousarray(full_, dtype=np.float64)
        check(lib.b200sp_crop_mask(xd.data_ptr(), N, o.ctypes.data, f.ctypes.data, valid.data_ptr(),
                                   xyz_offset.data_ptr() if write else None, count.data_ptr(), st), "crop_mask")
        return int(count.item())

    if curr_scale > point_range:
        crop_scale = math.sqrt(point_range / curr_scale)
        full = np.minimum(full, np.array([crop_scale * room_range[0], crop_scale * room_range[1], room_range[2]]))
        n_valid = mask(zero, full, False)
    while n_valid > max_npoint:
        offset = np.clip(full - room_range + 0.001, None, 0) * np.random.rand(3)
        n_valid = mask(offset, full, True)
        full[:2] -= 32
    return xyz_offset, valid.bool()
