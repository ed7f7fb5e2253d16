"""Augmentation row on the device, through the C-ABI (doda_b200/augment.py -> csrc/augment.cu), against
the reference-made golden vectors, the numpy oracle on fresh seeded inputs, and -- when it travelled to this box -- the
staged, unmodified dataset/augmentor/augmentor_utils.py.  Coordinates are float64 like the reference's: the bar is
1e-9 absolute on values of O(100) (double rounding in a different summation order); masks and counts are exact.
"""
import os
import warnings

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _golden():
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "augment_golden.npz"))


def test_elastic_matches_reference_golden(cuda_dev):
    from doda_b200 import augment
    g = _golden()
    x = g["elastic_x"]
    for i in range(2):
        gran, mag, seed = g["elastic_arg%d" % i]
        np.random.seed(int(seed))  # the device version draws the noise from numpy's stream like the reference
        out = augment.elastic(x, int(gran), float(mag))
        assert out.dtype == torch.float64 and out.is_cuda
        assert np.abs(out.cpu().numpy() - g["elastic_out%d" % i]).max() <= TOL


@pytest.mark.parametrize("N,gran,mag,dtype", [(1, 6, 40.0, np.float32), (5000, 6, 40.0, np.float32), (5000, 20, 160.0, np.float64),
                                              (300000, 6, 40.0, np.float32), (40000, 3, 12.5, np.float32)])
def test_elastic_matches_oracle(cuda_dev, N, gran, mag, dtype):
    from doda_b200 import augment
    from oracle.augment import elastic_ref
    rng = np.random.RandomState(N + gran)
    x = (rng.rand(N, 3) * np.array([420, 310, 140]) - np.array([210, 155, 0])).astype(dtype)
    np.random.seed(7); ref = elastic_ref(x, gran, mag)
    np.random.seed(7); out = augment.elastic(torch.from_numpy(x).to(cuda_dev), gran, mag)
    state_after = np.random.rand()
    np.random.seed(7); elastic_ref(x, gran, mag)
    assert np.random.rand() == state_after, "the device version must consume numpy's random stream like the reference"
    assert np.abs(out.cpu().numpy() - ref).max() <= TOL
    assert np.abs(ref - x).max() > 1.0  # the distortion is not a no-op


def test_elastic_points_on_grid_nodes_and_outside(cuda_dev):
    """points exactly on grid nodes (searchsorted's left edge) and an explicit noise field with points outside it"""
    from doda_b200 import augment
    from oracle.augment import elastic_ref
    gran = 6
    x = np.array([[0, 0, 0], [12, -24, 36], [59.999, 60, -60], [-72, 72, 0.5]], dtype=np.float64)
    np.random.seed(1); ref = elastic_ref(x, gran, 40.0)
    np.random.seed(1); out = augment.elastic(x, gran, 40.0)
    assert np.abs(out.cpu().numpy() - ref).max() <= TOL


def test_elastic_matches_staged_reference(cuda_dev):
    from doda_b200 import augment
    from oracle import stage_ref
    au = stage_ref.load_augmentor_utils()
    if au is None:
        pytest.skip("reference not staged on this box")
    warnings.simplefilter("ignore")
    rng = np.random.RandomState(2)
    x = (rng.rand(20000, 3) * np.array([400, 300, 150]) - np.array([200, 150, 0])).astype(np.float32)
    for gran, mag in ((6, 40), (20, 160)):
        np.random.seed(21); ref = au.elastic(x, gran, mag)
        np.random.seed(21); out = augment.elastic(x, gran, mag)
        assert np.abs(out.cpu().numpy() - ref).max() <= TOL


def test_crop_matches_reference_golden(cuda_dev):
    from doda_b200 import augment
    g = _golden()
    xyz = g["crop_xyz"]
    for i in range(3):
        f0, f1, pr, mx, seed = g["crop_arg%d" % i]
        np.random.seed(int(seed))
        xo, valid = augment.crop(xyz, [int(f0), int(f1)], float(pr), int(mx))
        assert valid.dtype == torch.bool
        assert np.array_equal(valid.cpu().numpy(), g["crop_valid%d" % i])
        assert np.array_equal(xo.cpu().numpy(), g["crop_off%d" % i])  # one double add per coordinate: exact


@pytest.mark.parametrize("N,fs,pr,mx", [(300000, [128, 512], 2e9, 100000), (300000, [128, 512], 2e7, 250000),
                                        (1000, [128, 512], 2e9, 250000), (0, [128, 512], 2e9, 10)])
def test_crop_matches_oracle(cuda_dev, N, fs, pr, mx):
    from doda_b200 import augment
    from oracle.augment import crop_ref
    rng = np.random.RandomState(N % 97)
    xyz = (rng.rand(N, 3) * np.array([900, 700, 150])).astype(np.float64)
    if N == 0:
        xo, valid = augment.crop(xyz, fs, pr, mx)
        assert xo.shape == (0, 3) and valid.shape == (0,)
        return
    np.random.seed(5); r0, r1 = crop_ref(xyz, fs, pr, mx)
    np.random.seed(5); xo, valid = augment.crop(torch.from_numpy(xyz).to(cuda_dev), fs, pr, mx)
    assert np.array_equal(valid.cpu().numpy(), r1)
    assert np.array_equal(xo.cpu().numpy(), r0)
    assert int(valid.sum()) <= max(mx, 0) or int(valid.sum()) == N


def test_scene_aug_matches_reference_golden(cuda_dev):
    from doda_b200 import augment
    from oracle.stage_ref import _EasyDict
    g = _golden()
    aug = _EasyDict({"jitter": True, "flip": {"p": 0.5}, "rotation": {"p": 1.0, "value": [0.0, 0.0, 1.0]}})
    np.random.seed(300)
    out = augment.scene_aug(aug, g["elastic_x"])
    # np.matmul may fuse or reorder the three products per output; float64 on values of O(100)
    assert np.abs(out.cpu().numpy() - g["scene_out"]).max() <= 1e-10
