"""Golden vectors for the augmentation row, produced by the REFERENCE's own functions
(dataset/augmentor/augmentor_utils.py: elastic 61-80, crop 449-472, scene_aug 85-104) imported from /root/reference
through oracle/stage_ref.py.  Run in the build container:  python tests/golden/make_augment_golden.py
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import stage_ref  # noqa: E402


def main():
    warnings.simplefilter("ignore")
    au = stage_ref.load_augmentor_utils()
    assert au is not None, "reference not staged"
    rng = np.random.RandomState(11)
    x = (rng.rand(1200, 3) * np.array([400, 300, 150]) - np.array([200, 150, 0])).astype(np.float32)
    out = {"elastic_x": x}
    for i, (gran, mag) in enumerate(((6, 40), (20, 160))):
        np.random.seed(100 + i)
        out["elastic_out%d" % i] = au.elastic(x, gran, mag)
        out["elastic_arg%d" % i] = np.array([gran, mag, 100 + i], dtype=np.float64)
    xyz = (rng.rand(2500, 3) * np.array([900, 700, 150])).astype(np.float64)
    out["crop_xyz"] = xyz
    for i, (fs, pr, mx) in enumerate((([128, 512], 2e9, 800), ([128, 512], 2e7, 2000), ([128, 1024], 2e9, 3000))):
        np.random.seed(200 + i)
        xo, valid = au.crop(xyz, fs, pr, mx)
        out["crop_off%d" % i] = xo
        out["crop_valid%d" % i] = valid
        out["crop_arg%d" % i] = np.array([fs[0], fs[1], pr, mx, 200 + i], dtype=np.float64)
    aug = stage_ref._EasyDict({"jitter": True, "flip": {"p": 0.5}, "rotation": {"p": 1.0, "value": [0.0, 0.0, 1.0]}})
    np.random.seed(300)
    out["scene_out"] = au.scene_aug(aug, x)
    np.savez_compressed(os.path.join(HERE, "augment_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "augment_golden.npz"))


if __name__ == "__main__":
    main()
