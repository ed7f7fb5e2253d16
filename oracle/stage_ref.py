"""Stage the reference's UNCHANGED Python files of the hot path into oracle/_ref/src/ -- TEST INFRASTRUCTURE.

`/root/reference` does not exist on the GPU box, but the drop-in boundary is only proven when the reference's own
`model/unet.py`, `model/unet_block.py`, `model/dsnorm.py`, `lib/*/functions/*.py` and `util/model_utils.py` run, as they
are, on top of compat/ (spconv, PG_OP, pointops2_cuda) on a B200.  This script copies those files byte for byte from
/root/reference into oracle/_ref/src/ (git-ignored like the compiled reference extensions next to it, NOT
gpurun-ignored, so the copy travels to the GPU box and never enters the repository's history).  Nothing is edited;
`manifest.json` records the sha256 of every staged file so a test can prove the copy is the reference's.

  python -m oracle.stage_ref            (build() in __graft_entry__.py calls stage() when /root/reference is present)

`activate()` puts oracle/_ref/src and compat/ on sys.path (plus tiny stand-ins for third-party imports that are absent
from this image: easydict, tensorboardX, SharedArray, open3d, plyfile -- SURVEY.md Appendix C.1) and returns the staged
root, or None when nothing is staged.  Only tests/ and bench.py's `--model reference` switch call it.
"""
import hashlib
import json
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
DST = os.path.join(HERE, "_ref", "src")
REF = "/root/reference"

FILES = [
    "model/__init__.py", "model/unet.py", "model/unet_block.py", "model/dsnorm.py",
    "lib/pointgroup_ops/functions/pointgroup_ops.py",
    "lib/pointops2/__init__.py", "lib/pointops2/functions/__init__.py", "lib/pointops2/functions/pointops2.py",
    "util/__init__.py", "util/loss_utils.py", "util/lovasz_loss.py", "util/model_utils.py", "util/common_utils.py",
    "util/config.py", "util/pseudo_labels_util.py",
    "cfgs/da_front3d_scannet/spconv.yaml", "cfgs/dataset_cfgs/scannet/scannet_cfg.yaml",
    "cfgs/dataset_cfgs/front3d/front3d_cfg.yaml",
    "dataset/augmentor/augmentor_utils.py",   # elastic / crop / scene_aug: the f3 row's parity target
]


def _sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def stage(force=False):
    """copy FILES from /root/reference (if present) -> oracle/_ref/src/; returns the staged root or None"""
    if not os.path.isdir(REF):
        return DST if os.path.exists(os.path.join(DST, "manifest.json")) else None
    man = {}
    for rel in FILES:
        src = os.path.join(REF, rel)
        if not os.path.exists(src):
            continue
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if force or not os.path.exists(dst) or _sha(dst) != _sha(src):
            shutil.copyfile(src, dst)
        man[rel] = _sha(src)
    # namespace directories the reference imports through (`lib.pointgroup_ops.functions`): plain dirs are enough
    with open(os.path.join(DST, "manifest.json"), "w") as f:
        json.dump({"source": "CVMI-Lab/DODA @ /root/reference, copied unmodified", "sha256": man}, f, indent=1)
    return DST


def staged_root():
    return DST if os.path.exists(os.path.join(DST, "manifest.json")) else None


def verify():
    """every staged file still has the sha256 recorded when it was copied from /root/reference"""
    root = staged_root()
    if root is None:
        return False
    man = json.load(open(os.path.join(root, "manifest.json")))["sha256"]
    return all(os.path.exists(os.path.join(root, r)) and _sha(os.path.join(root, r)) == h for r, h in man.items())


class _EasyDict(dict):
    """what the reference uses of `easydict.EasyDict` (util/config.py:9): attribute access on nested dicts"""

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, _EasyDict):
            v = _EasyDict(v)
        super().__setitem__(k, v)

    __setattr__ = __setitem__

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


def _stub_third_party():
    if "easydict" not in sys.modules:
        try:
            import easydict  # noqa: F401
        except ImportError:
            m = types.ModuleType("easydict")
            m.EasyDict = _EasyDict
            sys.modules["easydict"] = m
    for name in ("open3d", "SharedArray", "plyfile"):
        if name not in sys.modules:
            try:
                __import__(name)
            except ImportError:
                sys.modules[name] = types.ModuleType(name)
    if "tensorboardX" not in sys.modules:
        try:
            import tensorboardX  # noqa: F401
        except ImportError:
            m = types.ModuleType("tensorboardX")

            class SummaryWriter(object):
                def __init__(self, *a, **k):
                    pass

                def add_scalar(self, *a, **k):
                    pass

                def close(self):
                    pass

            m.SummaryWriter = SummaryWriter
            sys.modules["tensorboardX"] = m


def activate():
    """-> staged root (on sys.path together with compat/), or None.  /root/reference itself is used when it is present
    and nothing has been staged (this container before build())."""
    root = staged_root()
    if root is None and os.path.isdir(REF):
        root = stage()
    if root is None:
        return None
    compat = os.path.join(ROOT, "compat")
    for p in (root, compat):
        if p not in sys.path:
            sys.path.insert(0, p)
    _stub_third_party()
    return root


def make_cfg(mid_channel=16, n_classes=11, block_residual=True, voxel_mode=4, use_xyz=False, loss="cross_entropy"):
    """the keys of cfgs/da_front3d_scannet/spconv.yaml + dataset cfg that model/unet.py reads (unet.py:19-27,88-91,
    105-114), as the attribute-dict the reference's util/config.py builds"""
    return _EasyDict({
        "MODEL": {"BACKBONE": {"in_channel": 6 if use_xyz else 3, "mid_channel": mid_channel, "block_reps": 2,
                               "block_residual": block_residual, "use_xyz": use_xyz}},
        "COMMON_CLASSES": {"n_classes": n_classes},
        "DATA_CONFIG": {"DATA_CLASS": {"n_classes": n_classes, "ignore_label": 255},
                        "DATA_PROCESSOR": {"voxel_mode": voxel_mode}},
        "OPTIMIZATION": {"loss": loss},
    })


def load_augmentor_utils():
    """the staged dataset/augmentor/augmentor_utils.py as a module, loaded from its file (the `dataset` package's own
    __init__ pulls the whole data pipeline in); None when nothing is staged.  cv2 / open3d are stubbed when absent."""
    import importlib.util
    root = activate()
    if root is None:
        return None
    path = os.path.join(root, "dataset", "augmentor", "augmentor_utils.py")
    if not os.path.exists(path):
        return None
    if "cv2" not in sys.modules:
        try:
            import cv2  # noqa: F401
        except ImportError:
            sys.modules["cv2"] = types.ModuleType("cv2")
    spec = importlib.util.spec_from_file_location("_ref_augmentor_utils", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(stage(force="--force" in sys.argv), "verified" if verify() else "NOT verified")
