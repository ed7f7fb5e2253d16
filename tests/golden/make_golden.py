"""Generate the golden fixtures under tests/golden/ by running the REFERENCE's own code in this container:
  * the reference PG_OP extension compiled unmodified from /root/reference/lib/pointgroup_ops (oracle/build_ref.py)
    for the CPU entry points voxelize_idx (voxelize.cpp:11-155) and bfs_cluster (bfs_cluster.cpp:28-111);
  * the reference model/unet.py + model/unet_block.py imported from /root/reference on top of the compat/ shims
    (spconv, PG_OP, pointops2_cuda), for the state_dict key names / shapes a checkpoint must match;
  * the reference's metric epilogue `intersectionAndUnionGPU` (util/common_utils.py:233-247), imported from the
    reference file itself (open3d / SharedArray stubbed, `.cuda()` made a no-op: its histc path runs on CPU).
Run:  python tests/golden/make_golden.py [--only metrics]   (needs /root/reference; the fixtures are committed)
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import build_ref  # noqa: E402


def voxelize_cases():
    rng = np.random.RandomState(0)
    cases = {}
    # the 6-point example of SURVEY.md §8c
    cases["six"] = (np.array([[0, 1, 1, 1], [0, 1, 1, 1], [0, 2, 2, 2], [1, 1, 1, 1], [0, 2, 2, 2], [0, 1, 1, 1]],
                             dtype=np.int64), 2, 4)
    for mode in (1, 2, 3, 4):
        pts = rng.randint(0, 9, size=(500, 3))
        b = np.sort(rng.randint(0, 3, size=(500, 1)), axis=0)
        cases["rand4_m%d" % mode] = (np.concatenate([b, pts], 1).astype(np.int64), 3, mode)
    # NB: 3-column input is undefined behaviour in the reference (voxelize_outputmap strides rows by dimension+1,
    # voxelize.cpp:44-49, and overruns the [M,3] output) -> not a usable golden case; DODA always passes 4 columns.
    uniq = np.stack(np.unravel_index(rng.choice(1000, 200, replace=False), (10, 10, 10)), 1)
    cases["uniq_m0"] = (np.concatenate([np.zeros((200, 1), dtype=np.int64), uniq], 1).astype(np.int64), 1, 0)
    cases["big_m4"] = (np.concatenate([np.sort(rng.randint(0, 2, size=(20000, 1)), axis=0),
                                       rng.randint(0, 40, size=(20000, 3))], 1).astype(np.int64), 2, 4)
    return cases


def metric_fixtures():
    """inputs + outputs of the reference's intersectionAndUnionGPU, run here from /root/reference/util/common_utils.py"""
    import importlib.util
    for name in ("open3d", "SharedArray"):
        sys.modules.setdefault(name, types.ModuleType(name))
    spec = importlib.util.spec_from_file_location("ref_common_utils", "/root/reference/util/common_utils.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    real_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    out = {}
    try:
        rng = np.random.RandomState(5)
        for i, (K, shape) in enumerate([(11, (20000,)), (13, (200, 30)), (20, (6, 8, 10)), (5, (64,))]):
            pred = rng.randint(0, K + 2, size=shape)           # K, K+1: outside the histogram range
            lab = rng.randint(0, K, size=shape)
            same = rng.rand(*shape) < 0.5
            pred = np.where(same, lab, pred)
            lab = np.where(rng.rand(*shape) < 0.1, 255, lab)
            ai, au, at = ref.intersectionAndUnionGPU(torch.from_numpy(pred), torch.from_numpy(lab), K, 255)
            out["c%d/pred" % i], out["c%d/label" % i] = pred.astype(np.int64), lab.astype(np.int64)
            out["c%d/K" % i] = np.array([K], dtype=np.int64)
            out["c%d/intersection" % i], out["c%d/union" % i], out["c%d/target" % i] = ai.numpy(), au.numpy(), at.numpy()
    finally:
        torch.Tensor.cuda = real_cuda
    np.savez_compressed(os.path.join(HERE, "iou_metrics.npz"), **out)


def main():
    if "--only" in sys.argv and sys.argv[sys.argv.index("--only") + 1] == "metrics":
        metric_fixtures()
        print("iou_metrics.npz written to", HERE)
        return
    metric_fixtures()
    build_ref.build_all()
    ref = build_ref.load("PG_OP")
    out = {}
    for name, (coords, bs, mode) in voxelize_cases().items():
        c = torch.from_numpy(coords).contiguous()
        oc = c.new_empty(0)
        imap = torch.zeros(c.shape[0], dtype=torch.int32)
        omap = torch.zeros(0, dtype=torch.int32)
        ref.voxelize_idx(c, oc, imap, omap, bs, mode)
        out[name + "/coords"] = coords
        out[name + "/args"] = np.array([bs, mode], dtype=np.int64)
        out[name + "/out_coords"] = oc.numpy().copy()
        out[name + "/input_map"] = imap.numpy().copy()
        out[name + "/output_map"] = omap.numpy().copy()
    np.savez_compressed(os.path.join(HERE, "voxelize_idx.npz"), **out)

    # bfs_cluster on the example adjacency of lib/pointgroup_ops/functions/pointgroup_ops.py:397-403
    idx = np.array([0, 1, 2, 3, 4, 5, 1, 11, 2, 0, 3, 6, 7, 3, 0, 4, 8, 4, 9, 5, 10, 6, 2, 12, 13, 7, 2, 8, 3, 9, 4, 5,
                    10, 1, 11, 12, 13, 6, 13, 6, 14, 13, 14, 15, 14, 15, 16, 15, 16, 17, 18, 17, 18, 19], dtype=np.int32)
    start_len = np.array([[0, 6], [6, 2], [8, 5], [13, 4], [17, 2], [19, 2], [21, 4], [25, 2], [27, 2], [29, 2], [31, 2],
                          [33, 2], [35, 3], [38, 3], [41, 3], [44, 3], [47, 2], [49, 2], [51, 2], [53, 1]], dtype=np.int32)
    sem = np.array([0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2], dtype=np.int32)
    bfs = {"idx": idx, "start_len": start_len, "sem": sem}
    for thr in (1, 2, 4):
        ci, co = torch.zeros(0, dtype=torch.int32), torch.zeros(0, dtype=torch.int32)
        ref.bfs_cluster(torch.from_numpy(sem), torch.from_numpy(idx), torch.from_numpy(start_len), ci, co, 20, thr)
        bfs["thr%d/cluster_idxs" % thr] = ci.numpy().copy()
        bfs["thr%d/cluster_offsets" % thr] = co.numpy().copy()
    np.savez_compressed(os.path.join(HERE, "bfs_cluster.npz"), **bfs)

    # state_dict layout of the reference model built on the compat shims
    sys.path.insert(0, os.path.join(ROOT, "compat"))
    sys.path.insert(0, "/root/reference")

    class _NS(types.SimpleNamespace):
        pass
    res = {}
    from model.unet import SparseConvNet as RefNet  # the reference's file, unchanged
    for m in (16, 32):
        cfg = _NS(MODEL=_NS(BACKBONE=_NS(in_channel=3, mid_channel=m, block_reps=2, block_residual=True)),
                  COMMON_CLASSES=_NS(n_classes=11))
        net = RefNet(cfg)
        res["m%d" % m] = {k: list(v.shape) for k, v in net.state_dict().items()}
    with open(os.path.join(HERE, "unet_state_dict.json"), "w") as f:
        json.dump(res, f, indent=0, sort_keys=True)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
