"""CPU: pin the oracle.  (1) the spconv restatement (oracle/rulebook.py, oracle/conv.py) against an INDEPENDENT dense
formulation (torch conv3d on the densified tensor) -- spconv v1.2 itself is not in /root/reference, so
this is the strongest pin available ("parity unpinned" by reference fixtures); (2) autograd consistency of the
restated backward; (3) canonical-form helpers."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import rel_err, random_coords
from oracle.rulebook import get_indice_pairs_ref, canonicalize, conv_out_shape
from oracle.conv import indice_conv_ref, indice_conv_backward_ref


def _densify(feats, coords, batch, shape):
    d = torch.zeros((batch, feats.shape[1]) + tuple(shape), dtype=feats.dtype)
    c = torch.as_tensor(coords, dtype=torch.int64)
    d[c[:, 0], :, c[:, 1], c[:, 2], c[:, 3]] = feats
    return d


def _rows(dense, coords):
    c = torch.as_tensor(coords, dtype=torch.int64)
    return dense[c[:, 0], :, c[:, 1], c[:, 2], c[:, 3]]


@pytest.mark.parametrize("shape,n", [((9, 8, 7), 150), ((6, 6, 6), 216), ((5, 4, 3), 1)])
def test_subm_matches_dense_conv3d(shape, n):
    torch.manual_seed(0)
    batch, Cin, Cout = 2, 5, 4
    coords = random_coords(0, n, batch, shape)
    feats = torch.randn(coords.shape[0], Cin, dtype=torch.float64)
    W = torch.randn(3, 3, 3, Cin, Cout, dtype=torch.float64)
    outids, pairs, pairnum, oshape = get_indice_pairs_ref(coords, batch, list(shape), 3, 1, 1, 1, subm=True)
    assert np.array_equal(outids, coords) and list(oshape) == list(shape)
    out = indice_conv_ref(feats, W, pairs, pairnum, coords.shape[0], subm=True)
    dense = F.conv3d(_densify(feats, coords, batch, shape), W.permute(4, 3, 0, 1, 2), padding=1)
    assert rel_err(out, _rows(dense, coords)) < 1e-12
    assert int(pairnum[13]) == coords.shape[0]  # the centre offset pairs every site with itself
    assert np.array_equal(pairs[0, 13, :pairnum[13]], pairs[1, 13, :pairnum[13]])


@pytest.mark.parametrize("shape", [(8, 8, 8), (9, 7, 5), (3, 2, 2)])
def test_down_k2s2_and_inverse_match_dense(shape):
    torch.manual_seed(1)
    batch, Cin, Cout = 2, 4, 6
    n = max(1, int(np.prod(shape)) // 3)
    coords = random_coords(1, n, batch, shape)
    feats = torch.randn(coords.shape[0], Cin, dtype=torch.float64)
    W = torch.randn(2, 2, 2, Cin, Cout, dtype=torch.float64)
    outids, pairs, pairnum, oshape = get_indice_pairs_ref(coords, batch, list(shape), 2, 2, 0, 1, subm=False)
    assert list(oshape) == conv_out_shape(list(shape), [2] * 3, [2] * 3, [0] * 3, [1] * 3)
    out = indice_conv_ref(feats, W, pairs, pairnum, outids.shape[0])
    dense = F.conv3d(_densify(feats, coords, batch, shape), W.permute(4, 3, 0, 1, 2), stride=2)
    assert tuple(dense.shape[2:]) == tuple(oshape)
    assert rel_err(out, _rows(dense, outids)) < 1e-12
    # active outputs = cells with >= 1 active input; inputs on a dropped odd plane have no pair
    occ = F.max_pool3d(_densify(torch.ones(coords.shape[0], 1, dtype=torch.float64), coords, batch, shape), 2, 2)
    assert int(occ.sum()) == outids.shape[0]
    kept = np.all(coords[:, 1:] // 2 < np.asarray(oshape), axis=1)
    assert int(pairnum.sum()) == int(kept.sum())
    # output rows are in ascending flattened index (A.4)
    key = ((outids[:, 0].astype(np.int64) * oshape[0] + outids[:, 1]) * oshape[1] + outids[:, 2]) * oshape[2] + outids[:, 3]
    assert np.all(np.diff(key) > 0)
    # inverse conv = conv_transpose3d restricted to the original sites, weight un-flipped
    Wi = torch.randn(2, 2, 2, Cout, Cin, dtype=torch.float64)
    up = indice_conv_ref(out, Wi, pairs, pairnum, coords.shape[0], inverse=True)
    dense_up = F.conv_transpose3d(_densify(out, outids, batch, oshape), Wi.permute(3, 4, 0, 1, 2), stride=2)
    ref = torch.zeros_like(up)
    c = torch.as_tensor(coords, dtype=torch.int64)
    k = torch.as_tensor(kept)
    ref[k] = dense_up[c[k, 0], :, c[k, 1], c[k, 2], c[k, 3]]
    assert rel_err(up, ref) < 1e-12


@pytest.mark.parametrize("ks,st,pd,dl", [(3, 2, 1, 1), (3, 1, 0, 1), ((3, 1, 2), (2, 1, 1), (1, 0, 0), 1), (3, 1, 2, 2)])
def test_generic_conv_matches_dense(ks, st, pd, dl):
    torch.manual_seed(2)
    shape, batch, Cin, Cout = (9, 8, 7), 2, 3, 2
    coords = random_coords(2, 120, batch, shape)
    feats = torch.randn(coords.shape[0], Cin, dtype=torch.float64)
    t = lambda v: tuple(v) if isinstance(v, tuple) else (v,) * 3
    W = torch.randn(*t(ks), Cin, Cout, dtype=torch.float64)
    outids, pairs, pairnum, oshape = get_indice_pairs_ref(coords, batch, list(shape), ks, st, pd, dl, subm=False)
    out = indice_conv_ref(feats, W, pairs, pairnum, outids.shape[0])
    dense = F.conv3d(_densify(feats, coords, batch, shape), W.permute(4, 3, 0, 1, 2), stride=t(st), padding=t(pd),
                     dilation=t(dl))
    assert tuple(dense.shape[2:]) == tuple(oshape)
    assert rel_err(out, _rows(dense, outids)) < 1e-12


@pytest.mark.parametrize("subm,inverse", [(True, False), (False, False), (False, True)])
def test_backward_restatement_matches_autograd(subm, inverse):
    torch.manual_seed(3)
    shape, batch = (7, 6, 6), 2
    coords = random_coords(3, 80, batch, shape)
    if subm:
        outids, pairs, pairnum, _ = get_indice_pairs_ref(coords, batch, list(shape), 3, 1, 1, 1, subm=True)
        kshape = (3, 3, 3)
    else:
        outids, pairs, pairnum, _ = get_indice_pairs_ref(coords, batch, list(shape), 2, 2, 0, 1, subm=False)
        kshape = (2, 2, 2)
    n_in, n_out = (outids.shape[0], coords.shape[0]) if inverse else (coords.shape[0], outids.shape[0])
    f = torch.randn(n_in, 4, dtype=torch.float64, requires_grad=True)
    W = torch.randn(*kshape, 4, 3, dtype=torch.float64, requires_grad=True)
    g = torch.randn(n_out, 3, dtype=torch.float64)
    out = indice_conv_ref(f, W, pairs, pairnum, n_out, inverse=inverse, subm=subm)
    out.backward(g)
    din, dW = indice_conv_backward_ref(f.detach(), W.detach(), g, pairs, pairnum, inverse, subm)
    assert rel_err(din, f.grad) < 1e-12 and rel_err(dW, W.grad) < 1e-12
    assert torch.autograd.gradcheck(lambda a, b: indice_conv_ref(a, b, pairs, pairnum, n_out, inverse, subm),
                                    (f.detach()[:, :2].clone().requires_grad_(True),
                                     W.detach()[..., :2, :2].clone().requires_grad_(True)), eps=1e-6, atol=1e-5)


def test_canonicalize_is_idempotent_and_order_free():
    shape, batch = (9, 9, 9), 2
    coords = random_coords(4, 200, batch, shape)
    outids, pairs, pairnum, oshape = get_indice_pairs_ref(coords, batch, list(shape), 2, 2, 0, 1, subm=False)
    rng = np.random.RandomState(0)
    # scramble: permute output numbering and the pair order inside each offset (what spconv's CUDA path may emit)
    perm = rng.permutation(outids.shape[0])
    inv = np.empty_like(perm); inv[perm] = np.arange(perm.shape[0])
    sp = pairs.copy()
    for k in range(pairs.shape[1]):
        n = pairnum[k]
        o = rng.permutation(n)
        sp[0, k, :n] = pairs[0, k, :n][o]
        sp[1, k, :n] = inv[pairs[1, k, :n][o]]
    o2, p2, n2 = canonicalize(outids[perm], sp, pairnum, oshape)
    assert np.array_equal(o2, outids) and np.array_equal(p2, pairs) and np.array_equal(n2, pairnum)
    o3, p3, n3 = canonicalize(o2, p2, n2, oshape)
    assert np.array_equal(o3, o2) and np.array_equal(p3, p2)


def test_empty_input():
    outids, pairs, pairnum, _ = get_indice_pairs_ref(np.zeros((0, 4), dtype=np.int32), 1, [8, 8, 8], 3, 1, 1, 1, True)
    assert pairs.shape == (2, 27, 0) and pairnum.sum() == 0
    out = indice_conv_ref(torch.zeros(0, 4), torch.zeros(3, 3, 3, 4, 2), pairs, pairnum, 0, subm=False)
    assert out.shape == (0, 2)


def test_unet_oracle_runs_and_is_deterministic():
    """whole-net oracle on a tiny scene: forward is finite, reproducible, and fp32 agrees with fp64."""
    from doda_b200 import scenes
    from doda_b200.unet import SparseConvNet
    from oracle.unet_ref import model_step_ref
    torch.manual_seed(0)
    batch = scenes.collate([scenes.scene_with_voxels(0, 600), scenes.scene_with_voxels(1, 600)], dup_max=2)
    sd = SparseConvNet(mid_channel=16).state_dict()
    l1, s1 = model_step_ref(sd, batch, training=True)
    l2, s2 = model_step_ref(sd, batch, training=True)
    assert torch.isfinite(s1).all() and torch.equal(s1, s2)
    sd64 = {k: v.double() if v.is_floating_point() else v for k, v in sd.items()}
    b64 = dict(batch); b64["feats"] = batch["feats"].double()
    l64, s64 = model_step_ref(sd64, b64, training=True)
    assert rel_err(s1, s64) < 1e-3


def test_metric_restatement_matches_reference_golden():
    """tests/golden/iou_metrics.npz holds inputs and outputs of the reference's own intersectionAndUnionGPU
    (util/common_utils.py:233-247, run from the reference file by tests/golden/make_golden.py); the restatement the
    GPU parity test compares against must reproduce them exactly"""
    import os
    import numpy as np
    import torch
    from doda_b200 import metrics
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "iou_metrics.npz"))
    n = len({k.split("/")[0] for k in z.files})
    assert n == 4
    for i in range(n):
        K = int(z["c%d/K" % i][0])
        ai, au, at = metrics.intersection_and_union_ref(torch.from_numpy(z["c%d/pred" % i]), torch.from_numpy(z["c%d/label" % i]), K, 255)
        assert np.array_equal(ai.numpy(), z["c%d/intersection" % i])
        assert np.array_equal(au.numpy(), z["c%d/union" % i]) and np.array_equal(at.numpy(), z["c%d/target" % i])


@pytest.mark.parametrize("name", ["six", "rand4_m3", "rand4_m4", "big_m4"])
def test_numpy_voxelizer_matches_reference_golden(name):
    """oracle/voxelize.py (collate step of bench.py's reference arm) against fixtures produced by the reference's own
    compiled voxelize_idx (tests/golden/make_golden.py)"""
    import os
    from oracle.voxelize import voxelize_idx_ref
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "voxelize_idx.npz"))
    bs, mode = g[name + "/args"]
    oc, im, om = voxelize_idx_ref(g[name + "/coords"], int(bs), int(mode))
    assert np.array_equal(oc.numpy(), g[name + "/out_coords"])
    assert np.array_equal(im.numpy(), g[name + "/input_map"])
    assert np.array_equal(om.numpy(), g[name + "/output_map"])


def test_gpu_native_baseline_is_the_oracles_algorithm():
    """baseline/gpu_native.py (stock torch ops, its own sort + searchsorted rulebooks) and oracle/unet_ref.py (numpy
    rulebooks) are two independent statements of spconv v1.2's native algorithm: identical in fp64 on CPU tensors,
    forward and backward, residual net"""
    import json
    import os
    from baseline.gpu_native import NativeUNet, voxelize_mean
    from oracle.voxelize import voxelize_idx_ref, init_state_dict
    from oracle.unet_ref import model_step_ref
    from doda_b200 import scenes
    shapes = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "unet_state_dict.json")))["m16"]
    sd = init_state_dict(shapes, seed=1)
    batch = scenes.collate([scenes.scene_with_voxels(0, 2500), scenes.scene_with_voxels(1, 2000)], dup_max=2,
                           voxelize=voxelize_idx_ref)
    sd64 = {k: (v.double().clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in sd.items()}
    b64 = dict(batch)
    b64["feats"] = batch["feats"].double()
    l64, s64 = model_step_ref(sd64, b64, True)
    l64.backward()
    net = NativeUNet({k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}, torch.device("cpu"))
    net.build_rulebooks(batch["voxel_locs"], batch["spatial_shape"])
    loss, scores = net.step(voxelize_mean(batch["feats"].double(), batch["v2p_map"]), batch["p2v_map"].long(),
                            batch["labels"])
    assert abs(float(loss.detach()) - float(l64.detach())) <= 1e-12
    assert float((scores.detach() - s64.detach()).abs().max()) <= 1e-10
    for k in ("input_conv.0.weight", "unet.u.u.conv.2.weight", "unet.u.deconv.2.weight", "linear.weight"):
        r = sd64[k].grad
        assert float((net.sd[k].grad - r).abs().max() / r.abs().max()) <= 1e-9, k


def test_encoder_decoder_oracle_restores_the_active_set_and_is_linear_without_bn():
    from helpers import surface_coords
    from oracle.unet_ref import encoder_decoder_ref
    coords, shape = surface_coords(7, 3000, 1)
    planes = [4, 8, 12]
    torch.manual_seed(0)
    wd = [torch.randn(2, 2, 2, planes[i], planes[i + 1], dtype=torch.float64) for i in range(2)]
    wu = [torch.randn(2, 2, 2, planes[i + 1], planes[i], dtype=torch.float64) for i in range(2)]
    x1 = torch.randn(coords.shape[0], 4, dtype=torch.float64)
    x2 = torch.randn(coords.shape[0], 4, dtype=torch.float64)
    f = lambda x: encoder_decoder_ref(wd, wu, None, None, x, coords, shape, 1)
    y = f(2 * x1 - 3 * x2)
    assert y.shape == x1.shape
    assert float((y - (2 * f(x1) - 3 * f(x2))).abs().max()) <= 1e-9 * float(y.abs().max())


# ---------------------------------------------------------------------------------------------
# augmentation row: numpy restatement vs the reference's own outputs
# ---------------------------------------------------------------------------------------------
def _aug_golden():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "augment_golden.npz"))


def test_augment_oracle_matches_reference_golden():
    """oracle/augment.py against tests/golden/augment_golden.npz (made by the reference's elastic / crop with seeded
    np.random, tests/golden/make_augment_golden.py): bit-exact, the restatement consumes the random stream identically"""
    from oracle import augment
    g = _aug_golden()
    x = g["elastic_x"]
    for i in range(2):
        gran, mag, seed = g["elastic_arg%d" % i]
        np.random.seed(int(seed))
        out = augment.elastic_ref(x, int(gran), float(mag))
        assert out.dtype == np.float64 and np.array_equal(out, g["elastic_out%d" % i])
    xyz = g["crop_xyz"]
    for i in range(3):
        f0, f1, pr, mx, seed = g["crop_arg%d" % i]
        np.random.seed(int(seed))
        xo, valid = augment.crop_ref(xyz, [int(f0), int(f1)], float(pr), int(mx))
        assert np.array_equal(valid, g["crop_valid%d" % i]) and np.array_equal(xo, g["crop_off%d" % i])
        assert valid.sum() <= mx


def test_augment_oracle_matches_staged_reference():
    """the same on fresh inputs against the staged, unmodified augmentor_utils.py when it is present"""
    import warnings
    from oracle import augment, stage_ref
    au = stage_ref.load_augmentor_utils()
    if au is None:
        pytest.skip("reference not staged")
    warnings.simplefilter("ignore")
    rng = np.random.RandomState(4)
    x = (rng.rand(4000, 3) * np.array([300, 500, 120]) - np.array([150, 250, 0])).astype(np.float32)
    for gran, mag in ((6, 40), (20, 160)):
        np.random.seed(9); ref = au.elastic(x, gran, mag)
        np.random.seed(9); mine = augment.elastic_ref(x, gran, mag)
        assert np.array_equal(ref, mine)
    xyz = (rng.rand(50000, 3) * np.array([900, 700, 150])).astype(np.float64)
    for fs, pr, mx in (([128, 512], 2e9, 20000), ([128, 512], 2e7, 40000)):
        np.random.seed(5); r0, r1 = au.crop(xyz, fs, pr, mx)
        np.random.seed(5); m0, m1 = augment.crop_ref(xyz, fs, pr, mx)
        assert np.array_equal(r1, m1) and np.array_equal(r0, m0)
