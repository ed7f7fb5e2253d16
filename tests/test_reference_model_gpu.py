"""The drop-in boundary, proven at run time: the reference's OWN model files (model/unet.py, model/unet_block.py,
model/dsnorm.py, lib/pointgroup_ops/functions/pointgroup_ops.py, lib/pointops2/functions/pointops2.py,
util/model_utils.py, util/common_utils.py) run UNCHANGED on a B200 through compat/ (spconv, PG_OP, pointops2_cuda)
and give the oracle's results.

The files are the byte-for-byte copies staged under oracle/_ref/src/ by oracle/stage_ref.py (sha256 manifest checked
here); /root/reference itself is never read at run time."""
import os
import sys

import numpy as np
import pytest
import torch

from helpers import (rel_err, oracle_step, grad_report, assert_grad_parity, capture_relu_masks, pinned_grad_report,
                     assert_pinned_grad_parity)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref(cuda_dev):
    from oracle import stage_ref
    root = stage_ref.activate()
    if root is None:
        pytest.skip("reference model files are not staged (python -m oracle.stage_ref needs /root/reference)")
    assert stage_ref.verify(), "oracle/_ref/src differs from the manifest written when it was copied"
    import model.unet as mu
    import model.unet_block as mb
    import spconv
    assert os.path.abspath(mu.__file__).startswith(os.path.abspath(root)), mu.__file__
    assert os.path.abspath(mb.__file__).startswith(os.path.abspath(root)), mb.__file__
    assert spconv.__name__ == "doda_b200.spconv"
    return stage_ref


def _batch(target, seeds=(0, 1)):
    from doda_b200 import scenes
    return scenes.collate([scenes.scene_with_voxels(s, target) for s in seeds], dup_max=2)


def test_reference_model_fn_runs_unchanged_and_matches_oracle_and_mirror(ref, cuda_dev):
    """ref: model/unet.py:58-99 (SparseConvNet.forward, test_model_feat), 154-198 (model_fn), model/unet_block.py:32-38,
    87-100 -- executed from the staged files, on the engine, forward + backward, 2 x 20 k voxels."""
    from model.unet import SparseConvNet as RefNet, model_fn_decorator
    from doda_b200.unet import SparseConvNet as Mirror, model_step
    cfg = ref.make_cfg(mid_channel=16)
    batch = _batch(20000)
    torch.manual_seed(0)
    net = RefNet(cfg)
    sd0 = {k: v.detach().clone() for k, v in net.state_dict().items()}
    loss64, scores64, sd64 = oracle_step(sd0, batch, torch.float64)
    _, _, sd32 = oracle_step(sd0, batch, torch.float32)
    net = net.to(cuda_dev).train()
    model_fn = model_fn_decorator(cfg, 2)
    masks = capture_relu_masks(net)
    ret = model_fn(batch, net, 0)
    ret["loss"].backward()
    torch.cuda.synchronize()
    assert len(masks) == 65
    assert set(ret) >= {"loss", "output", "preds", "labels"}
    # vs the fp64 oracle: the boundary's tolerance for activations (north_star: 1e-4 rel)
    e_scores = rel_err(ret["output"], scores64)
    assert e_scores <= 1e-4, e_scores
    assert abs(float(ret["loss"]) - float(loss64)) <= 1e-4 * max(1.0, abs(float(loss64)))
    assert torch.equal(ret["preds"].cpu(), ret["output"].max(1)[1].cpu())
    rep = grad_report([(n, p.grad) for n, p in net.named_parameters()], sd64, sd32)
    print("reference model on the engine, grads vs fp64 oracle:", rep)
    assert_grad_parity(rep, "reference model")
    _, _, sd64p = oracle_step(sd0, batch, torch.float64, relu_masks=masks)
    prep = pinned_grad_report([(n, p.grad) for n, p in net.named_parameters()], sd64p)
    print("reference model on the engine, grads vs fp64 oracle with pinned gates:", prep)
    assert_pinned_grad_parity(prep, "reference model")
    assert rel_err(net.linear.weight.grad, sd64["linear.weight"].grad) <= 1e-4
    # vs the mirror (doda_b200/unet.py): same weights, same batch -> same activations (the mirror only swaps in the
    # engine's devoxelize gather and cross-entropy, which do not change the forward values)
    mirror = Mirror(mid_channel=16)
    mirror.load_state_dict(sd0)
    mirror = mirror.to(cuda_dev).train()
    loss_m, scores_m = model_step(mirror, batch, device=cuda_dev)
    loss_m.backward()
    assert rel_err(ret["output"], scores_m) <= 1e-6, rel_err(ret["output"], scores_m)
    assert abs(float(loss_m) - float(ret["loss"])) <= 1e-6
    for k, v in net.state_dict().items():  # BatchNorm running statistics took the same update
        if "running_" in k:
            assert rel_err(v, mirror.state_dict()[k]) <= 1e-6, k
    g_ref = dict((n, p.grad) for n, p in net.named_parameters())
    worst = max(rel_err(p.grad, g_ref[n]) for n, p in mirror.named_parameters())
    rep_m = grad_report([(n, p.grad) for n, p in mirror.named_parameters()], sd64, sd32)
    print("mirror grads vs fp64 oracle:", rep_m, "worst mirror-vs-reference-model grad diff:", worst)
    assert_grad_parity(rep_m, "mirror")


def test_reference_vggblock_net_matches_oracle(ref, cuda_dev):
    """block_residual: False -> ref: model/unet_block.py:41-52 VGGBlock (unused by the shipped cfgs, built anyway)"""
    from model.unet import SparseConvNet as RefNet, model_fn_decorator
    from doda_b200.unet import SparseConvNet as Mirror, model_step
    cfg = ref.make_cfg(mid_channel=16, block_residual=False)
    batch = _batch(12000, seeds=(5, 6))
    torch.manual_seed(1)
    net = RefNet(cfg)
    sd0 = {k: v.detach().clone() for k, v in net.state_dict().items()}
    assert any(".conv_layers.2.weight" in k for k in sd0)
    loss64, scores64, sd64 = oracle_step(sd0, batch, torch.float64)
    _, _, sd32 = oracle_step(sd0, batch, torch.float32)
    net = net.to(cuda_dev).train()
    masks = capture_relu_masks(net)
    ret = model_fn_decorator(cfg, 2)(batch, net, 0)
    ret["loss"].backward()
    assert rel_err(ret["output"], scores64) <= 1e-4
    rep = grad_report([(n, p.grad) for n, p in net.named_parameters()], sd64, sd32)
    print("VGG net grads:", rep)
    assert_grad_parity(rep, "vgg")
    _, _, sd64p = oracle_step(sd0, batch, torch.float64, relu_masks=masks)
    prep = pinned_grad_report([(n, p.grad) for n, p in net.named_parameters()], sd64p)
    print("VGG net grads vs fp64 oracle with pinned gates:", prep)
    assert_pinned_grad_parity(prep, "vgg")
    mirror = Mirror(mid_channel=16, block_residual=False)
    mirror.load_state_dict(sd0)
    mirror = mirror.to(cuda_dev).train()
    _, scores_m = model_step(mirror, batch, device=cuda_dev)
    assert rel_err(scores_m, scores64) <= 1e-4


def test_reference_test_model_fn_pseudo_labels_and_knn_broadcast(ref, cuda_dev):
    """ref: model/unet.py:115-152 test_model_fn: softmax, per-class confidence threshold, pseudo labels, and the
    crop branch's pointops.knnquery(1, ...) label broadcast (lib/pointops2/functions/pointops2.py:54-69) on the
    engine's pointops2_cuda.  Checked against a plain torch restatement on the same logits."""
    from model.unet import SparseConvNet as RefNet, model_fn_decorator
    cfg = ref.make_cfg(mid_channel=16)
    batch = _batch(6000, seeds=(2, 3))
    torch.manual_seed(2)
    net = RefNet(cfg).to(cuda_dev).eval()
    test_fn = model_fn_decorator(cfg, 2, test=True)
    thres = [0.05 + 0.01 * i for i in range(11)]
    with torch.no_grad():
        r = test_fn(batch, net, 0, thres=thres)
    out = r["output"].double().cpu()
    sm = torch.softmax(out, 1)
    conf, lab = sm.max(1)
    mask = conf > torch.tensor(thres, dtype=torch.float64)[lab]
    pl = lab.clone()
    pl[~mask] = 255
    assert torch.equal(r["pseudo_labels"].cpu(), pl)
    w = torch.zeros_like(conf)
    w[mask] = conf[mask]
    assert rel_err(r["weight"], w) <= 1e-5
    # crop branch: the batch holds a crop (every second point); labels are broadcast back to all points by 1-NN
    N = batch["locs_float"].shape[0]
    off = batch["offsets"]
    keep = torch.zeros(N, dtype=torch.bool)
    keep[::2] = True
    crop_off = torch.tensor([0] + [int(keep[:int(o)].sum()) for o in off[1:]], dtype=torch.int32)
    b2 = dict(batch)
    b2["locs_float_all"], b2["offsets_all"], b2["labels_all"] = batch["locs_float"], off, batch["labels"]
    from doda_b200 import scenes as _sc  # re-collate the cropped points so that p2v / v2p stay consistent
    from doda_b200 import pointgroup_ops as _pg
    locs_c = batch["locs"][keep].contiguous()
    vl, p2v, v2p = _pg.voxelization_idx(locs_c, 2, 4)
    b2.update(locs=locs_c, voxel_locs=vl, p2v_map=p2v, v2p_map=v2p, locs_float=batch["locs_float"][keep].contiguous(),
              feats=batch["feats"][keep].contiguous(), labels=batch["labels"][keep].contiguous(), offsets=crop_off)
    with torch.no_grad():
        r2 = test_fn(b2, net, 0, thres=0.0, with_crop=True)
    assert r2["output"].shape[0] == N and r2["labels"].shape[0] == N
    # 1-NN restatement (brute force, per scene)
    xf, xa = b2["locs_float"].double(), b2["locs_float_all"].double()
    nn_idx = torch.empty(N, dtype=torch.int64)
    for b in range(2):
        s0, s1, a0, a1 = int(crop_off[b]), int(crop_off[b + 1]), int(off[b]), int(off[b + 1])
        for c0 in range(a0, a1, 4096):
            d = torch.cdist(xa[c0:min(c0 + 4096, a1)], xf[s0:s1])
            nn_idx[c0:min(c0 + 4096, a1)] = d.argmin(1) + s0
    with torch.no_grad():
        r1 = test_fn(dict(b2, offsets_all=crop_off), net, 0, thres=0.0, with_crop=True)  # no broadcast: offsets equal
    d_same = (xa - xf[nn_idx]).norm(dim=1)
    got = r2["preds"].cpu()
    exp = r1["preds"].cpu()[nn_idx]
    # ties between equidistant neighbours may resolve differently: compare where the nearest neighbour is unique
    agree = (got == exp).double().mean()
    assert float(agree) >= 0.999, float(agree)
    assert float(d_same.max()) < 8.0  # every kept point is within a few voxels of a dropped one


def test_reference_dsnorm_convert_and_domain_statistics(ref, cuda_dev):
    """ref: model/dsnorm.py:63-84 (DSNorm.forward), 178-214 (convert_dsnorm), 335-344 (set_ds_source / set_ds_target):
    the engine's fused BN kernels are routed by duck typing; compare with the reference module's OWN forward
    (F.batch_norm), same weights, same inputs, source and target passes, running statistics per domain."""
    import copy
    from model.dsnorm import DSNorm, set_ds_source, set_ds_target
    from model.unet import SparseConvNet as RefNet, model_fn_decorator
    from doda_b200.spconv import modules as spm
    cfg = ref.make_cfg(mid_channel=16)
    torch.manual_seed(3)
    net = DSNorm.convert_dsnorm(RefNet(cfg))
    n_ds = sum(1 for m in net.modules() if m.__class__.__name__ == "DSNorm")
    assert n_ds == 65
    net = net.to(cuda_dev).train()
    twin = copy.deepcopy(net)
    model_fn = model_fn_decorator(cfg, 2)
    b_src, b_tgt = _batch(5000, seeds=(10, 11)), _batch(5000, seeds=(12, 13))
    outs = {}
    for name, model, fused in (("engine", net, True), ("torch", twin, False)):
        spm.fuse_bn = fused  # False: SparseSequential calls the reference module's own forward
        try:
            model.apply(set_ds_source)
            r_s = model_fn(b_src, model, 0)
            r_s["loss"].backward()
            model.apply(set_ds_target)
            r_t = model_fn(b_tgt, model, 0)
            r_t["loss"].backward()
        finally:
            spm.fuse_bn = True
        outs[name] = (r_s["output"].detach(), r_t["output"].detach())
    assert rel_err(outs["engine"][0], outs["torch"][0]) <= 1e-4
    assert rel_err(outs["engine"][1], outs["torch"][1]) <= 1e-4
    sa, sb = net.state_dict(), twin.state_dict()
    n_src = n_tgt = 0
    for k in sa:
        if "running_mean" in k or "running_var" in k:
            assert rel_err(sa[k], sb[k]) <= 1e-4, k
            n_src += "_source" in k
            n_tgt += "_target" in k
        if "num_batches_tracked" in k:
            assert int(sa[k]) == int(sb[k]), (k, int(sa[k]), int(sb[k]))
    assert n_src == 130 and n_tgt == 130
    # source and target statistics really are separate buffers after a state_dict round trip (convert_dsnorm aliases
    # them until the first load, model/dsnorm.py:205-206)
    # two fp32 evaluations of a deep ReLU net: gate flips bound how close their gradients can be (helpers.py)
    g1 = [rel_err(p.grad, q.grad) for p, q in zip(net.parameters(), twin.parameters())]
    assert float(np.median(g1)) <= 2e-2, float(np.median(g1))


def test_reference_checkpoint_round_trip(ref, cuda_dev, tmp_path):
    """ref: util/model_utils.py:20-94: save_params (state_dict to CPU, optimizer state) -> load_params_from_ckpt with
    `module.` prefixes stripped, into a freshly built model on the engine; the loaded model reproduces the saved
    model's outputs exactly and the optimizer state comes back."""
    import types
    import util.common_utils as cu
    import util.model_utils as mu
    from model.unet import SparseConvNet as RefNet, model_fn_decorator
    cfg = ref.make_cfg(mid_channel=16)
    torch.manual_seed(4)
    net = RefNet(cfg).to(cuda_dev).train()
    opt = torch.optim.SGD(net.parameters(), lr=0.005, momentum=0.9, weight_decay=1e-4)
    model_fn = model_fn_decorator(cfg, 2)
    batch = _batch(4000, seeds=(20, 21))
    r = model_fn(batch, net, 0)
    r["loss"].backward()
    opt.step()
    mu.get_git_commit_id = lambda: "test"  # the reference shells out to `git rev-parse` (util/common_utils.py)
    path = str(tmp_path / "train_epoch_1.pth")
    mu.save_params(path, net, opt, 1, metric=0.5)
    # a DDP-style checkpoint: keys carry the `module.` prefix (update_checkpoint strips it)
    ck = torch.load(path, weights_only=False)
    ck["state_dict"] = type(ck["state_dict"])(("module." + k, v) for k, v in ck["state_dict"].items())
    path2 = str(tmp_path / "ddp.pth")
    torch.save(ck, path2)
    real_load = torch.load
    torch.load = lambda *a, **k: real_load(*a, **dict(k, weights_only=False))
    try:
        torch.manual_seed(99)
        net2 = RefNet(cfg).to(cuda_dev).train()
        opt2 = torch.optim.SGD(net2.parameters(), lr=0.005, momentum=0.9, weight_decay=1e-4)
        net2, opt2, epoch = mu.load_params_from_ckpt(path2, True, net2, opt2)
        metric, _ = mu.load_metric_from_ckpt(path2, True)
    finally:
        torch.load = real_load
    assert epoch == 1 and metric == 0.5
    for (k, a), (_, b) in zip(net.state_dict().items(), net2.state_dict().items()):
        assert torch.equal(a.cpu(), b.cpu()), k
    st1, st2 = opt.state_dict()["state"], opt2.state_dict()["state"]
    assert len(st1) == len(st2) > 0
    for i in st1:
        assert torch.equal(st1[i]["momentum_buffer"].cpu(), st2[i]["momentum_buffer"].cpu())
    net.eval(), net2.eval()
    test_fn = model_fn_decorator(cfg, 2, test=True)
    with torch.no_grad():
        o1 = test_fn(batch, net, 0)["output"]
        o2 = test_fn(batch, net2, 0)["output"]
    assert rel_err(o2, o1) <= 1e-6
    # the same checkpoint loads into the engine's own mirror (identical key names and shapes)
    from doda_b200.unet import SparseConvNet as Mirror
    m = Mirror(mid_channel=16)
    m.load_state_dict(mu.update_checkpoint(torch.load(path2, weights_only=False))["state_dict"])


def test_reference_update_meter_matches_engine_metrics(ref, cuda_dev):
    """ref: util/common_utils.py:233-256 (intersectionAndUnionGPU + update_meter) run from the staged file against
    doda_b200.metrics (one device pass, one packed all-reduce, one host read)"""
    import util.common_utils as cu
    from doda_b200 import metrics
    rng = np.random.RandomState(0)
    K = 11
    pred = torch.from_numpy(rng.randint(0, K, size=50000)).to(cuda_dev)
    lab = torch.from_numpy(rng.randint(0, K, size=50000))
    lab[rng.rand(50000) < 0.05] = 255
    lab = lab.to(cuda_dev)
    im, um, tm = cu.AverageMeter(), cu.AverageMeter(), cu.AverageMeter()
    r = cu.update_meter(im, um, tm, pred, lab, K, 255, False)
    im2, um2, tm2 = cu.AverageMeter(), cu.AverageMeter(), cu.AverageMeter()
    r2 = metrics.update_meter(im2, um2, tm2, pred, lab, K, 255, False)
    assert abs(r[3] - r2[3]) <= 1e-9
    for a, b in zip(r[4:], r2[4:]):
        assert np.array_equal(np.asarray(a), np.asarray(b))
    assert np.array_equal(im.sum, im2.sum) and np.array_equal(um.sum, um2.sum) and np.array_equal(tm.sum, tm2.sum)
    m1, m2 = cu.calc_metrics(im, um, tm), cu.calc_metrics(im2, um2, tm2)
    assert abs(m1[0] - m2[0]) <= 1e-12 and abs(m1[2] - m2[2]) <= 1e-12


def test_reference_model_with_attached_tape_is_the_same_computation(ref, cuda_dev):
    """doda_b200.tape.attach(model) on the reference's OWN classes (optional one-liner of INTEGRATION.md): the U-Net
    runs as one autograd node, and loss / scores / BatchNorm running statistics / BatchNorm gradients are bit-identical
    to the unchanged module-by-module execution, conv weight gradients (float atomics on both sides) within 1e-5"""
    from model.unet import SparseConvNet as RefNet, model_fn_decorator
    from doda_b200 import tape
    cfg = ref.make_cfg(mid_channel=16)
    batch = _batch(8000)
    torch.manual_seed(3)
    net = RefNet(cfg).to(cuda_dev).train()
    sd0 = {k: v.detach().clone() for k, v in net.state_dict().items()}
    model_fn = model_fn_decorator(cfg, 2)
    out = {}
    for attached in (False, True):
        net.load_state_dict(sd0)
        for p in net.parameters():
            p.grad = None
        if attached:
            assert tape.attach(net) >= 1
        runs0 = tape.runs
        ret = model_fn(batch, net, 0)
        ret["loss"].backward()
        assert (tape.runs - runs0 == 1) == attached
        out[attached] = (ret["loss"].detach().clone(), ret["output"].detach().clone(),
                         {n: p.grad.detach().clone() for n, p in net.named_parameters()},
                         {n: b.detach().clone() for n, b in net.named_buffers()})
    assert torch.equal(out[True][0], out[False][0]) and torch.equal(out[True][1], out[False][1])
    for n, b in out[False][3].items():
        assert torch.equal(b, out[True][3][n]), n
    for n, g in out[False][2].items():
        if g.dim() >= 3:
            assert rel_err(out[True][2][n], g) <= 1e-5, n
        else:
            assert torch.equal(out[True][2][n], g), n


def test_attached_tape_follows_convert_dsnorm_and_domain_switches(ref, cuda_dev):
    """tape.attach BEFORE convert_dsnorm: the cached plan must notice that every BatchNorm module was swapped for a
    DSNorm (model/dsnorm.py:178-214) and the taped node must use the domain's own running statistics
    (set_ds_source / set_ds_target, 335-344): bit-identical to the module-by-module execution of the same DSNorm net,
    source and target passes"""
    import copy
    from model.dsnorm import DSNorm, set_ds_source, set_ds_target
    from model.unet import SparseConvNet as RefNet, model_fn_decorator
    from doda_b200 import tape
    cfg = ref.make_cfg(mid_channel=16)
    torch.manual_seed(4)
    base = RefNet(cfg)
    assert tape.attach(base) >= 1          # attached while the net still holds nn.BatchNorm1d modules
    tape.cached_plan(base.unet)            # ... and a plan referring to them
    net = DSNorm.convert_dsnorm(base).to(cuda_dev).train()
    assert all(m.__class__.__name__ == "DSNorm" for m in tape.cached_plan(net.unet)["_bns"])
    plain = copy.deepcopy(net)
    for m in plain.modules():              # the twin runs module by module
        m.__dict__.pop("forward", None)
    model_fn = model_fn_decorator(cfg, 2)
    b_src, b_tgt = _batch(4000, seeds=(20, 21)), _batch(4000, seeds=(22, 23))
    outs = {}
    for name, model, taped in (("taped", net, True), ("modules", plain, False)):
        runs0 = tape.runs
        model.apply(set_ds_source)
        r_s = model_fn(b_src, model, 0)
        r_s["loss"].backward()
        model.apply(set_ds_target)
        r_t = model_fn(b_tgt, model, 0)
        r_t["loss"].backward()
        assert (tape.runs - runs0 == 2) == taped
        outs[name] = (r_s["output"].detach().clone(), r_t["output"].detach().clone(),
                      {k: v.detach().clone() for k, v in model.state_dict().items()},
                      {n: p.grad.detach().clone() for n, p in model.named_parameters()})
    assert torch.equal(outs["taped"][0], outs["modules"][0]) and torch.equal(outs["taped"][1], outs["modules"][1])
    for k, v in outs["modules"][2].items():
        assert torch.equal(v, outs["taped"][2][k]), k
    for n, g in outs["modules"][3].items():
        if g.dim() >= 3:
            assert rel_err(outs["taped"][3][n], g) <= 1e-5, n
        else:
            assert torch.equal(outs["taped"][3][n], g), n
