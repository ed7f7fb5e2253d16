"""Oracle comparison AT the BASELINE.json sizes: the sm_100a path against the fp64 CPU
oracle on configs[1] (2 x 150 k voxels, full SparseConvNet, forward + backward) and configs[3] (one 400 k-voxel scene,
stride-2 encoder-decoder).  At these sizes BatchNorm runs over >= 45 rows even at the deepest level, so the whole-net
gradient bars are FIXED numbers (no "x times the fp32 oracle's own error" rule)."""
import numpy as np
import pytest
import torch

from helpers import rel_err, surface_coords

pytestmark = pytest.mark.gpu


def test_full_size_unet_matches_oracle_2x150k(cuda_dev):
    from doda_b200 import scenes
    from doda_b200.unet import SparseConvNet, model_step
    from oracle.unet_ref import model_step_ref
    torch.manual_seed(0)
    batch = scenes.collate([scenes.scene_with_voxels(i, 150000) for i in range(2)], seed=0, dup_max=2)
    assert batch["voxel_locs"].shape[0] == 300000
    model = SparseConvNet(mid_channel=16)
    sd64 = {k: (v.detach().double().clone().requires_grad_(True) if v.is_floating_point() else v.clone())
            for k, v in model.state_dict().items()}
    b64 = dict(batch)
    b64["feats"] = batch["feats"].double()
    torch.set_num_threads(max(torch.get_num_threads(), 8))
    loss64, scores64 = model_step_ref(sd64, b64, training=True)
    loss64.backward()
    model = model.to(cuda_dev).train()
    loss, scores = model_step(model, batch, device=cuda_dev)
    loss.backward()
    e = rel_err(scores, scores64)
    print("full size 2x150k: scores rel err %.3e, loss %.7f vs %.7f" % (e, float(loss), float(loss64)))
    assert e <= 1e-4, e
    assert abs(float(loss) - float(loss64)) <= 1e-5 * max(1.0, abs(float(loss64)))
    errs, num, den = {}, 0.0, 0.0
    for name, p in model.named_parameters():
        r = sd64[name].grad
        errs[name] = rel_err(p.grad, r)
        num += float((p.grad.double().cpu() - r).pow(2).sum())
        den += float(r.pow(2).sum())
    v = np.array(list(errs.values()))
    worst = max(errs, key=errs.get)
    print("full size grads: median %.2e p90 %.2e max %.2e (%s) l2 %.2e" %
          (np.median(v), np.percentile(v, 90), v.max(), worst, (num / den) ** 0.5))
    # fixed bars (well-conditioned BN at this size)
    assert np.median(v) <= 2e-4, np.median(v)
    assert np.percentile(v, 90) <= 2e-3, np.percentile(v, 90)
    assert (num / den) ** 0.5 <= 1e-3


def test_cfg4_encoder_decoder_matches_oracle_400k(cuda_dev):
    from doda_b200 import spconv
    from oracle.unet_ref import encoder_decoder_ref
    from test_parity_gpu import _encoder_decoder
    torch.manual_seed(5)
    coords, shape = surface_coords(7, 400000, 1)
    n = coords.shape[0]
    assert n == 400000
    planes = [16 * i for i in range(1, 8)]
    net = _encoder_decoder(planes, with_bn=True)
    mods = list(net._modules.values())
    L = len(planes) - 1
    bns = [m for m in mods if isinstance(m, torch.nn.BatchNorm1d)]
    convs = [m for m in mods if hasattr(m, "indice_key")]
    with torch.no_grad():
        for b in bns:  # non-trivial affine
            b.weight.uniform_(0.5, 1.5)
            b.bias.uniform_(-0.2, 0.2)
    x = torch.randn(n, planes[0])
    wd = [c.weight.detach().double().clone().requires_grad_(True) for c in convs[:L]]
    wu_rev = [c.weight.detach().double().clone().requires_grad_(True) for c in convs[L:]]  # level L-1 .. 0
    wu = list(reversed(wu_rev))
    bd = [(b.weight.detach().double(), b.bias.detach().double()) for b in bns[:L]]
    bu = list(reversed([(b.weight.detach().double(), b.bias.detach().double()) for b in bns[L:]]))
    x64 = x.double().requires_grad_(True)
    ref = encoder_decoder_ref(wd, wu, bd, bu, x64, coords, shape, 1)
    g = torch.randn(n, planes[0])
    ref.backward(g.double())
    net = net.to(cuda_dev).train()
    xd = x.to(cuda_dev).requires_grad_(True)
    y = net(spconv.SparseConvTensor(xd, torch.from_numpy(coords).to(cuda_dev), shape, 1))
    y.features.backward(g.to(cuda_dev))
    e = rel_err(y.features, ref)
    ex = rel_err(xd.grad, x64.grad)
    print("cfg4 400k encoder-decoder: out rel err %.3e, dx rel err %.3e" % (e, ex))
    assert e <= 1e-4 and ex <= 1e-3, (e, ex)
    ew = [rel_err(c.weight.grad, w.grad) for c, w in zip(convs[:L], wd)] + \
         [rel_err(c.weight.grad, w.grad) for c, w in zip(convs[L:], wu_rev)]
    print("cfg4 dW rel errs:", ["%.1e" % v for v in ew])
    assert max(ew) <= 1e-3, ew


def test_step_is_bit_reproducible_where_claimed(cuda_dev):
    """INTEGRATION.md "Reproducibility": scores, loss, every activation gradient and the BatchNorm / linear parameter
    gradients are bit-identical run to run (no float atomics on any activation: the deep levels' split-K partial sums
    are added in a fixed order, the devoxelize backward is a segmented sum); only the conv WEIGHT gradients, whose
    per-CTA partial blocks meet through vector atomics, may differ in the last bits."""
    from doda_b200 import scenes
    from doda_b200.unet import SparseConvNet, model_step
    torch.manual_seed(0)
    batch = scenes.collate([scenes.scene_with_voxels(0, 9000), scenes.scene_with_voxels(1, 7000)], dup_max=2)
    model = SparseConvNet(mid_channel=16).to(cuda_dev).train()
    runs = []
    for _ in range(3):
        for p in model.parameters():
            p.grad = None
        loss, scores = model_step(model, batch, device=cuda_dev)
        loss.backward()
        runs.append((loss.detach().clone(), scores.detach().clone(),
                     {n: p.grad.detach().clone() for n, p in model.named_parameters()}))
    for r in runs[1:]:
        assert torch.equal(r[0], runs[0][0]) and torch.equal(r[1], runs[0][1])
        for n, g in r[2].items():
            if g.dim() == 5:  # conv weights: atomics between CTAs
                assert rel_err(g, runs[0][2][n]) <= 1e-5, n
            else:
                assert torch.equal(g, runs[0][2][n]), n
