"""GPU parity: the sm_100a path (through the C ABI) against the CPU oracle on identical seeded inputs.
Integer / index results are compared bit-exactly; fp32 results with the SURVEY.md A.8 metric
max|a-b| / max|b| <= 1e-4 (north_star's tolerance) against an fp64 evaluation of the oracle."""
import numpy as np
import pytest
import torch

from helpers import rel_err, random_coords, surface_coords

pytestmark = pytest.mark.gpu

TOL = 1e-4


def _rb_oracle(coords, batch, shape, ks, st, pd, dl, subm):
    from oracle.rulebook import get_indice_pairs_ref
    return get_indice_pairs_ref(coords, batch, shape, ks, st, pd, dl, subm=subm)


def _check_rulebook(cuda_dev, coords, batch, shape, ks, st, pd, dl, subm):
    from doda_b200 import ops
    rb = ops.build_rulebook(torch.from_numpy(coords).to(cuda_dev), batch, shape, ks, st, pd, dl, subm=subm)
    outids, pairs, pairnum, oshape = _rb_oracle(coords, batch, shape, ks, st, pd, dl, subm)
    assert list(rb.out_spatial_shape) == list(oshape)
    assert np.array_equal(rb.outids.cpu().numpy(), outids)
    assert np.array_equal(rb.pairnum.cpu().numpy(), pairnum)
    assert np.array_equal(rb.pairs.cpu().numpy(), pairs)  # includes the -1 padding
    return rb, (outids, pairs, pairnum, oshape)


@pytest.mark.parametrize("n,shape,batch", [(2000, (40, 37, 29), 2), (1, (5, 5, 5), 1), (300, (7, 7, 7), 3),
                                           (20000, (128, 128, 128), 2)])
def test_rulebook_subm_exact(cuda_dev, n, shape, batch):
    coords = random_coords(1, n, batch, shape)
    rb, (_, pairs, pairnum, _) = _check_rulebook(cuda_dev, coords, batch, list(shape), 3, 1, 1, 1, True)
    # the engine's out->in table is consistent with the pairs: nbr[out, k] == in
    nbr = rb.nbr.cpu().numpy()
    K = 27
    ref = np.full_like(nbr, -1)
    for k in range(K):
        n_k = pairnum[k]
        ref[pairs[1, k, :n_k], k] = pairs[0, k, :n_k]
    assert np.array_equal(nbr, ref)


def test_rulebook_subm_surface_scene(cuda_dev):
    coords, shape = surface_coords(0, 10000, 2)
    _check_rulebook(cuda_dev, coords, 2, shape, 3, 1, 1, 1, True)


@pytest.mark.parametrize("shape", [(40, 37, 29), (33, 33, 33), (8, 9, 2)])
def test_rulebook_down_k2s2_exact(cuda_dev, shape):
    # odd extents: voxels on the last odd plane are dropped (SURVEY.md §7.2 "odd spatial shapes")
    coords = random_coords(2, min(3000, int(np.prod(shape)) // 3), 2, shape)
    _check_rulebook(cuda_dev, coords, 2, list(shape), 2, 2, 0, 1, False)


@pytest.mark.parametrize("ks,st,pd,dl", [(3, 2, 1, 1), (3, 1, 1, 1), (3, 1, 0, 1), ((3, 1, 2), (2, 1, 1), (1, 0, 0), 1),
                                         (3, 1, 2, 2)])
def test_rulebook_generic_conv_exact(cuda_dev, ks, st, pd, dl):
    shape = (21, 18, 16)
    coords = random_coords(3, 900, 2, shape)
    _check_rulebook(cuda_dev, coords, 2, list(shape), ks, st, pd, dl, False)


def test_rulebook_speculative_down_build(cuda_dev):
    """the strided builder's first half started early (when the level's SubM table is built) gives the same rulebook
    as the blocking build, whether the guessed geometry was right (k2 s2) or wrong (k3 s2 -> guess dropped)"""
    from doda_b200 import ops
    shape = [40, 37, 29]
    coords = torch.from_numpy(random_coords(5, 4000, 2, shape)).to(cuda_dev)

    def build(spec, ks, st, pd):
        ops.speculate_down = spec
        try:
            ops._spec.clear()
            ops._spec_hint[cuda_dev.index or 0] = ([2, 2, 2], [2, 2, 2], [0, 0, 0], [1, 1, 1])
            ops.build_rulebook(coords, 2, shape, 3, 1, 1, 1, subm=True)
            started = len(ops._spec) == 1
            rb = ops.build_rulebook(coords, 2, shape, ks, st, pd, 1)
            assert len(ops._spec) == 0
            return started, rb
        finally:
            ops.speculate_down = True

    for ks, st, pd in ((2, 2, 0), (3, 2, 1)):
        s1, a = build(True, ks, st, pd)
        s0, b = build(False, ks, st, pd)
        assert s1 and not s0
        for name in ("outids", "fwd", "bwd", "pairs", "pairnum"):
            assert torch.equal(getattr(a, name), getattr(b, name)), name
    torch.cuda.synchronize()


def test_rulebook_empty(cuda_dev):
    from doda_b200 import ops
    coords = torch.zeros((0, 4), dtype=torch.int32, device=cuda_dev)
    rb = ops.build_rulebook(coords, 1, [8, 8, 8], 3, 1, 1, 1, subm=True)
    assert rb.pairs.shape == (2, 27, 0) and int(rb.pairnum.sum()) == 0


def _conv_case(cuda_dev, kind, Cin, Cout, seed=0, n=1500, shape=(24, 22, 20)):
    from doda_b200 import ops
    from oracle.conv import indice_conv_ref, indice_conv_backward_ref
    torch.manual_seed(seed)
    coords = random_coords(seed, n, 2, shape)
    dev_coords = torch.from_numpy(coords).to(cuda_dev)
    if kind == "subm":
        rb = ops.build_rulebook(dev_coords, 2, list(shape), 3, 1, 1, 1, subm=True)
        outids, pairs, pairnum, _ = _rb_oracle(coords, 2, list(shape), 3, 1, 1, 1, True)
        kshape, fn, n_in, n_out, inv, subm = (3, 3, 3), ops.SubMConvFunction, coords.shape[0], coords.shape[0], False, True
    else:
        rb = ops.build_rulebook(dev_coords, 2, list(shape), 2, 2, 0, 1, subm=False)
        outids, pairs, pairnum, _ = _rb_oracle(coords, 2, list(shape), 2, 2, 0, 1, False)
        kshape = (2, 2, 2)
        if kind == "down":
            fn, n_in, n_out, inv, subm = ops.SparseConvFunction, coords.shape[0], outids.shape[0], False, False
        else:
            fn, n_in, n_out, inv, subm = ops.SparseInverseConvFunction, outids.shape[0], coords.shape[0], True, False
    feats = torch.randn(n_in, Cin)
    W = torch.randn(*kshape, Cin, Cout) * 0.2
    gout = torch.randn(n_out, Cout)
    ref = indice_conv_ref(feats.double(), W.double(), pairs, pairnum, n_out, inverse=inv, subm=subm)
    din_ref, dW_ref = indice_conv_backward_ref(feats.double(), W.double(), gout.double(), pairs, pairnum, inv, subm)
    f = feats.to(cuda_dev).requires_grad_(True)
    w = W.to(cuda_dev).requires_grad_(True)
    out = fn.apply(f, w, rb)
    out.backward(gout.to(cuda_dev))
    assert rel_err(out, ref) <= TOL, ("fwd", rel_err(out, ref))
    assert rel_err(f.grad, din_ref) <= TOL, ("dgrad", rel_err(f.grad, din_ref))
    assert rel_err(w.grad, dW_ref) <= TOL, ("wgrad", rel_err(w.grad, dW_ref))
    # raw spconv-style entry points on the same rulebook
    out2 = ops.indice_conv(f.detach(), w.detach(), rb.pairs, rb.pairnum, n_out, inverse=inv, subm=subm)
    din2, dW2 = ops.indice_conv_backward(f.detach(), w.detach(), gout.to(cuda_dev), rb.pairs, rb.pairnum, inv, subm)
    assert rel_err(out2, ref) <= TOL and rel_err(din2, din_ref) <= TOL and rel_err(dW2, dW_ref) <= TOL


@pytest.mark.parametrize("Cin,Cout", [(16, 16), (3, 16), (32, 16), (32, 32), (48, 48), (64, 32), (112, 112), (5, 7),
                                      (96, 48)])
def test_subm_conv_fwd_bwd(cuda_dev, Cin, Cout):
    _conv_case(cuda_dev, "subm", Cin, Cout)


@pytest.mark.parametrize("Cin,Cout", [(16, 32), (32, 48), (96, 112), (6, 10)])
def test_down_conv_fwd_bwd(cuda_dev, Cin, Cout):
    _conv_case(cuda_dev, "down", Cin, Cout, shape=(25, 22, 21))


@pytest.mark.parametrize("Cin,Cout", [(32, 16), (48, 32), (112, 96), (10, 6)])
def test_inverse_conv_fwd_bwd(cuda_dev, Cin, Cout):
    _conv_case(cuda_dev, "inverse", Cin, Cout, shape=(25, 22, 21))


def test_subm_conv_large_tiles(cuda_dev):
    # enough rows to take the 128-row tile path (>= 128*148 rows)
    _conv_case(cuda_dev, "subm", 16, 16, n=12000, shape=(64, 64, 32))
    _conv_case(cuda_dev, "subm", 32, 32, n=12000, shape=(64, 64, 32))
    _conv_case(cuda_dev, "subm", 64, 64, n=12000, shape=(64, 64, 32))


@pytest.mark.parametrize("kind,Cin,Cout", [("subm", 16, 16), ("subm", 32, 48), ("down", 16, 32), ("inverse", 32, 16)])
def test_conv_many_tiles_per_cta(cuda_dev, kind, Cin, Cout):
    """> 2 x 148 x 128 rows: the persistent conv kernel walks several tiles per CTA (cross-tile pipelining, both
    accumulator sets, empty tiles in the pair-grouped mode); checked against the fp64 oracle"""
    _conv_case(cuda_dev, kind, Cin, Cout, seed=7, n=45000, shape=(96, 96, 64))


@pytest.mark.parametrize("Cin,Cout", [(16, 16), (32, 16), (16, 32), (32, 32)])
def test_register_gather_kernel_matches_tcgen05_and_oracle(cuda_dev, Cin, Cout):
    """the two tensor-path kernels (conv_direct.cu for the narrow layers, conv_tc.cu for the rest) on the same
    inputs: SubM forward + dgrad (row order and mask order), strided conv and its inverse, each against the oracle"""
    from doda_b200 import ops
    from oracle.conv import indice_conv_ref
    torch.manual_seed(4)
    shape = (26, 23, 21)
    coords = random_coords(4, 2500, 2, shape)
    c = torch.from_numpy(coords).to(cuda_dev)
    n = coords.shape[0]
    rb = ops.build_rulebook(c, 2, list(shape), 3, 1, 1, 1, subm=True)
    rd = ops.build_rulebook(c, 2, list(shape), 2, 2, 0, 1)
    nd = rd.outids.shape[0]
    feat, g = torch.randn(n, Cin), torch.randn(n, Cout)
    W3, W8 = torch.randn(27, Cin, Cout) * 0.2, torch.randn(8, Cin, Cout) * 0.3
    fd, gd, W3d, W8d = feat.to(cuda_dev), g.to(cuda_dev), W3.to(cuda_dev), W8.to(cuda_dev)
    _, pairs, pairnum, _ = _rb_oracle(coords, 2, list(shape), 3, 1, 1, 1, True)
    ref_fwd = indice_conv_ref(feat.double(), W3.double(), pairs, pairnum, n, subm=True)
    Wt = torch.flip(W3, [0]).transpose(1, 2).contiguous()
    ref_dg = indice_conv_ref(g.double(), Wt.double(), pairs, pairnum, n, subm=True)
    outs = {}
    try:
        for on in (1, 0):
            ops.set_conv_direct(on)
            assert ops._direct_covers(27, Cin, Cout) == bool(on)
            outs[on] = [ops.gather_gemm(fd, W3d, rb.nbr, n),
                        ops.gather_gemm(fd, W3d, rb.nbr_perm, n, orow=rb.order, rowmask=rb.rowmask),
                        ops.gather_gemm(gd, W3d, rb.nbr_perm, n, wflags=ops.W_T_MIRROR, orow=rb.order, rowmask=rb.rowmask),
                        ops.conv_forward_raw("conv", fd, W8d, rd, None)]
            coarse = outs[on][3]
            outs[on].append(ops.conv_forward_raw("inverse", coarse, W8d.transpose(1, 2).contiguous(), rd, None))
            outs[on].append(ops.conv_backward_raw("conv", fd, W8d, torch.ones(nd, Cout, device=cuda_dev), rd, None, True, False)[0])
            # weight gradients: table form (wgrad_direct.cu) vs pair lists, all four conv kinds
            gcoarse = torch.sin(torch.arange(nd * Cout, device=cuda_dev, dtype=torch.float32)).view(nd, Cout)
            outs[on].append(ops.conv_backward_raw("subm", fd, W3d, gd, rb, None, False, True)[1].clone())
            outs[on].append(ops.conv_backward_raw("conv", fd, W8d, gcoarse, rd, None, False, True)[1].clone())
            outs[on].append(ops.conv_backward_raw("inverse", gcoarse, W8d.transpose(1, 2).contiguous(), fd, rd, None, False, True)[1].clone())
            outs[on].append(ops.conv_backward_raw("dense", fd, W3d[:1], gd, None, None, False, True)[1].clone())
    finally:
        ops.set_conv_direct(1)
    assert rel_err(outs[1][0], ref_fwd) <= TOL and rel_err(outs[1][1], ref_fwd) <= TOL and rel_err(outs[1][2], ref_dg) <= TOL
    from oracle.conv import indice_conv_backward_ref
    _, ref_dw = indice_conv_backward_ref(feat.double(), W3.double(), g.double(), pairs, pairnum, subm=True)
    assert rel_err(outs[1][6], ref_dw.view(27, Cin, Cout)) <= TOL
    assert rel_err(outs[1][9].view(Cin, Cout), feat.double().t() @ g.double()) <= TOL
    for i, (a, b) in enumerate(zip(outs[1], outs[0])):
        assert rel_err(a, b) <= TOL, i


def test_prepared_weight_images_follow_updates(cuda_dev):
    """module path: weight images are prepared ahead (one launch for all layers) and must track in-place updates"""
    from doda_b200 import spconv, ops
    from oracle.conv import indice_conv_ref
    torch.manual_seed(11)
    shape = (20, 18, 16)
    coords = random_coords(11, 900, 2, shape)
    _, pairs, pairnum, _ = _rb_oracle(coords, 2, list(shape), 3, 1, 1, 1, True)
    # 80 channels: a shape of the persistent tcgen05 kernel (the register-gather kernel of the narrow layers reads the
    # raw weights and has no image)
    conv = spconv.SubMConv3d(16, 80, 3, padding=1, bias=False, indice_key="k").to(cuda_dev)
    other = spconv.SubMConv3d(80, 16, 3, padding=1, bias=False, indice_key="k").to(cuda_dev)
    feats = torch.randn(coords.shape[0], 16)
    for it in range(3):
        x = spconv.SparseConvTensor(feats.to(cuda_dev).requires_grad_(True), torch.from_numpy(coords).to(cuda_dev),
                                    list(shape), 2)
        y = other(conv(x))
        ref = indice_conv_ref(feats.double(), conv.weight.detach().double().cpu(), pairs, pairnum, coords.shape[0], subm=True)
        ref = indice_conv_ref(ref, other.weight.detach().double().cpu(), pairs, pairnum, coords.shape[0], subm=True)
        assert rel_err(y.features, ref) <= TOL, it
        y.features.sum().backward()
        with torch.no_grad():  # an optimizer step: in-place update bumps the version, images must be rebuilt
            conv.weight.add_(0.05 * torch.randn_like(conv.weight))
            other.weight.mul_(1.1)
    assert ops.prepared_weights(conv) is not None


def _encoder_decoder(planes, with_bn):
    """BASELINE configs[3]: the k2 s2 SparseConv3d / SparseInverseConv3d encoder-decoder of UBlock
    (model/unet_block.py:67-79) without the SubM blocks: 6 down + 6 up"""
    from doda_b200 import spconv
    mods = []
    for l in range(len(planes) - 1):
        if with_bn:
            mods += [torch.nn.BatchNorm1d(planes[l], eps=1e-4, momentum=0.1), torch.nn.ReLU()]
        mods.append(spconv.SparseConv3d(planes[l], planes[l + 1], 2, stride=2, bias=False, indice_key="spconv%d" % l))
    for l in reversed(range(len(planes) - 1)):
        if with_bn:
            mods += [torch.nn.BatchNorm1d(planes[l + 1], eps=1e-4, momentum=0.1), torch.nn.ReLU()]
        mods.append(spconv.SparseInverseConv3d(planes[l + 1], planes[l], 2, bias=False, indice_key="spconv%d" % l))
    return spconv.SparseSequential(*mods)


def test_cfg4_encoder_decoder_400k(cuda_dev):
    """BASELINE configs[3] size (one 400 k-voxel scene, stride-2 encoder-decoder, 1 x B200), properties that need no CPU
    oracle: the inverse convs restore the input's active set; the net without BN is linear in x and in each layer's
    weight, so <g, y> = <dL/dx, x> = <dL/dW_l, W_l> for every layer (forward, dgrad and wgrad of the strided and the
    inverse conv are mutually adjoint, table-form and pair-list kernels alike); with BN+ReLU it runs and is finite."""
    from doda_b200 import spconv
    torch.manual_seed(5)
    coords, shape = surface_coords(7, 400000, 1)
    c = torch.from_numpy(coords).to(cuda_dev)
    n = c.shape[0]
    assert n == 400000
    planes = [16 * i for i in range(1, 8)]
    net = _encoder_decoder(planes, with_bn=False).to(cuda_dev)
    x = torch.randn(n, planes[0], device=cuda_dev, requires_grad=True)
    t = spconv.SparseConvTensor(x, c, shape, 1)
    y = net(t)
    assert y.indices is c or torch.equal(y.indices, c)
    assert y.features.shape == (n, planes[0])
    levels = [t.indice_dict["spconv%d" % l].outids.shape[0] for l in range(6)]
    assert all(a > b > 0 for a, b in zip([n] + levels, levels))  # each level strictly coarser, none empty
    g = torch.randn_like(y.features)
    s = (g * y.features).sum()
    s.backward()
    ref = float(s.detach())
    assert abs(float((x.grad * x.detach()).sum()) - ref) <= 2e-4 * abs(ref)
    for name, p in net.named_parameters():
        assert abs(float((p.grad * p.detach()).sum()) - ref) <= 2e-4 * abs(ref), name
    net_bn = _encoder_decoder(planes, with_bn=True).to(cuda_dev).train()
    xb = torch.randn(n, planes[0], device=cuda_dev, requires_grad=True)
    yb = net_bn(spconv.SparseConvTensor(xb, c, shape, 1))
    yb.features.square().mean().backward()
    assert torch.isfinite(yb.features).all() and torch.isfinite(xb.grad).all()
    assert all(torch.isfinite(p.grad).all() for p in net_bn.parameters())


def test_full_size_adjoint_and_linearity(cuda_dev):
    """BASELINE configs[1] size (2 x 150 k voxels): properties that need no CPU oracle.
    <g, conv(x)> = <dgrad(g), x> = <W, wgrad(x, g)> (the three kernels are mutually adjoint), linearity in x, and
    rulebook sanity (symmetric SubM table, centre offset = identity, order is a permutation)."""
    from doda_b200 import ops
    torch.manual_seed(3)
    coords, shape = surface_coords(0, 150000, 2)
    c = torch.from_numpy(coords).to(cuda_dev)
    n = c.shape[0]
    assert n == 300000
    rb = ops.build_rulebook(c, 2, shape, 3, 1, 1, 1, subm=True)
    nbr = rb.nbr
    assert torch.equal(nbr[:, 13], torch.arange(n, device=cuda_dev, dtype=torch.int32))
    # symmetry: j = nbr[i, k]  <=>  i = nbr[j, 26 - k]
    for k in (0, 5, 12, 20):
        i = torch.nonzero(nbr[:, k] >= 0).squeeze(1)
        j = nbr[i, k].long()
        assert torch.equal(nbr[j, 26 - k].long(), i)
    assert torch.equal(torch.sort(rb.order.long()).values, torch.arange(n, device=cuda_dev))
    assert int(rb.pairnum.sum()) == int((nbr >= 0).sum())
    for Cin, Cout in ((16, 16), (32, 16)):
        x = torch.randn(n, Cin, device=cuda_dev, requires_grad=True)
        x2 = torch.randn(n, Cin, device=cuda_dev)
        W = (torch.randn(3, 3, 3, Cin, Cout, device=cuda_dev) * 0.2).requires_grad_(True)
        g = torch.randn(n, Cout, device=cuda_dev)
        y = ops.SubMConvFunction.apply(x, W, rb)
        y.backward(g)
        a = float((g.double() * y.detach().double()).sum())
        b = float((x.grad.double() * x.detach().double()).sum())
        cc = float((W.grad.double() * W.detach().double()).sum())
        scale = float(g.double().norm() * y.detach().double().norm())
        assert abs(a - b) <= 1e-5 * scale and abs(a - cc) <= 1e-5 * scale, (a, b, cc, scale)
        with torch.no_grad():
            y2 = ops.SubMConvFunction.apply(x2, W, rb)
            y12 = ops.SubMConvFunction.apply(2.0 * x.detach() - 3.0 * x2, W, rb)
        assert rel_err(y12, 2.0 * y.detach().double() - 3.0 * y2.double()) <= TOL


def test_conv1x1_fwd_bwd(cuda_dev):
    from doda_b200 import ops
    torch.manual_seed(0)
    x = torch.randn(3001, 64)
    W = torch.randn(1, 1, 1, 64, 32) * 0.2
    g = torch.randn(3001, 32)
    xd, wd = x.to(cuda_dev).requires_grad_(True), W.to(cuda_dev).requires_grad_(True)
    y = ops.DenseConvFunction.apply(xd, wd)
    y.backward(g.to(cuda_dev))
    W2 = W.view(64, 32).double()
    assert rel_err(y, x.double() @ W2) <= TOL
    assert rel_err(xd.grad, g.double() @ W2.t()) <= TOL
    assert rel_err(wd.grad.view(64, 32), x.double().t() @ g.double()) <= TOL


@pytest.mark.parametrize("M,C", [(5000, 16), (777, 48), (20000, 32), (64, 112), (3, 224), (1000, 10),
                                 (3000, 320), (900, 384), (500, 1024),  # 320 / 384: the concat BN of the m=32 net
                                 # the deep U-Net levels: one-launch cluster form up to 16 x 64 KB (3000 x 64 forward
                                 # fits, its backward with x and dy does not: a mixed pair of forms)
                                 (6149, 64), (3000, 64), (1381, 80), (1381, 160), (223, 96), (45, 112), (5, 16)])
@pytest.mark.parametrize("relu", [True, False])
def test_bn_relu_fwd_bwd(cuda_dev, M, C, relu):
    from doda_b200 import ops
    torch.manual_seed(M + C)
    x = torch.randn(M, C) * 2 + 0.5
    g = torch.randn(M, C)
    bn_ref = torch.nn.BatchNorm1d(C, eps=1e-4, momentum=0.1).double()
    with torch.no_grad():
        bn_ref.weight.uniform_(0.5, 1.5)
        bn_ref.bias.uniform_(-0.5, 0.5)
    bn = torch.nn.BatchNorm1d(C, eps=1e-4, momentum=0.1).to(cuda_dev)
    bn.load_state_dict({k: v.float() if v.is_floating_point() else v for k, v in bn_ref.state_dict().items()})
    xr = x.double().requires_grad_(True)
    yr = bn_ref(xr)
    yr = torch.relu(yr) if relu else yr
    yr.backward(g.double())
    xd = x.to(cuda_dev).requires_grad_(True)
    y = ops.batch_norm_relu(xd, bn, relu=relu)
    y.backward(g.to(cuda_dev))
    assert rel_err(y, yr) <= TOL
    assert rel_err(xd.grad, xr.grad) <= 2e-4
    assert rel_err(bn.weight.grad, bn_ref.weight.grad) <= TOL
    assert rel_err(bn.bias.grad, bn_ref.bias.grad) <= TOL
    assert rel_err(bn.running_mean, bn_ref.running_mean) <= TOL
    assert rel_err(bn.running_var, bn_ref.running_var) <= TOL
    assert int(bn.num_batches_tracked) == 1
    # eval mode uses the running statistics
    bn.eval(); bn_ref.eval()
    with torch.no_grad():
        ye = ops.batch_norm_relu(x.to(cuda_dev), bn, relu=relu)
        yer = bn_ref(x.double())
        yer = torch.relu(yer) if relu else yer
    assert rel_err(ye, yer) <= TOL


def test_bn_statistics_survive_large_mean_over_std(cuda_dev):
    """single-pass E[x^2] - mean^2 in fp32 cancels when |mean| >> std; the kernel accumulates around row 0"""
    from doda_b200 import ops
    torch.manual_seed(0)
    M, C = 40000, 32
    x = (torch.randn(M, C, dtype=torch.float64) * 0.5 + 1.0e3).float()
    bn = torch.nn.BatchNorm1d(C, eps=1e-4, momentum=0.1).to(cuda_dev)
    y = ops.batch_norm_relu(x.to(cuda_dev), bn, relu=False)
    xd = x.double()
    ref = (xd - xd.mean(0)) / torch.sqrt(xd.var(0, unbiased=False) + 1e-4)
    # the input itself carries 1e3 * 2^-24 = 6e-5 of rounding per element (std 0.5): 1e-3 of the output scale
    assert rel_err(y, ref) <= 2e-3
    assert rel_err(bn.running_var, 0.9 + 0.1 * xd.var(0, unbiased=True)) <= 1e-3


class _DSNormLike(torch.nn.Module):
    """Stand-in with the attribute surface of DODA's DSNorm (model/dsnorm.py:20-84): one affine pair, separate
    running statistics per domain selected by `domain_label`.  forward() is the reference computation."""

    def __init__(self, C, eps=1e-4, momentum=0.1):
        super().__init__()
        self.num_features, self.eps, self.momentum, self.affine, self.track_running_stats = C, eps, momentum, True, True
        self.weight = torch.nn.Parameter(torch.rand(C) + 0.5)
        self.bias = torch.nn.Parameter(torch.rand(C) - 0.5)
        for dom in ("source", "target"):
            self.register_buffer("running_mean_" + dom, torch.zeros(C))
            self.register_buffer("running_var_" + dom, torch.ones(C))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))
        self.domain_label = 0

    def forward(self, x):
        if self.training:
            self.num_batches_tracked += 1
        rm = self.running_mean_target if self.domain_label else self.running_mean_source
        rv = self.running_var_target if self.domain_label else self.running_var_source
        return torch.nn.functional.batch_norm(x, rm, rv, self.weight, self.bias, self.training, self.momentum, self.eps)


def test_dsnorm_domain_statistics(cuda_dev):
    """SparseSequential routes DSNorm through the fused BN kernels; each domain updates only its own running stats"""
    from doda_b200 import spconv
    torch.manual_seed(0)
    C = 32
    mine = _DSNormLike(C).to(cuda_dev)
    ref = _DSNormLike(C).double()
    ref.load_state_dict({k: (v.double() if v.is_floating_point() else v) for k, v in mine.state_dict().items()})
    seq = spconv.SparseSequential(mine, torch.nn.ReLU())
    coords = torch.zeros(500, 4, dtype=torch.int32, device=cuda_dev)
    for dom in (0, 1, 1, 0):
        mine.domain_label = ref.domain_label = dom
        x = torch.randn(500, C) * 3 + dom
        t = spconv.SparseConvTensor(x.to(cuda_dev), coords, [8, 8, 8], 1)
        y = seq(t).features
        yr = torch.relu(ref(x.double()))
        assert rel_err(y, yr) <= TOL
    for k, v in mine.state_dict().items():
        assert rel_err(v.float(), ref.state_dict()[k].float()) <= TOL, k
    mine.eval(); ref.eval()
    mine.domain_label = ref.domain_label = 1
    x = torch.randn(300, C)
    with torch.no_grad():
        y = seq(spconv.SparseConvTensor(x.to(cuda_dev), coords[:300], [8, 8, 8], 1)).features
    assert rel_err(y, torch.relu(ref(x.double()))) <= TOL


def test_fused_bn_relu_conv_matches_unfused(cuda_dev):
    """the one-node [BN, ReLU, conv] path of SparseSequential gives the same activations, gradients, running
    statistics and `input.features` side effect as the three separate nodes"""
    from doda_b200 import spconv
    from doda_b200.spconv import modules as spm
    torch.manual_seed(2)
    shape = (24, 22, 20)
    coords = torch.from_numpy(random_coords(2, 1500, 2, shape)).to(cuda_dev)
    feats = torch.randn(coords.shape[0], 16)
    res = []
    for fuse in (True, False):
        torch.manual_seed(9)
        seq = spconv.SparseSequential(torch.nn.BatchNorm1d(16, eps=1e-4, momentum=0.1), torch.nn.ReLU(),
                                      spconv.SubMConv3d(16, 32, 3, padding=1, bias=False, indice_key="s"),
                                      torch.nn.BatchNorm1d(32, eps=1e-4, momentum=0.1), torch.nn.ReLU(),
                                      spconv.SparseConv3d(32, 48, 2, stride=2, bias=False, indice_key="d")).to(cuda_dev)
        spm.fuse_conv = fuse
        try:
            f = feats.to(cuda_dev).requires_grad_(True)
            x = spconv.SparseConvTensor(f, coords, list(shape), 2)
            y = seq(x)
            side = x.features  # rebound to the first BN+ReLU activation
            (y.features.pow(2).sum() + side.sum()).backward()
        finally:
            spm.fuse_conv = True
        res.append((y.features.detach(), side.detach(), f.grad, [p.grad for p in seq.parameters()],
                    [b.clone() for b in seq.buffers()]))
    a, b = res
    assert rel_err(a[0], b[0]) <= 1e-5 and rel_err(a[1], b[1]) <= 1e-6 and rel_err(a[2], b[2]) <= 1e-4
    for ga, gb in zip(a[3], b[3]):
        assert rel_err(ga, gb) <= 1e-4
    for ba, bb in zip(a[4], b[4]):
        assert rel_err(ba.float(), bb.float()) <= 1e-6


@pytest.mark.parametrize("C,weighted", [(11, False), (13, True), (40, False)])
def test_cross_entropy_matches_torch(cuda_dev, C, weighted):
    """loss epilogue (model/unet.py:168-170): value and gradient against torch's fp64 CPU cross_entropy, with
    ignored rows and class weights; all-ignored input gives nan like torch"""
    from doda_b200 import ops
    torch.manual_seed(C)
    N = 20000
    logits = torch.randn(N, C) * 3
    labels = torch.randint(0, C, (N,))
    labels[torch.rand(N) < 0.07] = 255
    w = torch.rand(C) + 0.5 if weighted else None
    ref_in = logits.double().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(ref_in, labels, weight=w.double() if weighted else None, ignore_index=255)
    (ref * 1.7).backward()
    x = logits.to(cuda_dev).requires_grad_(True)
    out = ops.cross_entropy(x, labels.to(cuda_dev), w.to(cuda_dev) if weighted else None, ignore_index=255)
    (out * 1.7).backward()
    assert abs(float(out) - float(ref)) <= 1e-5 * abs(float(ref))
    assert rel_err(x.grad, ref_in.grad) <= 1e-5
    assert torch.all(x.grad[labels.to(cuda_dev) == 255] == 0)
    none = ops.cross_entropy(x.detach(), torch.full((N,), 255, device=cuda_dev), None, ignore_index=255)
    assert torch.isnan(none)
    mod = ops.CrossEntropyLoss(ignore_index=255).to(cuda_dev)
    assert float(mod(x.detach(), labels.to(cuda_dev))) == float(ops.cross_entropy(x.detach(), labels.to(cuda_dev), None, 255))


@pytest.mark.parametrize("K,shape", [(11, (50000,)), (13, (300, 40)), (20, (7, 9, 11)), (11, (0,))])
def test_intersection_and_union_matches_reference_formula(cuda_dev, K, shape):
    """metric epilogue (util/common_utils.py:233-247): per-class intersection / union / target counts, exact, with
    ignored labels, out-of-range predictions and an empty input"""
    from doda_b200 import metrics
    g = torch.Generator().manual_seed(K)
    pred = torch.randint(0, K + 2, shape, generator=g)  # K, K+1: out of the histogram range
    lab = torch.randint(0, K, shape, generator=g)
    if pred.numel():
        lab.view(-1)[torch.rand(lab.numel(), generator=g) < 0.1] = 255
        same = torch.rand(shape, generator=g) < 0.5
        pred = torch.where(same & (lab != 255), lab, pred)
    ri, ru, rt = metrics.intersection_and_union_ref(pred, lab, K, 255)
    for dtype in (torch.int64, torch.int32):
        i, u, t = metrics.intersectionAndUnionGPU(pred.to(cuda_dev, dtype), lab.to(cuda_dev, dtype), K, 255)
        assert i.is_cuda and i.dtype == torch.float32 and i.shape == (K,)
        assert torch.equal(i.cpu(), ri) and torch.equal(u.cpu(), ru) and torch.equal(t.cpu(), rt)


def test_intersection_and_union_matches_reference_golden(cuda_dev):
    """device metric epilogue against outputs of the reference's own intersectionAndUnionGPU (golden fixture)"""
    import os
    from doda_b200 import metrics
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "iou_metrics.npz"))
    for i in range(4):
        K = int(z["c%d/K" % i][0])
        a = metrics.intersectionAndUnionGPU(torch.from_numpy(z["c%d/pred" % i]).to(cuda_dev),
                                            torch.from_numpy(z["c%d/label" % i]).to(cuda_dev), K, 255)
        for got, name in zip(a, ("intersection", "union", "target")):
            assert np.array_equal(got.cpu().numpy(), z["c%d/%s" % (i, name)]), (i, name)


@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4])
def test_voxelize_idx_gpu_is_bit_identical_to_cpu(cuda_dev, mode):
    """device voxelizer against the CPU entry point, which is pinned to the reference's compiled
    code by the golden vectors: voxel order, p2v and v2p maps identical, for every mode, 3- and 4-column coordinates,
    heavy duplication, and the 6-point known answer"""
    from doda_b200 import pointgroup_ops
    rng = np.random.RandomState(10 + mode)
    cases = []
    if mode == 0:  # unique coordinates only
        cases.append(np.stack(np.unravel_index(rng.permutation(4096)[:3000], (16, 16, 16)), 1))
        cases.append(np.concatenate([rng.randint(0, 3, (3000, 1)), cases[0]], 1))
    else:
        for n, side, ncol in ((5000, 12, 4), (3000, 40, 3), (20000, 6, 4), (1, 3, 4), (257, 2, 4)):
            c = rng.randint(0, side, size=(n, 3))
            if ncol == 4:
                c = np.concatenate([np.sort(rng.randint(0, 3, size=(n, 1)), 0), c], 1)
            cases.append(c)
        cases.append(np.array([[0, 1, 1, 1], [0, 1, 1, 1], [0, 2, 2, 2], [1, 1, 1, 1], [0, 2, 2, 2], [0, 1, 1, 1]]))
    for c in cases:
        c = torch.from_numpy(np.ascontiguousarray(c.astype(np.int64)))
        oc, im, om = pointgroup_ops.voxelization_idx(c, 3, mode)
        goc, gim, gom = pointgroup_ops.voxelization_idx_gpu(c.to(cuda_dev), 3, mode)
        assert goc.is_cuda and gim.dtype == torch.int32 and gom.dtype == torch.int32
        assert torch.equal(goc.cpu(), oc) and torch.equal(gim.cpu(), im) and torch.equal(gom.cpu(), om), (mode, tuple(c.shape))


def test_voxelize_idx_gpu_full_size_and_errors(cuda_dev):
    from doda_b200 import pointgroup_ops, scenes
    batch = scenes.collate([scenes.scene_with_voxels(i, 150000) for i in range(2)], seed=0, dup_max=2)
    locs = batch["locs"].contiguous() if "locs" in batch else None
    if locs is None:  # rebuild point coordinates from the maps: every point carries its voxel's coordinates
        locs = batch["voxel_locs"][batch["p2v_map"].long()].contiguous()
    oc, im, om = pointgroup_ops.voxelization_idx(locs, 2, 4)
    goc, gim, gom = pointgroup_ops.voxelization_idx_gpu(locs.to(cuda_dev), 2, 4)
    assert oc.shape[0] == 300000
    assert torch.equal(goc.cpu(), oc) and torch.equal(gim.cpu(), im) and torch.equal(gom.cpu(), om)
    e = pointgroup_ops.voxelization_idx_gpu(torch.zeros((0, 4), dtype=torch.int64, device=cuda_dev), 1, 4)
    assert e[0].shape == (0, 4) and e[1].numel() == 0 and e[2].shape[0] == 0
    with pytest.raises(RuntimeError):
        pointgroup_ops.voxelization_idx_gpu(torch.tensor([[0, -1, 0, 0]], device=cuda_dev), 1, 4)
    # batch indices past 15 (round-1 limit): the key gives the batch whatever bits the grid leaves -- 64 scenes on a
    # 1024-wide grid, and batch 40000 on a small one; only a 2^20-wide grid still holds the batch below 15
    rng = np.random.RandomState(3)
    for nb, side in ((64, 1024), (40000, 30), (14, (1 << 20) - 1)):
        c = np.concatenate([np.sort(rng.randint(0, nb, size=(6000, 1)), 0), rng.randint(0, 8, size=(6000, 3)) * (side // 8)], 1)
        c[0, 0], c[-1, 0], c[5, 1:] = 0, nb - 1, side - 1 if side < (1 << 20) - 1 else side
        c = torch.from_numpy(np.ascontiguousarray(c.astype(np.int64)))
        oc, im, om = pointgroup_ops.voxelization_idx(c, nb, 4)
        goc, gim, gom = pointgroup_ops.voxelization_idx_gpu(c.to(cuda_dev), nb, 4)
        assert torch.equal(goc.cpu(), oc) and torch.equal(gim.cpu(), im) and torch.equal(gom.cpu(), om), (nb, side)
    with pytest.raises(RuntimeError):
        pointgroup_ops.voxelization_idx_gpu(torch.tensor([[15, (1 << 20) - 1, 0, 0]], device=cuda_dev), 16, 4)
    with pytest.raises(RuntimeError):
        pointgroup_ops.voxelization_idx_gpu(torch.tensor([[0, 1 << 20, 0, 0]], device=cuda_dev), 1, 4)
    with pytest.raises(RuntimeError):
        pointgroup_ops.voxelization_idx_gpu(torch.tensor([[0, 1, 1, 1], [0, 1, 1, 1]], device=cuda_dev), 1, 0)


def test_voxelize_and_devoxelize(cuda_dev):
    from doda_b200 import pointgroup_ops, ops
    from oracle.unet_ref import voxelize_mean_ref
    rng = np.random.RandomState(0)
    pts = rng.randint(0, 12, size=(4000, 3))
    locs = torch.from_numpy(np.concatenate([rng.randint(0, 2, size=(4000, 1)), pts], 1)).long().contiguous()
    vl, p2v, v2p = pointgroup_ops.voxelization_idx(locs, 2, 4)
    feats = torch.randn(4000, 3)
    ref = voxelize_mean_ref(feats.double(), v2p.numpy())
    fd = feats.to(cuda_dev).requires_grad_(True)
    out = pointgroup_ops.voxelization(fd, v2p.to(cuda_dev), 4)
    assert rel_err(out, ref) <= 1e-6
    g = torch.randn_like(out)
    out.backward(g)
    fr = feats.double().requires_grad_(True)
    voxelize_mean_ref(fr, v2p.numpy()).backward(g.double().cpu())
    assert rel_err(fd.grad, fr.grad) <= 1e-6
    # devoxelize gather + scatter-add backward (model/unet.py:62)
    for C, dt in ((16, torch.int32), (16, torch.int64), (5, torch.int32)):
        src = torch.randn(vl.shape[0], C, device=cuda_dev, requires_grad=True)
        idx = p2v.to(cuda_dev).to(dt)
        y = ops.gather_rows(src, idx)
        assert torch.equal(y, src[idx.long()])
        gy = torch.randn_like(y)
        y.backward(gy)
        ref_g = torch.zeros_like(src).index_add_(0, idx.long(), gy)
        assert rel_err(src.grad, ref_g) <= 1e-5
        # the same with the voxel -> points map: backward = atomics-free segmented sum, bit-identical run to run
        grads = []
        for _ in range(2):
            s2 = src.detach().clone().requires_grad_(True)
            y2 = ops.devoxelize(s2, idx, v2p.to(cuda_dev))
            assert torch.equal(y2, y)
            y2.backward(gy)
            grads.append(s2.grad)
        assert torch.equal(grads[0], grads[1]) and rel_err(grads[0], ref_g) <= 1e-5


def _unet_parity(cuda_dev, target, mid, tol_scores, tol_grad):
    """Whole net (71 convs + 65 BN) against the fp64 oracle.  The deep levels of a small scene hold a handful of
    rows, so BatchNorm (eps 1e-4) and the ReLU gates make some weight gradients ill-conditioned IN FP32 ITSELF: the
    fp32 CPU oracle (the reference algorithm, fp32 SGEMM) is evaluated too and bounds what fp32 can deliver.  The
    sm_100a path must match fp64 to `tol_grad`, or be as close to fp64 as the fp32 reference algorithm is."""
    from doda_b200 import scenes
    from doda_b200.unet import SparseConvNet, model_step
    from oracle.unet_ref import model_step_ref
    torch.manual_seed(0)
    batch = scenes.collate([scenes.scene_with_voxels(0, target), scenes.scene_with_voxels(1, target)], dup_max=2)
    model = SparseConvNet(mid_channel=mid)
    sd64 = {k: (v.detach().double().clone().requires_grad_(True) if v.is_floating_point() else v.clone())
            for k, v in model.state_dict().items()}
    sd32 = {k: (v.detach().clone().requires_grad_(True) if v.is_floating_point() else v.clone())
            for k, v in model.state_dict().items()}
    b64 = dict(batch)
    b64["feats"] = batch["feats"].double()
    loss_ref, scores_ref = model_step_ref(sd64, b64, training=True)
    loss_ref.backward()
    loss32, scores32 = model_step_ref(sd32, batch, training=True)
    loss32.backward()
    model = model.to(cuda_dev).train()
    from helpers import capture_relu_masks, pinned_grad_report, assert_pinned_grad_parity, oracle_step
    sd0 = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    masks = capture_relu_masks(model)
    loss, scores = model_step(model, batch, device=cuda_dev)
    loss.backward()
    assert rel_err(scores, scores_ref) <= tol_scores, ("scores", rel_err(scores, scores_ref))
    _, _, sd64p = oracle_step(sd0, batch, torch.float64, relu_masks=masks)
    prep = pinned_grad_report([(n, p.grad) for n, p in model.named_parameters()], sd64p)
    print("unet grads vs fp64 oracle with pinned gates:", prep)
    assert_pinned_grad_parity(prep, "target %d m %d" % (target, mid))
    assert abs(float(loss.detach()) - float(loss_ref.detach())) <= tol_scores * max(1.0, abs(float(loss_ref.detach())))
    from helpers import grad_report, assert_grad_parity
    report = grad_report([(n, p.grad) for n, p in model.named_parameters()], sd64, sd32)
    print("unet grad parity:", report)
    assert_grad_parity(report, "target %d m %d" % (target, mid))
    assert rel_err(model.linear.weight.grad, sd64["linear.weight"].grad) <= 1e-4


def test_staged_batch_and_host_batch_give_the_same_step(cuda_dev):
    """ops.stage_batch (H2D copies + coordinate cast on the index stream, ahead of the step) feeds model_step the
    same values as the plain host batch: identical loss and scores, two steps in a row (event hand-off both ways)"""
    from doda_b200 import scenes, ops
    from doda_b200.unet import SparseConvNet, model_step
    torch.manual_seed(0)
    batch = scenes.collate([scenes.scene_with_voxels(3, 6000), scenes.scene_with_voxels(4, 5000)], dup_max=2)
    pinned = dict(batch)
    for k in ("voxel_locs", "p2v_map", "v2p_map", "feats", "labels"):
        pinned[k] = batch[k].pin_memory()
    model = SparseConvNet(mid_channel=16).to(cuda_dev).eval()  # eval: no running-stat updates between the passes
    with torch.no_grad():
        ref_loss, ref_scores = model_step(model, batch, device=cuda_dev)
        for _ in range(2):
            staged = ops.stage_batch(pinned, cuda_dev)
            assert staged["voxel_locs"].dtype == torch.int32 and staged["feats"].is_cuda
            loss, scores = model_step(model, staged, device=cuda_dev)
            # split-K layers add partial sums with atomics: equal up to summation order
            assert rel_err(scores, ref_scores) <= 1e-5 and abs(float(loss) - float(ref_loss)) <= 1e-5


def test_unet_fwd_bwd_small_scene(cuda_dev):
    # whole-net check: 71 convs + 65 BN deep, so the per-op 1e-4 bound compounds; 1e-3 on logits / grads
    _unet_parity(cuda_dev, 20000, 16, 1e-3, 5e-3)


def test_unet_fwd_bwd_m32(cuda_dev):
    _unet_parity(cuda_dev, 8000, 32, 1e-3, 5e-3)


@pytest.mark.parametrize("impl", ["tc", "fp32"])
def test_conv_impls_agree(cuda_dev, impl):
    """both conv implementations (tcgen05 3xTF32 and fp32 CUDA cores), with and without the mask-sorted processing
    order, produce the same SubM conv within the fp32 tolerance"""
    from doda_b200 import ops
    torch.manual_seed(5)
    coords, shape = surface_coords(3, 6000, 2)
    c = torch.from_numpy(coords).to(cuda_dev)
    rb = ops.build_rulebook(c, 2, shape, 3, 1, 1, 1, subm=True)
    n = c.shape[0]
    # order is a permutation, nbr_perm is nbr in that order, masks are non-decreasing along it
    order = rb.order.cpu().numpy()
    assert np.array_equal(np.sort(order), np.arange(n))
    nbr = rb.nbr.cpu().numpy()
    assert np.array_equal(rb.nbr_perm.cpu().numpy(), nbr[order])
    masks = ((nbr >= 0).astype(np.int64) << np.arange(27)).sum(1)
    assert np.all(np.diff(masks[order]) >= 0)
    assert np.array_equal(rb.rowmask.cpu().numpy().astype(np.int64), masks[order])
    feat = torch.randn(n, 32, device=cuda_dev)
    W3 = torch.randn(27, 32, 48, device=cuda_dev) * 0.2
    ref = torch.zeros(n, 48, dtype=torch.float64)
    fd, Wd, t = feat.double().cpu(), W3.double().cpu(), rb.nbr.cpu().long()
    for k in range(27):
        m = t[:, k] >= 0
        ref[m] += fd[t[m, k]] @ Wd[k]
    ops.set_conv_impl(impl)
    try:
        a = ops.gather_gemm(feat, W3, rb.nbr, n)
        b = ops.gather_gemm(feat, W3, rb.nbr_perm, n, orow=rb.order)
        b2 = ops.gather_gemm(feat, W3, rb.nbr_perm, n, orow=rb.order, rowmask=rb.rowmask)
        assert torch.equal(b, b2) or impl == "tc" and rel_err(b2, b) <= 1e-5
    finally:
        ops.set_conv_impl("tc")
    assert rel_err(a, ref) <= TOL and rel_err(b, ref) <= TOL


@pytest.mark.parametrize("C", [(48, 48), (64, 32), (16, 80), (112, 112)])
def test_tma_gather4_loader_matches_the_cp_async_loader(cuda_dev, C):
    """k_conv_tc's two gather engines -- per-thread cp.async into the un-swizzled canonical layout, and TMA
    cp.async.bulk.tensor tile::gather4 into 128/64/32-byte-swizzled K-major tiles (row coordinates straight from the
    rulebook table, -1 = out of bounds = zeros) -- feed the same MMAs: identical results bit for bit, table and pair
    mode, forward and dgrad, and both within the fp32 bar of the oracle"""
    from doda_b200 import ops
    torch.manual_seed(7)
    Cin, Cout = C
    coords, shape = surface_coords(5, 9000, 2)
    c = torch.from_numpy(coords).to(cuda_dev)
    rb = ops.build_rulebook(c, 2, shape, 3, 1, 1, 1, subm=True)
    rbs = ops.build_rulebook(c, 2, shape, 2, 2, 0, 1, subm=False)
    n = c.shape[0]
    x = torch.randn(n, Cin, device=cuda_dev)
    g = torch.randn(n, Cout, device=cuda_dev)
    xc = torch.randn(rbs.outids.shape[0], Cin, device=cuda_dev)
    W3 = torch.randn(27, Cin, Cout, device=cuda_dev) * 0.2
    W8 = torch.randn(8, Cin, Cout, device=cuda_dev) * 0.2
    outs = {}
    try:
        for tma in (0, 1):
            ops.set_conv_tma(tma)
            outs[tma] = (ops.gather_gemm(x, W3, rb.nbr_perm, n, orow=rb.order, rowmask=rb.rowmask),
                         ops.gather_gemm(g, W3, rb.nbr_perm, n, wflags=ops.W_T_MIRROR, orow=rb.order, rowmask=rb.rowmask),
                         ops.gather_gemm_pairs(xc, W8, rbs.pairs[1], rbs.pairs[0], rbs.pairnum, n, n))
    finally:
        ops.set_conv_tma(0)
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)
    ref = torch.zeros(n, Cout, dtype=torch.float64)
    fd, Wd, t = x.double().cpu(), W3.double().cpu(), rb.nbr.cpu().long()
    for k in range(27):
        m = t[:, k] >= 0
        ref[m] += fd[t[m, k]] @ Wd[k]
    assert rel_err(outs[1][0], ref) <= TOL


def test_spconv_surface_sequential_semantics(cuda_dev):
    """SparseSequential mutates .features of the SAME object for dense modules."""
    from doda_b200 import spconv
    coords = torch.from_numpy(random_coords(0, 500, 1, (16, 16, 16))).to(cuda_dev)
    x = spconv.SparseConvTensor(torch.randn(500, 8, device=cuda_dev), coords, [16, 16, 16], 1)
    keep = x.features
    seq = spconv.SparseSequential(torch.nn.BatchNorm1d(8, eps=1e-4, momentum=0.1), torch.nn.ReLU()).to(cuda_dev)
    y = seq(x)
    assert y is x and x.features is not keep and float(x.features.min()) >= 0.0
    conv = spconv.SparseSequential(spconv.SubMConv3d(8, 4, 3, padding=1, bias=True, indice_key="k")).to(cuda_dev)
    z = conv(x)
    assert z is not x and z.indice_dict is x.indice_dict and "k" in x.indice_dict
    assert z.features.shape == (500, 4)
    d = z.dense()
    assert d.shape == (1, 4, 16, 16, 16)


@pytest.mark.gpu
@pytest.mark.parametrize("raw", [True, False])
@pytest.mark.parametrize("residual", [True, False])
def test_taped_ublock_matches_module_path(cuda_dev, residual, raw):
    """doda_b200/tape.py runs the U-Net sub-tree as ONE autograd node (host cost); it must be the same computation as
    the module-by-module path: identical loss, scores, BatchNorm running statistics and BatchNorm / activation
    gradients bit for bit; conv weight gradients (float atomics in both paths) to 1e-5"""
    from doda_b200 import scenes, tape
    from doda_b200.unet import SparseConvNet, model_step
    torch.manual_seed(0)
    batch = scenes.collate([scenes.scene_with_voxels(0, 6000), scenes.scene_with_voxels(1, 5000)], dup_max=2)
    model = SparseConvNet(mid_channel=16, block_residual=residual).to(cuda_dev).train()
    sd0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    res = {}
    raw0 = tape.raw
    try:
        tape.raw = raw  # raw device buffers from a per-step arena (default) / torch tensors per layer
        for on in (False, True):
            tape.enabled = on
            model.load_state_dict(sd0)
            for p in model.parameters():
                p.grad = None
            runs0 = tape.runs
            loss, scores = model_step(model, batch, device=cuda_dev)
            loss.backward()
            assert (tape.runs - runs0 == 1) == on
            res[on] = (loss.detach().clone(), scores.detach().clone(),
                       {n: p.grad.detach().clone() for n, p in model.named_parameters()},
                       {n: b.detach().clone() for n, b in model.named_buffers()})
    finally:
        tape.enabled = True
        tape.raw = raw0
    assert torch.equal(res[True][0], res[False][0]) and torch.equal(res[True][1], res[False][1])
    for n, b in res[False][3].items():
        assert torch.equal(b, res[True][3][n]), n
    for n, g in res[False][2].items():
        if g.dim() >= 3:  # conv filters
            assert rel_err(res[True][2][n], g) <= 1e-5, n
        else:
            assert torch.equal(res[True][2][n], g), n


@pytest.mark.gpu
def test_taped_ublock_falls_back_and_frees(cuda_dev):
    """eval mode / no_grad / a second backward: module path or a clear error, never a wrong result"""
    from doda_b200 import scenes, tape
    from doda_b200.unet import SparseConvNet, model_step
    torch.manual_seed(1)
    batch = scenes.collate([scenes.scene_with_voxels(0, 3000)], dup_max=2)
    model = SparseConvNet(mid_channel=16).to(cuda_dev).train()
    loss, _ = model_step(model, batch, device=cuda_dev)
    loss.backward(retain_graph=True)
    with pytest.raises(RuntimeError, match="ONE backward"):
        loss.backward()
    runs0 = tape.runs
    model.eval()
    with torch.no_grad():
        _, s_eval = model_step(model, batch, device=cuda_dev)
    assert tape.runs == runs0 and torch.isfinite(s_eval).all()
    model.train()
    model.unet.u.blocks.block0.conv_branch[0].eval()  # one BatchNorm of level 2 frozen: levels 1 and 2 run module by
    _, s1 = model_step(model, batch, device=cuda_dev)  # module, the level-3 sub-tree still qualifies and is taped
    assert tape.runs == runs0 + 1 and torch.isfinite(s1).all()


@pytest.mark.gpu
def test_fused_residual_and_gradient_adds_match_separate_adds(cuda_dev):
    """b200sp_gather_gemm_res (out = conv + res in the epilogue), b200sp_bn_bwd_add (dx = BN backward + add) and
    b200sp_copy_cols against the unfused sequences: bit-identical (the same fp32 adds, just not as a kernel of their own)"""
    from doda_b200 import ops
    from doda_b200._lib import lib, check
    torch.manual_seed(0)
    st = ops._stream()
    for M, C, Co in ((20000, 16, 16), (9000, 32, 32), (3000, 48, 48), (300, 96, 96), (2000, 32, 16)):
        coords, shape = surface_coords(1, M // 2, 2)
        c = torch.from_numpy(coords).to(cuda_dev)
        rb = ops.build_rulebook(c, 2, shape, 3, 1, 1, 1, subm=True)
        n = c.shape[0]
        x = torch.randn(n, C, device=cuda_dev)
        W3 = torch.randn(27, C, Co, device=cuda_dev) * 0.2
        res = torch.randn(n, Co, device=cuda_dev)
        ref = ops.gather_gemm(x, W3, rb.nbr_perm, n, orow=rb.order, rowmask=rb.rowmask) + res
        out = torch.empty(n, Co, device=cuda_dev)
        ws = torch.zeros(int(lib.b200sp_conv_ws_bytes(27, C, Co)), dtype=torch.uint8, device=cuda_dev)
        check(lib.b200sp_gather_gemm_res(x.data_ptr(), n, C, W3.data_ptr(), 0, rb.nbr_perm.data_ptr(), rb.order.data_ptr(),
                                         rb.rowmask.data_ptr(), 27, out.data_ptr(), n, Co, res.data_ptr(), ws.data_ptr(),
                                         ws.numel(), st), "gather_gemm_res")
        assert torch.equal(out, ref), (M, C, Co)
    for M, C in ((30000, 16), (777, 48), (45, 112), (1000, 10)):
        x = torch.randn(M, C, device=cuda_dev) * 2 + 0.3
        dy = torch.randn(M, C, device=cuda_dev)
        add = torch.randn(M, C, device=cuda_dev)
        w = torch.rand(C, device=cuda_dev) + 0.5
        b = torch.randn(C, device=cuda_dev) * 0.1
        mean, var = x.mean(0), x.var(0, unbiased=False)
        inv = torch.rsqrt(var + 1e-4)
        ws = torch.zeros(int(lib.b200sp_bn_ws_bytes(M, C)), dtype=torch.uint8, device=cuda_dev)
        dx0, dx1 = torch.empty_like(x), torch.empty_like(x)
        g0, g1 = torch.empty(2, C, device=cuda_dev), torch.empty(2, C, device=cuda_dev)
        check(lib.b200sp_bn_bwd(x.data_ptr(), dy.data_ptr(), M, C, w.data_ptr(), b.data_ptr(), mean.data_ptr(), inv.data_ptr(), 1,
                                dx0.data_ptr(), g0.data_ptr(), g0.data_ptr() + 4 * C, ws.data_ptr(), ws.numel(), st), "bn_bwd")
        check(lib.b200sp_bn_bwd_add(x.data_ptr(), dy.data_ptr(), M, C, w.data_ptr(), b.data_ptr(), mean.data_ptr(), inv.data_ptr(),
                                    1, dx1.data_ptr(), g1.data_ptr(), g1.data_ptr() + 4 * C, add.data_ptr(), ws.data_ptr(),
                                    ws.numel(), st), "bn_bwd_add")
        assert torch.equal(dx1, dx0 + add) and torch.equal(g0, g1), (M, C)
    a = torch.randn(5000, 48, device=cuda_dev)
    bsrc = torch.randn(5000, 32, device=cuda_dev)
    cat = torch.empty(5000, 80, device=cuda_dev)
    check(lib.b200sp_copy_cols(a.data_ptr(), 5000, 48, 0, 48, cat.data_ptr(), 80, 0, st), "copy_cols")
    check(lib.b200sp_copy_cols(bsrc.data_ptr(), 5000, 32, 0, 32, cat.data_ptr(), 80, 48, st), "copy_cols")
    assert torch.equal(cat, torch.cat((a, bsrc), 1))
    back = torch.empty(5000, 32, device=cuda_dev)
    check(lib.b200sp_copy_cols(cat.data_ptr(), 5000, 80, 48, 32, back.data_ptr(), 32, 0, st), "copy_cols")
    assert torch.equal(back, bsrc)
    with pytest.raises(RuntimeError):
        check(lib.b200sp_copy_cols(a.data_ptr(), 10, 48, 40, 16, cat.data_ptr(), 80, 0, st), "copy_cols")
