"""GPU parity of the PG_OP / pointops2_cuda replacement kernels against the REFERENCE'S OWN kernels: the reference
sources (lib/pointgroup_ops, lib/pointops2) compiled unmodified for sm_100a by oracle/build_ref.py into oracle/_ref/
(the prebuilt .so files travel to the GPU box; /root/reference is not needed at run time).  Same inputs on both
sides, outputs allocated by the caller exactly as the reference's Python wrappers do
(lib/pointgroup_ops/functions/pointgroup_ops.py, lib/pointops2/functions/pointops2.py).
Integer results are compared exactly (neighbour lists as sets where the reference's order is thread-timing
dependent), float results to 1e-5 (atomics change the summation order)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def refs(cuda_dev):
    from oracle import build_ref
    pg, po = build_ref.load("PG_OP"), build_ref.load("pointops2_cuda")
    if pg is None or po is None:
        pytest.skip("oracle/_ref/*.so not built (python -m oracle.build_ref)")
    from doda_b200 import pg_op, pointops2_cuda
    return pg, po, pg_op, pointops2_cuda


def _sync():
    torch.cuda.synchronize()  # the reference launches on the legacy default stream


def _close(a, b, tol=1e-5):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max()) <= tol * max(1.0, float(b.abs().max()))


def _batched_points(n_per, B, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    xyz = torch.rand(n_per * B, 3, generator=g) * scale
    batch_idxs = torch.arange(B).repeat_interleave(n_per).int()
    offsets = torch.arange(0, (B + 1) * n_per, n_per).int()
    return xyz.cuda(), batch_idxs.cuda(), offsets.cuda()


@pytest.mark.parametrize("mode", [3, 4])
def test_voxelize_fp_bp_and_point_recover(refs, mode):
    pg, _, mine, _ = refs
    from doda_b200 import pointgroup_ops
    rng = np.random.RandomState(0)
    locs = torch.from_numpy(np.concatenate([np.sort(rng.randint(0, 2, size=(3000, 1)), 0),
                                            rng.randint(0, 10, size=(3000, 3))], 1)).long().contiguous()
    _, p2v, v2p = pointgroup_ops.voxelization_idx(locs, 2, mode)
    M, A, C = v2p.shape[0], v2p.shape[1] - 1, 5
    feats = torch.randn(3000, C).cuda()
    v2p = v2p.cuda()
    outs = []
    for mod in (pg, mine):
        o = torch.zeros(M, C, device="cuda")
        mod.voxelize_fp(feats, o, v2p, mode, M, A, C)
        d = torch.zeros(3000, C, device="cuda")
        g = torch.arange(M * C, device="cuda", dtype=torch.float32).view(M, C) / 100
        mod.voxelize_bp(g.contiguous(), d, v2p, mode, M, A, C)
        r = torch.zeros(3000, C, device="cuda")
        mod.point_recover_fp(o.contiguous(), r, v2p, M, A, C)
        rb = torch.zeros(M, C, device="cuda")
        mod.point_recover_bp(torch.ones(3000, C, device="cuda"), rb, v2p, M, A, C)
        _sync()
        outs.append((o, d, r, rb))
    for a, b in zip(outs[1], outs[0]):
        assert _close(a, b)


def test_sec_mean_min_max(refs):
    pg, _, mine, _ = refs
    torch.manual_seed(0)
    for N, C, offs in ((100, 8, [0, 10, 20, 50, 100]),  # the example of functions/pointgroup_ops.py:410-414
                       (5000, 33, sorted(set([0, 5000] + list(np.random.RandomState(1).randint(1, 5000, 40)))))):
        inp = torch.randn(N, C).cuda()
        offsets = torch.tensor(offs, dtype=torch.int32).cuda()
        P = offsets.numel() - 1
        res = []
        for mod in (pg, mine):
            o1, o2, o3 = (torch.zeros(P, C, device="cuda") for _ in range(3))
            mod.sec_mean(inp, offsets, o1, P, C)
            mod.sec_min(inp, offsets, o2, P, C)
            mod.sec_max(inp, offsets, o3, P, C)
            d = torch.zeros(N, C, device="cuda")
            mod.sec_mean_bp(d, offsets, torch.arange(P * C, device="cuda", dtype=torch.float32).view(P, C).contiguous(), P, C)
            _sync()
            res.append((o1, o2, o3, d))
        for a, b in zip(res[1], res[0]):
            assert _close(a, b)


def test_roipool_and_get_iou(refs):
    pg, _, mine, _ = refs
    torch.manual_seed(1)
    N, C = 4000, 16
    offs = torch.tensor(sorted(set([0, N] + list(np.random.RandomState(2).randint(1, N, 30)))), dtype=torch.int32).cuda()
    P = offs.numel() - 1
    feats = torch.randn(N, C).cuda()
    res = []
    for mod in (pg, mine):
        o = torch.zeros(P, C, device="cuda")
        mi = torch.zeros(P, C, dtype=torch.int32, device="cuda")
        mod.roipool_fp(feats, offs, o, mi, P, C)
        d = torch.zeros(N, C, device="cuda")
        mod.roipool_bp(d, offs, mi, torch.randn(P, C, generator=torch.Generator().manual_seed(3)).cuda().contiguous(), P, C)
        _sync()
        res.append((o, mi, d))
    assert _close(res[1][0], res[0][0]) and torch.equal(res[1][1], res[0][1]) and _close(res[1][2], res[0][2])
    # get_iou
    rng = np.random.RandomState(4)
    nInst, Npts = 7, 3000
    labels = torch.from_numpy(rng.randint(-1, nInst, size=Npts)).long()
    labels[labels < 0] = -100
    pointnum = torch.tensor([(labels == i).sum() for i in range(nInst)], dtype=torch.int32)
    prop_idx = torch.from_numpy(rng.randint(0, Npts, size=2500)).int()
    prop_off = torch.tensor(sorted(set([0, 2500] + list(rng.randint(1, 2500, 12)))), dtype=torch.int32)
    nP = prop_off.numel() - 1
    ious = []
    for mod in (pg, mine):
        iou = torch.zeros(nP, nInst, device="cuda")
        mod.get_iou(prop_idx.cuda(), prop_off.cuda(), labels.cuda(), pointnum.cuda(), iou, nInst, nP)
        _sync()
        ious.append(iou)
    assert _close(ious[1], ious[0])


def test_ballquery_batch_p(refs):
    pg, _, mine, _ = refs
    xyz, bidx, boff = _batched_points(600, 2, 0)
    n, mean_active, radius = xyz.shape[0], 50, 0.12
    res = []
    for mod in (pg, mine):
        idx = torch.zeros(n * mean_active, dtype=torch.int32, device="cuda")
        sl = torch.zeros(n, 2, dtype=torch.int32, device="cuda")
        nact = mod.ballquery_batch_p(xyz, bidx, boff, idx, sl, n, mean_active, radius)
        _sync()
        res.append((int(nact), idx.cpu().numpy(), sl.cpu().numpy()))
    assert res[0][0] == res[1][0] and res[0][0] <= n * mean_active
    assert np.array_equal(res[0][2][:, 1], res[1][2][:, 1])  # neighbour counts per point
    for (_, idx, sl) in res:
        assert int(sl[:, 1].sum()) == res[0][0]
    for i in range(n):
        s0, l0 = res[0][2][i]
        s1, l1 = res[1][2][i]
        assert set(res[0][1][s0:s0 + l0].tolist()) == set(res[1][1][s1:s1 + l1].tolist()), i
    # the wrapper's retry protocol: too small a buffer reports the needed size
    idx = torch.zeros(n * 2, dtype=torch.int32, device="cuda")
    sl = torch.zeros(n, 2, dtype=torch.int32, device="cuda")
    assert mine.ballquery_batch_p(xyz, bidx, boff, idx, sl, n, 2, radius) == res[0][0]


def test_knn_batch(refs):
    pg, _, mine, _ = refs
    xyz, bidx, _ = _batched_points(500, 2, 1)
    qxyz, _, qoff = _batched_points(40, 2, 2)
    n, m, k = xyz.shape[0], qxyz.shape[0], 4
    out = []
    for mod in (pg, mine):
        idx = torch.zeros(n, k, dtype=torch.int32, device="cuda")
        mod.knn_batch(xyz, qxyz, bidx, qoff, idx, n, m, k)
        _sync()
        out.append(idx)
    assert torch.equal(out[0], out[1])


def _knn(po_mod, xyz, new_xyz, off, new_off, nsample):
    m = new_xyz.shape[0]
    idx = torch.zeros(m, nsample, dtype=torch.int32, device="cuda")
    d2 = torch.zeros(m, nsample, device="cuda")
    po_mod.knnquery_cuda(m, nsample, xyz, new_xyz, off[1:].contiguous(), new_off[1:].contiguous(), idx, d2)
    _sync()
    return idx, d2


def test_pointops2_knnquery_and_fps(refs):
    _, po, _, mine = refs
    xyz, _, off = _batched_points(700, 2, 3)
    new_xyz, _, new_off = _batched_points(90, 2, 4)
    i0, d0 = _knn(po, xyz, new_xyz, off, new_off, 8)
    i1, d1 = _knn(mine, xyz, new_xyz, off, new_off, 8)
    assert _close(d1, d0, 1e-6) and torch.equal(i0, i1)
    # furthest point sampling (deterministic): 2 batches of 700 -> 100 + 150 samples
    n_off = torch.tensor([100, 250], dtype=torch.int32).cuda()
    res = []
    for mod in (po, mine):
        idx = torch.zeros(250, dtype=torch.int32, device="cuda")
        tmp = torch.full((1400,), 1e10, device="cuda")
        mod.furthestsampling_cuda(2, 700, xyz, off[1:].contiguous(), n_off, tmp, idx)
        idx2 = torch.zeros(250, dtype=torch.int32, device="cuda")
        tmp2 = torch.full((1400,), 1e10, device="cuda")
        mod.furthestsampling_dim_cuda(2, 700, 3, xyz, off[1:].contiguous(), n_off, tmp2, idx2)
        _sync()
        res.append((idx, idx2))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])


def test_pointops2_grouping_interp_subtraction_aggregation(refs):
    _, po, _, mine = refs
    torch.manual_seed(5)
    xyz, _, off = _batched_points(400, 2, 6)
    n, ns, c, wc = xyz.shape[0], 8, 16, 4
    idx, _ = _knn(po, xyz, xyz, off, off, ns)
    inp = torch.randn(n, c).cuda()
    inp2 = torch.randn(n, c).cuda()
    pos = torch.randn(n, ns, c).cuda()
    w = torch.randn(n, ns, wc).cuda()
    gout3 = torch.randn(n, ns, c).cuda()
    gout2 = torch.randn(n, c).cuda()
    w3 = torch.rand(n, 3).cuda()
    idx3 = idx[:, :3].contiguous()
    res = []
    for mod in (po, mine):
        r = {}
        o = torch.zeros(n, ns, c, device="cuda"); mod.grouping_forward_cuda(n, ns, c, inp, idx, o); r["gf"] = o
        g = torch.zeros(n, c, device="cuda"); mod.grouping_backward_cuda(n, ns, c, gout3, idx, g); r["gb"] = g
        o = torch.zeros(n, c, device="cuda"); mod.interpolation_forward_cuda(n, c, 3, inp, idx3, w3, o); r["if"] = o
        g = torch.zeros(n, c, device="cuda"); mod.interpolation_backward_cuda(n, c, 3, gout2, idx3, w3, g); r["ib"] = g
        o = torch.zeros(n, ns, c, device="cuda"); mod.subtraction_forward_cuda(n, ns, c, inp, inp2, idx, o); r["sf"] = o
        g1, g2 = torch.zeros(n, c, device="cuda"), torch.zeros(n, c, device="cuda")
        mod.subtraction_backward_cuda(n, ns, c, idx, gout3, g1, g2); r["sb1"], r["sb2"] = g1, g2
        o = torch.zeros(n, c, device="cuda"); mod.aggregation_forward_cuda(n, ns, c, wc, inp, pos, w, idx, o); r["af"] = o
        gi, gp, gw = torch.zeros(n, c, device="cuda"), torch.zeros(n, ns, c, device="cuda"), torch.zeros(n, ns, wc, device="cuda")
        mod.aggregation_backward_cuda(n, ns, c, wc, inp, pos, w, idx, gout2, gi, gp, gw)
        r["abi"], r["abp"], r["abw"] = gi, gp, gw
        _sync()
        res.append(r)
    for key in res[0]:
        assert _close(res[1][key], res[0][key], 2e-5), key
