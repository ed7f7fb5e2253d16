"""`baseline_gpu_native`: spconv v1.2's NATIVE algorithm on the same GPU, written with stock torch ops only -- the
stand-in for "reference spconv-CUDA" that north_star's ">= 5x" is quoted against (SURVEY.md §8d:
spconv v1.2 itself cannot be built against torch 2.11 / CUDA 12.9 offline).

Per sparse conv, exactly what `indice_conv` / `indice_conv_backward` do: for every kernel offset a row
gather (`index_select`), an fp32 SGEMM (`torch.mm`, TF32 off) and a scatter-add (`index_add_`); the SubM centre offset
is one dense GEMM; backward = two gathers, two SGEMMs and a scatter-add per offset.  BatchNorm / ReLU / concat / the
devoxelize gather / cross-entropy are torch's own CUDA kernels, as in the reference.  Rulebooks are built with torch
ops on the device (sorted flat keys + searchsorted instead of spconv's dense grid).

None of the engine's code is on this path: no import of doda_b200, no libb200sparse.  Weights come in as a state_dict
with the reference's key names.  `NativeUNet.step()` runs one forward + backward eagerly; `capture()` records the same
step (rulebooks excluded: their sizes need host reads) into a CUDA graph so that the reported time is the algorithm's
GPU time without Python launch overhead -- the most favourable reading for the baseline.
"""
import numpy as np
import torch
import torch.nn.functional as F


def _flat(c, shape):
    return ((c[:, 0] * shape[0] + c[:, 1]) * shape[1] + c[:, 2]) * shape[2] + c[:, 3]


def subm_pairs(coords, shape):
    """coords int64 [M,4] (cuda) -> list of 27 (in_idx, out_idx) int64 pairs lists; offset order = SURVEY.md A.3"""
    M = coords.shape[0]
    keys = _flat(coords, shape)
    skeys, perm = torch.sort(keys)
    out = []
    ar = torch.arange(M, device=coords.device)
    for k0 in range(3):
        for k1 in range(3):
            for k2 in range(3):
                if (k0, k1, k2) == (1, 1, 1):
                    out.append(None)  # centre: dense GEMM
                    continue
                d = torch.tensor([0, k0 - 1, k1 - 1, k2 - 1], device=coords.device)
                n = coords + d
                ok = ((n[:, 1] >= 0) & (n[:, 1] < shape[0]) & (n[:, 2] >= 0) & (n[:, 2] < shape[1]) &
                      (n[:, 3] >= 0) & (n[:, 3] < shape[2]))
                nk = _flat(n, shape)
                pos = torch.searchsorted(skeys, nk).clamp_(max=M - 1)
                hit = ok & (skeys[pos] == nk)
                oi = ar[hit]                 # output site q
                ii = perm[pos[hit]]          # input site q + (k - 1)
                out.append((ii, oi))
    return out


def down_pairs(coords, shape):
    """k=2 s=2 p=0: -> (out_coords [M',4] ascending flat index, out_shape, list of 8 (in_idx, out_idx))"""
    oshape = [(s - 2) // 2 + 1 for s in shape]
    oc = coords.clone()
    oc[:, 1:] = coords[:, 1:] // 2
    keep = (oc[:, 1] < oshape[0]) & (oc[:, 2] < oshape[1]) & (oc[:, 3] < oshape[2])
    okeys = _flat(oc, oshape)
    uk = torch.unique(okeys[keep])  # sorted: spconv-CUDA's sort-unique output order
    orow = torch.searchsorted(uk, okeys).clamp_(max=max(uk.numel() - 1, 0))
    off = (coords[:, 1] % 2) * 4 + (coords[:, 2] % 2) * 2 + (coords[:, 3] % 2)
    ar = torch.arange(coords.shape[0], device=coords.device)
    pairs = []
    for k in range(8):
        m = keep & (off == k)
        pairs.append((ar[m], orow[m]))
    vol = oshape[0] * oshape[1] * oshape[2]
    b = uk // vol
    r = uk % vol
    out_coords = torch.stack([b, r // (oshape[1] * oshape[2]), (r // oshape[2]) % oshape[1], r % oshape[2]], 1)
    return out_coords, oshape, pairs


class _NativeConv(torch.autograd.Function):
    """indice_conv / indice_conv_backward of spconv v1.2 (ConvAlgo.Native) with torch ops"""

    @staticmethod
    def forward(ctx, feat, W, pairs, n_out, centre):
        K = W.shape[0]
        out = feat.new_zeros((n_out, W.shape[2]))
        if centre >= 0:
            torch.mm(feat, W[centre], out=out)
        for k in range(K):
            if k == centre or pairs[k] is None or pairs[k][0].numel() == 0:
                continue
            ii, oi = pairs[k]
            out.index_add_(0, oi, torch.mm(feat.index_select(0, ii), W[k]))
        ctx.save_for_backward(feat, W)
        ctx.pairs, ctx.centre = pairs, centre
        return out

    @staticmethod
    def backward(ctx, dout):
        feat, W = ctx.saved_tensors
        dout = dout.contiguous()
        pairs, centre = ctx.pairs, ctx.centre
        din = torch.zeros_like(feat)
        dW = torch.zeros_like(W)
        if centre >= 0:
            torch.mm(feat.t(), dout, out=dW[centre])
            torch.mm(dout, W[centre].t(), out=din)
        for k in range(W.shape[0]):
            if k == centre or pairs[k] is None or pairs[k][0].numel() == 0:
                continue
            ii, oi = pairs[k]
            a = feat.index_select(0, ii)
            g = dout.index_select(0, oi)
            torch.mm(a.t(), g, out=dW[k])
            din.index_add_(0, ii, torch.mm(g, W[k].t()))
        return din, dW, None, None, None


def _swap(pairs):
    return [None if p is None else (p[1], p[0]) for p in pairs]


class NativeUNet(object):
    """DODA's SparseConvNet (model/unet.py:58-69, model/unet_block.py:32-38,87-100) over the native algorithm"""

    def __init__(self, state_dict, device, nlevels=7, reps=2):
        self.dev = device
        self.sd = {k: v.detach().to(device).clone().requires_grad_(v.is_floating_point() and "running" not in k)
                   for k, v in state_dict.items()}
        self.nlevels, self.reps = nlevels, reps
        self.rb = None

    def params(self):
        return [v for v in self.sd.values() if v.requires_grad]

    def build_rulebooks(self, coords, shape):
        """13 rulebooks per forward like spconv: 7 SubM + 6 strided (each reads its sizes back to the host)"""
        rbs, c, s = [], coords.long(), [int(v) for v in shape]
        for l in range(self.nlevels):
            e = {"subm": subm_pairs(c, s), "n": c.shape[0]}
            if l + 1 < self.nlevels:
                oc, os_, dp = down_pairs(c, s)
                e["down"], e["n_down"] = dp, oc.shape[0]
                c, s = oc, os_
            rbs.append(e)
        self.rb = rbs
        return rbs

    def _w(self, name):
        w = self.sd[name]
        return w.reshape(-1, w.shape[-2], w.shape[-1])

    def _bn_relu(self, pre, x):
        return F.relu(F.batch_norm(x, None, None, self.sd[pre + ".weight"], self.sd[pre + ".bias"], True, 0.1, 1e-4))

    def _subm(self, name, x, l):
        return _NativeConv.apply(x, self._w(name), self.rb[l]["subm"], x.shape[0], 13)

    def _residual(self, pre, x, l):
        w_skip = self.sd.get(pre + ".i_branch.0.weight")
        skip = x if w_skip is None else torch.mm(x, w_skip.reshape(w_skip.shape[-2], w_skip.shape[-1]))
        h = self._subm(pre + ".conv_branch.2.weight", self._bn_relu(pre + ".conv_branch.0", x), l)
        h = self._subm(pre + ".conv_branch.5.weight", self._bn_relu(pre + ".conv_branch.3", h), l)
        return h + skip

    def _ublock(self, pre, x, l):
        for i in range(self.reps):
            x = self._residual("%s.blocks.block%d" % (pre, i), x, l)
        if l + 1 < self.nlevels:
            e = self.rb[l]
            d = _NativeConv.apply(self._bn_relu(pre + ".conv.0", x), self._w(pre + ".conv.2.weight"), e["down"],
                                  e["n_down"], -1)
            u = self._ublock(pre + ".u", d, l + 1)
            up = _NativeConv.apply(self._bn_relu(pre + ".deconv.0", u), self._w(pre + ".deconv.2.weight"),
                                   _swap(e["down"]), x.shape[0], -1)
            x = torch.cat((x, up), dim=1)
            for i in range(self.reps):
                x = self._residual("%s.blocks_tail.block%d" % (pre, i), x, l)
        return x

    def forward(self, voxel_feats, p2v, labels):
        x = _NativeConv.apply(voxel_feats, self._w("input_conv.0.weight"), self.rb[0]["subm"], voxel_feats.shape[0], 13)
        x = self._ublock("unet", x, 0)
        x = self._bn_relu("output_layer.0", x)
        scores = F.linear(x[p2v], self.sd["linear.weight"], self.sd["linear.bias"])
        return F.cross_entropy(scores, labels, ignore_index=255), scores

    def step(self, voxel_feats, p2v, labels):
        for p in self.params():
            p.grad = None
        loss, scores = self.forward(voxel_feats, p2v, labels)
        loss.backward()
        return loss, scores


def voxelize_mean(feats, v2p):
    """mode-4 voxelization with torch ops (the reference's voxelize_fp: mean of the voxel's points)"""
    cnt = v2p[:, 0].clamp(min=1).to(feats.dtype)
    out = feats.new_zeros((v2p.shape[0], feats.shape[1]))
    for j in range(1, v2p.shape[1]):
        m = v2p[:, 0] >= j
        out[m] += feats[v2p[m, j].long()]
    return out / cnt[:, None]


def measure(state_dict, batch, device, steps=5, warmup=2, graph=True):
    """-> dict(ms_eager, ms_rulebooks, ms_graph, launches?) for one fwd+bwd of the full net on `batch`"""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    net = NativeUNet(state_dict, device)
    coords = batch["voxel_locs"].to(device)
    feats = batch["feats"].to(device)
    v2p = batch["v2p_map"].to(device)
    p2v = batch["p2v_map"].to(device).long()
    labels = batch["labels"].to(device)
    shape = [int(s) for s in batch["spatial_shape"]]
    vf = voxelize_mean(feats, v2p)

    def ev():
        return torch.cuda.Event(enable_timing=True)

    # rulebooks (13 per forward, sizes read back to the host like spconv's strided builder does)
    for _ in range(max(1, warmup)):
        net.build_rulebooks(coords, shape)
    torch.cuda.synchronize()
    s, e = ev(), ev()
    s.record()
    for _ in range(steps):
        net.build_rulebooks(coords, shape)
    e.record()
    torch.cuda.synchronize()
    ms_rb = s.elapsed_time(e) / steps
    for _ in range(max(1, warmup)):
        loss, scores = net.step(vf, p2v, labels)
    torch.cuda.synchronize()
    s, e = ev(), ev()
    s.record()
    for _ in range(steps):
        net.step(vf, p2v, labels)
    e.record()
    torch.cuda.synchronize()
    ms_eager = s.elapsed_time(e) / steps
    out = {"ms_eager": ms_eager + ms_rb, "ms_rulebooks": ms_rb, "ms_graph": None,
           "loss": float(loss.detach()), "scores": scores.detach()}
    if graph:
        try:
            # fresh leaves whose first use (and so their AccumulateGrad nodes) is on the capture stream itself, warm-up and
            # capture on that one stream: a gradient accumulation that has to hop to another stream breaks the capture
            g = torch.cuda.CUDAGraph()
            net_g = NativeUNet(state_dict, device)
            net_g.rb = net.rb
            side = torch.cuda.Stream()
            torch.cuda.synchronize()
            with torch.cuda.stream(side):
                net_g.step(vf, p2v, labels)
            torch.cuda.synchronize()
            for p in net_g.params():
                p.grad = None
            # relaxed: other threads of the process (clock sampler, copy stream) may call the CUDA API during the capture
            with torch.cuda.graph(g, stream=side, capture_error_mode="relaxed"):
                l2, _ = net_g.forward(vf, p2v, labels)
                l2.backward()
            for _ in range(max(1, warmup)):
                g.replay()
            torch.cuda.synchronize()
            s, e = ev(), ev()
            s.record()
            for _ in range(steps):
                g.replay()
            e.record()
            torch.cuda.synchronize()
            out["ms_graph"] = s.elapsed_time(e) / steps + ms_rb
        except Exception as ex:  # graph capture is a favour to the baseline, not a requirement
            import traceback
            out["graph_error"] = repr(ex)[:200] + " | " + traceback.format_exc()[-700:]
    return out
