"""every PG_OP / pointops2_cuda replacement kernel timed NEXT TO the reference's own kernel
(the reference sources compiled unmodified for sm_100a into oracle/_ref/ by oracle/build_ref.py), same inputs, same
caller-allocated outputs, CUDA events around `reps` back-to-back calls after warm-up (device-wide sync on both sides:
the reference launches on the legacy default stream).  Prints a markdown table: time per call and algorithmic GB/s.

usage (GPU box): python tools/bench_refkernels.py > gpurun_out/refkernels.md"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import build_ref
from doda_b200 import pg_op as mine_pg, pointops2_cuda as mine_po, pointgroup_ops

pg, po = build_ref.load("PG_OP"), build_ref.load("pointops2_cuda")
assert pg is not None and po is not None, "oracle/_ref/*.so not built"
dev = torch.device("cuda")


def timeit(fn, reps=10):
    fn(); fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps * 1e3  # us


rows = []


def row(name, cite, size, nbytes, f_ref, f_mine, reps=10):
    t_ref, t_mine = timeit(f_ref, reps), timeit(f_mine, reps)
    rows.append("| `%s` | %s | %s | %.1f | %.0f | %.1f | %.0f | %.1fx |" % (
        name, cite, size, t_ref, nbytes / t_ref / 1e3 if nbytes else 0, t_mine, nbytes / t_mine / 1e3 if nbytes else 0, t_ref / t_mine))
    print(rows[-1], flush=True)


def batched_points(n_per, B, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    xyz = torch.rand(n_per * B, 3, generator=g) * scale
    return xyz.cuda(), torch.arange(B).repeat_interleave(n_per).int().cuda(), torch.arange(0, (B + 1) * n_per, n_per).int().cuda()


print("# PG_OP / pointops2_cuda replacement kernels next to the reference's own kernels (round 2)\n")
print("Reference = `oracle/_ref/PG_OP.so`, `oracle/_ref/pointops2_cuda.so`: the reference sources compiled unmodified for sm_100a. "
      "GB/s = algorithmic bytes (inputs read once + outputs written once) / time.\n")
print("| kernel | reference file:line | size | reference us | GB/s | engine us | GB/s | speed-up |")
print("|---|---|---|---:|---:|---:|---:|---:|")

# ---- voxelize_fp / bp, point_recover: 450 k points -> 300 k voxels, C = 3 (DODA's input) and C = 16
rng = np.random.RandomState(0)
locs = torch.from_numpy(np.concatenate([np.sort(rng.randint(0, 2, size=(450000, 1)), 0), rng.randint(0, 110, size=(450000, 3))], 1)).long().contiguous()
_, p2v, v2p = pointgroup_ops.voxelization_idx(locs, 2, 4)
M, A = v2p.shape[0], v2p.shape[1] - 1
v2p_d = v2p.cuda()
for C in (3, 16):
    feats = torch.randn(450000, C).cuda()
    o = torch.zeros(M, C, device=dev)
    d = torch.zeros(450000, C, device=dev)
    g = torch.randn(M, C).cuda()
    nb = 4 * 450000 * C + 4 * M * C + 4 * M * (1 + A)
    row("voxelize_fp C=%d" % C, "voxelize.cu:10-30", "N=450k M=%dk" % (M // 1000), nb,
        lambda: pg.voxelize_fp(feats, o, v2p_d, 4, M, A, C), lambda: mine_pg.voxelize_fp(feats, o, v2p_d, 4, M, A, C))
    row("voxelize_bp C=%d" % C, "voxelize.cu:35-52", "N=450k M=%dk" % (M // 1000), nb,
        lambda: pg.voxelize_bp(g, d, v2p_d, 4, M, A, C), lambda: mine_pg.voxelize_bp(g, d, v2p_d, 4, M, A, C))
r = torch.zeros(450000, 16, device=dev)
row("point_recover_fp C=16", "voxelize.cpp:185-194", "N=450k", 4 * 450000 * 16 + 4 * M * 16 + 4 * M * (1 + A),
    lambda: pg.point_recover_fp(g, r, v2p_d, M, A, 16), lambda: mine_pg.point_recover_fp(g, r, v2p_d, M, A, 16))

# ---- sec_mean / min / max: 200 k rows, C = 16, 2000 segments
N, C, P = 200000, 16, 2000
inp = torch.randn(N, C).cuda()
cuts = np.sort(rng.choice(np.arange(1, N), P - 1, replace=False))
offs = torch.from_numpy(np.concatenate([[0], cuts, [N]]).astype(np.int32)).cuda()
out = torch.zeros(P, C, device=dev)
nb = 4 * N * C + 4 * P * C + 4 * (P + 1)
for nm, cite in (("sec_mean", "sec_mean.cu:12-33"), ("sec_min", "sec_mean.cu:62-83"), ("sec_max", "sec_mean.cu:88-109")):
    row(nm, cite, "N=200k C=16 P=2000", nb, lambda nm=nm: getattr(pg, nm)(inp, offs, out, P, C), lambda nm=nm: getattr(mine_pg, nm)(inp, offs, out, P, C))
dinp = torch.zeros(N, C, device=dev)
# NB argument order of the reference binding: (d_inp [N,C] OUTPUT, offsets, d_out [P,C]) -- functions/pointgroup_ops.py:287
row("sec_mean_bp", "sec_mean.cu:37-57", "N=200k C=16 P=2000", nb, lambda: pg.sec_mean_bp(dinp, offs, out, P, C), lambda: mine_pg.sec_mean_bp(dinp, offs, out, P, C))

# ---- roipool / get_iou
feats = torch.randn(N, C).cuda()
o2 = torch.zeros(P, C, device=dev)
mi = torch.zeros(P, C, dtype=torch.int32, device=dev)
row("roipool_fp", "roipool.cu:12-31", "N=200k C=16 P=2000", nb, lambda: pg.roipool_fp(feats, offs, o2, mi, P, C), lambda: mine_pg.roipool_fp(feats, offs, o2, mi, P, C))

# ---- ballquery_batch_p: 2 x 20 k points
xyz, bidx, boff = batched_points(20000, 2, 0)
n, mean_active, radius = xyz.shape[0], 60, 0.03
idx = torch.zeros(n * mean_active, dtype=torch.int32, device=dev)
sl = torch.zeros(n, 2, dtype=torch.int32, device=dev)
row("ballquery_batch_p", "bfs_cluster.cu:15-89", "n=2x20k r=0.03", 12 * n + 8 * n + 4 * n * 20,
    lambda: pg.ballquery_batch_p(xyz, bidx, boff, idx, sl, n, mean_active, radius),
    lambda: mine_pg.ballquery_batch_p(xyz, bidx, boff, idx, sl, n, mean_active, radius), reps=5)

# ---- knn_batch (PG_OP): n = 2 x 20 k, m = 2 x 2 k queries, k = 8
qxyz, _, qoff = batched_points(2000, 2, 2)
m, k = qxyz.shape[0], 8
kidx = torch.zeros(n, k, dtype=torch.int32, device=dev)
row("knn_batch", "knn.cu:7-71", "n=40k m=4k k=8", 12 * (n + m) + 4 * n * k, lambda: pg.knn_batch(xyz, qxyz, bidx, qoff, kidx, n, m, k),
    lambda: mine_pg.knn_batch(xyz, qxyz, bidx, qoff, kidx, n, m, k), reps=5)

# ---- pointops2: knnquery (nsample 1 = DODA's label broadcast, and 16), FPS, grouping, interpolation, subtraction, aggregation
off1 = boff[1:].contiguous()
for ns in (1, 16):
    kid = torch.zeros(n, ns, dtype=torch.int32, device=dev)
    kd = torch.zeros(n, ns, device=dev)
    row("knnquery_cuda nsample=%d" % ns, "knnquery_cuda_kernel.cu:65-115", "n=m=2x20k", 24 * n + 8 * n * ns,
        lambda: po.knnquery_cuda(n, ns, xyz, xyz, off1, off1, kid, kd), lambda: mine_po.knnquery_cuda(n, ns, xyz, xyz, off1, off1, kid, kd), reps=3)
n_off = torch.tensor([5000, 10000], dtype=torch.int32).cuda()
fidx = torch.zeros(10000, dtype=torch.int32, device=dev)


def fps(mod):
    tmp = torch.full((n,), 1e10, device=dev)
    mod.furthestsampling_cuda(2, 20000, xyz, off1, n_off, tmp, fidx)


row("furthestsampling_cuda", "sampling_cuda_kernel.cu:15-", "2x20k -> 2x5k", 0, lambda: fps(po), lambda: fps(mine_po), reps=2)
ns, c, wc = 16, 32, 8
gi = torch.zeros(n, ns, dtype=torch.int32, device=dev)
gd = torch.zeros(n, ns, device=dev)
po.knnquery_cuda(n, ns, xyz, xyz, off1, off1, gi, gd)
torch.cuda.synchronize()
inp, inp2 = torch.randn(n, c).cuda(), torch.randn(n, c).cuda()
pos, w = torch.randn(n, ns, c).cuda(), torch.randn(n, ns, wc).cuda()
g3, g2 = torch.randn(n, ns, c).cuda(), torch.randn(n, c).cuda()
o3, o2 = torch.zeros(n, ns, c, device=dev), torch.zeros(n, c, device=dev)
o2b = torch.zeros(n, c, device=dev)
gp, gw = torch.zeros(n, ns, c, device=dev), torch.zeros(n, ns, wc, device=dev)
w3 = torch.rand(n, 3).cuda()
i3 = gi[:, :3].contiguous()
b3 = 4 * n * ns * c
row("grouping_forward", "grouping_cuda_kernel.cu:5-14", "n=40k ns=16 c=32", b3 * 2 + 4 * n * ns, lambda: po.grouping_forward_cuda(n, ns, c, inp, gi, o3), lambda: mine_po.grouping_forward_cuda(n, ns, c, inp, gi, o3))
row("grouping_backward", "grouping_cuda_kernel.cu:16-25", "n=40k ns=16 c=32", b3 * 2 + 4 * n * ns, lambda: po.grouping_backward_cuda(n, ns, c, g3, gi, o2), lambda: mine_po.grouping_backward_cuda(n, ns, c, g3, gi, o2))
row("interpolation_forward", "interpolation_cuda_kernel.cu:5-18", "n=40k k=3 c=32", 4 * n * c * 4 + 24 * n, lambda: po.interpolation_forward_cuda(n, c, 3, inp, i3, w3, o2), lambda: mine_po.interpolation_forward_cuda(n, c, 3, inp, i3, w3, o2))
row("interpolation_backward", "interpolation_cuda_kernel.cu:20-33", "n=40k k=3 c=32", 4 * n * c * 4 + 24 * n, lambda: po.interpolation_backward_cuda(n, c, 3, g2, i3, w3, o2), lambda: mine_po.interpolation_backward_cuda(n, c, 3, g2, i3, w3, o2))
row("subtraction_forward", "subtraction_cuda_kernel.cu:5-16", "n=40k ns=16 c=32", b3 * 2 + 4 * n * ns, lambda: po.subtraction_forward_cuda(n, ns, c, inp, inp2, gi, o3), lambda: mine_po.subtraction_forward_cuda(n, ns, c, inp, inp2, gi, o3))
row("subtraction_backward", "subtraction_cuda_kernel.cu:18-30", "n=40k ns=16 c=32", b3 * 3 + 4 * n * ns, lambda: po.subtraction_backward_cuda(n, ns, c, gi, g3, o2, o2b), lambda: mine_po.subtraction_backward_cuda(n, ns, c, gi, g3, o2, o2b))
row("aggregation_forward", "aggregation_cuda_kernel.cu:5-20", "n=40k ns=16 c=32 w_c=8", b3 * 2 + 4 * n * ns * wc, lambda: po.aggregation_forward_cuda(n, ns, c, wc, inp, pos, w, gi, o2), lambda: mine_po.aggregation_forward_cuda(n, ns, c, wc, inp, pos, w, gi, o2))
row("aggregation_backward", "aggregation_cuda_kernel.cu:22-39", "n=40k ns=16 c=32 w_c=8", b3 * 3 + 8 * n * ns * wc, lambda: po.aggregation_backward_cuda(n, ns, c, wc, inp, pos, w, gi, g2, o2, gp, gw), lambda: mine_po.aggregation_backward_cuda(n, ns, c, wc, inp, pos, w, gi, g2, o2, gp, gw))
