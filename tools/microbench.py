"""BASELINE configs[4]: rulebook-build + fused conv (fwd / dgrad / wgrad) microbench sweep over voxel counts, occupancy
and channel widths; L2 flushed between timed launches.  Writes a markdown table (algorithmic GB/s against the measured
HBM peak) -- usage (GPU box): python tools/microbench.py > gpurun_out/microbench.md"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from doda_b200 import ops, scenes

dev = torch.device("cuda")
peak = 6548.8
pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = float(json.load(open(pk))["hbm_gbs"])
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=5):
    fn(); fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts))


print("# Microbench sweep (round 1): SubM 3x3x3 rulebook build and fused conv kernels, L2 flushed between launches\n")
print("peak = " + ("%.1f" % peak) + " GB/s (MEASURED_PEAKS.json, of measured).  Algorithmic bytes: rulebook 16M + 8P + 4K + 8MK (two tables); "
      "conv 4(M Cin + M Cout) + 4 K Cin Cout + 8 P; wgrad P(8 + 4Cin + 4Cout) + 4 K Cin Cout.  Kernels as the engine "
      "dispatches them: C = 16, 32 register-gather conv (conv_direct.cu), C = 16 table-form wgrad (wgrad_direct.cu), "
      "the rest tcgen05 (conv_tc.cu, wgrad_tc.cu).\n")
print("| scene | M | P/M | rulebook ms | GB/s | frac | C | fwd ms | GB/s | frac | dgrad ms | wgrad ms | GB/s | frac |")
print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
cases = [("uniform %.1f%%" % (100 * occ), M, occ) for M in (10000, 100000, 1000000) for occ in (0.005, 0.02, 0.05)]
cases += [("surface", M, None) for M in (30000, 150000, 300000, 600000)]
for name, M, occ in cases:
    vox = scenes.uniform_scene(0, M, occ) if occ else scenes.scene_with_voxels(0, M)
    coords = np.concatenate([np.zeros((vox.shape[0], 1), dtype=np.int64), vox], 1).astype(np.int32)
    shape = (coords[:, 1:].max(0) + 1).tolist()
    c = torch.from_numpy(coords).to(dev)
    t_rb = timeit(lambda: ops.build_rulebook(c, 1, shape, 3, 1, 1, 1, subm=True))
    rb = ops.build_rulebook(c, 1, shape, 3, 1, 1, 1, subm=True)
    n = c.shape[0]
    P = int(rb.pairnum.sum())
    b_rb = 16 * n + 8 * P + 4 * 27 + 8 * n * 27
    first = True
    for C in (16, 32, 64):
        x = torch.randn(n, C, device=dev)
        g = torch.randn(n, C, device=dev)
        W3 = torch.randn(27, C, C, device=dev) * 0.1
        t_f = timeit(lambda: ops.gather_gemm(x, W3, rb.nbr_perm, n, orow=rb.order, rowmask=rb.rowmask))
        t_d = timeit(lambda: ops.gather_gemm(g, W3, rb.nbr_perm, n, orow=rb.order, wflags=ops.W_T_MIRROR, rowmask=rb.rowmask))
        # the engine's own dispatch: table form (wgrad_direct.cu) where it is used, pair lists (wgrad_tc.cu) elsewhere
        t_w = timeit(lambda: ops.conv_backward_raw("subm", x, W3, g, rb, None, False, True))
        b_c = 4 * (2 * n * C) + 4 * 27 * C * C + 8 * P
        b_w = P * (8 + 8 * C) + 4 * 27 * C * C
        gf, gw = b_c / t_f / 1e6, b_w / t_w / 1e6
        print("| %s | %d | %.1f | %s | %s | %s | %d | %.3f | %.0f | %.3f | %.3f | %.3f | %.0f | %.3f |" % (
            name if first else "", n, P / n, ("%.3f" % t_rb) if first else "", ("%.0f" % (b_rb / t_rb / 1e6)) if first else "",
            ("%.3f" % (b_rb / t_rb / 1e6 / peak)) if first else "", C, t_f, gf, gf / peak, t_d, t_w, gw, gw / peak))
        first = False
    sys.stdout.flush()
