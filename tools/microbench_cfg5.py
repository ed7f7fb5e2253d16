"""BASELINE configs[4]: rulebook-build + gather / scatter-add + fused-conv microbench sweep,
M in {10k, 30k, 100k, 300k, 1M} voxels x occupancy in {0.5, 1, 2, 5} % x C in {16, 32, 64}, L2 flushed between timed
launches, CUDA events on the launch stream.  Algorithmic bytes exactly as SURVEY.md 8(d):

  rulebook SubM : 16 M + 8 P + 4 K                      (+ 8 M K for the engine's two tables, reported separately)
  gather        : P (4 + 4 C) read + 4 P C written       buf[i] = feat[pairs_in[i]]          (k_gather_rows)
  scatter-add   : P (4 + 4 C) read + 8 P C RMW           out[pairs_out[i]] += buf[i]          (k_scatter_add_rows)
  fused conv    : 4 (M Cin + M Cout) + 4 K Cin Cout + 8 P                                     (k_conv_direct / k_conv_tc)

usage (GPU box):  python tools/microbench_cfg5.py [--quick] > gpurun_out/cfg5.md
                  python tools/microbench_cfg5.py --one gather|scatter|rulebook|conv M occ C    (one launch x 3, for ncu)
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from doda_b200 import ops, scenes
from doda_b200._lib import lib, check

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


def gather(src, idx, out):
    check(lib.b200sp_gather_rows(src.data_ptr(), idx.data_ptr(), 0, idx.shape[0], src.shape[1], out.data_ptr(), ops._stream()),
          "gather_rows")


def scatter_add(src, idx, dst):
    check(lib.b200sp_scatter_add_rows(src.data_ptr(), idx.data_ptr(), 0, idx.shape[0], src.shape[1], dst.data_ptr(),
                                      ops._stream()), "scatter_add_rows")


def setup(M, occ):
    vox = scenes.uniform_scene(0, M, occ)
    coords = np.concatenate([np.zeros((vox.shape[0], 1), dtype=np.int64), vox], 1).astype(np.int32)
    shape = (coords[:, 1:].max(0) + 1).tolist()
    c = torch.from_numpy(coords).to(dev)
    rb = ops.build_rulebook(c, 1, shape, 3, 1, 1, 1, subm=True)
    pn = rb.pairnum.cpu().numpy()
    # the rulebook's pair lists, all offsets back to back (what spconv's per-offset gather / scatter walk)
    pin = torch.cat([rb.pairs[0, k, :pn[k]] for k in range(27)]).contiguous()
    pout = torch.cat([rb.pairs[1, k, :pn[k]] for k in range(27)]).contiguous()
    return c, shape, rb, pin, pout


def one(kind, M, occ, C):
    c, shape, rb, pin, pout = setup(M, occ)
    n, P = c.shape[0], pin.shape[0]
    x = torch.randn(n, C, device=dev)
    buf = torch.empty(P, C, device=dev)
    out = torch.zeros(n, C, device=dev)
    W3 = torch.randn(27, C, C, device=dev) * 0.1
    for _ in range(3):
        if kind == "gather":
            gather(x, pin, buf)
        elif kind == "scatter":
            scatter_add(buf, pout, out)
        elif kind == "rulebook":
            ops.build_rulebook(c, 1, shape, 3, 1, 1, 1, subm=True)
        else:
            ops.gather_gemm(x, W3, rb.nbr_perm, n, orow=rb.order, rowmask=rb.rowmask)
    torch.cuda.synchronize()
    print("done", kind, n, P, C)


def main():
    if "--one" in sys.argv:
        i = sys.argv.index("--one")
        return one(sys.argv[i + 1], int(sys.argv[i + 2]), float(sys.argv[i + 3]), int(sys.argv[i + 4]))
    quick = "--quick" in sys.argv
    Ms = (10000, 100000, 1000000) if quick else (10000, 30000, 100000, 300000, 1000000)
    occs = (0.005, 0.05) if quick else (0.005, 0.01, 0.02, 0.05)
    print("# cfg-5 microbench sweep (round 2): SubM 3x3x3 rulebook build, per-pair gather / scatter-add, fused conv forward\n")
    print("peak = %.1f GB/s (MEASURED_PEAKS.json hbm_gbs, of measured); L2 (126 MB) flushed with a 256 MB write between timed "
          "launches; median of 5; uniform-random scenes (doda_b200/scenes.py:uniform_scene).  `frac` = algorithmic GB/s / peak; "
          "small cases sit at their launch floor (~3-6 us), not at a bandwidth.\n" % peak)
    print("| M | occ % | P/M | rulebook ms | GB/s | frac | C | gather ms | GB/s | frac | scatter-add ms | GB/s | frac | conv fwd ms | GB/s | frac |")
    print("|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
    for M in Ms:
        for occ in occs:
            c, shape, rb, pin, pout = setup(M, occ)
            n, P = c.shape[0], pin.shape[0]
            t_rb = timeit(lambda: ops.build_rulebook(c, 1, shape, 3, 1, 1, 1, subm=True))
            b_rb = 16 * n + 8 * P + 4 * 27
            first = True
            for C in (16, 32, 64):
                x = torch.randn(n, C, device=dev)
                buf = torch.empty(P, C, device=dev)
                out = torch.zeros(n, C, device=dev)
                W3 = torch.randn(27, C, C, device=dev) * 0.1
                t_g = timeit(lambda: gather(x, pin, buf))
                t_s = timeit(lambda: scatter_add(buf, pout, out))
                t_c = timeit(lambda: ops.gather_gemm(x, W3, rb.nbr_perm, n, orow=rb.order, rowmask=rb.rowmask))
                b_g = P * (4 + 4 * C) + 4 * P * C
                b_s = P * (4 + 4 * C) + 8 * P * C
                b_c = 4 * (2 * n * C) + 4 * 27 * C * C + 8 * P
                f = lambda b, t: (b / t / 1e6, b / t / 1e6 / peak)
                print("| %s | %s | %s | %s | %s | %s | %d | %.4f | %.0f | %.3f | %.4f | %.0f | %.3f | %.4f | %.0f | %.3f |" % (
                    n if first else "", ("%.1f" % (100 * occ)) if first else "", ("%.2f" % (P / n)) if first else "",
                    ("%.3f" % t_rb) if first else "", ("%.0f" % (b_rb / t_rb / 1e6)) if first else "",
                    ("%.3f" % (b_rb / t_rb / 1e6 / peak)) if first else "", C, t_g, *f(b_g, t_g), t_s, *f(b_s, t_s), t_c, *f(b_c, t_c)))
                first = False
                del x, buf, out
            sys.stdout.flush()
    # devoxelize-shaped gather / scatter-add (model/unet.py:62): N points <- M voxels, the engine's two backward forms
    print("\n## devoxelize (features[p2v], model/unet.py:62) and its backward, N points over M voxels, C = 16\n")
    print("| M voxels | N points | gather ms | GB/s | frac | atomic scatter-add ms | GB/s | frac | segmented sum (v2p) ms | GB/s | frac |")
    print("|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
    for M in ((30000, 300000) if quick else (30000, 100000, 300000, 1000000)):
        b = scenes.collate([scenes.uniform_scene(1, M, 0.02)], dup_max=2)
        N, C = b["p2v_map"].shape[0], 16
        p2v, v2p = b["p2v_map"].to(dev), b["v2p_map"].to(dev)
        src = torch.randn(M, C, device=dev)
        g = torch.randn(N, C, device=dev)
        outp = torch.empty(N, C, device=dev)
        d = torch.zeros(M, C, device=dev)
        t_g = timeit(lambda: gather(src, p2v, outp))
        t_s = timeit(lambda: scatter_add(g, p2v, d))
        A = v2p.shape[1] - 1
        t_v = timeit(lambda: check(lib.b200sp_voxelize_fp(g.data_ptr(), d.data_ptr(), v2p.data_ptr(), 0, M, A, C, ops._stream()), "seg"))
        b_g = 4 * N + 4 * N * C + 4 * M * C
        b_s = 4 * N + 4 * N * C + 8 * M * C
        b_v = 4 * M * (1 + A) + 4 * N * C + 8 * M * C
        print("| %d | %d | %.4f | %.0f | %.3f | %.4f | %.0f | %.3f | %.4f | %.0f | %.3f |" % (
            M, N, t_g, b_g / t_g / 1e6, b_g / t_g / 1e6 / peak, t_s, b_s / t_s / 1e6, b_s / t_s / 1e6 / peak,
            t_v, b_v / t_v / 1e6, b_v / t_v / 1e6 / peak))
        sys.stdout.flush()


if __name__ == "__main__":
    main()
