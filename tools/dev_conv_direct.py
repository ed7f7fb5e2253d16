"""dev helper (GPU box): register-gather kernel (conv_direct.cu) vs the persistent tcgen05 kernel -- accuracy against
an fp64 reference on small cases, time per launch on the U-Net's level shapes (mask-sorted tables, fwd and dgrad)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from helpers import rel_err, random_coords, surface_coords
from doda_b200 import ops
from doda_b200._lib import lib
from dev_conv_tc import ref_gather_gemm, timeit

dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit_cold(fn, n=10):
    for _ in range(2): fn()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return tot / n


def case(M, Cin, Cout, check, surface=True, K=27):
    torch.manual_seed(0)
    if surface:
        coords, shape = surface_coords(0, M // 2, 2)
    else:
        side = int(round((M / 2 / 0.03) ** (1 / 3))) + 2
        shape = [side] * 3
        coords = random_coords(0, M // 2, 2, shape)
    c = torch.from_numpy(coords).to(dev)
    rb = ops.build_rulebook(c, 2, shape, 3, 1, 1, 1, subm=True)
    n = c.shape[0]
    feat = torch.randn(n, Cin, device=dev)
    g = torch.randn(n, Cout, device=dev)
    W3 = torch.randn(27, Cin, Cout, device=dev) * 0.2
    line = "M=%6d %3d->%3d covers=%d" % (n, Cin, Cout, lib.b200sp_conv_direct_covers(27, Cin, Cout))
    outs = {}
    for on in (1, 0):
        lib.b200sp_set_conv_direct(on)
        f_row = lambda: ops.gather_gemm(feat, W3, rb.nbr, n)
        f_fwd = lambda: ops.gather_gemm(feat, W3, rb.nbr_perm, n, orow=rb.order, rowmask=rb.rowmask)
        f_dg = lambda: ops.gather_gemm(g, W3, rb.nbr_perm, n, wflags=ops.W_T_MIRROR, orow=rb.order, rowmask=rb.rowmask)
        o_row, o_fwd, o_dg = f_row(), f_fwd(), f_dg()
        torch.cuda.synchronize()
        outs[on] = (o_row, o_fwd, o_dg)
        tag = "direct" if on else "tcgen05"
        if check:
            Wt = torch.flip(W3, [0]).transpose(1, 2).contiguous()
            e1 = rel_err(o_row, ref_gather_gemm(feat, W3, rb.nbr)); e2 = rel_err(o_fwd, ref_gather_gemm(feat, W3, rb.nbr))
            e3 = rel_err(o_dg, ref_gather_gemm(g, Wt, rb.nbr))
            line += " | %s err row %.1e sorted %.1e dgrad %.1e" % (tag, e1, e2, e3)
        else:
            line += " | %s fwd %.1f us (cold %.1f) dgrad %.1f us row-order %.1f us" % (
                tag, timeit(f_fwd) * 1e3, timeit_cold(f_fwd) * 1e3, timeit(f_dg) * 1e3, timeit(f_row) * 1e3)
    d = max(float((a - b).abs().max() / b.abs().max()) for a, b in zip(outs[1], outs[0]))
    print(line + " | direct vs tcgen05 max diff %.1e" % d, flush=True)
    lib.b200sp_set_conv_direct(1)


if __name__ == "__main__":
    for (M, ci, co) in ((1000, 16, 16), (4000, 32, 32), (3000, 48, 48), (3000, 64, 64), (3000, 32, 16), (3000, 16, 48),
                        (3000, 64, 32), (2000, 48, 64)):
        case(M, ci, co, True, surface=False)
    for (M, ci, co) in ((300000, 16, 16), (300000, 32, 16), (118000, 32, 32), (118000, 64, 32), (26500, 48, 48),
                        (6100, 64, 64)):
        case(M, ci, co, False)
