"""dev helper (GPU box): tensor-core conv path vs fp64 reference + timing vs the fp32 CUDA-core kernel."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
from helpers import rel_err, random_coords, surface_coords
from doda_b200 import ops

dev = torch.device("cuda")


def ref_gather_gemm(feat, W3, tab):
    f = feat.double().cpu(); W = W3.double().cpu(); t = tab.cpu().long()
    out = torch.zeros(t.shape[0], W.shape[2], dtype=torch.float64)
    for k in range(W.shape[0]):
        m = t[:, k] >= 0
        out[m] += f[t[m, k]] @ W[k]
    return out


def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n


def case(M, Cin, Cout, check=True, surface=False):
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
    W3 = torch.randn(27, Cin, Cout, device=dev) * 0.2
    res = {}
    for impl in ("tc", "fp32"):
        ops.set_conv_impl(impl)
        try:
            out = ops.gather_gemm(feat, W3, rb.nbr, n)
            torch.cuda.synchronize()
        except Exception as ex:
            print("  %s FAILED: %s" % (impl, ex)); continue
        err = rel_err(out, ref_gather_gemm(feat, W3, rb.nbr)) if check else float("nan")
        # dgrad flavour
        g = torch.randn(n, Cout, device=dev)
        din = ops.gather_gemm(g, W3, rb.nbr, n, wflags=ops.W_T_MIRROR)
        if check:
            Wt = torch.flip(W3, [0]).transpose(1, 2).contiguous()
            err2 = rel_err(din, ref_gather_gemm(g, Wt, rb.nbr))
        else:
            err2 = float("nan")
        ms = timeit(lambda: ops.gather_gemm(feat, W3, rb.nbr, n))
        outm = ops.gather_gemm(feat, W3, rb.nbr_perm, n, orow=rb.order, rowmask=rb.rowmask)
        errm = float((outm - out).abs().max() / out.abs().max())
        msm = timeit(lambda: ops.gather_gemm(feat, W3, rb.nbr_perm, n, orow=rb.order, rowmask=rb.rowmask))
        res[impl] = (err, err2, ms, errm, msm)
    P = int((rb.nbr >= 0).sum())
    t = (rb.nbr_perm[: (n // 128) * 128].view(-1, 128, 27) >= 0).any(1).sum(1).float().mean().item()
    t0 = (rb.nbr[: (n // 128) * 128].view(-1, 128, 27) >= 0).any(1).sum(1).float().mean().item()
    print("   active offsets per 128-row tile: row order %.1f, morton %.1f" % (t0, t))
    print("M=%d Cin=%d Cout=%d P/M=%.1f  " % (n, Cin, Cout, P / n) +
          "  ".join("%s: fwd %.1e dgrad %.1e %.3f ms | morton: diff %.1e %.3f ms" % (k, *v) for k, v in res.items()), flush=True)
    ops.set_conv_impl("tc")


if __name__ == "__main__":
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    case(1000, 16, 16)
    case(4000, 16, 16)
    case(4000, 32, 32)
    case(4000, 3, 16)
    case(3000, 48, 48)
    case(3000, 5, 7)
    case(3000, 96, 48)
    case(3000, 112, 112)
    if not quick:
        case(300000, 16, 16, check=False, surface=True)
        case(300000, 32, 16, check=False, surface=True)
        case(120000, 32, 32, check=False, surface=True)
        case(120000, 64, 32, check=False, surface=True)
        case(26000, 48, 48, check=False, surface=True)
        case(26000, 96, 48, check=False, surface=True)
