"""dev helper (GPU box): the out-stationary tcgen05 weight gradient (wgrad_os.cu, through ops.wgrad_table) against an
fp64 reference and against the pair-list kernel (wgrad_tc.cu, through ops.wgrad), with per-launch times."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from helpers import rel_err, surface_coords
from doda_b200 import ops
from doda_b200._lib import lib
dev = torch.device("cuda")


def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3


def ref_wgrad(a, g, tab, K):
    a64, g64, t = a.double().cpu(), g.double().cpu(), tab.cpu().long()
    out = torch.zeros(K, a.shape[1], g.shape[1], dtype=torch.float64)
    for k in range(K):
        m = t[:, k] >= 0
        out[k] = a64[t[m, k]].t() @ g64[m]
    return out


cases = [(20000, 32, 32), (20000, 64, 32), (20000, 48, 48), (9000, 96, 48), (6149, 64, 64), (6149, 128, 64), (1381, 80, 80),
         (1381, 160, 80), (223, 96, 96), (223, 192, 96), (45, 112, 112), (20000, 16, 32)]
if len(sys.argv) > 1 and sys.argv[1] == "l1":
    cases = [(300000, 16, 16), (300000, 32, 16), (300000, 16, 32), (117656, 32, 32)]
elif len(sys.argv) > 1 and sys.argv[1] == "big":
    cases = [(117656, 32, 32), (117656, 64, 32), (26506, 48, 48), (26506, 96, 48), (300000, 32, 16)]
for M, Ca, Cb in cases:
    torch.manual_seed(0)
    coords, shape = surface_coords(0, max(M // 2, 20), 2)
    c = torch.from_numpy(coords).to(dev)
    rb = ops.build_rulebook(c, 2, shape, 3, 1, 1, 1, subm=True)
    n = c.shape[0]
    a = torch.randn(n, Ca, device=dev)
    g = torch.randn(n, Cb, device=dev)
    try:
        dW = torch.zeros(27, Ca, Cb, device=dev)
        ops.wgrad_table(a, g, rb.nbr_perm, n, 27, orow=rb.order, rowmask=rb.rowmask, out=dW)
        torch.cuda.synchronize()
        name = lib.b200sp_last_kernel().decode()
        # rows in processing order: tab row r <-> g row order[r]
        err = rel_err(dW, ref_wgrad(a, g[rb.order.long()], rb.nbr_perm, 27)) if n <= 30000 else float("nan")
        dW2 = torch.zeros(27, Ca, Cb, device=dev)
        ops.wgrad(a, g, rb.pairs[0], rb.pairs[1], rb.pairnum, n, 27, out=dW2)
        torch.cuda.synchronize()
        name2 = lib.b200sp_last_kernel().decode()
        scratch = torch.zeros(27, Ca, Cb, device=dev)
        t1 = timeit(lambda: ops.wgrad_table(a, g, rb.nbr_perm, n, 27, orow=rb.order, rowmask=rb.rowmask, out=scratch))
        t2 = timeit(lambda: ops.wgrad(a, g, rb.pairs[0], rb.pairs[1], rb.pairnum, n, 27, out=scratch))
        print("rows %6d %3dx%3d: %s err %.2e vs pair-list %.2e | %s %.1f us, %s %.1f us" %
              (n, Ca, Cb, name, err, rel_err(dW, dW2), name, t1, name2, t2), flush=True)
    except Exception as ex:
        print("rows %6d %3dx%3d FAILED: %s" % (n, Ca, Cb, ex), flush=True)
# strided (K = 8) and 1x1 tables
for M, Ca, Cb in ((20000, 32, 48), (6000, 64, 80)):
    coords, shape = surface_coords(1, M // 2, 2)
    c = torch.from_numpy(coords).to(dev)
    rb = ops.build_rulebook(c, 2, shape, 2, 2, 0, 1, subm=False)
    n, nc = c.shape[0], rb.outids.shape[0]
    a = torch.randn(n, Ca, device=dev); g = torch.randn(nc, Cb, device=dev)
    dW = torch.zeros(8, Ca, Cb, device=dev)
    ops.wgrad_table(a, g, rb.bwd, nc, 8, out=dW)
    name = lib.b200sp_last_kernel().decode()
    print("strided rows %d -> %d %dx%d: %s err %.2e" % (n, nc, Ca, Cb, name, rel_err(dW, ref_wgrad(a, g, rb.bwd, 8))), flush=True)
    a2 = torch.randn(nc, Cb, device=dev); g2 = torch.randn(n, Ca, device=dev)
    dW = torch.zeros(8, Cb, Ca, device=dev)
    ops.wgrad_table(a2, g2, rb.fwd, n, 8, out=dW)
    print("inverse %dx%d: %s err %.2e" % (Cb, Ca, lib.b200sp_last_kernel().decode(), rel_err(dW, ref_wgrad(a2, g2, rb.fwd, 8))), flush=True)
    dW = torch.zeros(1, Ca, Cb, device=dev)
    g3 = torch.randn(n, Cb, device=dev)
    ops.wgrad_table(a, g3, None, n, 1, out=dW)
    print("1x1 %dx%d: %s err %.2e" % (Ca, Cb, lib.b200sp_last_kernel().decode(), rel_err(dW[0], a.double().t() @ g3.double())), flush=True)
