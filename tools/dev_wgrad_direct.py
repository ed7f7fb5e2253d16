"""dev helper (GPU box): table-form wgrad (wgrad_direct.cu) vs the pair-list tcgen05 wgrad: time per launch on the
U-Net's level shapes, and the difference between the two."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from helpers import surface_coords
from doda_b200 import ops
from dev_conv_tc import timeit
dev = torch.device("cuda")
for (M, ca, cb) in ((300000, 16, 16), (300000, 32, 16), (118000, 32, 32), (118000, 16, 32)):
    coords, shape = surface_coords(0, M // 2, 2)
    c = torch.from_numpy(coords).to(dev)
    rb = ops.build_rulebook(c, 2, shape, 3, 1, 1, 1, subm=True)
    n = c.shape[0]
    a = torch.randn(n, ca, device=dev); g = torch.randn(n, cb, device=dev)
    f_t = lambda: ops.wgrad_table(a, g, rb.nbr_perm, n, 27, orow=rb.order, rowmask=rb.rowmask)
    f_p = lambda: ops.wgrad(a, g, rb.pairs[0], rb.pairs[1], rb.pairnum, n, 27)
    d1, d0 = f_t().clone(), f_p().clone()
    torch.cuda.synchronize()
    print("M=%6d %2dx%2d  table %.1f us  pairs(tcgen05) %.1f us  max diff %.1e" % (
        n, ca, cb, timeit(f_t) * 1e3, timeit(f_p) * 1e3, float((d1 - d0).abs().max() / d0.abs().max())), flush=True)
