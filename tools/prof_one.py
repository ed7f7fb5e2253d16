"""dev helper: run ONE conv shape a few times (for ncu captures).  usage: prof_one.py M Cin Cout [impl] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from helpers import surface_coords
from doda_b200 import ops
M, Cin, Cout = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
impl = sys.argv[4] if len(sys.argv) > 4 else "tc"
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
dev = torch.device("cuda")
coords, shape = surface_coords(0, M // 2, 2)
c = torch.from_numpy(coords).to(dev)
rb = ops.build_rulebook(c, 2, shape, 3, 1, 1, 1, subm=True)
n = c.shape[0]
feat = torch.randn(n, Cin, device=dev)
W3 = torch.randn(27, Cin, Cout, device=dev) * 0.2
g = torch.randn(n, Cout, device=dev)
ops.set_conv_impl(impl)
for _ in range(reps):
    out = ops.gather_gemm(feat, W3, rb.nbr_perm, n, orow=rb.order, rowmask=rb.rowmask)
    din = ops.gather_gemm(g, W3, rb.nbr_perm, n, wflags=ops.W_T_MIRROR, orow=rb.order, rowmask=rb.rowmask)
    dW = ops.wgrad(feat, g, rb.pairs[0], rb.pairs[1], rb.pairnum, n, 27)
    if ops._wgrad_table_covers(27, Cin, Cout):
        dW = ops.wgrad_table(feat, g, rb.nbr_perm, n, 27, orow=rb.order, rowmask=rb.rowmask)
torch.cuda.synchronize()
print("done", out.shape, dW.shape)
