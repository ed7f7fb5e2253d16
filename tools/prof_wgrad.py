"""dev helper: run ONE weight-gradient shape a few times through the table form (for ncu captures).
usage: prof_wgrad.py M Ca Cb [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from helpers import surface_coords
from doda_b200 import ops
M, Ca, Cb = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
dev = torch.device("cuda")
coords, shape = surface_coords(0, M // 2, 2)
c = torch.from_numpy(coords).to(dev)
rb = ops.build_rulebook(c, 2, shape, 3, 1, 1, 1, subm=True)
n = c.shape[0]
a = torch.randn(n, Ca, device=dev)
g = torch.randn(n, Cb, device=dev)
dW = torch.zeros(27, Ca, Cb, device=dev)
for _ in range(reps):
    ops.wgrad_table(a, g, rb.nbr_perm, n, 27, orow=rb.order, rowmask=rb.rowmask, out=dW)
torch.cuda.synchronize()
print("done", n)
