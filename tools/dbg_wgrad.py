import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from doda_b200 import ops
dev = torch.device("cuda")
torch.manual_seed(0)
n, Ca, Cb = 256, 16, 16
a = torch.randn(n, Ca, device=dev); b = torch.randn(n, Cb, device=dev)
ref = a.double().t() @ b.double()
dW = ops.wgrad(a, b, None, None, None, n, 1)[0].double()
torch.cuda.synchronize()
print("variant", os.environ.get("B200SP_WG_VARIANT"), "max|dW| %.3e max|ref| %.3e" % (dW.abs().max(), ref.abs().max()))
print("err vs ref %.2e, vs ref^T %.2e" % ((dW-ref).abs().max()/ref.abs().max(), (dW-ref.t()).abs().max()/ref.abs().max()))
print("nonzero frac", float((dW != 0).float().mean()))
# structured probe: a = one-hot rows to see which (pair, channel) lands where
a2 = torch.zeros(n, Ca, device=dev); b2 = torch.zeros(n, Cb, device=dev)
a2[3, 5] = 1.0; b2[3, 7] = 2.0
d2 = ops.wgrad(a2, b2, None, None, None, n, 1)[0]
nz = d2.nonzero().tolist()
print("probe a[3,5]=1,b[3,7]=2 -> nonzeros", nz[:8], [float(d2[i, j]) for i, j in nz[:8]])
a2.zero_(); b2.zero_(); a2[:, 0] = 1.0; b2[:, 0] = 1.0
d3 = ops.wgrad(a2, b2, None, None, None, n, 1)[0]
print("probe all-ones col0: dW[0,0]=%.1f (expect %d), sum=%.1f" % (float(d3[0, 0]), n, float(d3.sum())))
