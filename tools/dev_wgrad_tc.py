"""dev helper (GPU box): tensor-core wgrad vs fp64 reference + timing vs the fp32 CUDA-core kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
from helpers import rel_err, random_coords, surface_coords
from doda_b200 import ops
dev = torch.device("cuda")

def timeit(fn, n=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n

def case(M, Ca, Cb, check=True, surface=False, dense=False):
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
    a = torch.randn(n, Ca, device=dev)
    b = torch.randn(n, Cb, device=dev)
    res = {}
    for impl in ("tc", "fp32"):
        ops.set_conv_impl(impl)
        try:
            if dense:
                dW = ops.wgrad(a, b, None, None, None, n, 1)
            else:
                dW = ops.wgrad(a, b, rb.pairs[0], rb.pairs[1], rb.pairnum, n, 27)
            torch.cuda.synchronize()
        except Exception as ex:
            print("  %s FAILED: %s" % (impl, ex)); continue
        err = float("nan")
        if check:
            ad, bd = a.double().cpu(), b.double().cpu()
            if dense:
                ref = (ad.t() @ bd)[None]
            else:
                pr, pn = rb.pairs.cpu().long(), rb.pairnum.cpu()
                ref = torch.zeros(27, Ca, Cb, dtype=torch.float64)
                for k in range(27):
                    nk = int(pn[k])
                    ref[k] = ad[pr[0, k, :nk]].t() @ bd[pr[1, k, :nk]]
            err = rel_err(dW, ref)
        if dense:
            ms = timeit(lambda: ops.wgrad(a, b, None, None, None, n, 1))
        else:
            ms = timeit(lambda: ops.wgrad(a, b, rb.pairs[0], rb.pairs[1], rb.pairnum, n, 27))
        res[impl] = (err, ms)
    print("M=%d Ca=%d Cb=%d dense=%d  " % (n, Ca, Cb, dense) + "  ".join("%s: err %.1e %.3f ms" % (k, *v) for k, v in res.items()), flush=True)
    ops.set_conv_impl("tc")

if __name__ == "__main__":
    case(1000, 16, 16)
    case(4000, 16, 16)
    case(4000, 32, 32)
    case(4000, 3, 16)
    case(3000, 48, 48)
    case(3000, 5, 7)
    case(3000, 96, 48)
    case(3000, 112, 112)
    case(3000, 64, 32, dense=True)
    case(3000, 160, 160)
    case(300000, 16, 16, check=False, surface=True)
    case(300000, 32, 16, check=False, surface=True)
    case(120000, 32, 32, check=False, surface=True)
    case(120000, 64, 32, check=False, surface=True)
    case(26000, 48, 48, check=False, surface=True)
    case(26000, 96, 48, check=False, surface=True)
