"""dev helper (GPU box): deep-level conv shapes, the deterministic split-K of k_conv_tc vs no split (B200SP_TC_NOSPLIT
is read once per process -> run this script twice) -- per-launch time with CUDA events over 50 back-to-back launches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from helpers import surface_coords
from doda_b200 import ops
dev = torch.device("cuda")
print("NOSPLIT", os.environ.get("B200SP_TC_NOSPLIT"), "MAXSPLIT", os.environ.get("B200SP_TC_MAXSPLIT"))
for M, C in ((26506, 48), (6149, 64), (1381, 80), (223, 96), (45, 112)):
    coords, shape = surface_coords(0, max(M // 2, 20), 2)
    c = torch.from_numpy(coords).to(dev)
    rb = ops.build_rulebook(c, 2, shape, 3, 1, 1, 1, subm=True)
    n = c.shape[0]
    feat = torch.randn(n, C, device=dev)
    W3 = torch.randn(27, C, C, device=dev) * 0.2
    out = torch.empty(n, C, device=dev)
    # the weight image the training step prepares once per optimizer step (otherwise every call adds a k_prep_weights launch)
    import numpy as np
    from doda_b200._lib import lib, check
    img = torch.empty(int(lib.b200sp_conv_prepared_bytes(27, C, C)), dtype=torch.uint8, device=dev)
    desc = np.ascontiguousarray(np.asarray([(W3.data_ptr(), img.data_ptr(), 27, C, C, ops.W_FWD)], dtype=np.int64))
    scratch = torch.empty(4096, dtype=torch.uint8, device=dev)
    check(lib.b200sp_prep_weights_batch(desc.ctypes.data, 1, scratch.data_ptr(), scratch.numel(), ops._stream()), "prep")
    torch.cuda.synchronize()
    f = lambda: ops.gather_gemm(feat, W3, rb.nbr_perm, n, out=out, orow=rb.order, rowmask=rb.rowmask, wimg=img)
    for _ in range(5): f()
    torch.cuda.synchronize()
    # replayed from a CUDA graph: the Python + ctypes call (~15 us) would otherwise be the floor of the small shapes
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(50): f()
    gr.replay(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); gr.replay(); e.record(); torch.cuda.synchronize()
    a = f().clone(); b = f().clone()
    print("rows %6d C %3d: %.1f us / launch, bit-identical %s" % (n, C, s.elapsed_time(e) / 50 * 1e3, torch.equal(a, b)), flush=True)
