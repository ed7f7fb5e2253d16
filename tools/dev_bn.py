"""BatchNorm(+ReLU) forward / backward timings per U-Net level, cluster (one launch) vs two-kernel form.
usage (GPU box): python tools/dev_bn.py"""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import torch
    from doda_b200 import ops
    dev = torch.device("cuda")
    for M, C in ((166000, 16), (93000, 32), (26506, 48), (6149, 64), (6149, 128), (1381, 80), (1381, 160), (223, 96), (223, 192), (45, 112), (45, 224)):
        bn = torch.nn.BatchNorm1d(C, eps=1e-4).to(dev)
        x = torch.randn(M, C, device=dev, requires_grad=True)
        g = torch.randn(M, C, device=dev)
        def fwd():
            return ops.batch_norm_relu(x, bn, relu=True)
        y = fwd(); y.backward(g)
        def t(fn, n=40):
            # n launches replayed from a CUDA graph: the ctypes call (~8 us) would otherwise hide the GPU time
            for _ in range(3): fn()
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                for _ in range(n): fn()
            gr.replay(); torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); gr.replay(); e.record(); torch.cuda.synchronize()
            return s.elapsed_time(e) / n * 1e3
        from doda_b200._lib import lib
        mean = torch.empty(C, device=dev); inv = torch.empty(C, device=dev); yb = torch.empty_like(x); dx = torch.empty_like(x)
        dw = torch.empty(2, C, device=dev)
        ws = torch.zeros(int(lib.b200sp_bn_ws_bytes(M, C)), dtype=torch.uint8, device=dev)

        xd = x.detach()
        st = None
        f = lambda: lib.b200sp_bn_fwd_train(xd.data_ptr(), M, C, bn.weight.data_ptr(), bn.bias.data_ptr(), 1e-4, 1, yb.data_ptr(),
                                            mean.data_ptr(), inv.data_ptr(), 0, 0, 0.1, 0, ws.data_ptr(), ws.numel(), ops._stream())
        b = lambda: lib.b200sp_bn_bwd(xd.data_ptr(), g.data_ptr(), M, C, bn.weight.data_ptr(), bn.bias.data_ptr(), mean.data_ptr(),
                                      inv.data_ptr(), 1, dx.data_ptr(), dw.data_ptr(), dw.data_ptr() + 4 * C, ws.data_ptr(), ws.numel(), ops._stream())
        assert f() == 0 and b() == 0
        print("M %6d C %3d  fwd %6.1f us  bwd %6.1f us" % (M, C, t(f), t(b)), flush=True)
else:
    for on, kb in (("0", "64"), ("8", "64"), ("8", "200"), ("16", "64"), ("16", "128")):
        print("B200SP_BN_CLUSTER=%s _KB=%s" % (on, kb), flush=True)
        subprocess.run([sys.executable, __file__, "run"], env=dict(os.environ, B200SP_BN_CLUSTER=on, B200SP_BN_CLUSTER_KB=kb))
