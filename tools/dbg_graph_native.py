"""dev helper: which part of the GPU-native baseline step breaks CUDA-graph capture"""
import os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from doda_b200 import scenes
from doda_b200.unet import SparseConvNet
from baseline import gpu_native as gn
dev = torch.device("cuda")
torch.manual_seed(0)
NV = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
batch = scenes.collate([scenes.scene_with_voxels(0, NV), scenes.scene_with_voxels(1, NV)], dup_max=2)
model = SparseConvNet(mid_channel=16)
net = gn.NativeUNet(model.state_dict(), dev)
coords = batch["voxel_locs"].to(dev); feats = batch["feats"].to(dev); v2p = batch["v2p_map"].to(dev)
p2v = batch["p2v_map"].to(dev).long(); labels = batch["labels"].to(dev)
vf = gn.voxelize_mean(feats, v2p)
net.build_rulebooks(coords, [int(s) for s in batch["spatial_shape"]])
for _ in range(2): net.step(vf, p2v, labels)
torch.cuda.synchronize()


def attempt(name, fn, mode):
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side): fn()
    torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
    for p in net.params(): p.grad = None
    try:
        with torch.cuda.graph(g, capture_error_mode=mode): fn()
        g.replay(); torch.cuda.synchronize()
        print(name, mode, "OK", flush=True)
    except Exception as ex:
        print(name, mode, "FAILED", repr(ex)[:150], flush=True)
        torch.cuda.synchronize()


x0 = torch.randn(5000, 16, device=dev, requires_grad=True)
W0 = torch.randn(27, 16, 16, device=dev, requires_grad=True)
attempt("mm", lambda: torch.mm(x0, W0[0]), "global")
attempt("conv fwd", lambda: gn._NativeConv.apply(vf, net._w("input_conv.0.weight"), net.rb[0]["subm"], vf.shape[0], 13), "global")
attempt("conv fwd+bwd", lambda: gn._NativeConv.apply(vf, net._w("input_conv.0.weight"), net.rb[0]["subm"], vf.shape[0], 13).sum().backward(), "global")
attempt("bn", lambda: net._bn_relu("output_layer.0", torch.randn(5000, 16, device=dev)), "global")
attempt("index", lambda: x0[p2v[:1000] % 5000].sum().backward(), "global")
attempt("ce", lambda: torch.nn.functional.cross_entropy(torch.randn(labels.shape[0], 11, device=dev, requires_grad=True), labels, ignore_index=255).backward(), "global")
attempt("forward", lambda: net.forward(vf, p2v, labels), "global")
attempt("forward", lambda: net.forward(vf, p2v, labels), "relaxed")
attempt("step", lambda: net.forward(vf, p2v, labels)[0].backward(), "relaxed")
attempt("step", lambda: net.forward(vf, p2v, labels)[0].backward(), "global")
# now with the engine's own steps before the capture, as in bench.py
from doda_b200.unet import model_step
from doda_b200 import ops
m2 = SparseConvNet(mid_channel=16).to(dev).train()
res = ops.stage_batch({k: (v.pin_memory() if hasattr(v, "pin_memory") else v) for k, v in batch.items()}, dev) if hasattr(ops, "stage_batch") else batch
for _ in range(3):
    for p in m2.parameters(): p.grad = None
    loss, _ = model_step(m2, batch, device=dev)
    loss.backward()
torch.cuda.synchronize()
attempt("step after engine steps", lambda: net.forward(vf, p2v, labels)[0].backward(), "relaxed")
torch.cuda.synchronize(); torch.cuda.empty_cache()
attempt("step after engine steps + empty_cache", lambda: net.forward(vf, p2v, labels)[0].backward(), "relaxed")
