"""dev helper (GPU box): host time per section of one training step, measured with perf_counter wrappers on a TINY scene
(2 x 2000 voxels: the GPU is never the bound, so wall time = host time).  usage: python tools/host_sections.py [voxels]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from doda_b200 import scenes, ops, tape
from doda_b200.unet import SparseConvNet, model_step
NV = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
dev = torch.device("cuda")
batch = scenes.collate([scenes.scene_with_voxels(i, NV) for i in range(2)], seed=0, dup_max=2)
for k in ("voxel_locs", "p2v_map", "v2p_map", "feats", "labels"):
    batch[k] = batch[k].to(dev)
model = SparseConvNet(mid_channel=16).to(dev).train()
params = list(model.parameters())
acc = {}


def wrap(obj, name, key):
    f = getattr(obj, name)

    def g(*a, **k):
        t = time.perf_counter()
        try:
            return f(*a, **k)
        finally:
            acc[key] = acc.get(key, 0.0) + time.perf_counter() - t
    setattr(obj, name, g)


wrap(ops, "build_rulebook", "rulebooks (13 per step)")
wrap(ops, "_prepare_all", "weight images (1 batched launch)")
wrap(ops, "_layer_fwd", "layer_fwd wrappers + C calls (71)")
wrap(ops, "_layer_bwd", "layer_bwd wrappers + C calls (71)")
wrap(tape, "_ublock", "taped forward, all levels (incl. rulebooks, layer_fwd)")
wrap(tape.UBlockTapeFunction, "backward", "taped backward (incl. layer_bwd)")


def step():
    for p in params:
        p.grad = None
    ops.invalidate_prepared_weights()
    t0 = time.perf_counter()
    loss, _ = model_step(model, batch, device=dev)
    t1 = time.perf_counter()
    loss.backward()
    t2 = time.perf_counter()
    acc["model_step forward (everything)"] = acc.get("model_step forward (everything)", 0.0) + t1 - t0
    acc["loss.backward() (everything)"] = acc.get("loss.backward() (everything)", 0.0) + t2 - t1


for _ in range(5):
    step()
torch.cuda.synchronize()
acc.clear()
N = 20
t0 = time.perf_counter()
for _ in range(N):
    step()
torch.cuda.synchronize()
tot = (time.perf_counter() - t0) / N * 1e3
print("voxels per scene %d: %.2f ms per step wall (host-bound when the scene is tiny)" % (NV, tot))
for k, v in sorted(acc.items(), key=lambda kv: -kv[1]):
    print("  %-58s %6.2f ms/step" % (k, v / N * 1e3))
