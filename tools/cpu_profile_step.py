"""dev helper (GPU box): where does the HOST time of one training step go?  cProfile over a few steps."""
import cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from doda_b200 import scenes, ops
from doda_b200.unet import SparseConvNet, model_step
dev = torch.device("cuda")
batch = scenes.collate([scenes.scene_with_voxels(i, 150000) for i in range(2)], seed=0, dup_max=2)
for k in ("voxel_locs", "p2v_map", "v2p_map", "feats", "labels"):
    batch[k] = batch[k].to(dev)
model = SparseConvNet(mid_channel=16).to(dev).train()
crit = torch.nn.CrossEntropyLoss(ignore_index=255)
params = list(model.parameters())
def step():
    for p in params: p.grad = None
    ops.invalidate_prepared_weights()
    loss, _ = model_step(model, batch, criterion=None, device=dev)
    loss.backward()
for _ in range(3): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5): step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("enqueue %.2f ms/step, with final sync %.2f ms/step, launches/step %d" % ((t1 - t0) / 5 * 1e3, (t2 - t0) / 5 * 1e3, 0))
pr = cProfile.Profile()
pr.enable()
for _ in range(3): step()
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(45)
print(s.getvalue()[:14000])
