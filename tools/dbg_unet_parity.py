"""dev helper: per-parameter gradient error of the whole net vs the fp64 oracle (run on the GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from helpers import rel_err
from doda_b200 import scenes
from doda_b200.unet import SparseConvNet, model_step
from oracle.unet_ref import model_step_ref

target = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
mid = int(sys.argv[2]) if len(sys.argv) > 2 else 16
torch.manual_seed(0)
batch = scenes.collate([scenes.scene_with_voxels(0, target), scenes.scene_with_voxels(1, target)], dup_max=2)
model = SparseConvNet(mid_channel=mid)
sd64 = {k: (v.detach().double().clone().requires_grad_(True) if v.is_floating_point() else v.clone())
        for k, v in model.state_dict().items()}
sd32 = {k: (v.detach().clone().requires_grad_(True) if v.is_floating_point() else v.clone())
        for k, v in model.state_dict().items()}
b64 = dict(batch); b64["feats"] = batch["feats"].double()
loss_ref, scores_ref = model_step_ref(sd64, b64, training=True); loss_ref.backward()
loss32, scores32 = model_step_ref(sd32, batch, training=True); loss32.backward()
model = model.cuda().train()
loss, scores = model_step(model, batch, device="cuda"); loss.backward()
print("scores: gpu %.2e  cpu-fp32-oracle %.2e" % (rel_err(scores, scores_ref), rel_err(scores32, scores_ref)))
rows = []
for name, p in model.named_parameters():
    rows.append((rel_err(p.grad, sd64[name].grad), rel_err(sd32[name].grad, sd64[name].grad), name,
                 float(sd64[name].grad.abs().max())))
rows.sort(reverse=True)
for r in rows[:25]:
    print("gpu %.2e  cpu32 %.2e  %-50s |g|max %.3e" % r)
