"""dev helper (GPU box): the smoke() comparison (engine vs fp64 CPU oracle) at several scene sizes, twice each --
how the whole-net gradient error depends on BatchNorm conditioning, and whether two runs print the same digits."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from doda_b200 import scenes
from doda_b200.unet import SparseConvNet, model_step
from oracle.unet_ref import model_step_ref
dev = torch.device("cuda:0")
for target in (2000, 8000, 20000):
    torch.manual_seed(0)
    batch = scenes.collate([scenes.scene_with_voxels(0, target), scenes.scene_with_voxels(1, target)], dup_max=2)
    model = SparseConvNet(mid_channel=16)
    sd = {k: (v.detach().double().clone().requires_grad_(True) if v.is_floating_point() else v.clone())
          for k, v in model.state_dict().items()}
    b64 = dict(batch); b64["feats"] = batch["feats"].double()
    loss_ref, scores_ref = model_step_ref(sd, b64, training=True)
    loss_ref.backward()
    model = model.to(dev).train()
    for rep in range(2):
        for p in model.parameters(): p.grad = None
        loss, scores = model_step(model, batch, device=dev)
        loss.backward()
        torch.cuda.synchronize()
        err = float((scores.double().cpu() - scores_ref).abs().max() / scores_ref.abs().max())
        def rel_l2(a, b): return float((a.double().cpu() - b).norm() / max(float(b.norm()), 1e-12))
        lin = rel_l2(model.linear.weight.grad, sd["linear.weight"].grad)
        first = rel_l2(model.input_conv[0].weight.grad, sd["input_conv.0.weight"].grad)
        print("target %d run %d: loss %.7f (oracle %.7f) scores %.3e linear %.3e input_conv %.3e" %
              (target, rep, float(loss), float(loss_ref), err, lin, first), flush=True)
