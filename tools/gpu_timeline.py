"""dev helper (GPU box): CUPTI timeline of a few training steps through torch.profiler -> per-kernel totals, GPU busy
time vs wall time per step (how much of the step the GPU sits idle waiting for the host)."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from doda_b200 import scenes
from doda_b200.unet import SparseConvNet, model_step
dev = torch.device("cuda")
batch = scenes.collate([scenes.scene_with_voxels(i, 150000) for i in range(2)], seed=0, dup_max=2)
for k in ("voxel_locs", "p2v_map", "v2p_map", "feats", "labels"):
    batch[k] = batch[k].to(dev)
model = SparseConvNet(mid_channel=16).to(dev).train()
crit = torch.nn.CrossEntropyLoss(ignore_index=255)
def step():
    for p in model.parameters(): p.grad = None
    loss, _ = model_step(model, batch, criterion=None, device=dev)
    loss.backward()
for _ in range(4): step()
torch.cuda.synchronize()
# shapes of the conv / wgrad launches of one step, in launch order (to put next to the CUPTI durations)
from doda_b200 import ops
ops.profile_begin(); step(); torch.cuda.synchronize(); recs = ops.profile_end()
shapes = collections.defaultdict(list)
for r in recs:
    shapes[r["kernel"]].append(r)
NS = 4
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(NS): step()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
tot = collections.defaultdict(float); cnt = collections.Counter()
ivs = []
for e in evs:
    name = e.name.replace("(anonymous namespace)::", "").split("(")[0][:70]
    tot[name] += e.device_time if hasattr(e, "device_time") else e.cuda_time
    cnt[name] += 1
    ivs.append((e.time_range.start, e.time_range.end))
ivs.sort()
busy = 0.0; cur_s, cur_e = ivs[0]
for s, e in ivs[1:]:
    if s > cur_e:
        busy += cur_e - cur_s; cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
busy += cur_e - cur_s
wall = ivs[-1][1] - ivs[0][0]
print("steps %d: wall %.2f ms/step, GPU busy (union of kernels) %.2f ms/step, sum of kernel durations %.2f ms/step, kernels/step %d"
      % (NS, wall / NS / 1e3, busy / NS / 1e3, sum(tot.values()) / NS / 1e3, len(evs) // NS))
try:  # busy time per CUDA stream (compute stream, weight-gradient side stream, index stream)
    kev = prof.profiler.kineto_results.events()
    bys = collections.defaultdict(list)
    for e in kev:
        if e.device_type() == torch.autograd.DeviceType.CUDA:
            bys[e.device_resource_id()].append((e.start_ns(), e.start_ns() + e.duration_ns()))
    for sid, iv in sorted(bys.items(), key=lambda kv: -sum(b - a for a, b in kv[1])):
        print("  stream %s: %5d kernels/step, busy %.2f ms/step" % (sid, len(iv) // NS, sum(b - a for a, b in iv) / NS / 1e6))
except Exception as ex:
    print("per-stream breakdown unavailable:", repr(ex)[:120])
for name, t in sorted(tot.items(), key=lambda kv: -kv[1])[:28]:
    print("%9.1f us/step %5d  %s" % (t / NS, cnt[name] // NS, name))

per = collections.defaultdict(list)
for e in sorted(evs, key=lambda e: e.time_range.start):
    n = e.name.replace("(anonymous namespace)::", "")
    key = "wgrad" if "k_wgrad" in n else "conv" if "k_conv_tc" in n or "k_conv_direct" in n or "k_gather_gemm" in n else "bn" if "k_bn_" in n or "k_affine" in n else None
    if key:
        per[key].append((n.split("(")[0][-28:], e.time_range.end - e.time_range.start))
L = per["bn"]; n1 = len(L) // NS
rk = [r for r in recs if r["kernel"] in ("bn_fwd", "bn_bwd")]
print("== bn: %d launches/step, %d bn calls recorded" % (n1, len(rk)))
i = 0
for r in rk:
    nk = 2
    names = [L[n1 + i + q] for q in range(nk)]
    print("  %-7s M %6d C %3d : %s" % (r["kernel"], r["M"], r["C"], "  ".join("%s %.1f us" % (n[-22:], d) for n, d in names)))
    i += nk
for key in ("conv", "wgrad"):
    L = per[key]; n1 = len(L) // NS
    print("== %s: %d launches/step; per-launch us (step 2 of the capture), with the recorded shapes" % (key, n1))
    rk = [r for r in recs if (r["kernel"] in ("k_gather_gemm",) and key == "conv") or (r["kernel"] == "k_wgrad" and key == "wgrad")]
    for i in range(n1):
        name, d = L[n1 + i]
        r = rk[i] if i < len(rk) else {}
        print("  %3d %-28s %8.1f us  %s" % (i, name, d, {k: v for k, v in r.items() if k not in ("kernel", "ms")}))
