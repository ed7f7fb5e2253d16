"""voxelize_idx at BASELINE configs[1] size: 2 scenes, 450 k points -> 300 k voxels.
CPU entry point (DataLoader side, one thread) vs the device version (identical output); prints one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from doda_b200 import scenes, pointgroup_ops

batch = scenes.collate([scenes.scene_with_voxels(i, 150000) for i in range(2)], seed=0, dup_max=2)
locs = batch["locs"].contiguous()
for _ in range(2):
    pointgroup_ops.voxelization_idx(locs, 2, 4)
t = time.perf_counter()
for _ in range(5):
    oc, im, om = pointgroup_ops.voxelization_idx(locs, 2, 4)
cpu_ms = (time.perf_counter() - t) / 5 * 1e3
dev = torch.device("cuda")
g = locs.to(dev)
for _ in range(3):
    goc, gim, gom = pointgroup_ops.voxelization_idx_gpu(g, 2, 4)
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(10):
    goc, gim, gom = pointgroup_ops.voxelization_idx_gpu(g, 2, 4)
torch.cuda.synchronize()
gpu_ms = (time.perf_counter() - t) / 10 * 1e3
same = bool(torch.equal(goc.cpu(), oc) and torch.equal(gim.cpu(), im) and torch.equal(gom.cpu(), om))
print(json.dumps({"points": int(locs.shape[0]), "voxels": int(oc.shape[0]), "max_active": int(om.shape[1] - 1),
                  "cpu_ms_one_thread": round(cpu_ms, 2), "gpu_ms_wall_incl_one_host_sync": round(gpu_ms, 3),
                  "identical": same}))
