"""dev helper (GPU box): per-step wall time next to the caching allocator's cudaMalloc/cudaFree counts -- are the
outlier steps allocator growth?"""
import os, sys, time, gc
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
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def step():
    for p in params: p.grad = None
    ops.invalidate_prepared_weights()
    loss, _ = model_step(model, batch, criterion=None, device=dev)
    loss.backward()
gc.collect(); gc.disable()
prev = None
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 16):
    torch.cuda.synchronize()
    flush.zero_()
    t0 = time.perf_counter()
    step()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    st = torch.cuda.memory_stats()
    cur = (st["segment.all.allocated"], st["segment.all.freed"], st["num_alloc_retries"])
    print("step %2d enqueue %6.2f ms total %6.2f ms  cudaMalloc %d cudaFree %d retries %d reserved %.2f GB allocated-peak %.2f GB"
          % (i, (t1 - t0) * 1e3, (t2 - t0) * 1e3, cur[0] - (prev[0] if prev else 0), cur[1] - (prev[1] if prev else 0),
             cur[2], st["reserved_bytes.all.current"] / 2**30, st["allocated_bytes.all.peak"] / 2**30))
    prev = cur
