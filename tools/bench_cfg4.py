"""BASELINE configs[3]: one 400 k-voxel scene, the k2 s2 SparseConv3d / SparseInverseConv3d encoder-decoder only
(6 down + 6 up with BN/ReLU, no SubM blocks), fwd+bwd on 1 x B200.  Prints one JSON line (scenes/s, ms per step,
per-kernel shares from CUDA events).  Not the headline bench (bench.py is configs[1]); a parity-size + timing case."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from helpers import surface_coords
from doda_b200 import spconv, ops
from test_parity_gpu import _encoder_decoder

dev = torch.device("cuda")
torch.manual_seed(0)
coords, shape = surface_coords(7, 400000, 1)
c = torch.from_numpy(coords).to(dev)
n = c.shape[0]
planes = [16 * i for i in range(1, 8)]
net = _encoder_decoder(planes, with_bn=True).to(dev).train()
x = torch.randn(n, planes[0], device=dev)
params = list(net.parameters())
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def step():
    for p in params:
        p.grad = None
    ops.invalidate_prepared_weights()
    xi = x.clone().requires_grad_(True)
    y = net(spconv.SparseConvTensor(xi, ops.stage_coords(c, dev), shape, 1))
    y.features.square().mean().backward()


for _ in range(8):
    step()
torch.cuda.synchronize()
K = 20
evs = []
for _ in range(K):
    flush.zero_()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); step(); e.record()
    evs.append((s, e))
torch.cuda.synchronize()
ms = [s.elapsed_time(e) for s, e in evs]
ops.profile_begin(); step(); torch.cuda.synchronize(); recs = ops.profile_end()
by = {}
for r in recs:
    by[r["kernel"]] = by.get(r["kernel"], 0.0) + r["ms"]
levels = [n] + [int(r["n_out"]) for r in recs if r["kernel"] == "k_gather_gemm"][:6]
print(json.dumps({"workload": "configs[3]: 1 x 400k-voxel scene, k2s2 encoder-decoder 6 down + 6 up, BN/ReLU, fwd+bwd",
                  "ms_per_step": sum(ms) / K, "scenes_per_s": K / (sum(ms) / 1e3), "steps": K, "rows_per_level": levels,
                  "event_ms_by_kernel": {k: round(v, 3) for k, v in by.items()}, "step_ms": [round(v, 3) for v in ms]}))
