#!/usr/bin/env python
"""bench.py -- scenes/s forward+backward of DODA's sparse U-Net on synthetic ScanNet-shaped scenes
(BASELINE.json configs[1]: 2 x 150k-voxel scenes, full unet.py fwd+bwd, bs=2 per GPU).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
  torchrun --nproc-per-node N bench.py --gpus N ...      (one rank per GPU, NCCL; whole scenes are sharded, the
                                                          gradient mean after backward is the only collective:
                                                          parallel.allreduce_grads, or DDP with --ddp; "scaling": "weak")

Prints ONE JSON line (rank 0).  `value` = scenes/s with the collated batch already resident in HBM;
`e2e` = the same step through the public API from pinned HOST buffers (H2D of every step's batch, staged one step
ahead on the index stream, + an async D2H of every step's loss read one step late, all inside the timed region);
`roofline` = the kernel with the largest total time of the step over ALL families (conv fwd/dgrad, weight gradient,
BatchNorm), its dominant launch shape timed live with CUDA events against the measured HBM peak, plus one row per
family in `roofline.families`; `baseline_gpu_native` / `vs_gpu_native` = spconv v1.2's native algorithm with stock
torch ops on the same GPU (baseline/gpu_native.py: the stand-in for reference spconv-CUDA); `m32` = the same step
at mid_channel 32; `cpu_baseline` (N = 1) = the CPU oracle (restatement of spconv v1.2's native algorithm) on the host
cores, with `parity_full_size` = the engine's loss / per-point scores against it at the full size.
`--impl reference` times that CPU implementation as the reference arm (no product code on that path).
`--model reference` runs the reference's UNCHANGED model/unet.py + unet_block.py + model_fn (staged copy under
oracle/_ref/src, through compat/) instead of the mirror doda_b200/unet.py.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "scenes/sec fwd+bwd Sparse U-Net @150k voxels"
UNIT = "scenes/s"


WORKLOAD = "2x150k-voxel ScanNet-shape scenes, full SparseConvNet fwd+bwd, bs=%d, m=%d"
SETTLE = 10  # extra untimed steps per timed loop (see main); reported in the JSON line's config


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--voxels", type=int, default=150000)
    ap.add_argument("--bs", type=int, default=2, help="scenes per GPU")
    ap.add_argument("--mid", type=int, default=16, help="MODEL.BACKBONE.mid_channel (16 as shipped)")
    ap.add_argument("--ddp", action="store_true", help="N>1: torch DistributedDataParallel instead of parallel.allreduce_grads")
    ap.add_argument("--model", default="mirror", choices=["mirror", "reference"],
                    help="mirror: doda_b200/unet.py; reference: the reference's own model/unet.py, unchanged, via compat/")
    ap.add_argument("--attach-tape", action="store_true", help="--model reference: doda_b200.tape.attach(model) (taped U-Net sub-trees)")
    ap.add_argument("--as-rank", type=int, default=-1, help="diagnosis at N=1: run the scenes rank R holds in a multi-GPU run")
    ap.add_argument("--no-top-tape", action="store_true", help="diagnosis: level 1 module by module, levels 2-7 taped (what N>1 runs)")
    ap.add_argument("--no-overlap", action="store_true", help="N>1: one gradient mean AFTER backward instead of the overlapped reducer")
    ap.add_argument("--no-allreduce", action="store_true", help="N>1 diagnosis: skip the collective (load imbalance only)")
    ap.add_argument("--no-gpu-native", action="store_true")
    ap.add_argument("--no-m32", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--detail", default="", help="write per-kernel detail JSON here")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler(object):
    """SM clock / throttle-reason samples during the timed region (B200_PROFILING.md recipe), taken in-process
    through NVML (nvidia_ml_py; a one-shot `nvidia-smi` query is the fallback).

    An NVML query holds a driver lock: CUDA launches of this process stall for 10-110 ms while it runs (first seen
    with an `nvidia-smi -lms` subprocess, then with a sampling thread: one or two outlier steps per timed loop).
    So there is no polling: the timed loops call sample() once, at their middle step, BETWEEN two steps (where the L2
    flush already sits, outside the per-step event intervals) while the GPU is still executing the steps queued
    before it -- the clocks are read under load, and the query's stall is not booked on a step."""

    def __init__(self, index):
        self.index, self.rows, self.nv, self.h, self.mx = index, [], None, None, None
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.h = nv.nvmlDeviceGetHandleByIndex(index)
            self.mx = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
            self.bits = [("hw_slowdown", getattr(nv, "nvmlClocksEventReasonHwSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8))),
                         ("hw_thermal_slowdown", getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40))),
                         ("sw_thermal_slowdown", getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20))),
                         ("sw_power_cap", getattr(nv, "nvmlClocksEventReasonSwPowerCap", getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)))]
            self.get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            self.nv = nv
        except Exception:
            self.nv = None

    def sample(self):
        try:
            if self.nv is not None:
                sm = float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = int(self.get_reasons(self.h))
                self.rows.append([str(sm), str(self.mx)] + ["Active" if r & b else "Not Active" for _, b in self.bits])
            else:
                q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                     "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=20).stdout
                for line in out.splitlines():
                    if line.strip():
                        self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def make_batch(rank, bs, voxels):
    from doda_b200 import scenes
    return scenes.collate([scenes.scene_with_voxels(1000 * rank + i, voxels) for i in range(bs)], seed=rank, dup_max=2)


def conv_layer_bytes(rec):
    """Algorithmic (compulsory) bytes of one conv launch, SURVEY.md §8(d):
    4*(M_in*Cin + M_out*Cout) + 4*K*Cin*Cout + 8*P   (P = rulebook pairs of the layer)."""
    return 4 * (rec["n_in"] * rec["Cin"] + rec["n_out"] * rec["Cout"]) + 4 * rec["K"] * rec["Cin"] * rec["Cout"] \
        + 8 * rec["pairs"]


def wgrad_bytes(rec):
    """compulsory bytes of one weight-gradient launch: both operands once, dW once, the pair list:
    4*(M_a*Ca + M_g*Cb) + 4*K*Ca*Cb + 8*P"""
    return 4 * (rec["n_rows"] * rec["Ca"] + rec.get("n_b", rec["n_rows"]) * rec["Cb"]) + 4 * rec["K"] * rec["Ca"] * rec["Cb"] \
        + 8 * rec["pairs"]


def record_bytes(rec):
    k = rec["kernel"]
    if k == "k_gather_gemm":
        return conv_layer_bytes(rec)
    if k == "k_wgrad":
        return wgrad_bytes(rec)
    if k == "bn_fwd":
        return 12 * rec["M"] * rec["C"]   # read for the statistics, read + write for the apply
    if k == "bn_bwd":
        return 20 * rec["M"] * rec["C"]   # reduce reads x, dy; apply reads x, dy, writes dx
    return 0


def record_name(rec):
    if rec["kernel"] in ("k_gather_gemm", "k_wgrad"):
        return rec.get("name") or rec["kernel"]
    return {"bn_fwd": "k_bn_reduce<.,0> + k_affine_relu", "bn_bwd": "k_bn_reduce<.,1> + k_bn_bwd_apply"}.get(rec["kernel"], rec["kernel"])


def record_shape(rec):
    k = rec["kernel"]
    if k == "k_gather_gemm":
        return ("rows", rec["n_out"], "Cin", rec["Cin"], "Cout", rec["Cout"], "K", rec["K"], "pairs_mode", rec.get("pairs_mode", 0))
    if k == "k_wgrad":
        return ("rows", rec["n_rows"], "Ca", rec["Ca"], "Cb", rec["Cb"], "K", rec["K"])
    return ("rows", rec["M"], "C", rec["C"])


def roofline_from_records(recs, peak, how):
    """One row per kernel family of the step (CUDA-event time per launch on the launch stream, algorithmic bytes of
    SURVEY.md 8(d) / DESIGN.md 3); `roofline` itself describes the family with the LARGEST total time -- the step's
    dominant kernel, whatever it is -- through its dominant launch shape."""
    if not recs:
        return None
    fams = {}
    for r in recs:
        fams.setdefault(record_name(r), []).append(r)
    all_ms = sum(r["ms"] for r in recs)
    rows = []
    for name, rs in fams.items():
        ms = sum(r["ms"] for r in rs)
        by = sum(record_bytes(r) for r in rs)
        rows.append({"kernel": name, "launches_per_step": len(rs), "ms_per_step": ms, "algorithmic_bytes": by,
                     "algorithmic_GBs": by / (ms * 1e-3) / 1e9 if ms > 0 else None,
                     "frac": by / (ms * 1e-3) / 1e9 / peak if ms > 0 else None,
                     "share_of_engine_kernel_time": ms / max(all_ms, 1e-9)})
    rows.sort(key=lambda r: -r["ms_per_step"])
    top = rows[0]["kernel"]
    groups = {}
    for r in fams[top]:
        groups.setdefault(record_shape(r), []).append(r)
    key, grp = max(groups.items(), key=lambda kv: sum(r["ms"] for r in kv[1]))
    g_ms = sum(r["ms"] for r in grp) / len(grp)
    g_bytes = record_bytes(grp[0])
    ach = g_bytes / (g_ms * 1e-3) / 1e9
    shape = dict(zip(key[0::2], key[1::2]))
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")  # dram bytes per launch from `ncu --set full` captures
    if os.path.exists(tpath):
        for t in json.load(open(tpath)).get(top.split(" ")[0], []):
            same = all(t.get(k) == v for k, v in shape.items() if k not in ("rows", "pairs_mode"))
            rows_t = t.get("rows", t.get("n_out", 0))
            if same and abs(rows_t - shape["rows"]) <= 0.1 * shape["rows"]:
                traffic = t["dram_bytes"]
    gg = [r for r in recs if r["kernel"] == "k_gather_gemm"]
    conv_ms = sum(r["ms"] for r in gg)
    return {"bound": "hbm", "kernel": top, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "traffic": traffic, "peak_source": how,
            "launch_shape": dict(shape, launches_per_step=len(grp), avg_launch_ms=g_ms, algorithmic_bytes=g_bytes),
            "families": rows,
            "all_conv_launches": {"launches_per_step": len(gg), "total_ms": conv_ms,
                                  "algorithmic_GBs": sum(conv_layer_bytes(r) for r in gg) / (conv_ms * 1e-3) / 1e9 if gg else None,
                                  "useful_dense_tflops": sum(2.0 * r["pairs"] * r["Cin"] * r["Cout"] for r in gg) / (conv_ms * 1e-3) / 1e12 if gg else None,
                                  "share_of_engine_kernel_time": conv_ms / max(all_ms, 1e-9)}}


def pairs_per_level(model, resident, dev, is_ref):
    """rows and rulebook pairs per U-Net level of THIS batch (P/M says how heavy the synthetic scenes are next to
    SURVEY.md Appendix B's 7.9-13.6)"""
    import torch
    from doda_b200 import spconv, pointgroup_ops
    with torch.no_grad():
        vf = pointgroup_ops.voxelization(resident["feats"], resident["v2p_map"], 4)
        x = spconv.SparseConvTensor(vf, resident["voxel_locs"].int(), resident["spatial_shape"],
                                    resident["offsets"].size(0) - 1)
        model(x, resident["p2v_map"])
        out = []
        for l in range(1, 8):
            rb = x.indice_dict.get("subm%d" % l)
            if rb is None or rb.pairnum is None:
                continue
            M = int(rb.indices.shape[0])
            P = int(rb.pairnum.sum())
            out.append({"level": l, "rows": M, "pairs": P, "P_over_M": round(P / max(M, 1), 2)})
    return out


def run_gpu_native(model, batch, dev, value_ms, bs):
    """spconv v1.2's native algorithm with stock torch ops on this GPU (baseline/gpu_native.py), same weights, same
    batch.  vs_gpu_native = its best time (CUDA-graph replay + eager rulebooks when the capture works, else eager) /
    the engine's ms_per_step."""
    import torch
    from baseline import gpu_native
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    try:
        r = gpu_native.measure(sd, batch, dev, steps=3, warmup=1, graph=True)
    except Exception as e:
        import traceback
        return {"error": repr(e)[:300], "traceback": traceback.format_exc()[-800:], "vs_gpu_native": None}
    finally:
        torch.cuda.empty_cache()
    best = min(v for v in (r["ms_eager"], r["ms_graph"]) if v is not None)
    return {"algorithm": "per offset index_select -> torch.mm (fp32, TF32 off) -> index_add_, torch BN/ReLU/CE; rulebooks by "
                         "sort + searchsorted (torch ops)", "ms_per_step_eager": r["ms_eager"],
            "ms_per_step_cuda_graph": r["ms_graph"], "ms_rulebooks": r["ms_rulebooks"],
            "graph_error": r.get("graph_error"),
            "scenes_per_s": bs / (best * 1e-3), "vs_gpu_native": best / value_ms,
            "note": "stand-in for reference spconv-CUDA (spconv v1.2 cannot be built offline against torch 2.11); "
                    "vs_gpu_native uses the baseline's FASTER reading"}


def run_other_width(mid, batch, dev, model_step, SparseConvNet, ops, flush, steps=8):
    """the same fwd+bwd step at another mid_channel ("also report m=32")"""
    import torch
    model = SparseConvNet(mid_channel=mid).to(dev).train()
    params = list(model.parameters())
    res = {k: (v.to(dev) if hasattr(v, "to") and k in ("voxel_locs", "p2v_map", "v2p_map", "feats", "labels") else v)
           for k, v in batch.items()}

    def step():
        for p in params:
            p.grad = None
        ops.invalidate_prepared_weights()
        loss, _ = model_step(model, res, device=dev)
        loss.backward()

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    evs = []
    for _ in range(steps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        step()
        e.record()
        evs.append((s, e))
    torch.cuda.synchronize()
    ms = sum(s.elapsed_time(e) for s, e in evs) / steps
    bs = int(batch["offsets"].shape[0] - 1)
    del model, params
    torch.cuda.empty_cache()
    return {"mid_channel": mid, "ms_per_step": ms, "value": bs / (ms * 1e-3), "unit": UNIT, "steps": steps}


def _oracle_batch(bs, voxels, rank=0):
    """the same collated batch as make_batch(), built WITHOUT the product library: scene generator (pure numpy) +
    the oracle's numpy voxelizer"""
    from doda_b200 import scenes  # numpy only: loads no native code
    from oracle.voxelize import voxelize_idx_ref
    return scenes.collate([scenes.scene_with_voxels(1000 * rank + i, voxels) for i in range(bs)], seed=rank, dup_max=2,
                          voxelize=voxelize_idx_ref)


def run_reference(args):
    """Reference arm: the CPU restatement of the reference's (spconv v1.2 native) algorithm on the host cores, the
    full workload every step.  Nothing of the product runs here: scenes are numpy, the collate uses the oracle's
    voxelizer, the weights are initialised from the reference model's state_dict layout (tests/golden)."""
    import torch
    from oracle.unet_ref import model_step_ref
    from oracle.voxelize import init_state_dict
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    batch = _oracle_batch(args.bs, args.voxels)
    shapes = json.load(open(os.path.join(ROOT, "tests", "golden", "unet_state_dict.json")))["m%d" % args.mid]
    sd = {k: (v.requires_grad_(True) if v.is_floating_point() and "running" not in k else v)
          for k, v in init_state_dict(shapes, seed=0).items()}

    def step():
        for v in sd.values():
            v.grad = None
        loss, _ = model_step_ref(sd, batch, training=True)
        loss.backward()
        return float(loss.detach())

    steps, warm = max(1, args.steps), max(1, args.warmup)  # same K / W as the engine's arm (~4 s of CPU work per step)
    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    val = args.bs / dt
    sample = "full step (bs=%d x %dk voxels, m=%d), %d timed steps" % (args.bs, args.voxels // 1000, args.mid, steps)
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
           "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": {"workload": WORKLOAD % (args.bs, args.mid), "voxels_per_scene": args.voxels,
                      "scenes_per_gpu": args.bs, "mid_channel": args.mid,
                      "device": "host CPU, rank 0 only (oracle port of spconv v1.2's native algorithm; no GPU work)"},
           "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                            "sample": sample},
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    import torch.distributed as dist
    from doda_b200 import ops
    from doda_b200._lib import lib
    from doda_b200.unet import SparseConvNet, model_step

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the sm_100a path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib.load()  # fail loudly if the extension is missing

    torch.manual_seed(0)
    batch = make_batch(args.as_rank if (args.as_rank >= 0 and world == 1) else rank, args.bs, args.voxels)
    ref_model_fn = None
    if args.model == "reference":
        # the reference's own files, unchanged (oracle/_ref/src staged by oracle/stage_ref.py), on compat/
        from oracle import stage_ref
        if stage_ref.activate() is None:
            raise SystemExit("--model reference: reference model files are not staged (python -m oracle.stage_ref)")
        from model.unet import SparseConvNet as RefNet, model_fn_decorator
        cfg = stage_ref.make_cfg(mid_channel=args.mid)
        model = RefNet(cfg).to(dev).train()
        ref_model_fn = model_fn_decorator(cfg, args.bs)
        if args.attach_tape:  # optional one-liner of INTEGRATION.md: the reference's UBlocks as one autograd node each
            from doda_b200 import tape as _tape
            _tape.attach(model)
    else:
        model = SparseConvNet(mid_channel=args.mid).to(dev).train()
    net = model
    if world > 1 and args.ddp:  # A/B: torch's DistributedDataParallel instead of the engine's in-place gradient mean
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], gradient_as_bucket_view=True)
    criterion = None  # model_step's default: the engine's one-pass cross-entropy (ignore_index=255, mean)
    tensor_keys = ["voxel_locs", "p2v_map", "v2p_map", "feats", "labels"]
    host = dict(batch)
    for k in tensor_keys:
        host[k] = batch[k].pin_memory()
    resident = dict(batch)
    for k in tensor_keys:
        resident[k] = batch[k].to(dev)
    h2d_bytes = sum(host[k].numel() * host[k].element_size() for k in tensor_keys)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    # grow the caching allocator's pool once (the step's live set is ~2 GB, more with the host a step ahead): a
    # cudaMalloc inside a timed step is a 20-100 ms stall; the JSON line reports how many happened anyway
    # (the caching allocator keeps one pool per stream: rulebook tensors live in the index stream's pool and, being
    # handed to the main stream, are only reusable once the GPU has finished the step that used them)
    grow = [torch.empty(1 << 30, dtype=torch.uint8, device=dev) for _ in range(6)]
    grow += [torch.empty(1 << 20, dtype=torch.uint8, device=dev) for _ in range(256)]  # small pool (<= 1 MB requests)
    with torch.cuda.stream(ops.index_stream_for(dev)):
        grow += [torch.empty(1 << 30, dtype=torch.uint8, device=dev) for _ in range(4)]
        grow += [torch.empty(1 << 20, dtype=torch.uint8, device=dev) for _ in range(128)]
    del grow

    params = [p for p in model.parameters()]
    if args.no_top_tape and hasattr(model, "unet"):
        model.unet.tape = False
    from doda_b200 import ops as _engine_ops
    from doda_b200 import parallel
    reducer = None
    if world > 1 and not args.ddp:
        parallel.broadcast_parameters(model)
        if not args.no_overlap:
            # the gradient mean starts as soon as backward leaves the sub-network below level 1 (97 % of the parameters)
            reducer = parallel.OverlappedGradReducer(params, world)
            reducer.attach(model.unet.u)
            # the hook fires when `unet.u`'s module forward is entered: level 1 runs module by module, levels 2-7 are
            # the taped sub-tree (one autograd node: all its gradients exist when the hook's tensor gets its gradient)
            model.unet.tape = False

    def step(b):
        for p in params:  # what optimizer.zero_grad(set_to_none=True) does
            p.grad = None
        # the metric is fwd+bwd (no optimizer step), so the weights never change here; in training they change every
        # step and the engine re-prepares its weight images once per step -- charge that launch to every timed step
        _engine_ops.invalidate_prepared_weights()
        if ref_model_fn is not None:  # ref: model/unet.py:154-198 model_fn (its own H2D copies, voxelization, CE loss)
            loss = ref_model_fn(b, net, 0)["loss"]
        else:
            loss, _ = model_step(net, b, criterion=criterion, device=dev)
        loss.backward()
        if reducer is not None:
            reducer.finish()  # the path's only collective: gradient mean over ranks (NCCL), started inside backward
        elif world > 1 and not args.ddp and not args.no_allreduce:
            parallel.allreduce_grads(params, world)
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    loss_host = torch.zeros(max(args.steps, 8) + 8, dtype=torch.float32).pin_memory()
    losses = []

    def timed(b, nsteps, read_loss):
        """read_loss: every step copies its loss to pinned host memory (async, 4 bytes) and the host reads the value
        of the PREVIOUS step, the last one after the loop -- all nsteps losses are read inside the timed region, but
        the host is not parked on the GPU once per step (a training loop that logs its loss one step late)."""
        import gc
        evs, copied = [], []
        st0 = torch.cuda.memory_stats(dev)
        mallocs0 = st0.get("num_device_alloc", 0)
        gc.collect()
        gc.disable()  # a gen-2 collection inside the loop shows up as a 50-100 ms outlier step
        barrier()
        nxt = b
        for i in range(nsteps):
            flush.zero_()  # L2 flush between timed iterations (not timed)
            if sampler is not None and i == nsteps // 2:
                sampler.sample()  # between two steps, GPU busy with the steps queued so far (see ClockSampler)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            if read_loss and i == 0 and ref_model_fn is None:
                nxt = ops.stage_batch(b, dev, tensor_keys)  # step 0 copies its own inputs inside its interval
            cur = nxt
            if read_loss and i + 1 < nsteps and ref_model_fn is None:
                nxt = ops.stage_batch(b, dev, tensor_keys)  # step i+1's H2D copies overlap step i (all K in the region)
            loss = step(cur)
            if read_loss:
                loss_host[i:i + 1].copy_(loss.detach().reshape(1), non_blocking=True)  # device -> host, every step
                copied.append(torch.cuda.current_stream().record_event())
                if i > 0:
                    copied[i - 1].synchronize()
                    losses.append(float(loss_host[i - 1]))
                if i == nsteps - 1:
                    copied[i].synchronize()
                    losses.append(float(loss_host[i]))
            e.record()
            evs.append((s, e))
        barrier()
        gc.enable()
        per_step = [s.elapsed_time(e) for s, e in evs]
        timed.last_steps = per_step
        st1 = torch.cuda.memory_stats(dev)
        timed.last_mallocs = st1.get("num_device_alloc", 0) - mallocs0
        if rank == 0 and timed.last_mallocs:
            sys.stderr.write("[bench] cudaMalloc in timed loop: %s\n" % {k: st1[k] - st0.get(k, 0) for k in (
                "segment.small_pool.allocated", "segment.large_pool.allocated", "reserved_bytes.all.allocated",
                "reserved_bytes.small_pool.allocated", "reserved_bytes.large_pool.allocated") if k in st1})
        tot = sum(per_step)
        t = torch.tensor([tot], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # NVML is initialised BEFORE the warm-up (its start-up stalls driver calls for up to a second)
    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(max(args.warmup, 3)):
        step(resident)
    # untimed passes through the timing harness itself: with the host running ahead of the GPU the caching allocator
    # needs a few steps to reach its steady state (cudaMalloc inside a timed step is a 20-60 ms outlier)
    # (each timed loop directly behind a settle loop of its OWN mode: switching between the resident and the host-fed
    # loop changes the allocation sequence, and the first ~8 steps after a switch ran 0.5-1 ms slow on some boxes)
    timed(resident, SETTLE, False)
    if sampler:
        sampler.rows.clear()  # keep only samples taken during the timed regions
    calls0 = ops.launch_count()
    ms_total = timed(resident, args.steps, False)
    steps_ms = [round(v, 3) for v in timed.last_steps]
    mallocs = timed.last_mallocs
    launches = ops.launch_count() - calls0
    rows_keep = list(sampler.rows) if sampler else []
    timed(host, SETTLE, True)
    if sampler:
        sampler.rows[:] = rows_keep
    ms_e2e = timed(host, args.steps, True)
    scenes_per_step = args.bs * world
    value = scenes_per_step * args.steps / (ms_total / 1e3)
    e2e_val = scenes_per_step * args.steps / (ms_e2e / 1e3)

    peak, how = peaks()
    roof = None
    detail = None
    if not args.no_roofline:
        # every rank runs this extra step (the gradient all-reduce is a collective); only rank 0 keeps the records
        ops.profile_begin()
        step(resident)
        torch.cuda.synchronize()
        recs = ops.profile_end()
        if rank != 0:
            recs = []
        roof = roofline_from_records(recs, peak, how)
        detail = recs
    pm_levels = None
    if rank == 0:
        pm_levels = pairs_per_level(model, resident, dev, ref_model_fn is not None)
    gpu_native = None
    if rank == 0 and world == 1 and not args.no_gpu_native:
        gpu_native = run_gpu_native(model, batch, dev, value_ms=ms_total / args.steps, bs=args.bs)
    m32 = None
    if rank == 0 and world == 1 and not args.no_m32 and args.mid != 32 and ref_model_fn is None:
        m32 = run_other_width(32, batch, dev, model_step, SparseConvNet, ops, flush)
    cpu_base = parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:  # reported at N=1 only (rank 0, the box's host cores)
        from oracle.unet_ref import model_step_ref
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        sd = {k: v.detach().cpu().clone().requires_grad_(v.dtype.is_floating_point)
              for k, v in model.state_dict().items()}
        t0 = time.perf_counter()
        loss_c, scores_c = model_step_ref(sd, batch, training=True)
        loss_c.backward()
        dt = time.perf_counter() - t0
        cpu_base = {"value": args.bs / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                    "sample": "1 full step (bs=%d x %dk voxels, m=%d) of the CPU oracle, %.1f s"
                              % (args.bs, args.voxels // 1000, args.mid, dt)}
        # parity AT the full size: the engine's forward on the same weights / batch against the fp32 CPU oracle
        with torch.no_grad():
            if ref_model_fn is not None:
                r = ref_model_fn(resident, net, 0)
                loss_g, scores_g = r["loss"], r["output"]
            else:
                loss_g, scores_g = model_step(net, resident, criterion=criterion, device=dev)
        sc = scores_c.detach().double()
        parity = {"scores_rel": float((scores_g.double().cpu() - sc).abs().max() / sc.abs().max()),
                  "loss_abs": abs(float(loss_g) - float(loss_c.detach())), "oracle": "fp32 CPU port (oracle/unet_ref.py)",
                  "points": int(sc.shape[0]), "bar": "scores_rel <= 1e-4 (north_star); tests/test_fullsize_gpu.py holds "
                                                     "the same comparison against the fp64 oracle"}
    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
               "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "config": {"workload": WORKLOAD % (args.bs, args.mid), "voxels_per_scene": args.voxels,
                          "scenes_per_gpu": args.bs,
                          "mid_channel": args.mid, "model": "doda_b200/unet.py (mirror of the reference model)" if
                          ref_model_fn is None else "reference model/unet.py + unet_block.py + model_fn, unchanged, via compat/",
                          "parallelism": "dp%d (whole scenes per rank, %s)"
                          % (world, "DDP grad all-reduce" if args.ddp else ("in-place NCCL gradient mean started inside backward (OverlappedGradReducer)" if reducer is not None else "one in-place NCCL gradient mean after backward")), "l2": "256 MB flush between timed steps",
                          "settle": "%d further untimed steps through the timing harness before each timed loop"
                          % SETTLE, "levels": pm_levels},
               "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                       "ms_per_step": ms_e2e / args.steps,
                       "h2d": "pinned host -> device on the engine's index stream (ops.stage_batch): step i+1's inputs are "
                              "copied while step i computes, step 0 copies its own; all K copies inside the timed region"
                       if ref_model_fn is None else "the reference's test_model_feat copies the pinned host batch itself "
                                                    "(.cuda(non_blocking=True), model/unet.py:79-84) every step",
                       "loss_read": "async 4-byte copy to pinned memory every step; the host reads step i-1's value "
                                    "during step i and the last one before the closing sync (all K inside the timed "
                                    "region)"},
               "gpu_launches": launches, "clocks": sampler.summary() if sampler else None,
               "step_ms_rank0": steps_ms, "cuda_mallocs_in_timed_steps": mallocs}
        if roof:
            out["roofline"] = roof
        if gpu_native:
            out["baseline_gpu_native"] = gpu_native
            out["vs_gpu_native"] = gpu_native["vs_gpu_native"]
        if m32:
            out["m32"] = m32
        if cpu_base:
            out["cpu_baseline"] = cpu_base
        if parity:
            out["parity_full_size"] = parity
        print(json.dumps(out))
        if args.detail and detail is not None:
            with open(args.detail, "w") as f:
                json.dump(detail, f)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
