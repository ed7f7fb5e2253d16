import numpy as np
import torch


def rel_err(a, b):
    """max|a-b| / max(max|b|, 1e-6)"""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-6))


def random_coords(seed, n, batch, shape):
    """n distinct voxels per batch element inside `shape`, in random (unsorted) order -> int32 [batch*n, 4]"""
    rng = np.random.RandomState(seed)
    out = []
    vol = int(np.prod(shape))
    for b in range(batch):
        flat = rng.choice(vol, size=min(n, vol), replace=False)
        c = np.stack(np.unravel_index(flat, shape), axis=1)
        out.append(np.concatenate([np.full((c.shape[0], 1), b), c], axis=1))
    return np.concatenate(out, 0).astype(np.int32)


def surface_coords(seed, target, batch):
    from doda_b200 import scenes
    out = []
    for b in range(batch):
        v = scenes.scene_with_voxels(seed + b, target)
        v = v[np.random.RandomState(seed + b).permutation(v.shape[0])]
        out.append(np.concatenate([np.full((v.shape[0], 1), b), v], axis=1))
    c = np.concatenate(out, 0).astype(np.int32)
    shape = (c[:, 1:].max(0) + 1).tolist()
    return c, shape


from oracle.gates import capture_relu_masks  # noqa: E402,F401  (kept importable from helpers)


def oracle_step(sd_f32, batch, dtype, relu_masks=None):
    """one forward + backward of the CPU oracle (oracle/unet_ref.py) in `dtype` -> (loss, scores, state_dict with .grad)"""
    from oracle.unet_ref import model_step_ref
    sd = {k: (v.detach().to(dtype).clone().requires_grad_(True) if v.is_floating_point() else v.clone())
          for k, v in sd_f32.items()}
    b = dict(batch)
    b["feats"] = batch["feats"].to(dtype)
    loss, scores = model_step_ref(sd, b, training=True, relu_masks=relu_masks)
    loss.backward()
    return loss.detach(), scores.detach(), sd


def grad_report(named_grads, sd64, sd32):
    """per-parameter max-norm errors and the global L2 error of `named_grads` against the fp64 oracle, next to the
    same numbers for the fp32 oracle (the reference algorithm in plain fp32 torch ops)"""
    e_gpu, e_f32, num, den, num32 = [], [], 0.0, 0.0, 0.0
    for name, g in named_grads:
        r = sd64[name].grad
        e_gpu.append(rel_err(g, r))
        e_f32.append(rel_err(sd32[name].grad, r))
        num += float((g.double().cpu() - r).pow(2).sum())
        num32 += float((sd32[name].grad.double() - r).pow(2).sum())
        den += float(r.pow(2).sum())
    return {"gpu_median": float(np.median(e_gpu)), "gpu_p90": float(np.percentile(e_gpu, 90)),
            "gpu_max": float(np.max(e_gpu)), "gpu_l2": float((num / den) ** 0.5),
            "f32_median": float(np.median(e_f32)), "f32_p90": float(np.percentile(e_f32, 90)),
            "f32_max": float(np.max(e_f32)), "f32_l2": float((num32 / den) ** 0.5)}


GRAD_CAP = 5e-2     # UNPINNED gates: absolute cap only (see assert_grad_parity); the fixed bars are the PINNED ones below


def pinned_grad_report(named_grads, sd64p):
    """errors against the fp64 oracle evaluated with the ReLU gates PINNED to the implementation's own (smooth net)"""
    e, num, den = [], 0.0, 0.0
    worst = ("", 0.0)
    for name, g in named_grads:
        r = sd64p[name].grad
        v = rel_err(g, r)
        e.append(v)
        if v > worst[1]:
            worst = (name, v)
        num += float((g.double().cpu() - r).pow(2).sum())
        den += float(r.pow(2).sum())
    return {"median": float(np.median(e)), "p90": float(np.percentile(e, 90)), "max": float(np.max(e)),
            "l2": float((num / den) ** 0.5), "worst": worst[0]}


PINNED_MEDIAN, PINNED_P90, PINNED_L2 = 1e-4, 2e-4, 1e-4   # fixed bars with pinned gates (measured 2-3e-5 / 3-4e-5 / 2-3e-5)


def assert_pinned_grad_parity(rep, what=""):
    assert rep["median"] <= PINNED_MEDIAN and rep["p90"] <= PINNED_P90 and rep["l2"] <= PINNED_L2, (what, rep)


def assert_grad_parity(rep, what=""):
    """Whole-net GRADIENTS of a 71-conv / 65-BN ReLU net cannot be held to a per-op bar with FREE gates: two fp32
    evaluations whose forward activations agree to ~1e-6 disagree on the sign of a ~1e-6 fraction of the ReLU
    pre-activations, and every flipped gate changes its gradient contribution by 100 % -- an L2 gradient difference
    of ~1e-3 per flip-carrying layer, up to ~1e-2 over the net.  Whether a given evaluation flips a gate is chance:
    the fp32 CPU oracle sits 5.7e-3 from its own fp64 evaluation at 2 x 150 k voxels and 1.4e-3 at 2 x 8 k, but
    7e-6 on a run where none of its gates happened to flip -- so a bound expressed as a multiple
    of the fp32 oracle's distance is a coin toss and is not used.  The engine is held (a) to FIXED tight bars against
    the fp64 oracle evaluated with the gates pinned to the engine's own (assert_pinned_grad_parity: the smooth part
    of the computation, i.e. the arithmetic), and (b) with free gates to an absolute cap that only a wrong kernel
    exceeds; the fp32 oracle's own numbers are printed next to the engine's for the record."""
    for k in ("median", "p90", "l2"):
        assert rep["gpu_" + k] <= GRAD_CAP, (what, k, rep)
