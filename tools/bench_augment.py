"""f3 row timings: elastic (two calls, as the data pipeline makes them: gran 6 / mag 40 then gran 20 / mag 160 at voxel
scale 50) and crop on one 250 k-point scene -- device kernels vs the reference's numpy / scipy functions on one host core
(staged copy; falls back to the numpy oracle when the reference did not travel).  usage: python tools/bench_augment.py"""
import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from doda_b200 import augment
from oracle import stage_ref, augment as oracle_aug

warnings.simplefilter("ignore")
au = stage_ref.load_augmentor_utils()
kind = "reference (staged augmentor_utils.py)" if au is not None else "numpy oracle"
ref_elastic = au.elastic if au is not None else oracle_aug.elastic_ref
ref_crop = au.crop if au is not None else oracle_aug.crop_ref
rng = np.random.RandomState(0)
N = 250000
x = (rng.rand(N, 3) * np.array([420, 310, 140]) - np.array([210, 155, 0])).astype(np.float32)
xyz = (rng.rand(N, 3) * np.array([900, 700, 150])).astype(np.float64)
xd, xyzd = torch.from_numpy(x).cuda(), torch.from_numpy(xyz).cuda()


def cpu(fn, n=3):
    ts = []
    for _ in range(n):
        t = time.perf_counter(); fn(); ts.append(time.perf_counter() - t)
    return min(ts) * 1e3


def gpu(fn, n=10):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        t = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t)
    return float(np.median(ts)) * 1e3


def both_elastic(f, a):
    return f(f(a, 6, 40.0), 20, 160.0)


print("| op (N = %d points) | %s, 1 core, ms | B200 incl. host noise draw + upload, ms | ratio |" % (N, kind))
print("|---|---:|---:|---:|")
c = cpu(lambda: both_elastic(ref_elastic, x)); g = gpu(lambda: both_elastic(augment.elastic, xd))
print("| elastic x2 (gran 6 / 20) | %.1f | %.2f | %.0fx |" % (c, g, c / g))
np.random.seed(1); c = cpu(lambda: ref_crop(xyz, [128, 512], 2e9, 100000)); g = gpu(lambda: augment.crop(xyzd, [128, 512], 2e9, 100000))
print("| crop to 100 k points | %.1f | %.2f | %.0fx |" % (c, g, c / g))
m = np.eye(3) + 0.1
c = cpu(lambda: np.matmul(x, m)); g = gpu(lambda: augment.affine(xd, m))
print("| scene_aug matmul | %.2f | %.3f | %.0fx |" % (c, g, c / g))
