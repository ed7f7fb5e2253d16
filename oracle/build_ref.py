"""Compile the reference's OWN native extensions (PG_OP, pointops2_cuda) for sm_100a, unmodified, from the
sources where they lie under /root/reference, into oracle/_ref/ -- TEST INFRASTRUCTURE (the checker for the
PG_OP / pointops2 replacement kernels; never imported by the product).

  python -m oracle.build_ref [--force]

Needs two shim headers (oracle/shim/: google/dense_hash_map -> std::unordered_map, empty THC/THC.h).  Outputs
oracle/_ref/PG_OP.so and oracle/_ref/pointops2_cuda.so (git-ignored, NOT gpurun-ignored: they travel to the GPU
box).  `load(name)` imports a prebuilt module from oracle/_ref/ without touching /root/reference.
"""
import glob
import importlib.util
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
SHIM = os.path.join(HERE, "shim")
REF_LIB = "/root/reference/lib"

EXTS = {
    "PG_OP": ("pointgroup_ops", ["src/pointgroup_ops_api.cpp", "src/pointgroup_ops.cpp", "src/cuda.cu"]),
    "pointops2_cuda": ("pointops2", [
        "src/pointops_api.cpp", "src/knnquery/knnquery_cuda.cpp", "src/knnquery/knnquery_cuda_kernel.cu",
        "src/sampling/sampling_cuda.cpp", "src/sampling/sampling_cuda_kernel.cu",
        "src/sampling/sampling_dim_cuda_kernel.cu", "src/grouping/grouping_cuda.cpp",
        "src/grouping/grouping_cuda_kernel.cu", "src/interpolation/interpolation_cuda.cpp",
        "src/interpolation/interpolation_cuda_kernel.cu", "src/subtraction/subtraction_cuda.cpp",
        "src/subtraction/subtraction_cuda_kernel.cu", "src/aggregation/aggregation_cuda.cpp",
        "src/aggregation/aggregation_cuda_kernel.cu"]),
}


def so_path(name):
    return os.path.join(OUT, name + ".so")


def available(name):
    return os.path.exists(so_path(name))


def build_one(name, force=False, verbose=False):
    sub, srcs = EXTS[name]
    dst = so_path(name)
    if os.path.exists(dst) and not force:
        return dst
    root = os.path.join(REF_LIB, sub)
    if not os.path.isdir(root):
        raise RuntimeError("reference sources not present at %s" % root)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("MAX_JOBS", "8")
    from torch.utils.cpp_extension import load
    bdir = os.path.join(OUT, "build_" + name)
    os.makedirs(bdir, exist_ok=True)
    load(name=name, sources=[os.path.join(root, s) for s in srcs], extra_include_paths=[SHIM],
         extra_cflags=["-O2", "-w"], extra_cuda_cflags=["-O2", "-w"], build_directory=bdir, verbose=verbose,
         is_python_module=False)
    built = glob.glob(os.path.join(bdir, name + "*.so"))
    if not built:
        raise RuntimeError("build of %s produced no .so" % name)
    shutil.copy2(built[0], dst)
    shutil.rmtree(bdir, ignore_errors=True)
    return dst


def build_all(force=False, verbose=False):
    os.makedirs(OUT, exist_ok=True)
    return [build_one(n, force, verbose) for n in EXTS]


def load(name):
    """Import the prebuilt reference extension `name` from oracle/_ref/ (None if it was never built)."""
    p = so_path(name)
    if not os.path.exists(p):
        return None
    import torch  # noqa: F401  (libtorch must be loaded before the extension)
    spec = importlib.util.spec_from_file_location(name, p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose="-v" in sys.argv))
