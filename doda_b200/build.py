"""Build libb200sparse.so (sm_100a) in-tree with nvcc.  Usage: python -m doda_b200.build [--force]"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libb200sparse.so")
FAST_SRC = os.path.join(CSRC, "fastcall.c")
FAST_OUT = os.path.join(HERE, "_b200fast.so")  # CPython module doda_b200._b200fast (fast-call binding of the hot entry points)
SOURCES = ["core.cu", "rulebook.cu", "conv.cu", "conv_tc.cu", "conv_direct.cu", "wgrad_tc.cu", "wgrad_direct.cu", "wgrad_os.cu", "layer.cu", "elementwise.cu", "loss.cu", "voxelize_gpu.cu", "pgops.cu", "pointops.cu", "augment.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", 
         "-Xcompiler", "-fPIC", "-DB200SP_BUILD"]
FLAGS = [f for f in FLAGS if f]


def _deps_mtime():
    m = 0.0
    for root, _, files in os.walk(CSRC):
        for f in files:
            m = max(m, os.path.getmtime(os.path.join(root, f)))
    m = max(m, os.path.getmtime(os.path.join(HERE, "..", "include", "b200sparse.h")))
    return m


def build_fast(force=False):
    """gcc: csrc/fastcall.c -> _b200fast.so, linked against libb200sparse.so next to it (rpath $ORIGIN)"""
    import sysconfig
    if not force and os.path.exists(FAST_OUT) and os.path.getmtime(FAST_OUT) >= max(_deps_mtime(), os.path.getmtime(OUT)):
        return FAST_OUT
    cmd = [os.environ.get("CC", "gcc"), "-O2", "-shared", "-fPIC", "-I" + sysconfig.get_paths()["include"], FAST_SRC,
           "-o", FAST_OUT, "-L" + HERE, "-lb200sparse", "-Wl,-rpath,$ORIGIN"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("gcc failed for fastcall.c:\n%s\n%s" % (r.stdout, r.stderr))
    return FAST_OUT


def build(force=False, verbose=False):
    out = _build_lib(force, verbose)
    build_fast(force)
    return out


def _build_lib(force=False, verbose=False):
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= _deps_mtime():
        return OUT
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)

    def cc(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [NVCC, *FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(cc, srcs))
    cmd = [NVCC, "-shared", "-o", OUT, *objs, "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
