"""Builds libaoslo_b200.so (hand-written sm_100a CUDA + the C ABI) in-tree with nvcc.

    python -m detecting_cones_aoslo_b200.build [--force]

particles.cu is compiled with -fmad=false (FP64 decisions must round like NumPy);
blobness.cu keeps FMA for its f32 path and uses explicit non-fused ops for f64.
"""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
BUILD = os.path.join(PKG, "build")
LIB = os.path.join(PKG, "libaoslo_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"),
          "-I", CSRC]
UNITS = [
    ("blobness.cu", []),
    ("particles.cu", ["-fmad=false"]),
    ("hungarian.cu", ["-fmad=false"]),
    ("api.cu", []),
    ("kdtree_host.cpp", []),
    ("io_host.cpp", []),
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(force=False, verbose=False):
    nvcc = _nvcc()
    os.makedirs(BUILD, exist_ok=True)
    headers = [os.path.join(CSRC, "common.cuh"), os.path.join(ROOT, "include", "aoslo_b200.h"),
               os.path.abspath(__file__)]
    objs = []
    for src, extra in UNITS:
        spath = os.path.join(CSRC, src)
        obj = os.path.join(BUILD, os.path.splitext(src)[0] + ".o")
        objs.append(obj)
        if force or _stale(obj, [spath] + headers):
            cmd = [nvcc] + ARCH + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", spath, "-o", obj]
            if verbose:
                print(" ".join(cmd))
            subprocess.check_call(cmd)
    if force or _stale(LIB, objs):
        cmd = [nvcc] + ARCH + ["-shared", "-o", LIB] + objs
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
