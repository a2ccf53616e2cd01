"""Builds latticeum_b200/lib/liblattice_ajtai.so (CUDA kernels + C ABI) for sm_100a with nvcc, in-tree.

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
LIB = os.path.join(LIBDIR, "liblattice_ajtai.so")
SOURCES = ["engine.cu", "ring_kernels.cu", "mac_kernels.cu", "ntt_pow2.cu"]
HEADERS = ["goldilocks.cuh", "ring24.cuh", "ring8.cuh", "ring96.cuh", "spin.cuh", "tma.cuh", "kernels.h", os.path.join("..", "..", "include", "lattice_ajtai.h")]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the engine is CUDA-only and cannot be built without it")


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    objs = []
    for src in SOURCES:
        obj = os.path.join(LIBDIR, src.replace(".cu", ".o"))
        cmd = [nvcc(), *ARCH, "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++",
               "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        subprocess.check_call(cmd)
        objs.append(obj)
    subprocess.check_call([nvcc(), *ARCH, "-shared", "-ccbin", "/usr/bin/g++", "-o", LIB, *objs])
    return LIB


def build_variant(tag: str, defines) -> str:
    """A tuning build with extra -D flags into lib/variants/<tag>/ (tools/tune_mac.py loads it through LAT_LIB)."""
    out = os.path.join(LIBDIR, "variants", tag)
    os.makedirs(out, exist_ok=True)
    lib = os.path.join(out, "liblattice_ajtai.so")
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    subprocess.check_call([nvcc(), *ARCH, "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++",
                           "-shared", *[f"-D{d}" for d in defines], "-o", lib, *srcs])
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
