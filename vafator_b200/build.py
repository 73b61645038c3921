"""In-tree build of the native libraries (no JIT cache: the .so files travel with the repo).

    libvfk.so    CUDA kernels + C ABI (include/vfk.h), nvcc, sm_100a only
    libvfkio.so  host companion (include/vfk_io.h): BGZF / BAM decode, INFO text, synthetic data
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
INCLUDE = os.path.join(ROOT, "include")
LIBVFK = os.path.join(PKG, "libvfk.so")
LIBVFKIO = os.path.join(PKG, "libvfkio.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fopenmp", "-shared", "-I", INCLUDE]
CXX_FLAGS = ["-O3", "-std=c++17", "-fPIC", "-shared", "-fopenmp", "-I", INCLUDE]


def _stale(target: str, sources) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA library cannot be built")


def build_cuda(force: bool = False, verbose: bool = False) -> str:
    srcs = sorted(glob.glob(os.path.join(CSRC, "vfk_*.cu")))
    deps = srcs + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(INCLUDE, "*.h"))
    if force or _stale(LIBVFK, deps):
        extra = os.environ.get("VFK_NVCC_EXTRA", "").split()
        cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIBVFK] + srcs + ["-lgomp"]
        subprocess.check_call(cmd)
    return LIBVFK


def build_host(force: bool = False) -> str:
    srcs = sorted(glob.glob(os.path.join(CSRC, "vfkio_*.cpp")))
    deps = srcs + glob.glob(os.path.join(INCLUDE, "*.h"))
    if force or _stale(LIBVFKIO, deps):
        subprocess.check_call(["g++"] + CXX_FLAGS + ["-o", LIBVFKIO] + srcs + ["-lz", "-lpthread"])
    return LIBVFKIO


def build_all(force: bool = False, verbose: bool = False):
    return build_cuda(force, verbose), build_host(force)


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose="-v" in sys.argv))
