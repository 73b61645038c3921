"""In-tree build of the native libraries (no JIT cache: the .so files travel with the repo).

    libvfk.so    CUDA kernels + C ABI (include/vfk.h), nvcc, sm_100a only
    libvfkio.so  host companion (include/vfk_io.h): BGZF / BAM decode, INFO text
    libvfksynth.so  synthetic benchmark / test data (include/vfk_synth.h): not product code
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
LIBVFKSYNTH = os.path.join(PKG, "libvfksynth.so")

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


_SYNTH_SRCS = ("vfkio_synth.cpp", "vfkio_util.cpp")


def build_host(force: bool = False) -> str:
    srcs = sorted(s for s in glob.glob(os.path.join(CSRC, "vfkio_*.cpp")) if os.path.basename(s) != "vfkio_synth.cpp")
    deps = srcs + glob.glob(os.path.join(INCLUDE, "*.h"))
    if force or _stale(LIBVFKIO, deps):
        subprocess.check_call(["g++"] + CXX_FLAGS + ["-o", LIBVFKIO] + srcs + ["-lz", "-lpthread"])
    return LIBVFKIO


def build_synth(force: bool = False) -> str:
    srcs = [os.path.join(CSRC, s) for s in _SYNTH_SRCS]
    deps = srcs + glob.glob(os.path.join(INCLUDE, "*.h"))
    if force or _stale(LIBVFKSYNTH, deps):
        subprocess.check_call(["g++"] + CXX_FLAGS + ["-o", LIBVFKSYNTH] + srcs + ["-lpthread"])
    return LIBVFKSYNTH


def build_all(force: bool = False, verbose: bool = False):
    libs = build_cuda(force, verbose), build_host(force)
    build_synth(force)
    return libs


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose="-v" in sys.argv))
