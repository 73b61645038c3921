#!/usr/bin/env python
"""Alternative builds of libvfk.so for A/B runs on the GPU box (same sources, different -D tuning macros), written to
build/variants/<name>.so; select one with VFK_LIBVFK=build/variants/<name>.so.

    python tools/build_variants.py prep8=-DVFK_PREP_MIN_CTAS=8 warp3=-DVFK_MIN_CTAS=3
"""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vafator_b200 import build as B  # noqa: E402

out_dir = os.path.join(ROOT, "build", "variants")
os.makedirs(out_dir, exist_ok=True)
srcs = sorted(glob.glob(os.path.join(B.CSRC, "vfk_*.cu")))
procs = []
for arg in sys.argv[1:]:
    name, _, flags = arg.partition("=")
    out = os.path.join(out_dir, name + ".so")
    cmd = [B._nvcc()] + B.NVCC_FLAGS + flags.split(",") + ["-o", out] + srcs + ["-lgomp"]
    procs.append((name, subprocess.Popen(cmd)))
for name, p in procs:
    if p.wait() != 0:
        raise SystemExit("variant %s failed to build" % name)
    print("built", name)
