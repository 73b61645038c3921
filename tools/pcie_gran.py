#!/usr/bin/env python
"""Experiment: bytes crossing PCIe per ONE-byte in-place read of page-locked host memory (the access the pileup kernels make for a
read's quality and base when the host entry point leaves those pools on the host), per load flavour and L2 fetch granularity.
Time per access at a fixed access count is the measure (the link is the bottleneck): bytes/access = time x link rate / accesses.
    python tools/pcie_gran.py            (needs a GPU; builds tools/pcie_gran.cu into build/variants/)"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "build", "variants", "pcie_gran.so")
if not os.path.exists(so):
    os.makedirs(os.path.dirname(so), exist_ok=True)
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-shared", "-Xcompiler", "-fPIC", "-o", so,
                           os.path.join(ROOT, "tools", "pcie_gran.cu")])
lib = C.CDLL(so)
lib.pcie_gather.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
torch.cuda.init()
n, stride = 4_000_000, 256
host = torch.empty(n * stride + 64, dtype=torch.uint8, pin_memory=True)
host.fill_(1)
perm = torch.randperm(n, dtype=torch.int32).cuda()
out = torch.zeros(1, dtype=torch.int64, device="cuda")
# bulk copy rate of the same link, for scale
dst = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
dst.copy_(host[:256 << 20], non_blocking=True)
torch.cuda.synchronize()
e0.record()
for _ in range(4):
    dst.copy_(host[:256 << 20], non_blocking=True)
e1.record()
torch.cuda.synchronize()
link = 4 * (256 << 20) / (e0.elapsed_time(e1) * 1e-3)
print("bulk H2D copy: %.1f GB/s" % (link / 1e9))
names = ["ld.global.nc (__ldg)", "ld.volatile", "ld.global.cg", "ld.global.cs", "ld.global.nc.L1::no_allocate", "ld.global.cv"]
for gran in (32, 64, 128):
    rc = lib.pcie_set_granularity(gran)
    for mode, name in enumerate(names):
        lib.pcie_gather(mode, host.data_ptr(), n, stride, perm.data_ptr(), out.data_ptr(), None)
        torch.cuda.synchronize()
        e0.record()
        lib.pcie_gather(mode, host.data_ptr(), n, stride, perm.data_ptr(), out.data_ptr(), None)
        e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1) * 1e-3
        print("L2 fetch granularity %3d (rc %d)  %-30s %7.2f ms  %6.1f M accesses/s  ~%5.1f B per access at the bulk link rate" % (
            gran, rc, name, t * 1e3, n / t / 1e6, t * link / n))
