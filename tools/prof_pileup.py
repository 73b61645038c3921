#!/usr/bin/env python
"""Minimal driver for profiling: one synthetic BAM (canonical batch resident in HBM), N launches of the pileup path (read preparation, locate, pileup kernels).
    python tools/prof_pileup.py [--tiles 200000] [--launches 6] [--workload wes|wgs|amplicon|multisample]"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from vafator_b200 import device, synth  # noqa: E402
from vafator_b200.batch import build_units  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--tiles", type=int, default=200_000)
ap.add_argument("--launches", type=int, default=6)
ap.add_argument("--workload", default="wes")
ap.add_argument("--snv-period", type=int, default=500, help="wes: one SNV per this many target bases (small = L2-resident reads, many units)")
a = ap.parse_args()
cfg = {"wes": lambda: synth.wes(n_tiles=a.tiles, snv_period=a.snv_period), "wgs": lambda: synth.wgs(contig_len=a.tiles * 4096 // 10, n_contigs=2),
       "amplicon": lambda: synth.amplicon(n_tiles=max(a.tiles // 100, 10)),
       "multisample": lambda: synth.multisample(0, n_tiles=a.tiles)}[a.workload]()  # one 60x sample of C4
t0 = time.time()
batch = synth.reads(cfg)
recs = synth.variant_records(cfg)
units = build_units(recs, cfg.contigs)
params = device.make_params(min_mapq=1, min_baseq=30)
dev = device.Device(0)
dreads = dev.upload_reads(batch)
dunits = dev.upload_units(units, None)
out = dev.alloc_counts(units.n_units)
print("setup %.1f s: %d reads, %d units" % (time.time() - t0, batch.n_reads, units.n_units), flush=True)
evs = []
for i in range(a.launches):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    dev.pileup(dreads, dunits, params, out=out)
    e1.record()
    evs.append((e0, e1))
torch.cuda.synchronize()
# the FP64 epilogue on the counts of the last launch
eaf = torch.full((units.n_units,), 0.4, dtype=torch.float64, device=dev.device)
stats = dev.alloc_stats(units.n_units)
eps = []
for i in range(a.launches):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    dev.epilogue(units.n_units, out, eaf, 5e-7, 1e-3, out=stats)
    e1.record()
    eps.append((e0, e1))
torch.cuda.synchronize()
print("epilogue ms per launch:", " ".join("%.3f" % x.elapsed_time(y) for x, y in eps))
ms = [x.elapsed_time(y) for x, y in evs]
c = dev.to_host(out, device.COUNTS_DTYPE, units.n_units)
print("pileup ms per launch:", " ".join("%.3f" % m for m in ms), "| mean n_cover %.2f n_cand %.2f dp %.2f" % (
    c["n_cover"].mean(), c["n_cand"].mean(), c["dp"].mean()))
st = (c["status"] >> 4)
print("units by status bits:", {int(k): int(v) for k, v in zip(*np.unique(st, return_counts=True))})
dev.set_timing(True)
dev.pileup(dreads, dunits, params, out=out)
print("phases ms:", dev.last_timing())
