#!/usr/bin/env python
"""bench.py -- variants x BAMs annotated per second on B200, kernel-only and end to end.

Both arms start from the SAME boundary: the canonical columnar read batch (include/vfk.h: start, flag, MAPQ, CIGAR, 4-bit
sequence, qualities, mate fields, name id -- what a BAM decoder hands over) plus the variant records.  Everything derived
from the reads -- reference spans, position index, CIGAR shapes, read filter, mate pairing -- is inside the timed regions,
on the device for the GPU arm, in the C restatement for the CPU arm.

Workloads (BASELINE.json configs; --config):
  C3 (default)  ONE synthetic 30x WGS tumour BAM: 24 contigs x 10 Mbp (240 Mbp: chr1-22 scaled down 12x, stated), 48 M reads,
                1 variant per ~720 bp (90 % SNV, 5 % INS, 5 % DEL) => ~334 k units.  --gpus N cuts this one input into N
                interval shards of equal unit count (vafator_b200/sharding.py unit cuts), one per rank: STRONG scaling, no
                collective on the data path; each rank materialises only the reads of its own shard.
  C2            30x WES tumour/normal pair, 200 000 x 250 bp targets, 2 x 14.4 M reads, ~100 k SNVs => ~200 k units.
  C4            8 samples x 60x over a shared target set (--tiles, default 40 000 = 10 Mbp: 1/5 of the 50 Mbp of SURVEY.md 8d,
                stated), per-sample purity + 1 000-segment copy-number BED, 2 % multiallelic sites.
  C5            1000x amplicon panel: 5 000 amplicons, ~9 M reads, ~50 k variants (the histogram path).
One *step* = one pass of the hot path over the rank's shard of every BAM.  Resident in HBM that is ONE call per BAM shard; end
to end the same shard is fed as a few interval sub-batches (slices of the same host arrays, sharding.plan_shards) so that the
copies of one overlap the kernels of the previous one.

  value     units/s with the canonical batches resident in HBM: read preparation + locate + pileup + epilogue kernels
            (CUDA events around the K steps, max over ranks)
  e2e       units/s through vfk_annotate_host_async(): page-locked HOST arrays in, HOST results out, every copy inside
  roofline  pileup_warp_kernel: SURVEY.md 8(d) algorithmic bytes / its own CUDA-event time (vfk_last_timing), against the
            measured HBM copy peak (MEASURED_PEAKS.json)
  cpu_baseline / --impl reference
            the oracle port (C restatement of the pysam/htslib pileup incl. mate linking, OpenMP on all host cores + the
            reference's scalar scipy epilogue on a process pool) on a bounded sample of the same workload.  pysam itself is
            not installable on either box.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

# torchrun pins OMP_NUM_THREADS=1 for its children; the synthetic generator (OpenMP) should use this rank's share of the
# host cores instead.  Must happen before any OpenMP runtime is loaded.
if int(os.environ.get("WORLD_SIZE", "1")) > 1:
    os.environ["OMP_NUM_THREADS"] = str(max(2, (os.cpu_count() or 2) // int(os.environ["WORLD_SIZE"])))

# Rank 0 prints exactly ONE line on stdout.  Libraries write there too (NCCL prints its version banner with
# printf at NCCL_DEBUG=VERSION/WARN): file descriptor 1 is pointed at stderr for the rest of the process and the
# JSON line goes through a private duplicate of the original stdout.
sys.stdout.flush()
_RESULT_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MIN_MAPQ, MIN_BASEQ = 1, 30          # CLI defaults, vafator/command_line.py:63,71
FPR, ERROR_RATE = 5e-7, 1e-3         # vafator/power.py:10-11


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def pin_to_gpu_numa_node(local_rank: int):
    """Multi-GPU runs: keep this rank's host threads -- and with them its pinned buffers (first touch) -- on the
    NUMA node its GPU hangs off.  Best effort: returns the node, or None when the topology cannot be read."""
    try:
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = local_rank
        if vis:
            ids = [x.strip() for x in vis.split(",") if x.strip()]
            if local_rank < len(ids) and ids[local_rank].isdigit():
                idx = int(ids[local_rank])
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(idx)],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if not bus:
            return None
        dom, rest = bus.split(":", 1)
        path = "/sys/bus/pci/devices/%s:%s/numa_node" % (dom[-4:], rest)
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = []
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        cpus = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------
# workloads
# --------------------------------------------------------------------------------------------
class Workload:
    """A config of BASELINE.json: the samples (one synthetic BAM each, same genome and planted sites), their expected-VAF
    model and how many interval sub-batches a rank cuts each of its BAM shards into."""

    def __init__(self, name, args):
        from vafator_b200 import synth
        self.name = name
        S = synth.BASE_SEED
        if name == "C2":
            tiles = args.tiles or 200_000
            self.samples = [("tumor", synth.wes(n_tiles=tiles, seed=S + 2, normal=False)), ("normal", synth.wes(n_tiles=tiles, seed=S + 2, normal=True))]
            self.purity = {"tumor": 0.8, "normal": 1.0}
            self.sub = args.sub_batches or 2
            self.text = "C2: synthetic 30x WES tumor/normal, %d x 250bp targets, 2x150 reads, ~%dk SNVs, MQ>=1 BQ>=30" % (tiles, tiles // 2000)
            self.cpu_units = 20_000
        elif name == "C3":
            n_contigs, clen = args.contigs or 24, args.contig_len or 10_000_000
            self.samples = [("tumor", synth.wgs(n_contigs=n_contigs, contig_len=clen, seed=S + 3))]
            self.purity = {"tumor": 0.8}
            # a rank's shard goes through the device as a few interval sub-batches (the copies of one overlap the kernels of the
            # previous one); with more ranks a shard is smaller and fewer, larger launches serve it better
            self.sub = args.sub_batches or max(2, 8 // max(1, int(os.environ.get("WORLD_SIZE", "1"))))
            self.text = ("C3: ONE synthetic 30x WGS tumor BAM, %d contigs x %.0f Mbp (chr1-22 scaled down to %.0f Mbp), 2x150 reads, 1 variant / ~720 bp "
                         "(90%% SNV, 5%% INS, 5%% DEL), MQ>=1 BQ>=30; cut into one interval shard per GPU") % (
                             n_contigs, clen / 1e6, n_contigs * clen / 1e6)
            self.cpu_units = 40_000
        elif name == "C4":
            tiles = args.tiles or 40_000
            self.samples = [("s%d" % k, synth.multisample(k, n_tiles=tiles, seed=S + 4)) for k in range(8)]
            self.purity = {"s%d" % k: round(0.2 + 0.75 * k / 7.0, 3) for k in range(8)}
            self.sub = args.sub_batches or 1
            self.text = ("C4: 8 synthetic 60x samples over %d shared 250bp targets (%.0f Mbp; 50 Mbp in SURVEY.md 8d), per-sample purity + "
                         "1000-segment copy-number BED, 2%% multiallelic SNV sites, indels, MQ>=1 BQ>=30") % (tiles, tiles * 250 / 1e6)
            self.cpu_units = 16_000
        elif name == "C5":
            tiles = args.tiles or 5_000
            self.samples = [("panel", synth.amplicon(n_tiles=tiles, depth_pairs=900, seed=S + 5))]
            self.purity = {"panel": 0.6}
            self.sub = args.sub_batches or 2
            self.text = "C5: synthetic 1000x amplicon panel, %d amplicons, ~%dk variants (1 per ~25 bp), MQ>=1 BQ>=30" % (tiles, tiles * 10 // 1000)
            self.cpu_units = 1_500
        else:
            raise SystemExit("unknown config " + name)
        self.cfg0 = self.samples[0][1]
        self.recs = synth.variant_records(self.cfg0)
        self.contigs = self.cfg0.contigs
        last = {}
        for c, p, r, a in self.recs:
            last[c] = p
        self.last = last
        from vafator_b200.batch import build_units
        self.units_all = build_units(self.recs, self.contigs)
        self.stop_all = np.asarray([last[self.recs[i][0]] for i in self.units_all.rec], np.int32)

    def eaf(self, sample, recs):
        """Expected VAF per record (vafator/power.py:52-82); C4: per-sample copy-number BED of 1 000 segments."""
        pur = self.purity[sample]
        if self.name != "C4":
            ploidy = np.full(len(recs), 2.0)
        else:
            # 1 000 equal segments over the concatenated target space, CN in [0.5, 6] (the lookup itself is host work of
            # the reference too: vafator/ploidies.py:36-50; here a vectorised searchsorted)
            k = int(self.samples[[s for s, _ in self.samples].index(sample)][1].read_salt) % 97
            cn = 0.5 + 5.5 * ((np.arange(1000) * 37 + k * 11) % 101) / 100.0
            tid_of = {c: i for i, c in enumerate(self.contigs)}
            key = np.asarray([tid_of[r[0]] * (1 << 31) + r[1] for r in recs], np.int64)
            span = len(self.contigs) * (1 << 31)
            ploidy = cn[np.minimum((key * 1000) // span, 999)]
        return pur / (pur * np.maximum(1, ploidy) + (1 - pur) * 2)


def unit_cuts(n: int, k: int):
    from vafator_b200 import sharding
    return sharding.unit_cuts(n, k)


def build_batches(w: Workload, rank: int, world: int, pinned: bool, cap_units: int = None, sub: int = 1, params=None):
    """This rank's share of the ONE input: units [cut[rank], cut[rank+1]) of every BAM with the canonical reads of the tiles their
    columns can reach -- one batch per BAM.  sub > 1: also the same shard cut into `sub` interval sub-batches (slices of the same
    host arrays; only the re-based pool offsets are new arrays), for the end-to-end loop.  cap_units: only the first units
    (CPU sample).  Returns (batches, sub_batches, n_units)."""
    from vafator_b200 import synth, device
    from dataclasses import replace
    from vafator_b200 import sharding
    units_all, stop_all = w.units_all, w.stop_all
    n = units_all.n_units if cap_units is None else min(cap_units, units_all.n_units)
    cuts = unit_cuts(n, world)
    lo, hi = cuts[rank], cuts[rank + 1]
    batches, subs = [], []
    t0 = time.time()
    for sample, cfg in w.samples:
        eaf_all = w.eaf(sample, w.recs)
        if hi <= lo:
            continue
        units = sharding.slice_units(units_all, lo, hi)  # `rec` keeps indexing the records of the whole input
        g0, gc = synth.tile_range(cfg, int(units.tid[0]), int(units.pos[0]), int(units.tid[-1]), int(units.pos[-1]))
        batch = synth.reads(replace(cfg, g_first=g0, g_count=gc), pinned=pinned)
        stop = np.ascontiguousarray(stop_all[lo:hi])
        eaf = np.ascontiguousarray(eaf_all[units.rec], np.float64)
        batches.append(dict(sample=sample, batch=batch, units=units, stop=stop, eaf=eaf, recs=w.recs))
        if sub > 1:
            for sh in sharding.plan_shards(batch, units, sub, device.read_passes(batch, params), stop):
                subs.append(dict(sample=sample, batch=sharding.slice_reads(batch, sh.read_lo, sh.read_hi),
                                 units=sharding.slice_units(units, sh.unit_lo, sh.unit_hi), stop=np.ascontiguousarray(stop[sh.unit_lo:sh.unit_hi]),
                                 eaf=np.ascontiguousarray(eaf[sh.unit_lo:sh.unit_hi]), recs=w.recs))
    log("[bench] rank %d: %s units [%d, %d) of %d, %d batches (%d sub-batches), %d reads, %.1f s" % (
        rank, w.name, lo, hi, n, len(batches), len(subs), sum(b["batch"].n_reads for b in batches), time.time() - t0))
    return batches, (subs if sub > 1 else batches), n


def host_view(b, pinned: bool):
    """HostReads of a sub-batch whose per-read arrays are slices of a page-locked parent batch: only the re-based pool offsets
    (new arrays) are copied into page-locked memory."""
    from vafator_b200 import device
    h = device.HostReads(b["batch"], pinned=False)
    if pinned and not h.implicit_offsets:  # (implicit offsets: the arrays stay on the host, nothing reads them)
        for k in ("cigar_off", "seq_off", "qual_off"):
            p = device._empty(h.arrays[k].shape[0], h.arrays[k].dtype, True)
            p[:] = h.arrays[k]
            h.arrays[k] = p
    return h


def _epilogue_chunk(args):
    """The reference's per-unit scalar epilogue (vafator/annotator.py:382-420) on one chunk."""
    from oracle import vafator_ref as vr
    rows, eafs = args
    power = vr.Power(fpr=FPR, error_rate=ERROR_RATE)
    out = 0
    for (dp, ac_alt, ac_ref, s2), eaf in zip(rows, eafs):
        power.pu(dp, ac_alt, eaf)
        power.pw(dp, eaf)
        for m in range(3):
            if ac_alt and ac_ref:
                # scipy.stats.ranksums on lists of that size (vafator/rank_sum_test.py:11); the lists are
                # rebuilt with the same sizes -- their content does not change the cost
                vr.rank_sum(list(range(ac_alt)), list(range(ac_ref)))
        out += 1
    return out


def run_cpu_arm(w: Workload, repeats: int = 1):
    """Oracle port timed on the host cores: C mate linking + pileup (OpenMP) + scalar scipy epilogue (process pool), from the
    canonical arrays -- the boundary the GPU arm starts at."""
    import multiprocessing as mp
    from oracle import c_oracle
    cores = os.cpu_count() or 1
    batches, _, n_units_sample = build_batches(w, 0, 1, pinned=False, cap_units=w.cpu_units)
    tid_of = {c: i for i, c in enumerate(w.contigs)}
    prm = c_oracle.params(min_mapq=MIN_MAPQ, min_baseq=MIN_BASEQ)
    c_oracle.lib()
    jobs = []
    for b in batches:
        u, recs = b["units"], b["recs"]
        ounits = [(tid_of[recs[i][0]], recs[i][1], recs[i][2], recs[i][3][j], recs[i][3][0], w.last[recs[i][0]]) for i, j in zip(u.rec, u.alt_index)]
        jobs.append((b["batch"].as_dict(), ounits, b["eaf"].tolist()))
    times = []
    n_units = 0
    with mp.get_context("fork").Pool(cores) as pool:
        pool.map(_epilogue_chunk, [([(10, 3, 5, [0, 0, 0])], [0.5])] * (2 * cores))  # workers import scipy before the clock starts
        for _ in range(repeats):
            t0 = time.perf_counter()
            n_units = 0
            for d, ounits, eafs in jobs:
                mate = c_oracle.link_mates(d, prm)
                out = c_oracle.pileup(d, ounits, prm, mate)
                rows = list(zip(out["dp"].tolist(), out["ac_alt"].tolist(), out["ac_ref"].tolist(), out["s2"].tolist()))
                step = max(1, len(rows) // (cores * 4))
                chunks = [(rows[i:i + step], eafs[i:i + step]) for i in range(0, len(rows), step)]
                n_units += sum(pool.map(_epilogue_chunk, chunks))
            times.append(time.perf_counter() - t0)
    n_reads = sum(len(d["pos"]) for d, _, _ in jobs)
    return n_units, times, cores, n_reads


# --------------------------------------------------------------------------------------------
# clocks, PCIe
# --------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v == "Active"})
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=reasons, samples=len(sm))


class PcieRxSampler:
    """Bytes the GPU receives over PCIe, from NVML's RX throughput counter (a 20 ms window per query), sampled while the
    end-to-end loop runs: covers the bulk copies AND what the kernels read in place from page-locked host memory."""

    def __init__(self, index):
        self.samples, self.ok, self._stop = [], False, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                ids = [x.strip() for x in vis.split(",") if x.strip()]
                if index < len(ids) and ids[index].isdigit():
                    index = int(ids[index])
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            pynvml.nvmlDeviceGetPcieThroughput(self.h, pynvml.NVML_PCIE_UTIL_RX_BYTES)
            self.ok = True
        except Exception:
            self.ok = False

    def start(self):
        if self.ok:
            self.th = threading.Thread(target=self._run, daemon=True)
            self.th.start()

    def _run(self):
        while not self._stop:
            try:
                self.samples.append(self.nv.nvmlDeviceGetPcieThroughput(self.h, self.nv.NVML_PCIE_UTIL_RX_BYTES))  # KB/s
            except Exception:
                break

    def stop(self):
        self._stop = True
        if self.ok:
            self.th.join(timeout=1.0)
        return float(np.mean(self.samples)) * 1024.0 if self.samples else None  # bytes/s


# --------------------------------------------------------------------------------------------
# the two arms
# --------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C3", choices=["C2", "C3", "C4", "C5"])
    ap.add_argument("--tiles", type=int, default=0, help="C2 / C4 / C5: targets (amplicons) per BAM; 0 = the config's default")
    ap.add_argument("--contigs", type=int, default=0, help="C3: contigs (default 24)")
    ap.add_argument("--contig-len", type=int, default=0, help="C3: bases per contig (default 10 Mbp)")
    ap.add_argument("--sub-batches", type=int, default=0, help="interval sub-batches a rank cuts each BAM shard into")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        if rank != 0:
            return
        w = Workload(args.config, args)
        n_units, times, cores, n_reads = run_cpu_arm(w, repeats=args.warmup + args.steps)
        timed = times[args.warmup:]
        dt = sum(timed)
        val = n_units * len(timed) / dt
        sample = ("the first %d units (%d reads) of the same workload: C mate linking + pileup restatement (OpenMP) + "
                  "reference-style scalar scipy epilogue (process pool), from the canonical read arrays; pysam not installable") % (
                      n_units, n_reads)
        print(json.dumps({
            "impl": "reference", "metric": "variants x BAMs annotated per second", "value": val, "unit": "units/s",
            "n_gpus": args.gpus, "steps": len(timed), "warmup": args.warmup, "ms_per_step": 1e3 * dt / len(timed),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32/f64", "data": "synthetic",
            "config": {"workload": w.text, "sample": sample},
            "cpu_baseline": {"value": val, "unit": "units/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "units/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}), file=_RESULT_OUT, flush=True)
        return

    numa = pin_to_gpu_numa_node(local_rank) if world > 1 else None
    import torch
    from vafator_b200 import device
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU path for --impl b200)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    if world > 1 and "VFK_HOST_THREADS" not in os.environ:  # the host threads of the fetch-list gather: the box's cores shared out
        os.environ["VFK_HOST_THREADS"] = str(max(2, min(16, len(os.sched_getaffinity(0)) // world)))
    dev = device.Device(local_rank)
    params = device.make_params(min_mapq=MIN_MAPQ, min_baseq=MIN_BASEQ, include_ambiguous=True)
    w = Workload(args.config, args)
    batches, subs, units_total = build_batches(w, rank, world, pinned=not args.no_e2e, sub=1 if args.no_e2e else w.sub, params=params)
    for b in batches:
        b["host"] = device.HostReads(b["batch"])           # the canonical arrays, as generated (page-locked)
        b["dreads"] = dev.upload_reads(b["host"])          # ... and the same arrays resident in HBM, verbatim
        b["dunits"] = dev.upload_units(b["units"], b["stop"])
        b["deaf"] = torch.from_numpy(b["eaf"]).to(dev.device)
        b["counts"] = dev.alloc_counts(b["units"].n_units)
        b["stats"] = dev.alloc_stats(b["units"].n_units)
    for b in subs:
        if "host" not in b:
            b["host"] = host_view(b, pinned=True)
    torch.cuda.synchronize()
    units_per_step = sum(b["units"].n_units for b in batches)
    reads_rank = sum(b["batch"].n_reads for b in batches)

    def step():
        for b in batches:
            dev.pileup(b["dreads"], b["dunits"], params, out=b["counts"])
            dev.epilogue(b["units"].n_units, b["counts"], b["deaf"], FPR, ERROR_RATE, out=b["stats"])

    for _ in range(warmup):
        step()
    barrier()
    clocks = ClockSampler(local_rank)
    clocks.start()
    l0 = dev.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    launches = dev.launch_count() - l0
    ms_total = e0.elapsed_time(e1)
    # keep the GPU busy a little longer so that the clock sampler sees it under load
    t_end = time.time() + 1.0
    while time.time() < t_end:
        step()
    torch.cuda.synchronize()
    clk = clocks.stop()

    # ---- per-phase device times of the same calls (CUDA events recorded by the library on the launching stream)
    dev.set_timing(True)
    phase = dict(prep=0.0, locate=0.0, pileup=0.0, lists=0.0)
    n_timed = 0
    for _ in range(max(3, min(args.steps, 10))):
        for b in batches:
            dev.pileup(b["dreads"], b["dunits"], params, out=b["counts"])
            t = dev.last_timing()
            for k in phase:
                phase[k] += t[k]
            n_timed += 1
    dev.set_timing(False)
    phase = {k: v / n_timed for k, v in phase.items()}  # ms per launch (one batch)

    # sanity on the benchmarked data itself
    c_host = [dev.to_host(b["counts"], device.COUNTS_DTYPE, b["units"].n_units) for b in batches]
    for c in c_host:
        assert int(c["status"].max()) & (device.ST_DEFERRED | device.ST_READLEN | device.ST_BADBATCH) == 0
    d_raw = float(np.concatenate([c["n_cover"] for c in c_host]).mean())
    n_cand = float(np.concatenate([c["n_cand"] for c in c_host]).mean())

    # ---- end to end: page-locked canonical host arrays -> results on the host, through vfk_annotate_host_async ----
    e2e = None
    if not args.no_e2e:
        # results land in page-locked arrays (pageable ones work too: the library then bounces them, one more host copy in vfk_wait)
        c_out = [device.pinned_empty(b["units"].n_units, device.COUNTS_DTYPE) for b in subs]
        s_out = [device.pinned_empty(b["units"].n_units, device.STATS_DTYPE) for b in subs]

        n_slots = dev.slots

        def e2e_step():
            # a few batches in flight (the library's staging slots): the copies of one batch overlap the host gather and the kernels of others
            tickets = []
            for k, b in enumerate(subs):
                if k >= n_slots:
                    dev.wait(tickets[k - n_slots])
                tickets.append(dev.annotate_host_async(b["host"], b["units"], b["stop"], params, b["eaf"], FPR, ERROR_RATE, c_out[k], s_out[k]))
            for t in tickets[-n_slots:]:
                dev.wait(t)
        for _ in range(2):
            e2e_step()
        barrier()
        ts0 = dev.transfer_stats()
        rx = PcieRxSampler(local_rank)
        rx.start()
        t0 = time.perf_counter()
        n_e2e = max(3, min(args.steps, 5))
        for _ in range(n_e2e):
            e2e_step()
        torch.cuda.synchronize()
        dt_e2e = (time.perf_counter() - t0) / n_e2e
        rx_rate = rx.stop()
        ts1 = dev.transfer_stats()
        # the sub-batches of a BAM shard, put back in unit order, must equal the resident call on the whole shard
        at = 0
        for c in c_host:
            n_here = 0
            while n_here < len(c):
                for f in ("dp", "ac_alt", "ac_ref", "med2"):
                    assert np.array_equal(c_out[at][f], c[f][n_here:n_here + len(c_out[at])]), f
                n_here += len(c_out[at])
                at += 1
        assert at == len(subs)
        staged = (ts1["staged_bytes"] - ts0["staged_bytes"]) / n_e2e
        zero_copy = (ts1["zero_copy_batches"] - ts0["zero_copy_batches"]) / (n_e2e * len(subs))
        pairs = (ts1["fetch_pairs"] - ts0["fetch_pairs"]) / n_e2e
        gather_ms = (ts1["fetch_ns"] - ts0["fetch_ns"]) / n_e2e / 1e6
        d2h = (ts1["d2h_bytes"] - ts0["d2h_bytes"]) / n_e2e + 16.0 * pairs  # results + the fetch list's requests (written by the kernel)
        if rx_rate is not None:
            h2d, h2d_how = rx_rate * dt_e2e, "NVML PCIe RX counter sampled over the timed region x step time"
        else:
            # fallback when NVML has no counter: the library's own count of copied bytes + per candidate read the in-place reads
            # (header fields 27 B, quality + base sectors 64 B)
            h2d, h2d_how = staged + zero_copy * units_per_step * n_cand * 91.0, "library-counted copies + 91 B per candidate read (no NVML counter)"
        transfer = ("page-locked canonical arrays: starts / CIGARs / read lengths copied; the quality and sequence byte of every (column, read) pair "
                    "named by the device on a fetch list (%.1f M pairs per step, 16 B each written to the host), gathered by %d host threads between "
                    "two kernels of the call and copied back (bulk copies: %.0f MB per step, counted by the library); flags / MAPQ / mate fields / name "
                    "ids read in place over PCIe for the reads that overlap a column; %d batches in flight"
                    % (pairs / 1e6, ts1["fetch_threads"], staged / 1e6, n_slots)) if zero_copy > 0.5 and pairs > 0 else \
                   ("page-locked canonical arrays: starts / CIGARs / read lengths copied (%.0f MB per step, counted by the library), every other "
                    "array read in place over PCIe for the reads that overlap a column; %d batches in flight" % (staged / 1e6, n_slots)) if zero_copy > 0.5 else \
                   "every array staged by async copies (%.0f MB per step); %d batches in flight" % (staged / 1e6, n_slots)
        e2e = dict(dt=dt_e2e, h2d=h2d, d2h=d2h, transfer=transfer, staged=staged, h2d_how=h2d_how, rx_gbs=None if rx_rate is None else rx_rate / 1e9,
                   pairs=pairs, gather_ms=gather_ms, threads=ts1["fetch_threads"])

    # ---- max over ranks ----
    if dist is not None:
        t = torch.tensor([ms_total, e2e["dt"] if e2e else 0.0, phase["pileup"], phase["prep"], phase["locate"], phase["lists"]],
                         device=dev.device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, e2e_dt = float(t[0]), float(t[1])
        phase = dict(pileup=float(t[2]), prep=float(t[3]), locate=float(t[4]), lists=float(t[5]))
        if e2e:
            e2e["dt"] = e2e_dt
        tot = torch.tensor([units_per_step, reads_rank, e2e["h2d"] if e2e else 0.0, e2e["d2h"] if e2e else 0.0], device=dev.device, dtype=torch.float64)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        units_step_all, reads_all = float(tot[0]), float(tot[1])
        if e2e:
            e2e["h2d"], e2e["d2h"] = float(tot[2]), float(tot[3])
    else:
        units_step_all, reads_all = float(units_per_step), float(reads_rank)

    if rank == 0:
        peak, peak_src = measured_peak()
        # B_unit = 16 (variant) + 8 (index bounds) + 32 * D_raw + 160 (outputs), biallelic
        b_unit = 16 + 8 + 32.0 * d_raw + 160
        units_per_launch = units_per_step / len(batches)
        # the dominant kernel of the call: the register kernel (30x / 60x columns) or the list path (deep columns: tile_kernel, with the
        # medium / tail kernels' share of that phase)
        dominant, dom_ms = ("pileup_warp_kernel", phase["pileup"]) if phase["pileup"] >= phase["lists"] else ("tile_kernel", phase["lists"])
        achieved = units_per_launch * b_unit / (dom_ms * 1e-3) / 1e9
        reads_per_launch = reads_rank / len(batches)
        cig = float(np.mean([len(b["batch"].cigar) / max(b["batch"].n_reads, 1) for b in batches]))
        prep_bytes = reads_per_launch * (4 + 4 + 8 + 4 + 4 * cig + 8)  # starts, CIGAR counts / offsets / words, read lengths in; end + shape out
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            with open(tp) as fh:
                traffic = json.load(fh).get("r02_%s_%s_dram_bytes_per_launch" % (w.name, dominant))
        cpu = None
        if not args.no_cpu_baseline:
            n_cpu, times, cores, n_reads_cpu = run_cpu_arm(w, repeats=1)
            cpu = {"value": n_cpu / times[0], "unit": "units/s", "cores": cores, "kind": "port",
                   "sample": "the first %d units (%d reads, %.1f s) of the same workload: C mate linking + pileup restatement (OpenMP) + "
                             "reference-style scalar scipy epilogue (process pool), from the canonical read arrays; pysam itself is not "
                             "installable here" % (n_cpu, n_reads_cpu, times[0])}
        line = {
            "metric": "variants x BAMs annotated per second", "value": units_step_all * args.steps / (ms_total * 1e-3),
            "unit": "units/s", "n_gpus": world, "steps": args.steps, "warmup": warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32/f64", "data": "synthetic",
            "config": {"workload": w.text, "units_total": int(units_step_all), "reads_total": int(reads_all),
                       "units_per_step_rank0": units_per_step, "batches_per_step_rank0": len(batches), "e2e_sub_batches_rank0": len(subs),
                       "boundary": "canonical columnar read batch (decoded BAM fields, untouched) + variant units in; counts + statistics out. "
                                   "Read preparation (spans, position index, CIGAR shapes, filter, mate pairing) runs on the device inside every step",
                       "mean_depth_raw": round(d_raw, 2), "mean_candidates": round(n_cand, 2),
                       "l2": "inputs (%.1f GB of canonical read arrays on rank 0) exceed L2 many times over" % (
                           sum(b["host"].nbytes() for b in batches) / 1e9),
                       "sharding": "ONE input; units cut into %d equal interval shards (sharding.unit_cuts), shard k -> rank k, each rank holds the "
                                   "reads its columns can reach; no collective on the data path" % world,
                       "numa_node_rank0": numa},
            "roofline": {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_unit": b_unit, "units_per_launch": units_per_launch, "kernel_ms": dom_ms,
                         "phases_ms_per_launch": {k: round(v, 5) for k, v in phase.items()},
                         "prep_kernels": {"bytes_per_launch": prep_bytes, "achieved": prep_bytes / (phase["prep"] * 1e-3) / 1e9 if phase["prep"] > 0 else None,
                                          "unit": "GB/s", "frac": prep_bytes / (phase["prep"] * 1e-3) / 1e9 / peak if phase["prep"] > 0 else None}},
            "cpu_baseline": cpu,
            "e2e": None if e2e is None else {"value": units_step_all / e2e["dt"], "unit": "units/s",
                                            "h2d_bytes_per_step": int(e2e["h2d"]), "d2h_bytes_per_step": int(e2e["d2h"]),
                                            "ms_per_step": 1e3 * e2e["dt"], "h2d_measured_by": e2e["h2d_how"],
                                            # read preparation (spans, shapes, position index: prep + scan kernels) runs on the device inside
                                            # every timed call; its device time per step, from the resident calls' phase events
                                            "prep_ms": round(phase["prep"] * len(batches), 4),
                                            "h2d_copied_bytes_per_step_rank0": int(e2e["staged"]), "pcie_rx_gbs_rank0": e2e["rx_gbs"],
                                            "fetch_pairs_per_step_rank0": int(e2e["pairs"]), "host_gather_ms_per_step_rank0": round(e2e["gather_ms"], 3),
                                            "host_gather_threads": e2e["threads"],
                                            "transfer": e2e["transfer"]},
            "gpu_launches": int(launches), "clocks": clk,
        }
        print(json.dumps(line), file=_RESULT_OUT, flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
