#!/usr/bin/env python
"""bench.py -- variants x BAMs annotated per second on B200, kernel-only and end to end.

Workload at N=1 (BASELINE.json configs[1], "C2"): synthetic 30x WES tumour/normal pair --
200 000 targets of 250 bp (50 Mbp), 2 x 14.4 M reads, ~100 k SNVs, CLI default filters
(MQ>=1, BQ>=30) => ~200 k (variant, BAM) units per step.  One *step* = one pass of the hot path
(pileup kernel + FP64 epilogue kernel) over both BAMs.  For N>1 every rank owns an equally
sized, independent shard (different genomic intervals: different seed) -- weak scaling, no
collective on the data path; torch.distributed is used for the barrier and the
max-over-ranks only.

  value     units/s with the read batches resident in HBM (CUDA events around the K steps)
  e2e       units/s through vfk_annotate_host(): pinned HOST buffers in, HOST results out,
            H2D + kernels + D2H inside the timed region
  roofline  pileup kernel: SURVEY.md 8(d) algorithmic bytes / CUDA-event time of that kernel,
            against the measured HBM copy peak (MEASURED_PEAKS.json)
  cpu_baseline / --impl reference
            the oracle port (C restatement of the pysam/htslib pileup, OpenMP on all host cores
            + the reference's scalar scipy epilogue on a process pool) on a bounded sample of the
            same workload.  pysam itself is not installable on either box.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

# torchrun pins OMP_NUM_THREADS=1 for its children; the host-side batch builder (OpenMP) should use
# this rank's share of the host cores instead.  Must happen before any OpenMP runtime is loaded.
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
PURITY, PLOIDY = {"tumor": 0.8, "normal": 1.0}, {"tumor": 2.0, "normal": 2.0}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def pin_to_gpu_numa_node(local_rank: int):
    """Multi-GPU runs: keep this rank's host threads -- and with them its pinned buffers (first touch) -- on the
    NUMA node its GPU hangs off, so that eight ranks do not all pull their batches through one memory
    controller / PCIe root.  Best effort: returns the node, or None when the topology cannot be read."""
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


def eaf_of(sample):
    pur = PURITY[sample]
    return pur / (pur * max(1, PLOIDY[sample]) + (1 - pur) * 2)


# --------------------------------------------------------------------------------------------
# workload
# --------------------------------------------------------------------------------------------
def build_workload(n_tiles: int, rank: int, pinned: bool, no_columns: bool = False):
    """Tumour + normal batch over the same targets / sites; returns packed host batches + units."""
    from vafator_b200 import device, synth
    from vafator_b200.batch import build_units
    params = device.make_params(min_mapq=MIN_MAPQ, min_baseq=MIN_BASEQ, include_ambiguous=True)
    seed = synth.BASE_SEED + 2 + 1000 * rank
    t0 = time.time()
    if not no_columns:
        # both layouts of the two BAMs are ~23 GB of (page-locked) host memory per rank, plus ~12 GB while a batch is
        # being built: on a host that cannot give every rank that much the run keeps to the read-major layout
        try:
            import psutil
            avail = psutil.virtual_memory().available / max(int(os.environ.get("WORLD_SIZE", "1")), 1)
            need = 40e9 * n_tiles / 200_000
            if avail < need:
                log("[bench] rank %d: %.0f GB of host memory per rank < %.0f GB: no column store" % (rank, avail / 1e9, need / 1e9))
                no_columns = True
        except ImportError:
            pass
    cfg_t = synth.wes(n_tiles=n_tiles, seed=seed, normal=False)
    recs = synth.variant_records(cfg_t)
    units = build_units(recs, cfg_t.contigs)
    last = {}
    for c, p, r, a in recs:
        last[c] = p
    stop = np.asarray([last[recs[i][0]] for i in units.rec], np.int32)
    packed = {}
    n_reads = 0
    for sample, normal in (("tumor", False), ("normal", True)):
        # same seed => same genome and sites; the normal sample differs in VAF spectrum, which also
        # re-draws which reads carry an allele
        cfg = synth.wes(n_tiles=n_tiles, seed=seed, normal=normal)
        batch = synth.reads(cfg)
        n_reads += batch.n_reads
        packed[sample] = device.pack_reads(batch, params, pinned=pinned, columns=not no_columns)
        del batch
    log("[bench] rank %d workload: %d tiles, %d reads, %d records, %d units/BAM, %.1f s" %
        (rank, n_tiles, n_reads, len(recs), units.n_units, time.time() - t0))
    return dict(params=params, recs=recs, units=units, stop=stop, packed=packed, n_reads=n_reads, columns=not no_columns)


def cpu_sample(n_tiles: int, rank: int):
    """Canonical (un-packed) batches of a bounded sample of the same workload, for the CPU arm."""
    from vafator_b200 import synth
    from vafator_b200.batch import build_units
    seed = synth.BASE_SEED + 2 + 1000 * rank
    cfg = synth.wes(n_tiles=n_tiles, seed=seed)
    recs = synth.variant_records(cfg)
    units = build_units(recs, cfg.contigs)
    tid_of = {c: i for i, c in enumerate(cfg.contigs)}
    last = {}
    for c, p, r, a in recs:
        last[c] = p
    ounits = [(tid_of[recs[i][0]], recs[i][1], recs[i][2], recs[i][3][j], recs[i][3][0], last[recs[i][0]])
              for i, j in zip(units.rec, units.alt_index)]
    batches = {s: synth.reads(synth.wes(n_tiles=n_tiles, seed=seed, normal=(s == "normal"))) for s in ("tumor", "normal")}
    return batches, ounits


def _epilogue_chunk(args):
    """The reference's per-unit scalar epilogue (vafator/annotator.py:382-420) on one chunk."""
    from oracle import vafator_ref as vr
    rows, eaf = args
    power = vr.Power(fpr=FPR, error_rate=ERROR_RATE)
    out = 0
    for dp, ac_alt, ac_ref, s2 in rows:
        power.pu(dp, ac_alt, eaf)
        power.pw(dp, eaf)
        for m in range(3):
            if ac_alt and ac_ref:
                # scipy.stats.ranksums on lists of that size (vafator/rank_sum_test.py:11); the lists are
                # rebuilt with the same sizes -- their content does not change the cost
                vr.rank_sum(list(range(ac_alt)), list(range(ac_ref)))
        out += 1
    return out


def run_cpu_arm(n_tiles: int, rank: int, repeats: int = 1):
    """Oracle port timed on the host cores: C pileup (OpenMP) + scalar scipy epilogue (process pool)."""
    import multiprocessing as mp
    from oracle import c_oracle
    cores = os.cpu_count() or 1
    batches, ounits = cpu_sample(n_tiles, rank)
    prm = c_oracle.params(min_mapq=MIN_MAPQ, min_baseq=MIN_BASEQ)
    c_oracle.lib()
    times = []
    n_units = 0
    with mp.get_context("fork").Pool(cores) as pool:
        for _ in range(repeats):
            t0 = time.perf_counter()
            n_units = 0
            for sample, batch in batches.items():
                d = batch.as_dict()
                mate = c_oracle.link_mates(d, prm)
                out = c_oracle.pileup(d, ounits, prm, mate)
                rows = list(zip(out["dp"].tolist(), out["ac_alt"].tolist(), out["ac_ref"].tolist(), out["s2"].tolist()))
                step = max(1, len(rows) // (cores * 4))
                chunks = [(rows[i:i + step], eaf_of(sample)) for i in range(0, len(rows), step)]
                n_units += sum(pool.map(_epilogue_chunk, chunks))
            times.append(time.perf_counter() - t0)
    return n_units, times, cores


# --------------------------------------------------------------------------------------------
# clocks
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


# --------------------------------------------------------------------------------------------
# the two arms
# --------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--tiles", type=int, default=200_000, help="WES targets per BAM (200000 = the 50 Mbp C2 config)")
    ap.add_argument("--cpu-tiles", type=int, default=20_000, help="targets in the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-columns", action="store_true", help="read-major kernels only (no column store in the batches)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    workload = "C2: synthetic 30x WES tumor/normal, %d x 250bp targets, 2x150 reads, ~%dk SNVs, MQ>=1 BQ>=30" % (
        args.tiles, args.tiles // 2000)

    if args.impl == "reference":
        if rank != 0:
            return
        n_units, times, cores = run_cpu_arm(args.cpu_tiles, 0, repeats=args.warmup + args.steps)
        timed = times[args.warmup:]
        dt = sum(timed)
        val = n_units * len(timed) / dt
        sample = ("%d of %d targets of the same workload (%d units/step): C pileup restatement (OpenMP) + "
                  "reference-style scalar scipy epilogue (process pool); pysam not installable") % (
                      args.cpu_tiles, args.tiles, n_units)
        print(json.dumps({
            "impl": "reference", "metric": "variants x BAMs annotated per second", "value": val, "unit": "units/s",
            "n_gpus": args.gpus, "steps": len(timed), "warmup": args.warmup, "ms_per_step": 1e3 * dt / len(timed),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32/f64", "data": "synthetic",
            "config": {"workload": workload, "sample": sample},
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

    dev = device.Device(local_rank)
    w = build_workload(args.tiles, rank, pinned=not args.no_e2e, no_columns=args.no_columns)
    units, stop, params = w["units"], w["stop"], w["params"]
    nu = units.n_units
    samples = ("tumor", "normal")
    dunits = dev.upload_units(units, stop)
    dreads = {s: dev.upload_reads(w["packed"][s]) for s in samples}
    eaf_host = {s: np.full(nu, eaf_of(s)) for s in samples}
    eaf_dev = {s: torch.from_numpy(eaf_host[s]).to(dev.device) for s in samples}
    counts = {s: dev.alloc_counts(nu) for s in samples}
    stats = {s: dev.alloc_stats(nu) for s in samples}
    torch.cuda.synchronize()
    units_per_step = nu * len(samples)

    k_ev = []  # (start, end) CUDA events around every pileup launch of the timed region

    def step(record):
        for s in samples:
            if record:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
            dev.pileup(dreads[s], dunits, params, out=counts[s])
            if record:
                b.record()
                k_ev.append((a, b))
            dev.epilogue(nu, counts[s], eaf_dev[s], FPR, ERROR_RATE, out=stats[s])

    for _ in range(warmup):
        step(False)
    barrier()
    clocks = ClockSampler(local_rank)
    clocks.start()
    l0 = dev.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step(True)
    e1.record()
    barrier()
    launches = dev.launch_count() - l0
    ms_total = e0.elapsed_time(e1)
    # keep the GPU busy a little longer so that the clock sampler sees it under load
    t_end = time.time() + 1.0
    while time.time() < t_end:
        step(False)
    torch.cuda.synchronize()
    clk = clocks.stop()
    k_ms = float(np.mean([a.elapsed_time(b) for a, b in k_ev]))

    # parity guard on the benchmarked data itself: depth statistics must be sane and reproducible
    c_host = dev.to_host(counts["tumor"], device.COUNTS_DTYPE, nu)
    assert int(c_host["status"].max()) & (device.ST_DEFERRED | device.ST_READLEN) == 0
    d_raw = float(c_host["n_cover"].mean())
    n_cand = float(c_host["n_cand"].mean())

    # ---- end to end: pinned host buffers -> results on the host, through vfk_annotate_host ----
    e2e = None
    if not args.no_e2e:
        c_out = {s: np.empty(nu, device.COUNTS_DTYPE) for s in samples}
        s_out = {s: np.empty(nu, device.STATS_DTYPE) for s in samples}

        def e2e_step():
            # both BAMs are submitted before either is waited for: the library double-buffers, so the
            # copies of the second batch overlap the kernels of the first
            tickets = [dev.annotate_host_async(w["packed"][s], units, stop, params, eaf_host[s], FPR, ERROR_RATE,
                                               c_out[s], s_out[s]) for s in samples]
            for t in tickets:
                dev.wait(t)
        for _ in range(2):
            e2e_step()
        barrier()
        ts0 = dev.transfer_stats()
        cb0 = dev.column_batches()
        t0 = time.perf_counter()
        n_e2e = max(3, min(args.steps, 5))
        for _ in range(n_e2e):
            e2e_step()
        torch.cuda.synchronize()
        dt_e2e = (time.perf_counter() - t0) / n_e2e
        ts1 = dev.transfer_stats()
        assert np.array_equal(c_out["tumor"]["dp"], c_host["dp"]) and np.array_equal(c_out["tumor"]["ac_alt"], c_host["ac_alt"])
        n_batches = n_e2e * len(samples)
        zc = (ts1["zero_copy_batches"] - ts0["zero_copy_batches"]) / n_batches
        zc_hot = (ts1["zero_copy_hot_batches"] - ts0["zero_copy_hot_batches"]) / n_batches
        host_loc = (ts1["host_located_batches"] - ts0["host_located_batches"]) / n_batches
        col = (dev.column_batches() - cb0) / n_batches
        # bytes the library copied in bulk (counted by the library) + what the kernels gathered from pinned host
        # memory: per candidate read one 32-byte payload sector plus the overlap partner's / cold record's share
        # (~64 B), and its 20-byte header when the hot / mate arrays are gathered too
        h2d = (ts1["staged_bytes"] - ts0["staged_bytes"]) / n_e2e
        if col > 0.5:
            # column store: one contiguous run of 32-byte sectors per unit (n_cand = sectors of the column's window);
            # the few units left to the read-major kernels (~2 %) gather ~84 B per candidate read as before
            # (staged VFK_COL_TRANSPOSED=1 layout: one 8-byte header pair and one 2-byte entry per sector of the run)
            per_sector = 10.0 if os.environ.get("VFK_COL_TRANSPOSED", "")[:1] == "1" else 32.0
            if os.environ.get("VFK_HOST_GATHER", "")[:1] == "1":
                per_sector = 0.0  # staged: the runs are gathered on the host and bulk-copied (counted by the library)
            h2d += len(samples) * nu * n_cand * (per_sector + 0.02 * 84.0)
        else:
            h2d += len(samples) * nu * n_cand * (64.0 * zc + 20.0 * zc_hot)
        transfer = ("column store: SNV units read one contiguous sector run each from pinned host memory; " if col > 0.5 else "") + (
            "units located %s; headers %s; payload + cold + CIGAR %s; two batches in flight (double buffered)" % (
            "on the host (32 B per unit copied, no index transfer)" if host_loc > 0.5 else "by the locate kernel (index staged)",
            "gathered by the kernel from pinned host memory" if zc_hot > 0.5 else "staged by async copies",
            "gathered by the kernel from pinned host memory (zero copy)" if zc > 0.5 else "staged by async copies"))
        d2h = len(samples) * nu * (device.COUNTS_DTYPE.itemsize + device.STATS_DTYPE.itemsize)
        e2e = (dt_e2e, h2d, d2h, transfer)

    # ---- max over ranks ----
    if dist is not None:
        t = torch.tensor([ms_total, e2e[0] if e2e else 0.0, k_ms], device=dev.device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, e2e_dt, k_ms = [float(x) for x in t.tolist()]
        if e2e:
            e2e = (e2e_dt, e2e[1], e2e[2], e2e[3])
        tot = torch.tensor([units_per_step], device=dev.device, dtype=torch.float64)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        units_total_per_step = float(tot.item())
    else:
        units_total_per_step = float(units_per_step)

    if rank == 0:
        peak, peak_src = measured_peak()
        # B_unit = 16 (variant) + 8 (index bounds) + 32 * D_raw + 160 (outputs), biallelic
        b_unit = 16 + 8 + 32.0 * d_raw + 160
        achieved = nu * b_unit / (k_ms * 1e-3) / 1e9
        top_kernel = "pileup_column_kernel" if w["columns"] else "pileup_warp_kernel"
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            with open(tp) as fh:
                traffic = json.load(fh).get("%s_dram_bytes_per_launch" % top_kernel)
        cpu = None
        if not args.no_cpu_baseline:
            n_cpu, times, cores = run_cpu_arm(args.cpu_tiles, 0, repeats=1)
            cpu = {"value": n_cpu / times[0], "unit": "units/s", "cores": cores, "kind": "port",
                   "sample": "%d of %d targets (%d units, %.1f s): C pileup restatement (OpenMP) + reference-style scalar "
                             "scipy epilogue (process pool); pysam itself is not installable here" % (
                                 args.cpu_tiles, args.tiles, n_cpu, times[0])}
        line = {
            "metric": "variants x BAMs annotated per second", "value": units_total_per_step * args.steps / (ms_total * 1e-3),
            "unit": "units/s", "n_gpus": world, "steps": args.steps, "warmup": warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32/f64", "data": "synthetic",
            "config": {"workload": workload, "units_per_step_per_gpu": units_per_step, "reads_per_gpu": w["n_reads"],
                       "mean_depth_raw": round(d_raw, 2), "mean_candidates": round(n_cand, 2),
                       "l2": "inputs (%.1f GB of read batches per GPU) exceed L2; tumour and normal batches alternate" % (
                           sum(w["packed"][s].nbytes() for s in samples) / 1e9),
                       "layout": "read-major + column store" if w["columns"] else "read-major",
                       "sharding": "one independent interval shard per GPU, no collective",
                       "numa_node_rank0": numa},
            "roofline": {"bound": "hbm", "kernel": top_kernel, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_unit": b_unit, "units_per_launch": nu, "kernel_ms": k_ms},
            "cpu_baseline": cpu,
            "e2e": None if e2e is None else {"value": units_total_per_step / e2e[0], "unit": "units/s",
                                            "h2d_bytes_per_step": int(e2e[1]), "d2h_bytes_per_step": int(e2e[2]),
                                            "ms_per_step": 1e3 * e2e[0],
                                            "transfer": e2e[3]},
            "gpu_launches": int(launches), "clocks": clk,
        }
        print(json.dumps(line), file=_RESULT_OUT, flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
