#!/usr/bin/env python
"""A/B of the staged kernel variants in ONE process on ONE workload: for every combination asked
for, parity of every unit against the C oracle on a sample, then the CUDA-event time of the pileup call.

    python tools/ab_staged.py [--tiles 200000] [--launches 20] [--oracle-tiles 3000]

Needs a GPU.  The variants are environment switches read per call by libvfk.so / device.pack_reads:
VFK_COL_TRANSPOSED (store layout: the batch is re-packed when it changes), VFK_LIST_SHALLOW, VFK_MQ_SELECT; --e2e adds
the host entry point with and without VFK_HOST_GATHER."""
import argparse
import itertools
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from oracle import c_oracle  # noqa: E402  (checker only)
from vafator_b200 import device, synth  # noqa: E402
from vafator_b200.batch import build_units  # noqa: E402

FLAGS = ("VFK_COL_TRANSPOSED", "VFK_LIST_SHALLOW", "VFK_MQ_SELECT")
FIELDS = ("dp", "n_amb", "ac_ref", "ac_alt", "med2", "n_col")


def setup(n_tiles):
    cfg = synth.wes(n_tiles=n_tiles)
    batch = synth.reads(cfg)
    recs = synth.variant_records(cfg)
    units = build_units(recs, cfg.contigs)
    last = {}
    for c, p, _r, _a in recs:
        last[c] = p
    stop = np.asarray([last[recs[i][0]] for i in units.rec], np.int32)
    ounits = [(cfg.contigs.index(recs[i][0]), recs[i][1], recs[i][2], recs[i][3][j], recs[i][3][0], int(s))
              for i, j, s in zip(units.rec, units.alt_index, stop)]
    return batch, units, stop, ounits


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tiles", type=int, default=200_000)
    ap.add_argument("--oracle-tiles", type=int, default=3000)
    ap.add_argument("--launches", type=int, default=20)
    ap.add_argument("--e2e", action="store_true", help="also time the host entry point: in place vs VFK_HOST_GATHER, both store forms")
    ap.add_argument("--only", default="", help="comma separated flag subsets, e.g. 'VFK_MQ_SELECT,VFK_COL_TRANSPOSED+VFK_MQ_SELECT'")
    a = ap.parse_args()
    params = device.make_params(min_mapq=1, min_baseq=30)
    dev = device.Device(0)
    small = setup(a.oracle_tiles)
    want = c_oracle.pileup(small[0].as_dict(), small[3], c_oracle.params(min_mapq=1, min_baseq=30))
    big = setup(a.tiles)
    combos = [c for r in range(len(FLAGS) + 1) for c in itertools.combinations(FLAGS, r)]
    if a.only:
        combos = [()] + [tuple(x.split("+")) for x in a.only.split(",")]
    packed = {}
    for combo in combos:
        for f in FLAGS:
            os.environ[f] = "1" if f in combo else "0"
        layout = "VFK_COL_TRANSPOSED" in combo
        if layout not in packed:  # the store layout is the only thing that needs a re-pack
            packed.clear()
            packed[layout] = (dev.upload_reads(device.pack_reads(small[0], params, columns=True)),
                              dev.upload_reads(device.pack_reads(big[0], params, columns=True)))
        d_small, d_big = packed[layout]
        buf = dev.pileup(d_small, dev.upload_units(small[1], small[2]), params)
        dev.sync()
        got = dev.to_host(buf, device.COUNTS_DTYPE, small[1].n_units)
        both = (want["ac_ref"] > 0) & (want["ac_alt"] > 0)
        bad = [f for f in FIELDS if not np.array_equal(got[f], want[f])]
        if not np.array_equal(got["s2"][both], want["s2"][both]):
            bad.append("s2")
        dunits = dev.upload_units(big[1], big[2])
        out = dev.alloc_counts(big[1].n_units)
        for _ in range(3):
            dev.pileup(d_big, dunits, params, out=out)
        evs = []
        for _ in range(a.launches):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dev.pileup(d_big, dunits, params, out=out)
            e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        ms = sorted(x.elapsed_time(y) for x, y in evs)
        print(json.dumps({"variant": "+".join(combo) or "default", "parity": "ok" if not bad else "DIFFERS: " + ",".join(bad),
                          "units_checked": int(small[1].n_units), "pileup_ms_median": round(ms[len(ms) // 2], 4),
                          "pileup_ms_min": round(ms[0], 4), "units": int(big[1].n_units)}), flush=True)
    for f in FLAGS:
        os.environ.pop(f, None)
    if a.e2e:  # the host entry point: store read in place (sector / transposed form) vs gathered per unit and bulk-copied
        eaf = np.full(big[1].n_units, 0.5)
        for layout in ("0", "1"):
            os.environ["VFK_COL_TRANSPOSED"] = layout
            pinned = device.pack_reads(big[0], params, pinned=True, columns=True)
            ref = None
            for gather in ("0", "1"):
                os.environ["VFK_HOST_GATHER"] = gather
                for _ in range(2):
                    got = dev.annotate_host(pinned, big[1], big[2], params, eaf, 5e-7, 1e-3)
                ts = []
                for _ in range(a.launches):
                    t0 = time.perf_counter()
                    got = dev.annotate_host(pinned, big[1], big[2], params, eaf, 5e-7, 1e-3)
                    ts.append(time.perf_counter() - t0)
                ts.sort()
                same = "n/a"
                if ref is None:
                    ref = got
                else:
                    same = "ok" if all(np.array_equal(got[0][f], ref[0][f]) for f in FIELDS) and got[1].tobytes() == ref[1].tobytes() else "DIFFERS"
                print(json.dumps({"e2e": "annotate_host", "VFK_COL_TRANSPOSED": layout, "VFK_HOST_GATHER": gather,
                                  "equal_to_in_place": same, "ms_median": round(1e3 * ts[len(ts) // 2], 3), "ms_min": round(1e3 * ts[0], 3),
                                  "units": int(big[1].n_units)}), flush=True)
        for f in ("VFK_COL_TRANSPOSED", "VFK_HOST_GATHER"):
            os.environ.pop(f, None)


if __name__ == "__main__":
    t = time.time()
    main()
    print("total %.0f s" % (time.time() - t), file=sys.stderr)
