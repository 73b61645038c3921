#!/usr/bin/env python
"""Test infrastructure (not collected by pytest): randomised GPU-vs-oracle parity campaign: many seeds x read / variant mixes x filter thresholds.
    python tests/fuzz_parity.py [--seconds 300] [--seed0 1]
Every integer field of every unit must match the C oracle bit for bit; exits non-zero on the first
mismatch and prints the configuration that produced it."""
import argparse
import os
import random
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import c_oracle  # noqa: E402  (checker only)
from vafator_b200 import device, synth  # noqa: E402
from vafator_b200.batch import build_units  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=300)
ap.add_argument("--seed0", type=int, default=1)
ap.add_argument("--host", action="store_true", help="through vfk_annotate_host (host arrays in, results out) instead of vfk_pileup")
a = ap.parse_args()
dev = device.Device(0)
t_end = time.time() + a.seconds
n_cases = n_units = 0
seed = a.seed0
while time.time() < t_end:
    rng = random.Random(seed)
    kind = rng.choice(["wes", "wgs", "amplicon", "noisy", "short"])
    common = dict(seed=seed * 7919, multiallelic_pct=rng.choice([0, 5, 30]), p_softclip=rng.choice([0.03, 0.3]),
                  p_del=rng.choice([0.025, 0.2]), p_ins=rng.choice([0.02, 0.2]), p_skip=rng.choice([0.005, 0.08]),
                  p_dup=rng.choice([0.02, 0.2]), p_supp=rng.choice([0.005, 0.1]), p_orphan=rng.choice([0.01, 0.15]),
                  p_mapq60=rng.choice([0.85, 0.4]), p_n=rng.choice([0.001, 0.05]), base_error=rng.choice([0.002, 0.05]))
    if kind == "wes":
        cfg = synth.wes(n_tiles=rng.choice([1100, 3000]), indel_period=rng.choice([0, 300]), snv_period=rng.choice([40, 500]), **common)
    elif kind == "wgs":
        cfg = synth.wgs(contig_len=rng.choice([300_000, 900_000]), n_contigs=rng.choice([1, 3]), snv_period=rng.choice([100, 800]),
                        indel_period=rng.choice([400, 7200]), **{k: v for k, v in common.items()})
    elif kind == "amplicon":
        cfg = synth.amplicon(n_tiles=rng.choice([4, 12]), depth_pairs=rng.choice([200, 900, 2500]), **common)
    elif kind == "noisy":
        cfg = synth.wes(n_tiles=1200, indel_period=30, snv_period=20, frag_mean=rng.choice([160.0, 220.0]), frag_sd=40.0,
                        pairs_per_tile=rng.choice([10, 36, 80]), **common)
    else:
        rl = rng.choice([36, 75, 250])
        cfg = synth.wes(n_tiles=1100, read_len=rl, frag_mean=1.5 * rl, frag_sd=0.3 * rl, tile_len=max(2 * rl, 64), var_lo=rl // 2,
                        var_hi=3 * rl, snv_period=25, indel_period=200, **common)
    mq, bq, amb = rng.choice([0, 1, 20, 41]), rng.choice([0, 13, 30, 41]), rng.choice([True, False])
    batch = synth.reads(cfg)
    recs = synth.variant_records(cfg)
    units = build_units(recs, cfg.contigs)
    last = {}
    for c, p, r, al in recs:
        last[c] = p
    stop = np.asarray([last[recs[i][0]] for i in units.rec], np.int32)
    tid_of = {c: i for i, c in enumerate(cfg.contigs)}
    ou = [(tid_of[recs[i][0]], recs[i][1], recs[i][2], recs[i][3][j], recs[i][3][0], last[recs[i][0]])
          for i, j in zip(units.rec, units.alt_index)]
    want = c_oracle.pileup(batch.as_dict(), ou, c_oracle.params(min_mapq=mq, min_baseq=bq))
    params = device.make_params(min_mapq=mq, min_baseq=bq, include_ambiguous=amb)
    if a.host:
        got = dev.annotate_host(device.host_reads(batch, pinned=bool(seed & 1)), units, stop, params, np.full(units.n_units, 0.5), 5e-7, 1e-3)[0]
    else:
        buf = dev.pileup(dev.upload_reads(batch), dev.upload_units(units, stop), params)
        dev.sync()
        got = dev.to_host(buf, device.COUNTS_DTYPE, units.n_units)
    snv = (got["status"] & 3) == 0
    dp = want["dp"] - (0 if amb else np.where(snv, want["n_amb"], 0))
    both = (want["ac_ref"] > 0) & (want["ac_alt"] > 0)
    bad = (got["dp"] != dp) | (got["n_amb"] != want["n_amb"]) | (got["ac_ref"] != want["ac_ref"]) | (got["ac_alt"] != want["ac_alt"]) | \
          (got["n_col"] != want["n_col"]) | np.any(got["med2"] != want["med2"], axis=(1, 2)) | \
          (both & np.any(got["s2"][:, [0, 2]] != want["s2"][:, [0, 2]], axis=1)) | (both & snv & (got["s2"][:, 1] != want["s2"][:, 1])) | \
          ((got["status"] & 0x30) != 0)
    if bad.any():
        u = int(np.nonzero(bad)[0][0])
        print("MISMATCH seed", seed, kind, cfg, "mq/bq/amb", mq, bq, amb, "unit", u, ou[u], "\n got", got[u], "\nwant", want[u])
        sys.exit(1)
    n_cases += 1
    n_units += units.n_units
    seed += 1
print("fuzz ok: %d cases, %d units, seeds %d..%d" % (n_cases, n_units, a.seed0, seed - 1))
