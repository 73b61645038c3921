"""Shared test helpers: golden fixtures -> batches, expected integer records."""
import glob
import json
import os
from collections import OrderedDict

import numpy as np

from vafator_b200.batch import ReadBatch, variant_kind, KIND_SNV

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
AMBIGUOUS = ("N", "M", "R", "W", "S", "Y", "K", "V", "H", "D", "B")


def golden_scenarios():
    return sorted(glob.glob(os.path.join(GOLDEN, "pileup_*.json")))


def load_scenario(path):
    with open(path) as fh:
        return json.load(fh, object_pairs_hook=OrderedDict)


def scenario_batches(sc):
    return [ReadBatch.from_records(b, sc["contigs"]) for b in sc["bams"]]


def scenario_units(sc):
    """[(variant index, alt index, tid, pos1, ref, alt_j, alt0)] for every unit the device evaluates."""
    out = []
    for i, v in enumerate(sc["variants"]):
        kind = variant_kind(v["ref"], v["alts"][0])
        if kind is None:
            continue
        n_alt = len(v["alts"]) if kind == KIND_SNV else 1
        for j in range(n_alt):
            out.append((i, j, sc["contigs"].index(v["chrom"]), v["pos"], v["ref"], v["alts"][j], v["alts"][0]))
    return out


def midrank_s2(alt, ref):
    """Twice the sum of the ALT mid-ranks in the pooled sample (brute force)."""
    s2 = 0
    for a in alt:
        below = sum(1 for x in alt if x < a) + sum(1 for x in ref if x < a)
        ties = sum(1 for x in alt if x == a) + sum(1 for x in ref if x == a)
        s2 += 2 * below + ties + 1
    return s2


def med2(xs):
    if not xs:
        return -1
    s = sorted(xs)
    return s[(len(s) - 1) // 2] + s[len(s) // 2]


def expected_record(metrics, ref, alt_j, alt0, include_ambiguous):
    """Integer record implied by one golden CoverageMetrics JSON (or None = EMPTY_METRICS)."""
    kind = variant_kind(ref, alt0)
    rec = dict(dp=0, n_amb=0, ac_ref=0, ac_alt=0, med2=[[-1, -1]] * 3, s2=[0, 0, 0], n_val=[[0, 0]] * 3)
    if metrics is None:
        return rec
    if kind == KIND_SNV:
        rkey, akey = ref, alt_j
    else:
        rkey, akey = ref, alt0.upper()
    lists = [metrics["all_mqs"], metrics["all_bqs"], metrics["all_positions"]]
    rec["n_amb"] = sum(metrics["ac"].get(b, 0) for b in AMBIGUOUS) if kind == KIND_SNV else 0
    rec["dp"] = metrics["dp"] + (0 if include_ambiguous or kind != KIND_SNV else rec["n_amb"])
    rec["ac_ref"] = len(metrics["all_mqs"].get(rkey, []))
    rec["ac_alt"] = len(metrics["all_mqs"].get(akey, []))
    rec["med2"] = [[med2(l.get(rkey, [])), med2(l.get(akey, []))] for l in lists]
    rec["n_val"] = [[len(l.get(rkey, [])), len(l.get(akey, []))] for l in lists]
    rec["s2"] = [midrank_s2(l.get(akey, []), l.get(rkey, [])) for l in lists]
    return rec
