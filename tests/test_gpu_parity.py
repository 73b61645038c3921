"""GPU parity proper: the CUDA path, called through the C ABI, against the golden fixtures
(unmodified reference code over the pileup restatement) and against the oracle's statistics.

Bars: counts / depths / medians / rank sums (integers): bit-exact.
      z, p, pu, pw: relative 1e-6 (BASELINE.json north_star), k exact."""
import json
import math
import os

import numpy as np
import pytest

import helpers as H
from oracle import vafator_ref as vr
from vafator_b200 import device
from vafator_b200.batch import build_units, variant_kind, KIND_SNV

pytestmark = pytest.mark.gpu
SCENARIOS = H.golden_scenarios()
REL = 1e-6


@pytest.fixture(scope="module")
def dev():
    d = device.Device(0)
    yield d
    d.close()


def _close(got, want, rel=REL, abs_floor=1e-300):
    if want is None or (isinstance(want, float) and math.isnan(want)):
        return math.isnan(got)
    return abs(got - want) <= rel * abs(want) + abs_floor


def _units_and_stop(sc):
    recs = [(v["chrom"], v["pos"], v["ref"], v["alts"]) for v in sc["variants"]]
    units = build_units(recs, sc["contigs"])
    last = {}
    for c, p, r, a in recs:
        last[c] = max(last.get(c, 0), p)
    stop = np.asarray([last[recs[i][0]] for i in units.rec], np.int32)
    return recs, units, stop


def _check_counts(sc, run, b, recs, units, counts):
    prm = run["params"]
    for u in range(units.n_units):
        i, j = int(units.rec[u]), int(units.alt_index[u])
        chrom, pos1, ref, alts = recs[i]
        key = "%s:%d:%s:%s" % (chrom, pos1, ref, alts[0])
        want = H.expected_record(run["metrics"][b].get(key), ref, alts[j], alts[0], prm["include_ambiguous"])
        got = counts[u]
        ctx = (sc["name"], prm, b, recs[i], j)
        want_dp = want["dp"] - (0 if prm["include_ambiguous"] or variant_kind(ref, alts[0]) != KIND_SNV else want["n_amb"])
        assert got["status"] & (device.ST_DEFERRED | device.ST_READLEN) == 0, ctx
        assert got["dp"] == want_dp, ctx
        assert got["n_amb"] == want["n_amb"], ctx
        assert got["ac_ref"] == want["ac_ref"], ctx
        assert got["ac_alt"] == want["ac_alt"], ctx
        assert got["med2"].tolist() == want["med2"], ctx
        if want["ac_ref"] and want["ac_alt"]:
            s2 = [int(x) for x in got["s2"]]
            if variant_kind(ref, alts[0]) != KIND_SNV:
                s2[1] = want["s2"][1] = 0
            assert s2 == want["s2"], ctx


@pytest.mark.parametrize("path", SCENARIOS, ids=[os.path.basename(p) for p in SCENARIOS])
def test_pileup_kernel_matches_golden(dev, path):
    sc = H.load_scenario(path)
    recs, units, stop = _units_and_stop(sc)
    dunits = dev.upload_units(units, stop)
    for run in sc["runs"]:
        prm = run["params"]
        params = device.make_params(min_mapq=prm["min_mq"], min_baseq=prm["min_bq"],
                                    include_ambiguous=prm["include_ambiguous"])
        for b, batch in enumerate(H.scenario_batches(sc)):
            dreads = dev.upload_reads(batch)
            buf = dev.pileup(dreads, dunits, params)
            dev.sync()
            counts = dev.to_host(buf, device.COUNTS_DTYPE, units.n_units)
            _check_counts(sc, run, b, recs, units, counts)


@pytest.mark.parametrize("path", SCENARIOS[:3], ids=[os.path.basename(p) for p in SCENARIOS[:3]])
def test_host_entry_point_and_epilogue(dev, path):
    """vfk_annotate_host (host buffers in / out) + the FP64 epilogue against scipy as the reference calls it."""
    sc = H.load_scenario(path)
    recs, units, stop = _units_and_stop(sc)
    run = sc["runs"][0]
    prm = run["params"]
    params = device.make_params(min_mapq=prm["min_mq"], min_baseq=prm["min_bq"], include_ambiguous=prm["include_ambiguous"])
    power = vr.Power(purity={"s": 0.8}, ploidy={"s": 3.0})
    eaf = np.full(units.n_units, power.eaf("s"))
    for b, batch in enumerate(H.scenario_batches(sc)):
        packed = device.host_reads(batch, pinned=True)
        counts, stats = dev.annotate_host(packed, units, stop, params, eaf, power.fpr, power.error_rate)
        _check_counts(sc, run, b, recs, units, counts)
        for u in range(units.n_units):
            i, j = int(units.rec[u]), int(units.alt_index[u])
            chrom, pos1, ref, alts = recs[i]
            m = run["metrics"][b].get("%s:%d:%s:%s" % (chrom, pos1, ref, alts[0]))
            c, s = counts[u], stats[u]
            ctx = (sc["name"], recs[i], j)
            assert _close(s["pu"], float(power.pu_raw(int(c["dp"]), int(c["ac_alt"]), eaf[u]))), ctx
            pw, k = power.pw_raw(int(c["dp"]), eaf[u])
            assert s["k"] == k, ctx
            assert _close(s["pw"], float(pw)), ctx
            if m is None:
                continue
            kind = variant_kind(ref, alts[0])
            akey = alts[j] if kind == KIND_SNV else alts[0].upper()
            for mi, name in enumerate(("all_mqs", "all_bqs", "all_positions")):
                z, p = vr.rank_sum_raw(m[name].get(akey, []), m[name].get(ref, []))
                assert _close(s["z"][mi], z, abs_floor=1e-12), (ctx, name, s["z"][mi], z)
                assert _close(s["p"][mi], p), (ctx, name, s["p"][mi], p)


def test_epilogue_against_reference_golden(dev):
    """tests/golden/epilogue.json: raw z/p from scipy.stats.ranksums and pu/pw/k from the reference's
    PowerCalculator (vafator/power.py), generated by oracle/gen_golden.py."""
    import torch
    with open(os.path.join(H.GOLDEN, "epilogue.json")) as fh:
        g = json.load(fh)
    rows = g["ranksum"]
    counts = np.zeros(len(rows), device.COUNTS_DTYPE)
    for i, r in enumerate(rows):
        counts["ac_alt"][i], counts["ac_ref"][i] = len(r["alt"]), len(r["ref"])
        counts["dp"][i] = len(r["alt"]) + len(r["ref"])
        s2 = H.midrank_s2(r["alt"], r["ref"])
        counts["s2"][i] = (s2, s2, s2)
    eaf = np.full(len(rows), 0.5)
    buf = dev.epilogue(len(rows), torch.from_numpy(counts.view(np.uint8)).to(dev.device),
                       torch.from_numpy(eaf).to(dev.device), 5e-7, 1e-3)
    dev.sync()
    st = dev.to_host(buf, device.STATS_DTYPE, len(rows))
    for i, r in enumerate(rows):
        for m in range(3):
            assert _close(st["z"][i][m], r["z"], abs_floor=1e-12), (i, st["z"][i][m], r["z"])
            assert _close(st["p"][i][m], r["p"]), (i, st["p"][i][m], r["p"])
        assert round(float(st["z"][i][0]), 3) == r["z_round"] and round(float(st["p"][i][0]), 5) == r["p_round"]
    # power: one launch per distinct (fpr, error_rate)
    prow = g["power"]
    groups = {}
    for i, r in enumerate(prow):
        groups.setdefault((r["fpr"], r["error_rate"]), []).append(i)
    for (fpr, err), idx in groups.items():
        counts = np.zeros(len(idx), device.COUNTS_DTYPE)
        eaf = np.zeros(len(idx))
        for n, i in enumerate(idx):
            counts["dp"][n], counts["ac_alt"][n], eaf[n] = prow[i]["dp"], prow[i]["ac"], prow[i]["eaf"]
        buf = dev.epilogue(len(idx), torch.from_numpy(counts.view(np.uint8)).to(dev.device),
                           torch.from_numpy(eaf).to(dev.device), fpr, err)
        dev.sync()
        st = dev.to_host(buf, device.STATS_DTYPE, len(idx))
        for n, i in enumerate(idx):
            r = prow[i]
            assert st["k"][n] == r["k"], r
            assert _close(st["pu"][n], r["pu_raw"]), (r, st["pu"][n])
            if r["pw_raw"] is not None:
                assert _close(st["pw"][n], r["pw_raw"], abs_floor=1e-15), (r, st["pw"][n])
            assert round(float(st["pu"][n]), 5) == r["pu"] and round(float(st["pw"][n]), 5) == r["pw"], (r, st[n])
