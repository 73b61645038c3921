"""Pins the oracle: (1) the Python restatement of the vafator layers against the golden
fixtures produced by the unmodified reference code (and against the reference itself when
/root/reference is present); (2) the C restatement against the same fixtures."""
import math
import os
import sys
from collections import OrderedDict

import numpy as np
import pytest

import helpers as H
from oracle import c_oracle, htslib_pileup as hp, vafator_ref as vr

SCENARIOS = H.golden_scenarios()


class V:
    def __init__(self, d):
        self.CHROM, self.POS, self.REF, self.ALT = d["chrom"], d["pos"], d["ref"], d["alts"]


def _files(sc):
    return [hp.AlignmentFile([hp.Rec(r["qname"], r["flag"], r["tid"], r["pos"], r["mapq"], r["cigar"], r["seq"],
                                     bytes(ord(c) - 33 for c in r["qual"]), r["mtid"], r["mpos"], r["isize"], index=i)
                              for i, r in enumerate(b)], sc["contigs"]) for b in sc["bams"]]


def _power(sc, prm):
    pl = {s: ([tuple(r) for r in p] if isinstance(p, list) else p) for s, p in sc["ploidies"].items()}
    return vr.Power(ploidy=pl, purity=dict(sc["purities"]), normal_ploidy=prm.get("normal_ploidy", 2),
                    fpr=prm.get("fpr", 5e-7), error_rate=prm.get("error_rate", 1e-3))


@pytest.mark.parametrize("path", SCENARIOS, ids=[os.path.basename(p) for p in SCENARIOS])
def test_python_restatement_matches_reference_output(path):
    sc = H.load_scenario(path)
    files = _files(sc)
    for run in sc["runs"]:
        prm = run["params"]
        power = _power(sc, prm)
        by_chrom = OrderedDict()
        for i, v in enumerate(sc["variants"]):
            by_chrom.setdefault(v["chrom"], []).append((i, V(v)))
        per_bam = []
        for f in files:
            res = {}
            for chrom, vs in by_chrom.items():
                res[chrom] = vr.collect_metrics_for_chrom(chrom, [v for _, v in vs], f, prm["min_bq"], prm["min_mq"],
                                                          prm["include_ambiguous"])
            per_bam.append(res)
        for chrom, vs in by_chrom.items():
            for i, v in vs:
                samples = OrderedDict()
                for s, idx in sc["samples"].items():
                    samples[s] = [per_bam[b][chrom].get((v.POS, v.REF, v.ALT[0])) for b in idx]
                got = vr.info_fields(v.CHROM, v.POS, v.REF, v.ALT, samples, power)
                want = OrderedDict((k, val) for k, val in run["info"][i])
                assert list(got.keys()) == list(want.keys()), (sc["name"], i)
                for k in want:
                    assert str(got[k]) == str(want[k]), (sc["name"], sc["variants"][i], k, got[k], want[k])


@pytest.mark.parametrize("path", SCENARIOS, ids=[os.path.basename(p) for p in SCENARIOS])
def test_c_restatement_matches_golden(path):
    sc = H.load_scenario(path)
    batches = H.scenario_batches(sc)
    units = H.scenario_units(sc)
    for run in sc["runs"]:
        prm = run["params"]
        cp = c_oracle.params(min_mapq=prm["min_mq"], min_baseq=prm["min_bq"])
        for b, batch in enumerate(batches):
            got = c_oracle.pileup(batch.as_dict(), [(u[2], u[3], u[4], u[5], u[6]) for u in units], cp)
            for rec, (vi, aj, tid, pos1, ref, alt, alt0) in zip(got, units):
                v = sc["variants"][vi]
                key = "%s:%d:%s:%s" % (v["chrom"], pos1, ref, alt0)
                want = H.expected_record(run["metrics"][b].get(key), ref, alt, alt0, prm["include_ambiguous"])
                ctx = (sc["name"], prm, b, v, aj)
                assert rec["supported"] == 1
                assert rec["dp"] == want["dp"], ctx
                assert rec["n_amb"] == want["n_amb"], ctx
                assert rec["ac_ref"] == want["ac_ref"], ctx
                assert rec["ac_alt"] == want["ac_alt"], ctx
                assert rec["med2"].tolist() == want["med2"], ctx
                assert rec["n_val"].tolist() == want["n_val"], ctx
                assert [int(x) for x in rec["s2"]] == want["s2"], ctx


def test_overlap_tweak_by_inspection():
    """SURVEY.md A.2 consequences: two agreeing Q20 mates -> one Q40 + one Q0; two disagreeing
    Q37 mates -> 29 and 0."""
    def pair(seq_b, qa, qb):
        a = hp.Rec("p", 99, 0, 100, 60, "10M", "ACGTACGTAC", [qa] * 10, 0, 105, 15)
        b = hp.Rec("p", 147, 0, 105, 60, "10M", seq_b, [qb] * 10, 0, 100, -15)
        f = hp.AlignmentFile([a, b], ["c"])
        cols = {c.reference_pos: c for c in f.pileup("c", 100, 115, min_base_quality=0, max_depth=10 ** 6)}
        col = cols[107]
        return sorted(pr.alignment.query_qualities[pr.query_position] for pr in col.pileups)
    assert pair("CGTACGGGGG", 20, 20) == [0, 40]
    assert pair("CGAACGGGGG", 37, 37) == [0, 29]   # position 107: T vs A
    assert pair("CGAACGGGGG", 37, 30) == [0, 29]
    assert pair("CGTACGGGGG", 150, 150) == [0, 200]


def test_hashes():
    assert hp.x31_hash("a") == 97
    assert hp.x31_hash("ab") == 97 * 31 + 98
    # Wang hash is a bijection on uint32; spot values are self-consistent with the C restatement
    assert hp.wang_hash(0) == hp.wang_hash(0)
    assert len({hp.wang_hash(i) for i in range(1000)}) == 1000


@pytest.mark.skipif(not os.path.isdir("/root/reference/vafator"), reason="reference checkout not present")
def test_reference_known_answers_run_here():
    """The reference's own statistics known answers, through the reference modules AND the restatement."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    from oracle.gen_golden import import_reference
    rp, ra, rw, rl, rr = import_reference()
    alt = [0, 7, 10, 17, 20, 21, 30, 34, 40, 45]
    ref = [20, 25, 26, 30, 32, 40, 47, 50, 53, 60]
    assert rr.calculate_rank_sum_test(alt, ref) == (-2.154, 0.03121)  # vafator/tests/test_rank_sum_test.py:27-35
    assert vr.rank_sum(alt, ref) == (-2.154, 0.03121)
    pc = rw.PowerCalculator(tumor_ploidies={"tumor": rl.PloidyManager(genome_wide_ploidy=2.5)}, purities={"tumor": 0.8})
    mine = vr.Power(ploidy={"tumor": 2.5}, purity={"tumor": 0.8})
    assert pc.calculate_power(dp=10, ac=0, sample="tumor", variant=None) == 0.01734  # test_power_calculator.py:16-18
    for dp in (0, 1, 2, 10, 50, 100, 1000):
        for ac in (0, 1, dp // 2, dp):
            assert pc.calculate_power(dp=dp, ac=ac, sample="tumor", variant=None) == mine.pu(dp, ac, mine.eaf("tumor"))
        assert tuple(pc.calculate_absolute_power("tumor", None, dp)) == tuple(mine.pw(dp, mine.eaf("tumor")))


def test_epilogue_golden_restatement():
    import json
    with open(os.path.join(H.GOLDEN, "epilogue.json")) as fh:
        g = json.load(fh)
    for row in g["ranksum"]:
        z, p = vr.rank_sum_raw(row["alt"], row["ref"])
        assert z == row["z"] and p == row["p"]
        assert vr.rank_sum(row["alt"], row["ref"]) == (row["z_round"], row["p_round"])
    for row in g["power"][:400]:
        pw = vr.Power(ploidy={"t": row["ploidy"]}, purity={"t": row["purity"]}, normal_ploidy=row["normal_ploidy"],
                      fpr=row["fpr"], error_rate=row["error_rate"])
        eaf = pw.eaf("t")
        assert eaf == row["eaf"]
        assert pw.pu(row["dp"], row["ac"], eaf) == row["pu"]
        if row["dp"] > 0:
            assert pw.pw(row["dp"], eaf) == (row["pw"], row["k"])
    # notebooks/power.tsv (reference's published table): ploidy column is the tumour ploidy, purity as given
    for row in g["power_tsv"][::8]:
        pw = vr.Power(ploidy={"t": row["ploidy"]}, purity={"t": row["purity"]}, error_rate=row["error_rate"])
        got, k = pw.pw(int(row["dp"]), pw.eaf("t"))
        assert k == int(row["k"]) and abs(got - row["power"]) < 1e-12, row


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_two_formulations_agree_on_random_reads(seed):
    """Beyond the golden scenarios: the streaming restatement (htslib_pileup.py push / next state machine under the
    restated vafator column functions) and the direct per-pair C restatement on seeded synthetic reads with heavy
    CIGAR noise, overlapping mates, duplicates / orphans / supplementary records -- every integer of every unit."""
    from collections import namedtuple
    from vafator_b200 import synth
    rng = np.random.default_rng(seed)
    cfg = synth.wes(n_tiles=10, seed=synth.BASE_SEED + 70 + seed, pairs_per_tile=int(rng.choice([8, 30])), snv_period=25, indel_period=60,
                    multiallelic_pct=15, p_softclip=0.15, p_del=0.12, p_ins=0.12, p_skip=0.04, p_supp=0.05, p_orphan=0.05,
                    p_n=0.02, frag_mean=float(rng.choice([170.0, 240.0])), frag_sd=40.0, read_len=int(rng.choice([75, 150])))
    batch = synth.reads(cfg)
    recs = synth.variant_records(cfg)
    soa = batch.as_dict()
    f = hp.AlignmentFile(hp.recs_from_soa(soa), batch.contigs)
    V = namedtuple("V", "CHROM POS REF ALT")
    mq, bq, amb = int(rng.choice([0, 1, 20])), int(rng.choice([0, 13, 30])), bool(rng.integers(0, 2))
    n_checked = 0
    for chrom in batch.contigs:
        mine = [r for r in recs if r[0] == chrom]
        if not mine:
            continue
        stream = vr.collect_metrics_for_chrom(chrom, [V(*r) for r in mine], f, bq, mq, amb)
        units = [(batch.contigs.index(chrom), r[1], r[2], a, r[3][0], mine[-1][1]) for r in mine for a in r[3]
                 if vr.variant_kind(r[2], r[3][0]) is not None and (vr.variant_kind(r[2], r[3][0]) == "snv" or a == r[3][0])]
        direct = c_oracle.pileup(soa, units, c_oracle.params(min_mapq=mq, min_baseq=bq))
        for rec, (tid, pos1, ref, alt, alt0, _stop) in zip(direct, units):
            m = stream.get((pos1, ref, alt0))
            as_json = None if m is None else dict(ac=dict(m.ac), dp=m.dp, all_bqs=m.dist["bq"], all_mqs=m.dist["mq"],
                                                  all_positions=m.dist["pos"])
            want = H.expected_record(as_json, ref, alt, alt0, amb)
            ctx = (seed, chrom, pos1, ref, alt)
            assert rec["dp"] == want["dp"] and rec["n_amb"] == want["n_amb"], ctx
            assert rec["ac_ref"] == want["ac_ref"] and rec["ac_alt"] == want["ac_alt"], ctx
            assert rec["med2"].tolist() == want["med2"] and rec["n_val"].tolist() == want["n_val"], ctx
            assert [int(x) for x in rec["s2"]] == want["s2"], ctx
            n_checked += 1
    assert n_checked > 100
