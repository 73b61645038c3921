"""End-to-end parity on the device: INFO strings of the B200 path == INFO strings produced by the
unmodified reference code (tests/golden), through Engine.annotate and through the `vafator` CLI on
real files (BAM written by vafator_b200.bamio, decoded by the in-tree BGZF/BAM decoder)."""
import json
import math
import os
from collections import OrderedDict

import numpy as np
import pytest

import helpers as H
from oracle import vafator_ref as vr
from vafator_b200 import bamio, engine
from vafator_b200.command_line import annotator as cli
from vafator_b200.pileup_utils import VariantRecord
from vafator_b200.ploidies import PloidyManager
from vafator_b200.power import PowerCalculator

pytestmark = pytest.mark.gpu
SCENARIOS = H.golden_scenarios()


def _eaf(sc, prm, records):
    pl = {s: ([tuple(r) for r in p] if isinstance(p, list) else p) for s, p in sc["ploidies"].items()}
    power = vr.Power(ploidy=pl, purity=dict(sc["purities"]), normal_ploidy=prm.get("normal_ploidy", 2))
    return {s: np.asarray([power.eaf(s, r[0], r[1]) for r in records]) for s in sc["samples"]}


@pytest.mark.parametrize("path", SCENARIOS, ids=[os.path.basename(p) for p in SCENARIOS])
def test_engine_info_strings_match_reference(path):
    sc = H.load_scenario(path)
    records = [(v["chrom"], v["pos"], v["ref"], v["alts"]) for v in sc["variants"]]
    batches = H.scenario_batches(sc)
    for run in sc["runs"]:
        prm = run["params"]
        eng = engine.Engine(0, min_mapq=prm["min_mq"], min_baseq=prm["min_bq"], include_ambiguous=prm["include_ambiguous"],
                            fpr=prm.get("fpr", 5e-7), error_rate=prm.get("error_rate", 1e-3))
        bams = OrderedDict((s, [batches[b] for b in idx]) for s, idx in sc["samples"].items())
        got = eng.annotate(records, bams, _eaf(sc, prm, records))
        eng.close()
        for i, (g, want) in enumerate(zip(got, run["info"])):
            w = OrderedDict((k, v) for k, v in want)
            assert list(g.keys()) == list(w.keys()), (sc["name"], prm, records[i])
            for k in w:
                assert str(g[k]) == str(w[k]), (sc["name"], prm, records[i], k, g[k], w[k])


@pytest.mark.parametrize("indexed", [False, True], ids=["whole_file", "bai_regions"])
@pytest.mark.parametrize("name", ["pileup_random_13.json", "pileup_handmade.json"])
def test_cli_on_files_matches_reference(tmp_path, name, indexed):
    """vafator --input-vcf ... --bam ... on BAM/VCF/BED files: every INFO key=value of the reference; with a
    .bai next to the BAMs only the regions the reference would fetch are decoded -- same output."""
    sc = H.load_scenario(os.path.join(H.GOLDEN, name))
    run = sc["runs"][3]  # thresholds 30/13, fpr/error-rate/normal-ploidy set, ambiguous bases excluded
    prm = run["params"]
    vcf = tmp_path / "in.vcf"
    with open(vcf, "w") as fh:
        fh.write("##fileformat=VCFv4.2\n" + "".join("##contig=<ID=%s>\n" % c for c in sc["contigs"]))
        fh.write("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n")
        for k, v in enumerate(sc["variants"]):
            fh.write("%s\t%d\t.\t%s\t%s\t.\tPASS\t%s\n" % (v["chrom"], v["pos"], v["ref"], ",".join(v["alts"]),
                                                              "." if k % 2 else "OLD=1;DP=7"))
    argv = ["--input-vcf", str(vcf), "--output-vcf", str(tmp_path / "out.vcf"), "--mapping-quality", str(prm["min_mq"]),
            "--base-call-quality", str(prm["min_bq"]), "--fpr", str(prm["fpr"]), "--error-rate", str(prm["error_rate"]),
            "--normal-ploidy", str(prm["normal_ploidy"])]
    if not prm["include_ambiguous"]:
        argv.append("--exclude-ambiguous-bases")
    for s, idx in sc["samples"].items():
        for b in idx:
            p = tmp_path / ("bam%d.bam" % b)
            bamio.write_bam(str(p), sc["bams"][b], sc["contigs"], index=indexed)
            assert os.path.exists(str(p) + ".bai") == indexed
            argv += ["--bam", s, str(p)]
    for s, pur in sc["purities"].items():
        argv += ["--purity", s, str(pur)]
    for s, pl in sc["ploidies"].items():
        if isinstance(pl, list):
            bed = tmp_path / ("%s.bed" % s)
            with open(bed, "w") as fh:
                for row in pl:
                    fh.write("\t".join(map(str, row)) + "\n")
            argv += ["--tumor-ploidy", s, str(bed)]
        else:
            argv += ["--tumor-ploidy", s, str(pl)]
    cli(argv)
    lines = [l.rstrip("\n") for l in open(tmp_path / "out.vcf")]
    header = [l for l in lines if l.startswith("##")]
    assert any(l.startswith("##vafator_command_line=") for l in header)
    cmd = json.loads([l for l in header if l.startswith("##vafator_command_line=")][0].split("=", 1)[1])
    assert cmd["name"] == "vafator" and cmd["mapping_quality_threshold"] == prm["min_mq"]
    first = next(iter(sc["samples"]))
    assert any(l.startswith("##INFO=<ID=%s_af," % first) for l in header)
    body = [l for l in lines if not l.startswith("#")]
    assert len(body) == len(sc["variants"])
    for k, (line, want) in enumerate(zip(body, run["info"])):
        info = line.split("\t")[7].split(";")
        if k % 2 == 0:
            assert info[:2] == ["OLD=1", "DP=7"]  # pre-existing INFO is kept, annotations are appended
            info = info[2:]
        assert info == ["%s=%s" % (key, val) for key, val in want], (k, sc["variants"][k])


def test_power_and_rank_sum_api_known_answers():
    """The reference's own unit-test values through the drop-in classes (device arithmetic)."""
    pc = PowerCalculator(tumor_ploidies={"tumor": PloidyManager(genome_wide_ploidy=2.5)}, purities={"tumor": 0.8})
    assert pc.calculate_power(dp=10, ac=0, sample="tumor", variant=None) == 0.01734   # test_power_calculator.py:16-18
    assert round(pc.calculate_power(dp=10, ac=3, sample="tumor", variant=None), 5) == 0.55926
    assert pc.calculate_power(dp=0, ac=0, sample="tumor", variant=None) == 1.0        # :71-78
    oracle = vr.Power(ploidy={"tumor": 2.5}, purity={"tumor": 0.8})
    for dp in (100, 50, 10, 2, 0):
        assert pc.calculate_absolute_power("tumor", None, dp) == oracle.pw(dp, oracle.eaf("tumor"))
    from vafator_b200.rank_sum_test import calculate_rank_sum_test, get_rank_sum_tests
    alt = [0, 7, 10, 17, 20, 21, 30, 34, 40, 45]
    ref = [20, 25, 26, 30, 32, 40, 47, 50, 53, 60]
    assert calculate_rank_sum_test(alt, ref) == (-2.154, 0.03121)                     # test_rank_sum_test.py:27-35
    d = list(range(1, 10))
    assert calculate_rank_sum_test(d, d) == (0.0, 1.0)
    s1, _ = calculate_rank_sum_test([1, 2, 3, 4, 5, 6], [4, 5, 6, 7, 8, 9])
    s2, _ = calculate_rank_sum_test([4, 5, 6, 7, 8, 9], [1, 2, 3, 4, 5, 6])
    assert s1 < 0 < s2 and s1 == -s2
    assert all(math.isnan(x) for x in calculate_rank_sum_test([], [1, 2, 3]))
    v = VariantRecord(CHROM="chr1", POS=100, REF="A", ALT=["T"])
    pv, st = get_rank_sum_tests({"A": [20, 25, 30, 35, 40], "T": [1, 5, 10, 15, 20]}, v)
    assert len(st) == 1 and float(st[0]) < 0
    assert get_rank_sum_tests({"A": [20, 25, 30]}, v) == ([], [])


@pytest.mark.parametrize("name,run_index", [("pileup_random_11.json", 0), ("pileup_random_14.json", 2), ("pileup_handmade.json", 1)])
def test_collect_metrics_for_chrom_call_site(tmp_path, name, run_index):
    """The reference's first call site, same signature and same return type, on a BAM file: every field of every
    CoverageMetrics -- allele counts of every observed base, depth, medians, and the full per-read lists in read order --
    equals what the unmodified reference function returned (goldens).  (The reference also keeps a "" entry for reads
    inside a deletion, which nothing downstream reads: not reproduced.)  The same objects drive the unmodified reference
    annotator in tests/test_reference_callsite.py."""
    from vafator_b200.pileups import collect_metrics_for_chrom
    sc = H.load_scenario(os.path.join(H.GOLDEN, name))
    run = sc["runs"][run_index]
    p = tmp_path / "a.bam"
    bamio.write_bam(str(p), sc["bams"][0], sc["contigs"])
    bam = bamio.BamFile(str(p))
    gold = run["metrics"][0]
    seen = set()
    for chrom in sc["contigs"]:
        variants = [VariantRecord(v["chrom"], v["pos"], v["ref"], v["alts"]) for v in sc["variants"] if v["chrom"] == chrom]
        if not variants:
            continue
        res = collect_metrics_for_chrom(chrom, variants, bam, run["params"]["min_bq"], run["params"]["min_mq"],
                                        run["params"]["include_ambiguous"])
        for (pos, ref, alt0), m in res.items():
            key = "%s:%d:%s:%s" % (chrom, pos, ref, alt0)
            seen.add(key)
            g = gold[key]
            strip = lambda d: {k: v for k, v in d.items() if k != ""}
            assert m.dp == g["dp"], key
            assert {k: v for k, v in m.ac.items() if v} == {k: v for k, v in strip(g["ac"]).items() if v}, key
            for ours, theirs in ((m.all_mqs, g["all_mqs"]), (m.all_bqs, g["all_bqs"]), (m.all_positions, g["all_positions"])):
                assert {k: list(v) for k, v in dict(ours).items()} == strip(theirs), key
            for ours, theirs in ((m.mqs, g["mqs"]), (m.bqs, g["bqs"]), (m.positions, g["positions"])):
                t = strip(theirs)
                assert set(dict(ours)) == set(t), key
                for k, v in t.items():
                    assert (v is None and math.isnan(ours[k])) or ours[k] == v, (key, k)
    assert seen == set(gold)


def test_nist_variant_set_on_the_device_equals_the_oracle_double(tmp_path_factory, monkeypatch):
    """BASELINE config 1 on the device: the files of tests/test_integration_nist.py through the `vafator` command
    line with the real engine == the same run with the device stages doubled by the oracle, line for line."""
    import test_host_logic as T
    import test_integration_nist as N
    from vafator_b200 import annotator as ann
    records, reads, bam, vcf, d = N.make_nist_files(tmp_path_factory)
    cli(["--input-vcf", vcf, "--output-vcf", str(d / "gpu.vcf"), "--bam", "my_normal", bam, "--bam", "my_tumor", bam])
    monkeypatch.setattr(ann, "Engine", T.FileOracleEngine)
    cli(["--input-vcf", vcf, "--output-vcf", str(d / "oracle.vcf"), "--bam", "my_normal", bam, "--bam", "my_tumor", bam])
    strip = lambda p: [l for l in open(p) if not l.startswith("##vafator_command_line")]
    assert strip(d / "gpu.vcf") == strip(d / "oracle.vcf")
