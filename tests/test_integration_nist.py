"""BASELINE config 1 (SURVEY.md 8(d) C1 ii): the variant set of the reference's integration test
(tests/integration_test_01.sh: 368 NIST records, 310 SNV / 24 INS / 34 DEL, REF up to 18 bases, ALT up to 40 --
tests/golden/nist_variants.json, extracted by oracle/gen_golden_nist.py) against simulated 2x100 paired reads that
carry those very alleles (the BAM of that test is absent from the reference checkout).  On CPU: the `vafator` command
line on files with the device stages doubled by the oracle, the assertions of integration_test_01.sh, and agreement
of the two oracle formulations on the real indel shapes.  The same files through the GPU: tests/test_gpu_annotate.py
(round 2)."""
import json
import os
from collections import namedtuple

import numpy as np
import pytest

import helpers as H
from oracle import c_oracle, htslib_pileup as hp, vafator_ref as vr
from vafator_b200 import bamio

REGION0, REGION1 = 1_000_000, 2_000_000
READ_LEN = 100


def nist_records():
    with open(os.path.join(H.GOLDEN, "nist_variants.json")) as fh:
        return [(c, p, r, a) for c, p, r, a in json.load(fh)["records"]]


def simulate(records, seed=20261020, pairs_per_variant=45):
    """Reads around every variant: a random reference carrying the REF alleles (half of the insertions followed by a
    copy of their inserted bases, the only shape the reference's shifted comparison counts: SURVEY.md A.5), fragments
    that carry each variant they span with probability 0.4."""
    rng = np.random.default_rng(seed)
    base0 = REGION0 - 1000
    ref = rng.choice(np.frombuffer(b"ACGT", np.uint8), REGION1 - REGION0 + 2000)
    by_pos = {}
    for k, (c, pos, r, alts) in enumerate(records):
        ref[pos - 1 - base0:pos - 1 - base0 + len(r)] = np.frombuffer(r.encode(), np.uint8)
        by_pos[pos - 1] = (r, alts[0])
    for k, (c, pos, r, alts) in enumerate(records):
        ins = alts[0][1:] if len(r) == 1 and len(alts[0]) > 1 else ""
        nxt = records[k + 1][1] if k + 1 < len(records) else 1 << 40
        if ins and k % 2 == 0 and nxt - pos > len(ins) + 5:
            ref[pos - base0:pos - base0 + len(ins)] = np.frombuffer(ins.encode(), np.uint8)

    def read_from(start, carried):
        """(cigar ops, sequence) of READ_LEN query bases from 0-based `start` on the haplotype carrying `carried`."""
        ops, seq, x = [], [], start

        def op(code, n):
            if ops and ops[-1][0] == code:
                ops[-1][1] += n
            else:
                ops.append([code, n])
        while len(seq) < READ_LEN:
            v = by_pos.get(x)
            room = READ_LEN - len(seq)
            if v is not None and x in carried and 5 < len(seq) and room > len(v[1]) + 5:
                r, a = v
                if len(r) == 1 and len(a) == 1:
                    seq.append(a)
                    op(0, 1)
                    x += 1
                elif len(r) == 1:  # insertion after the anchor
                    seq.append(r)
                    op(0, 1)
                    seq.extend(a[1:])
                    op(1, len(a) - 1)
                    x += 1
                else:  # deletion after the anchor
                    seq.append(r[0])
                    op(0, 1)
                    op(2, len(r) - 1)
                    x += len(r)
                continue
            seq.append(chr(ref[x - base0]))
            op(0, 1)
            x += 1
        return [(c, n) for c, n in ops], "".join(seq), x

    quals = np.asarray([12, 23, 30, 37, 41])
    out = []
    n = 0
    positions = sorted(by_pos)
    for k, (c, pos, r, alts) in enumerate(records):
        for _ in range(pairs_per_variant):
            frag = int(np.clip(rng.normal(240, 40), READ_LEN + 10, 400))
            start = pos - 1 - int(rng.integers(20, frag - 20))
            lo = np.searchsorted(positions, start)
            hi = np.searchsorted(positions, start + frag + 60)
            carried = {q for q in positions[lo:hi] if rng.random() < 0.4}
            c1, s1, e1 = read_from(start, carried)
            start2 = start + frag - READ_LEN
            c2, s2, e2 = read_from(start2, carried)
            err = lambda s: "".join(ch if rng.random() > 0.004 else "ACGT"[int(rng.integers(0, 4))] for ch in s)
            mapq = int(rng.choice([60, 60, 60, 60, 29, 0]))
            name = "f%d" % n
            n += 1
            isize = e2 - start
            out.append(dict(qname=name, flag=99, tid=0, pos=start, mapq=mapq, cigar=c1, seq=err(s1),
                            qual=rng.choice(quals, READ_LEN).tolist(), mtid=0, mpos=start2, isize=isize))
            out.append(dict(qname=name, flag=147, tid=0, pos=start2, mapq=mapq, cigar=c2, seq=err(s2),
                            qual=rng.choice(quals, READ_LEN).tolist(), mtid=0, mpos=start, isize=-isize))
    out.sort(key=lambda d: d["pos"])
    return out


def make_nist_files(tmp_path_factory):
    d = tmp_path_factory.mktemp("nist")
    records = nist_records()
    reads = simulate(records)
    bam = d / "sim.chr1_1000000_2000000.bam"
    bamio.write_bam(str(bam), reads, ["chr1"], lengths=[249_250_621], index=True)
    vcf = d / "project.NIST.vcf"
    with open(vcf, "w") as fh:
        fh.write("##fileformat=VCFv4.1\n##contig=<ID=chr1,length=249250621>\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n")
        for c, p, r, a in records:
            fh.write("%s\t%d\t.\t%s\t%s\t50\tPASS\tAC=2;AF=0.5;DP=11\n" % (c, p, r, ",".join(a)))
    return records, reads, str(bam), str(vcf), d


@pytest.fixture(scope="module")
def nist_files(tmp_path_factory):
    return make_nist_files(tmp_path_factory)


def test_variant_set_is_the_references():
    records = nist_records()
    kinds = [("snv" if len(r) == 1 and len(a[0]) == 1 else "ins" if len(r) == 1 else "del") for _, _, r, a in records]
    assert len(records) == 368 and kinds.count("snv") == 310 and kinds.count("ins") == 24 and kinds.count("del") == 34
    assert max(len(r) for _, _, r, _ in records) == 18 and max(len(a[0]) for _, _, _, a in records) == 40
    assert records[0][:2] == ("chr1", 1001177) and records[-1][1] == 1992977


def test_integration_test_01_on_cpu(nist_files, monkeypatch):
    """tests/integration_test_01.sh: the same BAM as my_normal and my_tumor; every record carries both samples' af / ac /
    dp.  Device stages doubled by the C oracle (tests/test_host_logic.py); sparse fetch, streaming and text formatting
    are the product's."""
    import test_host_logic as T
    from vafator_b200 import annotator as ann
    from vafator_b200.command_line import annotator as cli
    records, reads, bam, vcf, d = nist_files
    monkeypatch.setattr(ann, "Engine", T.FileOracleEngine)
    T.FileOracleEngine.calls = []
    out = d / "vafator.vcf"
    cli(["--input-vcf", vcf, "--output-vcf", str(out), "--bam", "my_normal", bam, "--bam", "my_tumor", bam])
    body = [l.rstrip("\n").split("\t") for l in open(out) if not l.startswith("#")]
    assert len(body) == len(records)
    n_alt = {"snv": 0, "ins": 0, "del": 0}
    for (c, p, r, a), f in zip(records, body):
        info = dict(kv.split("=", 1) for kv in f[7].split(";"))
        assert f[:2] == [c, str(p)] and info["AC"] == "2"  # the record's own INFO is kept
        for s in ("my_normal", "my_tumor"):
            for k in ("af", "ac", "dp", "n", "pu", "pw", "k", "eaf", "mq", "bq", "pos"):
                assert "%s_%s" % (s, k) in info, (p, s, k)
        assert all(info["my_normal_" + k] == info["my_tumor_" + k] for k in ("af", "ac", "dp", "mq", "bq", "pos"))
        assert int(info["my_normal_dp"]) > 3, (p, info["my_normal_dp"])
        kind = "snv" if len(r) == 1 and len(a[0]) == 1 else "ins" if len(r) == 1 else "del"
        n_alt[kind] += int(info["my_normal_ac"]) > 0
        if kind != "snv":
            assert info["my_normal_bq"] == "0,0"  # indel metrics carry no base qualities (pileups.py:301,304)
    assert n_alt["snv"] > 290 and n_alt["del"] > 30 and 4 <= n_alt["ins"] <= 16  # only tandem-duplication insertions match
    # one chromosome group, fetched as windows: far fewer reads than the span would not apply here (dense), but one call
    assert [c[0] for c in T.FileOracleEngine.calls] == ["chr1"]


def test_oracle_formulations_agree_on_the_real_indel_shapes(nist_files):
    """Streaming restatement vs direct C restatement on a stretch of the NIST set that holds long insertions and
    deletions, reads carrying them."""
    records, reads, bam, vcf, d = nist_files
    lens = np.asarray([max(len(r), len(a[0])) for _, _, r, a in records])
    k0 = int(np.argmax(np.convolve(lens > 4, np.ones(40, int), "valid")))  # the 40 consecutive records richest in long indels
    sub = records[k0:k0 + 40]
    lo, hi = sub[0][1] - 600, sub[-1][1] + 600
    near = [x for x in reads if lo <= x["pos"] < hi]
    from vafator_b200.batch import ReadBatch
    batch = ReadBatch.from_records(near, ["chr1"])
    soa = batch.as_dict()
    f = hp.AlignmentFile(hp.recs_from_soa(soa), batch.contigs)
    V = namedtuple("V", "CHROM POS REF ALT")
    checked = supported = 0
    for mq, bq in ((0, 0), (1, 30)):
        stream = vr.collect_metrics_for_chrom("chr1", [V(*r) for r in sub], f, bq, mq, True)
        units = [(0, r[1], r[2], r[3][0], r[3][0], sub[-1][1]) for r in sub]
        direct = c_oracle.pileup(soa, units, c_oracle.params(min_mapq=mq, min_baseq=bq))
        for rec, (tid, pos1, ref, alt, alt0, _stop) in zip(direct, units):
            m = stream.get((pos1, ref, alt0))
            as_json = None if m is None else dict(ac=dict(m.ac), dp=m.dp, all_bqs=m.dist["bq"], all_mqs=m.dist["mq"],
                                                  all_positions=m.dist["pos"])
            want = H.expected_record(as_json, ref, alt, alt0, True)
            ctx = (mq, bq, pos1, ref, alt)
            assert rec["dp"] == want["dp"] and rec["ac_ref"] == want["ac_ref"] and rec["ac_alt"] == want["ac_alt"], ctx
            assert rec["med2"].tolist() == want["med2"] and [int(x) for x in rec["s2"]] == want["s2"], ctx
            checked += 1
            supported += rec["ac_alt"] > 0 and len(ref) != len(alt)
    assert checked == 80 and supported >= 4
