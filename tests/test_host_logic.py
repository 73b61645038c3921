"""Host logic on CPU: the composition layer (vafator_b200/engine.py: replicate merge, Counter / dict
lookup semantics, rounding, INFO ordering) driven by ORACLE counts instead of the device, against
the INFO strings the unmodified reference produced (tests/golden).  The device stages are replaced
by test doubles here only -- the product's Engine always calls libvfk.so."""
import os
from collections import OrderedDict

import numpy as np
import pytest
import scipy.special

import helpers as H
from oracle import c_oracle, vafator_ref as vr
from vafator_b200 import device, engine

SCENARIOS = H.golden_scenarios()


class OracleEngine(engine.Engine):
    def __init__(self, min_mapq, min_baseq, include_ambiguous, fpr, error_rate, normal_ploidy=2):
        self.params = device.make_params(min_mapq=min_mapq, min_baseq=min_baseq, include_ambiguous=include_ambiguous)
        self.fpr, self.error_rate = fpr, error_rate
        self._prm = (min_mapq, min_baseq, include_ambiguous)
        self._records = None

    def pileup(self, batch, units, stop):
        recs = self._records
        ou = [(batch.contigs.index(recs[i][0]), recs[i][1], recs[i][2], recs[i][3][j], recs[i][3][0], int(s))
              for i, j, s in zip(units.rec, units.alt_index, stop)]
        o = c_oracle.pileup(batch.as_dict(), ou, c_oracle.params(min_mapq=self._prm[0], min_baseq=self._prm[1]))
        c = np.zeros(len(ou), device.COUNTS_DTYPE)
        snv = (units.meta & 3) == 0
        c["dp"] = o["dp"] - (0 if self._prm[2] else np.where(snv, o["n_amb"], 0))
        for f in ("n_amb", "ac_ref", "ac_alt", "med2", "s2", "n_col"):
            c[f] = o[f]
        c["status"] = units.meta & 3
        return c

    def submit(self, batch, units, stop, eaf_units):
        """Test double of the fused device call (vfk_annotate_host_async): oracle counts + scipy statistics per unit."""
        counts = self.pileup(batch, units, stop) if units.n_units else np.zeros(0, device.COUNTS_DTYPE)
        return (None, counts, self.epilogue(counts, eaf_units))

    def values(self, batch, units, stop):
        """Test double of vfk_pileup_values: the per-read lists from the Python restatement."""
        from oracle import htslib_pileup as hp
        recs = self._records
        f = hp.AlignmentFile(hp.recs_from_soa(batch.as_dict()), batch.contigs)
        out = []
        for u in range(units.n_units):
            chrom, pos1 = batch.contigs[units.tid[u]], int(units.pos[u]) + 1
            kind = int(units.meta[u]) & 3
            ref_code, alt_code = (int(units.meta[u]) >> 8) & 0xFF, (int(units.meta[u]) >> 16) & 0xFF
            words = []
            for col in f.pileup(chrom, pos1 - 1, pos1, min_base_quality=self._prm[1], min_mapping_quality=self._prm[0],
                                max_depth=10 ** 6):
                if kind != 0:  # indel unit: REF list = reads without an indel here, ALT list = supporting reads (no BQ)
                    _, _, ref, alts = recs[units.rec[u]]
                    m = vr.column_metrics(pos1, ref, alts, col)
                    for allele, bit in ((ref, 28), (alts[0].upper(), 29)):
                        words += [mq | (qp << 16) | (1 << bit) for mq, qp in zip(m.dist["mq"][allele], m.dist["pos"][allele])]
                    continue
                for pr in col.pileups:
                    if pr.is_del or pr.is_refskip:
                        continue
                    q = pr.query_position
                    code = hp.SEQ_CHARS.index(pr.alignment.query_sequence[q])
                    w = pr.alignment.mapping_quality | (pr.alignment.query_qualities[q] << 8) | (q << 16)
                    w |= (1 << 28) if code == ref_code else 0
                    w |= (1 << 29) if code == alt_code else 0
                    if w >> 28:
                        words.append(w)
            out.append(np.asarray(words, np.uint32))
        return out

    def ranksum_s2(self, alt_lists, ref_lists):
        out = np.zeros((len(alt_lists), 3), np.uint64)
        for p, (a, r) in enumerate(zip(alt_lists, ref_lists)):
            for m, (sh, mk) in enumerate(((0, 0xFF), (8, 0xFF), (16, 0xFFF))):
                out[p, m] = H.midrank_s2([(int(x) >> sh) & mk for x in a], [(int(x) >> sh) & mk for x in r])
        return out

    def epilogue(self, counts, eaf):
        st = np.zeros(len(counts), device.STATS_DTYPE)
        power = vr.Power(fpr=self.fpr, error_rate=self.error_rate)
        for u, c in enumerate(counts):
            n1, n2 = int(c["ac_alt"]), int(c["ac_ref"])
            for m in range(3):
                if n1 and n2 and (m != 1 or (int(c["status"]) & 3) == 0):
                    z = (float(c["s2"][m]) / 2.0 - n1 * (n1 + n2 + 1) / 2.0) / np.sqrt(n1 * n2 * (n1 + n2 + 1) / 12.0)
                    st["z"][u][m], st["p"][u][m] = z, 2 * scipy.special.ndtr(-abs(z))
                else:
                    st["z"][u][m] = st["p"][u][m] = np.nan
            st["pu"][u] = power.pu_raw(int(c["dp"]), n1, eaf[u])
            st["pw"][u], st["k"][u] = power.pw_raw(int(c["dp"]), eaf[u])
        return st


@pytest.mark.parametrize("path", SCENARIOS, ids=[os.path.basename(p) for p in SCENARIOS])
def test_info_strings_match_reference(path):
    sc = H.load_scenario(path)
    records = [(v["chrom"], v["pos"], v["ref"], v["alts"]) for v in sc["variants"]]
    batches = H.scenario_batches(sc)
    for run in sc["runs"]:
        prm = run["params"]
        eng = OracleEngine(prm["min_mq"], prm["min_bq"], prm["include_ambiguous"], prm.get("fpr", 5e-7),
                           prm.get("error_rate", 1e-3))
        eng._records = records
        pl = {s: ([tuple(r) for r in p] if isinstance(p, list) else p) for s, p in sc["ploidies"].items()}
        power = vr.Power(ploidy=pl, purity=dict(sc["purities"]), normal_ploidy=prm.get("normal_ploidy", 2))
        eaf = {s: np.asarray([power.eaf(s, r[0], r[1]) for r in records]) for s in sc["samples"]}
        bams = OrderedDict((s, [batches[b] for b in idx]) for s, idx in sc["samples"].items())
        got = eng.annotate(records, bams, eaf)
        for i, (g, want) in enumerate(zip(got, run["info"])):
            w = OrderedDict((k, v) for k, v in want)
            assert list(g.keys()) == list(w.keys()), (sc["name"], prm, records[i], list(g.keys()), list(w.keys()))
            for k in w:
                assert str(g[k]) == str(w[k]), (sc["name"], prm, records[i], k, g[k], w[k])


class FileOracleEngine(OracleEngine):
    """Constructor of engine.Engine, device stages of OracleEngine -- stands in for Engine inside Annotator."""
    calls = []

    def __init__(self, device_index=0, min_mapq=0, min_baseq=29, include_ambiguous=True, fpr=5e-7, error_rate=1e-3,
                 devices=None):
        OracleEngine.__init__(self, min_mapq, min_baseq, include_ambiguous, fpr, error_rate)
        self.released = 0

    def annotate(self, records, bams, eaf, as_text=False, span=None):
        self._records = records
        FileOracleEngine.calls.append((records[0][0], len(records), [b.n_reads for bs in bams.values() for b in bs]))
        return OracleEngine.annotate(self, records, bams, eaf, as_text=as_text, span=span)

    def release_reads(self, batches=None):
        self.released += 1
        engine.Engine.release_reads(self, batches)


def _write_inputs(tmp_path, sc, indexed):
    from vafator_b200 import bamio
    vcf = tmp_path / "in.vcf"
    with open(vcf, "w") as fh:
        fh.write("##fileformat=VCFv4.2\n" + "".join("##contig=<ID=%s>\n" % c for c in sc["contigs"]))
        fh.write("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n")
        for k, v in enumerate(sc["variants"]):
            fh.write("%s\t%d\t.\t%s\t%s\t.\tPASS\t%s\n" % (v["chrom"], v["pos"], v["ref"], ",".join(v["alts"]),
                                                              "." if k % 2 else "OLD=1;DP=7"))
    bams = OrderedDict()
    for s, idx in sc["samples"].items():
        for b in idx:
            p = tmp_path / ("bam%d.bam" % b)
            bamio.write_bam(str(p), sc["bams"][b], sc["contigs"], index=indexed)
            bams.setdefault(s, []).append(str(p))
    return str(vcf), bams


@pytest.mark.parametrize("chunk", [None, 3, 1], ids=["whole_groups", "chunks_of_3", "chunks_of_1"])
@pytest.mark.parametrize("indexed", [False, True], ids=["whole_file", "bai_regions"])
@pytest.mark.parametrize("name", ["pileup_random_13.json", "pileup_handmade.json", "pileup_testx_real.json"])
def test_annotator_streams_chromosome_groups(tmp_path, monkeypatch, name, indexed, chunk):
    """Annotator.run on files (VCF reader, BAM decoder, region fetch, chromosome-group streaming, INFO append, VCF
    writer) with the device stages doubled by the oracle: every INFO key=value of the unmodified reference, one
    engine call per chromosome group, region batches released between groups.  A group annotated in chunks (each with
    the reads its own windows fetch, inside the span and under the read bound of the whole group) prints the same."""
    from vafator_b200 import annotator as ann
    from vafator_b200.ploidies import PloidyManager
    if chunk:
        monkeypatch.setattr(ann, "GROUP_CHUNK", chunk)
    sc = H.load_scenario(os.path.join(H.GOLDEN, name))
    run = sc["runs"][3] if len(sc["runs"]) > 3 else sc["runs"][-1]
    prm = run["params"]
    vcf, bams = _write_inputs(tmp_path, sc, indexed)
    ploidies = {}
    for s, pl in sc["ploidies"].items():
        if isinstance(pl, list):
            bed = tmp_path / ("%s.bed" % s)
            with open(bed, "w") as fh:
                for row in pl:
                    fh.write("\t".join(map(str, row)) + "\n")
            ploidies[s] = PloidyManager(local_copy_numbers=str(bed))
        else:
            ploidies[s] = PloidyManager(genome_wide_ploidy=float(pl))
    monkeypatch.setattr(ann, "Engine", FileOracleEngine)
    FileOracleEngine.calls = []
    a = ann.Annotator(input_vcf=vcf, output_vcf=str(tmp_path / "out.vcf"), input_bams=bams, purities=dict(sc["purities"]),
                      mapping_qual_thr=prm["min_mq"], base_call_qual_thr=prm["min_bq"], tumor_ploidies=ploidies,
                      normal_ploidy=prm.get("normal_ploidy", 2), fpr=prm.get("fpr", 5e-7), error_rate=prm.get("error_rate", 1e-3),
                      include_ambiguous_bases=prm["include_ambiguous"])
    a.run()
    body = [l.rstrip("\n") for l in open(tmp_path / "out.vcf") if not l.startswith("#")]
    assert len(body) == len(sc["variants"])
    for k, (line, want) in enumerate(zip(body, run["info"])):
        info = line.split("\t")[7].split(";")
        if k % 2 == 0:
            assert info[:2] == ["OLD=1", "DP=7"]
            info = info[2:]
        assert info == ["%s=%s" % (key, val) for key, val in want], (k, sc["variants"][k])
    # one engine call per run of consecutive records of one chromosome (per chunk of it)
    chroms = [v["chrom"] for v in sc["variants"]]
    groups = [c for i, c in enumerate(chroms) if i == 0 or chroms[i - 1] != c]
    assert sum(c[1] for c in FileOracleEngine.calls) == len(chroms)
    if chunk:
        sizes = [len(list(g)) for _, g in __import__("itertools").groupby(chroms)]
        assert len(FileOracleEngine.calls) == sum(-(-n // chunk) for n in sizes) and max(c[1] for c in FileOracleEngine.calls) <= chunk
        return
    assert [c[0] for c in FileOracleEngine.calls] == groups
    assert a.engine.released == (len(groups) + 1 if indexed else 1)
    assert not a.engine.__dict__.get("_hreads") or not indexed  # the wrappers of a group's batches are gone with the group
    if indexed and len(groups) > 1:  # a group's batches hold that chromosome's region only
        whole = [len([r for r in sc["bams"][b] if r["tid"] >= 0]) for idx in sc["samples"].values() for b in idx]
        assert all(n <= w for c in FileOracleEngine.calls for n, w in zip(c[2], whole))
        assert any(n < w for c in FileOracleEngine.calls for n, w in zip(c[2], whole))


class RandomCountsEngine(engine.Engine):
    """Device stages replaced by seeded random records: exercises the formatting paths on combinations no BAM
    scenario reaches (empty lists next to full ones, NaN statistics, uncovered records)."""

    def __init__(self, seed):
        self.params = device.make_params(min_mapq=1, min_baseq=30)
        self.fpr, self.error_rate = 5e-7, 1e-3
        self.rng = np.random.default_rng(seed)

    def pileup(self, batch, units, stop):
        n, rng = units.n_units, self.rng
        c = np.zeros(n, device.COUNTS_DTYPE)
        c["dp"] = rng.integers(0, 40, n)
        c["n_amb"] = rng.integers(0, 3, n)
        c["ac_ref"] = rng.integers(0, 3, n) * rng.integers(0, 12, n)
        c["ac_alt"] = rng.integers(0, 3, n) * rng.integers(0, 12, n)
        c["n_col"] = rng.integers(0, 4, n)
        c["med2"] = rng.integers(0, 400, (n, 3, 2))
        c["s2"] = rng.integers(0, 500, (n, 3))
        c["status"] = units.meta & 3
        return c

    def submit(self, batch, units, stop, eaf_units):
        counts = self.pileup(batch, units, stop)
        return (None, counts, self.epilogue(counts, eaf_units))

    def epilogue(self, counts, eaf):
        n, rng = len(counts), self.rng
        st = np.zeros(n, device.STATS_DTYPE)
        st["z"] = np.where(rng.random((n, 3)) < 0.1, np.nan, rng.normal(size=(n, 3)) * 3)
        st["p"] = np.where(np.isnan(st["z"]), np.nan, rng.random((n, 3)))
        st["pu"], st["pw"], st["k"] = rng.random(n), rng.random(n) * 1.0001, rng.integers(1, 9, n)
        return st


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_columnwise_formatting_equals_record_by_record(seed):
    """SingleBamColumns (all records at once) == Engine._one (record by record), field for field."""
    rng = np.random.default_rng(100 + seed)
    alleles = ["A", "C", "G", "T", "a", "N", "R", "AT", "ACG", "TTTT", "at", "GGc"]
    records = []
    pos = 100
    for _ in range(1500):
        pos += int(rng.integers(0, 30))
        ref = alleles[rng.integers(0, len(alleles))]
        alts = [alleles[rng.integers(0, len(alleles))] for _ in range(int(rng.choice([1, 1, 1, 2, 3])))]
        records.append(("chr1" if pos < 20000 else "chr2", pos, ref, alts))

    class B:
        contigs = ["chr1", "chr2"]
    bams = OrderedDict([("tumor", [B()]), ("normal", [B()])])  # replicate samples always go record by record
    eaf = {s: rng.random(len(records)) for s in bams}
    outs = []
    for vectorised in (True, False):
        eng = RandomCountsEngine(seed)
        eng.vectorised_format = vectorised
        outs.append(eng.annotate(records, bams, eaf))
    for i, (a, b) in enumerate(zip(*outs)):
        assert list(a.items()) == list(b.items()), (records[i], a, b)
    text = RandomCountsEngine(seed).annotate(records, bams, eaf, as_text=True)  # ... == the finished INFO text
    assert text == [";".join("%s=%s" % kv for kv in a.items()) for a in outs[0]]
    if seed == 1:  # ... and == the same columns formatted in chunks by worker processes
        eng = RandomCountsEngine(seed)
        eng.format_processes, eng.format_parallel_from = 2, 0
        pooled = eng.annotate(records, bams, eaf)
        assert [list(a.items()) for a in pooled] == [list(a.items()) for a in outs[0]]
        eng = RandomCountsEngine(seed)  # (the double's random stream restarts with the engine)
        eng.format_processes, eng.format_parallel_from = 2, 0
        assert eng.annotate(records, bams, eaf, as_text=True) == text
        assert engine._FORMAT_POOL is not None and engine._FORMAT_POOL[0] == 2  # the pool did start
    assert any("tumor_rsmq" in a for a in outs[0]) and any("tumor_rsmq" not in a for a in outs[0])
    assert any("nan" in str(a["tumor_mq"]) for a in outs[0])


@pytest.mark.parametrize("indexed", [False, True], ids=["whole_file", "bai_windows"])
def test_collect_metrics_for_chrom_call_site_on_cpu(tmp_path, monkeypatch, indexed):
    """The reference's first call site (vafator/pileups.py:106) with the device stages doubled by the oracle: same
    keys and values as the unmodified reference produced (tests/golden), also when an indexed BAM is fetched as
    windows around the variants."""
    import math
    from vafator_b200 import bamio, pileups
    from vafator_b200.pileup_utils import VariantRecord
    sc = H.load_scenario(os.path.join(H.GOLDEN, "pileup_testx_real.json"))
    run = sc["runs"][0]
    prm = run["params"]
    p = tmp_path / "a.bam"
    bamio.write_bam(str(p), sc["bams"][0], sc["contigs"], index=indexed)
    bam = bamio.BamFile(str(p))
    gold = run["metrics"][0]
    n_keys = 0
    for chrom in sc["contigs"]:
        mine = [v for v in sc["variants"] if v["chrom"] == chrom]
        if not mine:
            continue
        variants = [VariantRecord(v["chrom"], v["pos"], v["ref"], v["alts"]) for v in mine]
        eng = OracleEngine(prm["min_mq"], prm["min_bq"], True, 5e-7, 1e-3)
        eng._records = pileups.callsite_records(chrom, variants)  # one unit per letter other than REF for an SNV record
        monkeypatch.setattr(pileups, "_engine", lambda *a, **k: eng)
        res = pileups.collect_metrics_for_chrom(chrom, variants, bam, prm["min_bq"], prm["min_mq"], True)
        assert {"%s:%d:%s:%s" % ((chrom,) + k) for k in res} == {k for k in gold if k.startswith(chrom + ":")}
        for (pos, ref, alt0), m in res.items():
            g = gold["%s:%d:%s:%s" % (chrom, pos, ref, alt0)]
            assert m.dp == g["dp"]
            for allele in m.mqs:
                want = g["mqs"].get(allele)
                assert (math.isnan(m.mqs[allele]) and want is None) or m.mqs[allele] == want
            n_keys += 1
    assert n_keys > 150


def test_native_float_text_is_pythons():
    """vfkio_py_repr (the float printing of the native INFO formatter) against repr(float): random doubles of every magnitude,
    values rounded to 3 / 5 decimals as the statistics are, halves, integers, the switch-overs to exponent notation, specials."""
    import ctypes as C
    lib = device.libio()
    lib.vfkio_py_repr.argtypes = [C.c_double, C.c_char_p]
    rng = np.random.default_rng(3)
    xs = [0.0, -0.0, 1.0, -1.0, 0.5, 1e-4, 9.999e-5, 1e-5, 1.5e-5, 1e16, 9999999999999998.0, 1e15, 123456789012345678.0, 5e-324, 1.7976931348623157e308,
          float("nan"), float("inf"), float("-inf"), 0.1, 0.30000000000000004, 2.0 / 3.0, 100.0, 1e22, 1e23]
    xs += (rng.random(2000) * 10.0 ** rng.integers(-12, 20, 2000) * rng.choice([-1, 1], 2000)).tolist()
    xs += np.round(rng.random(2000), 5).tolist() + np.round(rng.normal(0, 8, 2000), 3).tolist() + (rng.integers(0, 600, 500) / 2.0).tolist()
    xs += np.frombuffer(rng.bytes(8 * 2000), np.float64).tolist()
    buf = C.create_string_buffer(40)
    for x in xs:
        n = lib.vfkio_py_repr(C.c_double(x), buf)
        assert buf.raw[:n].decode() == repr(float(x)), (x, buf.raw[:n])


def test_native_info_text_equals_python_formatter():
    """vfkio_format_info (libvfkio: what Annotator.run prints with) against format_single_bam with the text templates, string for
    string, on random per-record / per-row arrays: SNVs, indels, records without a unit or a column, multiallelic and ALT-less
    records, lower-case ALTs, empty lists, NaN statistics, depths of zero, allele-frequency rounding ties."""
    rng = np.random.default_rng(5)
    for trial in range(6):
        n_rec = [1, 7, 300, 5000, 5000, 40000][trial]
        n_alt = rng.choice([0, 1, 1, 1, 1, 2, 3], n_rec)
        first = np.concatenate([[0], np.cumsum(n_alt)]).astype(np.int64)
        n_row = int(first[-1])
        rec_of = np.repeat(np.arange(n_rec), n_alt)
        kind = rng.choice([-1, 0, 0, 0, 1, 2], n_rec).astype(np.int64)
        has = rng.random(n_rec) < 0.9
        dp = np.where(rng.random(n_rec) < 0.05, 0, rng.integers(1, 3000, n_rec)).astype(np.int64)
        n_amb = rng.integers(0, 3, n_rec).astype(np.int64)
        ac_ref = np.where(rng.random(n_rec) < 0.2, 0, rng.integers(1, 200, n_rec)).astype(np.int64)
        med2_ref = rng.integers(0, 600, (n_rec, 3)).astype(np.int64)
        eaf = np.where(rng.random(n_rec) < 0.5, 0.5, rng.random(n_rec))
        alt_idx = (np.arange(n_row) - first[rec_of]).astype(np.int64)
        present = rng.random(n_row) < 0.9
        ac = np.where(rng.random(n_row) < 0.3, 0, rng.integers(1, 3000, n_row)).astype(np.int64)
        if trial == 4:  # ties of round(ac / dp, 5): k / 200000, k odd
            dp[:] = 200000
            ac[:] = rng.integers(0, 100000, n_row) * 2 + 1
        med2_alt = rng.integers(0, 600, (n_row, 3)).astype(np.int64)
        pu, pw = np.round(rng.random(n_row), 5), np.round(rng.random(n_row) ** 8, 5)
        k = rng.integers(1, 60, n_row).astype(np.int64)
        z = np.round(rng.normal(0, 6, (n_row, 3)), 3)
        p = np.round(rng.random((n_row, 3)) ** 6, 5)
        z[rng.random((n_row, 3)) < 0.05] = np.nan
        p[np.isnan(z)] = np.nan
        args = (first, kind, alt_idx, present, has, dp, n_amb, ac, ac_ref, med2_ref, med2_alt, pu, pw, k, z, p, eaf)
        for sample in ("tumor", "s%1"):
            want, _ = engine.format_single_bam(*args, templates=engine.text_templates(sample))
            got = engine.format_single_bam_native(sample, *args)
            assert len(got) == len(want) == n_rec
            for i, (g, w) in enumerate(zip(got, want)):
                assert g == w, (trial, i, g, w)
