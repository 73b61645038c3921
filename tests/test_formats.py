"""CPU tests of the formats and CLI surface around the hot path."""
import gzip
import os
import subprocess
import sys

import numpy as np
import pytest

import helpers as H
from vafator_b200 import bamio, vcf
from vafator_b200.batch import ReadBatch
from vafator_b200.command_line import _validate_alignment_file_extension, annotator as cli, build_parser
from vafator_b200.ploidies import PloidyManager

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_RES = "/root/reference/vafator/tests/resources"


def test_help_exits_zero():
    """BASELINE config 1, literally tests/integration_test_00.sh:4 -- `vafator --help`."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "vafator"), "--help"], capture_output=True, text=True)
    assert r.returncode == 0 and "--input-vcf" in r.stdout and "--bam" in r.stdout


def test_cli_defaults_and_validation():
    a = build_parser().parse_args(["--input-vcf", "x", "--output-vcf", "y"])
    assert (a.mapping_quality, a.base_call_quality, a.normal_ploidy, a.fpr, a.error_rate) == (1, 30, 2, 5e-7, 1e-3)
    assert a.exclude_ambiguous_bases is False and a.num_processes == 1
    with pytest.raises(ValueError, match="SAM input is not supported"):
        _validate_alignment_file_extension("reads.sam")
    with pytest.raises(ValueError, match="at least one bam"):
        cli(["--input-vcf", "x", "--output-vcf", "y"])
    with pytest.raises(ValueError, match="purity value for a sample for which no BAM"):
        cli(["--input-vcf", "x", "--output-vcf", "y", "--bam", "t", "a.bam", "--purity", "n", "0.5"])
    with pytest.raises(ValueError, match="Multiple purity"):
        cli(["--input-vcf", "x", "--output-vcf", "y", "--bam", "t", "a.bam", "--purity", "t", "0.5", "--purity", "t", "0.6"])
    with pytest.raises(ValueError, match="neither a copy number value or a BED"):
        cli(["--input-vcf", "x", "--output-vcf", "y", "--bam", "t", "a.bam", "--tumor-ploidy", "t", "/no/such.bed"])


def test_ploidy_manager_reference_cases(tmp_path):
    """vafator/tests/test_ploidy_manager.py:10-114 boundary cases (BED reproduced here)."""
    assert PloidyManager().ploidy == 2.0
    assert PloidyManager(genome_wide_ploidy=3.5).get_ploidies(["chr1"], [5])[0] == 3.5
    bed = tmp_path / "cn.bed"
    bed.write_text("chr1\t10000\t20000\t1.2\nchr1\t20000\t30000\t2.1\nchr2\t10000\t20000\t3.2\n")
    pm = PloidyManager(local_copy_numbers=str(bed))
    assert pm.get_ploidies(["chr1"] * 4 + ["chr2", "chr3"], [10000, 10001, 20000, 20001, 15000, 15000]).tolist() == \
        [2.0, 1.2, 1.2, 2.1, 3.2, 2.0]
    assert pm.report_value == str(bed)
    with pytest.raises(ValueError):
        PloidyManager(local_copy_numbers="/no/such/file.bed")
    # overlapping rows: the first one in file order wins; numeric-only chromosome names never match (pandas dtype)
    bed.write_text("chr1\t0\t100\t4\nchr1\t50\t200\t6\n")
    assert PloidyManager(local_copy_numbers=str(bed)).get_ploidies(["chr1"] * 2, [60, 150]).tolist() == [4.0, 6.0]
    bed.write_text("1\t0\t100\t4\n")
    assert PloidyManager(local_copy_numbers=str(bed)).get_ploidies(["1"], [60]).tolist() == [2.0]


def test_vcf_round_trip(tmp_path):
    src = tmp_path / "a.vcf"
    src.write_text("##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS1\n"
                   "chr1\t10\t.\tA\tT,G\t.\tPASS\tDP=3;X\tGT\t0/1\nchr1\t12\trs1\tAT\tA\t50\t.\t.\tGT\t1/1\n")
    for out_name in ("o.vcf", "o.vcf.gz"):
        r = vcf.VcfReader(str(src))
        r.add_to_header("##note=1")
        r.add_info_to_header({"ID": "t_af", "Number": "A", "Type": "Float", "Description": "d"})
        w = vcf.VcfWriter(str(tmp_path / out_name), r)
        recs = list(r)
        assert (recs[0].CHROM, recs[0].POS, recs[0].REF, recs[0].ALT) == ("chr1", 10, "A", ["T", "G"])
        recs[0].INFO["t_af"] = "0.5,0.1"
        recs[0].INFO["DP"] = 9
        recs[1].INFO["t_dp"] = 4
        for x in recs:
            w.write_record(x)
        w.close()
        opener = gzip.open if out_name.endswith(".gz") else open
        lines = opener(tmp_path / out_name, "rt").read().splitlines()
        assert lines[:4] == ["##fileformat=VCFv4.2", "##note=1", '##INFO=<ID=t_af,Number=A,Type=Float,Description="d">',
                             "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS1"]
        assert lines[4].split("\t")[7] == "DP=9;X;t_af=0.5,0.1" and lines[4].endswith("GT\t0/1")
        assert lines[5].split("\t")[7] == "t_dp=4"
        assert [x.POS for x in vcf.VcfReader(str(tmp_path / out_name))] == [10, 12]


def test_bam_round_trip(tmp_path):
    sc = H.load_scenario(H.golden_scenarios()[0])
    p = tmp_path / "t.bam"
    bamio.write_bam(str(p), sc["bams"][0], sc["contigs"])
    b = bamio.BamFile(str(p))
    got, want = b.read_batch(), ReadBatch.from_records(sc["bams"][0], sc["contigs"])
    assert b.contigs == sc["contigs"]
    for k in ("tid", "pos", "flag", "mapq", "l_qseq", "n_cigar", "cigar", "seq", "qual", "mtid", "mpos", "isize",
              "qname_id", "qname_hash"):
        assert np.array_equal(getattr(got, k), getattr(want, k)), k
    (tmp_path / "bad.bam").write_bytes(b"not a bam")
    with pytest.raises(ValueError):
        bamio.BamFile(str(tmp_path / "bad.bam"))
    with pytest.raises(FileNotFoundError):
        bamio.BamFile(str(tmp_path / "missing.bam"))
    unsorted = list(reversed(sc["bams"][0]))
    bamio.write_bam(str(tmp_path / "u.bam"), unsorted, sc["contigs"])
    with pytest.raises(ValueError, match="not coordinate sorted"):
        bamio.BamFile(str(tmp_path / "u.bam")).read_batch()


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_RES, "TESTX_S1_L001.bam")), reason="reference checkout not present")
def test_decoder_on_the_references_real_bams():
    """The two real BAMs the reference ships (bwa, 101 bp, 4 contigs): record census from SURVEY.md 8(c)."""
    b = bamio.BamFile(os.path.join(REF_RES, "TESTX_S1_L001.bam"))
    r = b.read_batch()
    assert b.contigs == ["chr1", "chr2", "chr3", "chr4"] and list(b.lengths) == [1000000] * 4
    ops = np.bincount(r.cigar & 15, minlength=9)
    assert (ops[2], ops[1], ops[4]) == (34, 18, 124)  # D / I / S operations
    assert int(np.count_nonzero((r.flag & 4) == 0)) == 638 and set(np.unique(r.l_qseq)) == {101}
    r.check_sorted()


def _overlapping(whole, tid, beg, end):
    span = np.zeros(whole.n_reads, np.int64)
    for i in range(whole.n_reads):
        c = whole.cigar[whole.cigar_off[i]:whole.cigar_off[i] + whole.n_cigar[i]]
        span[i] = sum(int(w) >> 4 for w in c if (int(w) & 15) in (0, 2, 3, 7, 8))
    return (whole.tid == tid) & (whole.pos < end) & (whole.pos + np.maximum(span, 1) > beg)


def _same_reads(sub, whole, keep):
    assert sub.n_reads == int(keep.sum())
    for k in ("tid", "pos", "flag", "mapq", "l_qseq", "n_cigar", "mtid", "mpos", "isize", "qname_id", "qname_hash"):
        assert np.array_equal(getattr(sub, k), getattr(whole, k)[keep]), k
    idx = np.nonzero(keep)[0]
    for j, i in enumerate(idx[:200]):
        assert np.array_equal(sub.cigar[sub.cigar_off[j]:sub.cigar_off[j] + sub.n_cigar[j]],
                              whole.cigar[whole.cigar_off[i]:whole.cigar_off[i] + whole.n_cigar[i]])
        assert np.array_equal(sub.qual[sub.qual_off[j]:sub.qual_off[j] + sub.l_qseq[j]],
                              whole.qual[whole.qual_off[i]:whole.qual_off[i] + whole.l_qseq[i]])


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_RES, "TESTX_S1_L001.bam.bai")), reason="reference checkout not present")
def test_indexed_fetch_on_the_references_real_bam_and_samtools_index():
    """Region reads through the samtools-built .bai the reference ships == the whole-file decode, filtered."""
    path = os.path.join(REF_RES, "TESTX_S1_L001.bam")
    whole = bamio.BamFile(path, use_index=False).read_batch()
    b = bamio.BamFile(path)
    assert b.index_path and b.contigs == ["chr1", "chr2", "chr3", "chr4"]
    rng = np.random.default_rng(3)
    for tid, name in enumerate(b.contigs):
        pos = whole.pos[whole.tid == tid]
        spans = [(0, 1 << 29), (int(pos.min()), int(pos.min()) + 1), (int(pos.max()) + 5000, int(pos.max()) + 6000)]
        spans += [tuple(sorted(int(x) for x in rng.integers(0, 1_000_000, 2))) for _ in range(12)]
        for a, e in spans:
            if e > a:
                _same_reads(b.read_batch([(name, a, e)]), whole, _overlapping(whole, tid, a, e))
    # several regions in one call, given out of order and with an unknown contig
    got = b.read_batch([("chr3", 500_000, 700_000), ("chrZ", 1, 2), ("chr1", 0, 400_000)])
    _same_reads(got, whole, _overlapping(whole, 0, 0, 400_000) | _overlapping(whole, 2, 500_000, 700_000))
    assert b.read_batch().n_reads == whole.n_reads  # no regions: the whole file


def test_indexed_fetch_with_the_in_tree_index_writer(tmp_path):
    """write_bam(index=True) -> region reads == whole-file decode, filtered; long reads spanning several 16-kbp
    windows, empty windows and a contig without reads exercise the linear index."""
    rng = np.random.default_rng(5)
    recs = []
    for tid, n in ((0, 1500), (2, 600)):
        pos = np.sort(np.concatenate([rng.integers(1000, 60_000, n // 2), rng.integers(400_000, 420_000, n - n // 2)]))
        for k, p in enumerate(pos):
            if k % 97 == 0:
                cigar, l = "50M30000N50M", 100  # spans two or three windows
            elif k % 5 == 0:
                cigar, l = "10S80M2D10M", 100
            else:
                cigar, l = "100M", 100
            recs.append(dict(qname="r%d_%d" % (tid, k), flag=0, tid=tid, pos=int(p), mapq=60, cigar=cigar,
                             seq="".join("ACGT"[(k + j) & 3] for j in range(l)), qual="I" * l, mtid=-1, mpos=-1, isize=0))
    recs.sort(key=lambda r: (r["tid"], r["pos"]))
    path = str(tmp_path / "x.bam")
    bamio.write_bam(path, recs, ["c1", "c2", "c3"], [1 << 20] * 3, index=True)
    whole = bamio.BamFile(path, use_index=False).read_batch()
    assert whole.n_reads == len(recs)
    b = bamio.BamFile(path)
    assert b.index_path == path + ".bai"
    for tid, name in ((0, "c1"), (1, "c2"), (2, "c3")):
        spans = [(0, 1 << 20), (30_500, 30_501), (100_000, 300_000), (399_000, 400_500), (59_990, 405_000)]
        spans += [tuple(sorted(int(x) for x in rng.integers(0, 450_000, 2))) for _ in range(15)]
        for a, e in spans:
            if e > a:
                _same_reads(b.read_batch([(name, a, e)]), whole, _overlapping(whole, tid, a, e))
    with pytest.raises(ValueError):
        b._lib  # regions are merged per contig by read_batch; the C entry rejects unsorted input
        import ctypes as C
        names = (C.c_char_p * 2)(b"c3", b"c1")
        beg, end = np.asarray([0, 0], np.int32), np.asarray([10, 10], np.int32)
        h = C.c_void_p()
        from vafator_b200 import device as _dev
        _dev._io_check(b._lib.vfkio_bam_open_regions(path.encode(), (path + ".bai").encode(), 2, names,
                                                     C.c_void_p(beg.ctypes.data), C.c_void_p(end.ctypes.data), C.byref(h)))


# ---- multiallelics-filter against the reference run on the same inputs ----
def test_multiallelic_filter_matches_reference_golden(tmp_path):
    """tests/golden/multiallelic.json: outputs of the UNMODIFIED reference filter (oracle/gen_golden_multiallelic.py)."""
    import json
    import os
    from vafator_b200.multiallelic_filter import MultiallelicFilter
    golden = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "multiallelic.json")))
    assert len(golden) >= 9
    for name, case in golden.items():
        src, dst = tmp_path / (name + ".vcf"), tmp_path / (name + ".out.vcf")
        src.write_text(case["input"])
        MultiallelicFilter(input_vcf=str(src), output_vcf=str(dst), tumor_sample_name=case["sample"]).run()
        got = [l for l in dst.read_text().split("\n") if l and not l.startswith("##vafator_command_line=")]
        assert got == case["output"], name
        assert any(l.startswith("##vafator_command_line=") for l in dst.read_text().split("\n"))
    # the float32 widening and the leading comma of a first annotation are what the reference emits
    pairs = [l for l in golden["pairs"]["output"] if l.startswith("chr4\t1235")]
    assert len(pairs) == 1 and pairs[0].endswith("multiallelic=,T,0.10000000149011612")


def test_multiallelic_filter_cli_and_empty_input(tmp_path):
    from vafator_b200.command_line import multiallelics_filter
    src, dst = tmp_path / "in.vcf", tmp_path / "out.vcf.gz"
    src.write_text("##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n"
                   "1\t5\t.\tA\tC\t.\t.\tt_af=0.5\n1\t5\t.\tA\tG\t.\t.\tt_af=0.75\n")
    multiallelics_filter(["--input-vcf", str(src), "--output-vcf", str(dst), "--tumor-sample-name", "t"])
    import gzip
    body = [l for l in gzip.open(dst, "rt").read().split("\n") if l and not l.startswith("#")]
    assert body == ["1\t5\t.\tA\tG\t.\t.\tt_af=0.75;multiallelic=,C,0.5"]
    empty = tmp_path / "empty.vcf"
    empty.write_text("##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n")
    multiallelics_filter(["--input-vcf", str(empty), "--output-vcf", str(tmp_path / "e.out.vcf")])
    assert [l for l in (tmp_path / "e.out.vcf").read_text().split("\n") if l and not l.startswith("#")] == []


def test_decoder_rejects_damaged_files(tmp_path):
    """Neither entry point of the decoder may crash on a file that is not a sound BAM: CRAM magic, plain text,
    a file cut inside a BGZF member, a member whose deflate payload is damaged, an index of another file."""
    import helpers as H
    sc = H.load_scenario(os.path.join(H.GOLDEN, "pileup_random_11.json"))
    good = tmp_path / "good.bam"
    bamio.write_bam(str(good), sc["bams"][0], sc["contigs"], index=True)
    raw = open(good, "rb").read()
    n_good = bamio.BamFile(str(good), use_index=False).read_batch().n_reads
    assert n_good > 0

    def attempt(name, data, index=None):
        p = tmp_path / name
        p.write_bytes(data)
        if index is not None:
            (tmp_path / (name + ".bai")).write_bytes(index)
        outcomes = []
        for use_index in (False, True) if index is not None else (False,):
            try:
                f = bamio.BamFile(str(p), use_index=use_index)
                b = f.read_batch([(sc["contigs"][0], 0, 1 << 28)] if use_index else None)
                outcomes.append(b.n_reads)
            except ValueError as e:
                outcomes.append(str(e))
        return outcomes

    bai = open(str(good) + ".bai", "rb").read()
    assert all("CRAM" in str(o) for o in attempt("x.cram.bam", b"CRAM" + raw[4:], bai))
    assert all(isinstance(o, str) for o in attempt("text.bam", b"@HD\tVN:1.6\n" * 50, bai))
    for o in attempt("cut.bam", raw[:len(raw) // 2], bai):  # ends inside a member: an error, never a short batch
        assert isinstance(o, str), o
    damaged = bytearray(raw)
    damaged[len(raw) // 2] ^= 0xFF
    damaged[len(raw) // 2 + 1] ^= 0xFF
    outcomes = attempt("flip.bam", bytes(damaged), bai)
    assert isinstance(outcomes[0], str), outcomes  # whole file: inflate fails or the member's CRC32 does not match
    for o in outcomes:  # region fetch: an error, or the damaged member lies outside the region
        assert isinstance(o, str) or o <= n_good
    # stored (level 0) members: DEFLATE itself cannot notice a flipped quality byte, the BGZF CRC32 must (htslib fails such a block)
    b = bamio.BamFile(str(good), use_index=False).read_batch()
    stored = tmp_path / "stored.bam"
    bamio.write_bam_batch(str(stored), b, level=0)
    assert bamio.BamFile(str(stored), use_index=False).read_batch().n_reads == n_good
    raw0 = bytearray(open(stored, "rb").read())
    raw0[len(raw0) // 2] ^= 0x01
    out = attempt("stored_flip.bam", bytes(raw0))
    assert isinstance(out[0], str), out
    assert all(isinstance(o, str) for o in attempt("noindex.bam", raw, b"BAI\1" + b"\0" * 3)[1:])


def test_inflate_matches_zlib():
    """vfkio_inflate_raw (the decoder's own DEFLATE) against zlib: every compression level and strategy (stored,
    fixed and dynamic codes, single-distance-code streams), text-like / run-heavy / random / BAM-like inputs, empty
    input, sizes around the word-copy slack; wrong sizes, cut and damaged streams must be refused or decoded like
    zlib -- never a write outside the buffer."""
    import ctypes as C
    import zlib
    import helpers as H
    from vafator_b200 import device
    lib = device.libio()
    lib.vfkio_inflate_raw.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t]
    rng = np.random.default_rng(7)
    GUARD = 64

    def mine(stream, size):
        buf = np.full(size + 2 * GUARD, 0xA5, np.uint8)
        rc = lib.vfkio_inflate_raw(stream, len(stream), buf.ctypes.data + GUARD, size)
        assert (buf[:GUARD] == 0xA5).all() and (buf[GUARD + size:] == 0xA5).all(), "wrote outside the buffer"
        return rc, buf[GUARD:GUARD + size].tobytes()

    def inputs():
        yield b""
        yield b"a"
        yield b"abcabcabc" * 7000
        yield bytes(rng.integers(0, 256, 70000, dtype=np.uint8))
        yield bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), 65000))
        yield bytes(np.repeat(rng.integers(30, 42, 900, dtype=np.uint8), rng.integers(1, 140, 900)))[:65280]
        for n in (1, 7, 8, 9, 257, 258, 259, 265, 273, 274, 275, 300, 1000):
            yield bytes(rng.integers(0, 4, n, dtype=np.uint8))
        sc = H.load_scenario(os.path.join(H.GOLDEN, "pileup_random_12.json"))
        yield b"".join((r["qname"] + r["seq"]).encode() + bytes(r["qual"] if not isinstance(r["qual"], str) else r["qual"].encode())
                       for r in sc["bams"][0])[:65280]

    n_ok = n_fallback = 0
    for data in inputs():
        for level in (0, 1, 4, 6, 9):
            for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_RLE, zlib.Z_HUFFMAN_ONLY, zlib.Z_FILTERED):
                c = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
                stream = c.compress(data) + c.flush()
                rc, got = mine(stream, len(data))
                if rc == 0:
                    assert got == data
                    n_ok += 1
                else:
                    n_fallback += 1  # allowed (zlib takes over), but must stay the exception
                if len(data) > 0:
                    assert mine(stream, len(data) - 1)[0] != 0 and mine(stream, len(data) + 1)[0] != 0
                    assert mine(stream[:len(stream) // 2], len(data))[0] != 0
    assert n_ok > 20 * n_fallback, (n_ok, n_fallback)
    # damaged streams: whatever zlib says about them, a success here must carry zlib's bytes
    data = bytes(rng.choice(np.frombuffer(b"ACGTN#<>AAAA", np.uint8), 40000))
    stream = bytearray(zlib.compress(data, 6)[2:-4])
    for _ in range(300):
        bad = bytearray(stream)
        for _k in range(int(rng.integers(1, 4))):
            bad[int(rng.integers(0, len(bad)))] ^= 1 << int(rng.integers(0, 8))
        rc, got = mine(bytes(bad), len(data))
        if rc == 0:
            d = zlib.decompressobj(-15)
            try:
                ref = d.decompress(bytes(bad))
            except zlib.error:
                ref = None
            assert ref is not None and ref[:len(data)] == got and len(ref) == len(data)


def test_multi_region_fetch_equals_filter_of_the_whole_file(tmp_path):
    """Many regions in one call (sparse windows, touching windows, a window inside a long reference skip): the
    batch is exactly the reads of the whole file that overlap any region, each once, in file order -- for the
    parallel per-region scan (many regions) and the sequential one (few)."""
    from vafator_b200 import synth
    from vafator_b200.sharding import _ref_span
    cfg = synth.wgs(n_contigs=2, contig_len=300_000, p_skip=0.05, p_del=0.05)
    batch = synth.reads(cfg)
    p = tmp_path / "m.bam"
    bamio.write_bam_batch(str(p), batch, [301_000, 301_000])
    whole = bamio.BamFile(str(p), use_index=False).read_batch()
    assert whole.n_reads == batch.n_reads
    end = whole.pos.astype(np.int64) + np.maximum(_ref_span(whole), 1)
    bam = bamio.BamFile(str(p))
    rng = np.random.default_rng(5)
    for n_regions in (1, 3, 40, 400):
        regs = []
        for _ in range(n_regions):
            c = int(rng.integers(0, 2))
            a = int(rng.integers(0, 299_000))
            regs.append((bam.contigs[c], a, a + int(rng.choice([1, 50, 700, 5000]))))
        regs.append((bam.contigs[0], 100_000, 100_500))
        regs.append((bam.contigs[0], 100_500, 101_000))  # touches the one before
        got = bam.read_batch(regs)
        keep = np.zeros(whole.n_reads, bool)
        for c, a, e in regs:
            keep |= (whole.tid == bam.contigs.index(c)) & (whole.pos < e) & (end > a)
        assert got.n_reads == int(keep.sum()), n_regions
        for f in ("tid", "pos", "flag", "mapq", "qname_id", "mpos", "isize", "l_qseq", "n_cigar"):
            assert np.array_equal(getattr(got, f), getattr(whole, f)[keep]), (n_regions, f)
        assert np.array_equal(got.qual, np.concatenate([whole.qual[o:o + l] for o, l in zip(whole.qual_off[keep], whole.l_qseq[keep])]))
    assert bam.read_batch([("no_such_contig", 0, 100)]).n_reads == 0


def test_vcf_record_info_as_text():
    """Annotations handed over as finished text are appended, or -- when the record already carries one of the keys --
    replace the old values in place, exactly like the {key: value} form."""
    base = ["chr1", "100", ".", "A", "T", ".", "PASS"]
    new = {"tumor_af": "0.5", "tumor_dp": 12, "tumor_bq": "30.0,31.5"}
    text = ";".join("%s=%s" % kv for kv in new.items())
    for old in (".", "", "DP=7;SOMATIC", "tumor_dp=3;DP=7", "SOMATIC;tumor_af=0.1;tumor_bq=1,2;X=tumor_dp", "Xtumor_dp=4"):
        a, b = vcf.VcfRecord(base + [old, "GT", "0/1"]), vcf.VcfRecord(base + [old, "GT", "0/1"])
        a.INFO, b.INFO = dict(new), text
        assert a.line() == b.line(), old
    r = vcf.VcfRecord(base + ["DP=7"])
    r.INFO = text
    assert r.line().split("\t")[7] == "DP=7;" + text


def test_bgzf_vcf_writer_across_blocks(tmp_path):
    """A .vcf.gz larger than several BGZF blocks reads back (gzip: concatenated members) line for line."""
    src = tmp_path / "in.vcf"
    lines = ["chr1\t%d\t.\tA\tT\t.\tPASS\tDP=%d;NOTE=%s" % (100 + i, i % 97, "x" * (i % 50)) for i in range(6000)]
    src.write_text("##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n" + "\n".join(lines) + "\n")
    rd = vcf.VcfReader(str(src))
    out = tmp_path / "o.vcf.gz"
    wr = vcf.VcfWriter(str(out), rd)
    for rec in rd:
        rec.INFO = "tumor_dp=%d" % rec.POS
        wr.write_record(rec)
    wr.close()
    rd.close()
    raw = out.read_bytes()
    assert raw.count(b"\x1f\x8b\x08\x04") >= 4 and raw.endswith(bamio._BGZF_EOF)
    got = [l for l in gzip.open(out, "rt").read().split("\n") if l and not l.startswith("#")]
    assert got == ["%s;tumor_dp=%d" % (l, 100 + i) for i, l in enumerate(lines)]
