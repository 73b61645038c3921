"""Sparse fetch (vafator_b200/fetch.py): the windows around the variants of a chromosome group must hand the
pileup every read its columns depend on.  Checked with the C oracle: every integer field of every unit computed
from the windowed batch == computed from the span the reference fetches, on the reference's real BAMs and on
synthetic low-coverage reads with long reference skips and deletions (where the first read pushed after a column
can be a distant mate, so that windows must be extended)."""
import os

import numpy as np
import pytest

import helpers as H
from oracle import c_oracle
from vafator_b200 import bamio, device, fetch, synth

FIELDS = ("dp", "n_amb", "ac_ref", "ac_alt", "med2", "s2", "n_col")


def _oracle(batch, contigs, records, stop, min_mapq, min_baseq):
    units = [(contigs.index(c), p, ref, a, alts[0], stop) for c, p, ref, alts in records for a in alts]
    return c_oracle.pileup(batch.as_dict(), units, c_oracle.params(min_mapq=min_mapq, min_baseq=min_baseq))


def _check_group(bam, chrom, records, min_mapq, min_baseq, pad, gap):
    params = device.make_params(min_mapq=min_mapq, min_baseq=min_baseq)
    positions = [r[1] for r in records]
    first, last = positions[0], positions[-1]
    full = bam.read_batch([(chrom, first - 1, last)])
    part = fetch.fetch_group(bam, chrom, first, last, positions, params, pad=pad, merge_gap=gap)
    a = _oracle(full, bam.contigs, records, last, min_mapq, min_baseq)
    b = _oracle(part, bam.contigs, records, last, min_mapq, min_baseq)
    for f in FIELDS:
        assert np.array_equal(a[f], b[f]), (chrom, f, pad, gap, np.nonzero(np.any(np.atleast_2d(a[f] != b[f]).reshape(len(a[f]), -1), axis=1))[0][:5])
    return full.n_reads, part.n_reads


@pytest.mark.parametrize("pad,gap", [(0, 1), (150, 300), (fetch.PAD, fetch.MERGE_GAP)])
def test_windows_on_the_references_real_bams(tmp_path, pad, gap):
    sc = H.load_scenario(os.path.join(H.GOLDEN, "pileup_testx_real.json"))
    fewer = 0
    for b, recs in enumerate(sc["bams"]):
        p = tmp_path / ("b%d.bam" % b)
        bamio.write_bam(str(p), recs, sc["contigs"], index=True)
        bam = bamio.BamFile(str(p))
        for chrom in sc["contigs"]:
            records = [(v["chrom"], v["pos"], v["ref"], v["alts"]) for v in sc["variants"] if v["chrom"] == chrom]
            if not records:
                continue
            for mq, bq in ((0, 0), (1, 30)):
                n_full, n_part = _check_group(bam, chrom, records, mq, bq, pad, gap)
                assert n_part <= n_full
                fewer += n_part < n_full
    if pad < fetch.PAD:
        assert fewer > 0  # the windows did leave reads out


@pytest.mark.parametrize("seed", [5, 6])
def test_windows_on_low_coverage_reads_with_long_skips(tmp_path, seed):
    cfg = synth.wgs(n_contigs=2, contig_len=400_000, seed=synth.BASE_SEED + 40 + seed, pairs_per_tile=12, p_skip=0.25, p_del=0.15,
                    p_ins=0.05, p_softclip=0.05, frag_mean=260.0, frag_sd=80.0, snv_period=300, indel_period=900)
    batch = synth.reads(cfg)
    p = tmp_path / "s.bam"
    bamio.write_bam_batch(str(p), batch, [401_000] * 2)
    bam = bamio.BamFile(str(p))
    assert bam.read_batch([(c, 0, 401_000) for c in bam.contigs]).n_reads == batch.n_reads
    recs = synth.variant_records(cfg)
    rng = np.random.default_rng(seed)
    extended = 0
    for chrom in bam.contigs:
        mine = [r for r in recs if r[0] == chrom]
        keep = sorted(rng.choice(len(mine), size=min(len(mine), 120), replace=False).tolist())
        records = [mine[k] for k in keep]
        for pad, gap in ((0, 1), (40, 100), (fetch.PAD, fetch.MERGE_GAP)):
            n_full, n_part = _check_group(bam, chrom, records, 1, 20, pad, gap)
            assert n_part <= n_full
        # the tight windows cannot all have been enough: some column's next pushed read is a mate beyond a skip
        params = device.make_params(min_mapq=1, min_baseq=20)
        positions = [r[1] for r in records]
        cols = [q - 1 for q in positions]
        windows = fetch.plan_windows(cols, positions[0] - 1, positions[-1], 0, 1)
        tight = bam.read_batch([(chrom, a, b) for a, b in windows])
        extended += len(fetch.missing_reads(tight, device.read_passes(tight, params), cols, windows, positions[-1],
                                            bam.contigs.index(chrom)))
    assert extended > 0


@pytest.mark.parametrize("seed", [7, 8])
def test_chunks_of_a_group_fetch_what_their_columns_need(tmp_path, seed):
    """A chromosome group annotated in chunks (annotator.GROUP_CHUNK): each chunk fetches the windows of its own columns inside
    the GROUP's span and keeps the group's read bound.  Every unit of every chunk, from the chunk's batch, equals the unit computed
    from the whole span -- on low-coverage reads with long skips, sparse chunks, dense chunks, chunks of one record."""
    cfg = synth.wgs(n_contigs=1, contig_len=400_000, seed=synth.BASE_SEED + 60 + seed, pairs_per_tile=12, p_skip=0.25, p_del=0.15,
                    p_ins=0.05, p_softclip=0.05, frag_mean=260.0, frag_sd=80.0, snv_period=300, indel_period=900)
    batch = synth.reads(cfg)
    p = tmp_path / "s.bam"
    bamio.write_bam_batch(str(p), batch, [401_000])
    bam = bamio.BamFile(str(p))
    chrom = bam.contigs[0]
    recs = [r for r in synth.variant_records(cfg) if r[0] == chrom]
    rng = np.random.default_rng(seed)
    group = [recs[k] for k in sorted(rng.choice(len(recs), size=min(len(recs), 400), replace=False).tolist())]
    first, last = group[0][1], group[-1][1]
    params = device.make_params(min_mapq=1, min_baseq=20)
    full = bam.read_batch([(chrom, first - 1, last)])
    want = _oracle(full, bam.contigs, group, last, 1, 20)
    n_units = [len(r[3]) for r in group]
    unit0 = np.concatenate([[0], np.cumsum(n_units)])
    fewer = 0
    for size, pad, gap in ((1, 0, 1), (7, 40, 100), (50, fetch.PAD, fetch.MERGE_GAP), (150, 0, 1)):
        for at in range(0, len(group), size):
            chunk = group[at:at + size]
            part = fetch.fetch_group(bam, chrom, first, last, [r[1] for r in chunk], params, pad=pad, merge_gap=gap)
            got = _oracle(part, bam.contigs, chunk, last, 1, 20)
            u0, u1 = int(unit0[at]), int(unit0[min(at + size, len(group))])
            for f in FIELDS:
                assert np.array_equal(got[f], want[f][u0:u1]), (size, at, f)
            fewer += part.n_reads < full.n_reads
    assert fewer > 0


def test_plan_windows():
    assert fetch.plan_windows([100, 105, 5000, 20000], 50, 19000, pad=10, merge_gap=20) == [[90, 116], [4990, 5011]]
    assert fetch.plan_windows([100, 130], 0, 1000, pad=10, merge_gap=20) == [[90, 141]]
    assert fetch.plan_windows([], 0, 10) == []


def test_next_pushed_mate_beyond_a_skip_is_fetched(tmp_path):
    """Read `a` hangs in a 500-bp reference skip at the column; the base-quality filter looks at its next base
    (Q20), which the overlap adjustment has raised to Q40 iff mate `b` -- 300 bp beyond the column, the next read
    the reference pushes -- has arrived.  A window around the column alone misses `b`; fetch_group extends it."""
    contigs = ["chrS"]
    seq_a, seq_b = "ACGTACGTAC" + "G" * 40, "G" * 50
    found = None
    for k in range(40):  # the name hash decides which mate keeps the summed quality: find a name where `a` does
        name = "pair%d" % k
        recs = [dict(qname=name, flag=99, tid=0, pos=100, mapq=60, cigar="10M500N40M", seq=seq_a, qual=[20] * 50,
                     mtid=0, mpos=605, isize=555),
                dict(qname=name, flag=147, tid=0, pos=605, mapq=60, cigar="50M", seq=seq_b, qual=[20] * 50,
                     mtid=0, mpos=100, isize=-555),
                dict(qname="far", flag=0, tid=0, pos=5000, mapq=60, cigar="50M", seq="A" * 50, qual=[40] * 50)]
        p = tmp_path / ("n%d.bam" % k)
        bamio.write_bam(str(p), recs, contigs, lengths=[10000], index=True)
        bam = bamio.BamFile(str(p))
        records = [("chrS", 301, "A", ["AT"]), ("chrS", 5001, "A", ["T"])]
        full = bam.read_batch([("chrS", 300, 5001)])
        want = _oracle(full, contigs, records, 5001, 0, 30)
        if want["dp"][0] == 1:
            found = (bam, records, want)
            break
    assert found, "no read name gave read `a` the summed quality"
    bam, records, want = found
    params = device.make_params(min_mapq=0, min_baseq=30)
    tight = bam.read_batch([("chrS", 300, 301), ("chrS", 5000, 5001)])
    assert tight.n_reads == 2 and _oracle(tight, contigs, records, 5001, 0, 30)["dp"][0] == 0  # `b` missing: Q20, dropped
    got = fetch.fetch_group(bam, "chrS", 301, 5001, [301, 5001], params, pad=0, merge_gap=1)
    assert got.n_reads == 3
    for f in FIELDS:
        assert np.array_equal(_oracle(got, contigs, records, 5001, 0, 30)[f], want[f]), f
