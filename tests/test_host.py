"""CPU-side checks of the product's host code: the canonical batch as it crosses the C ABI, unit encoding, C-ABI surface."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import helpers as H
from oracle import c_oracle
from vafator_b200 import build, device
from vafator_b200.batch import ReadBatch, build_units, SEQ_CHARS, CODE_NONE

SCENARIOS = H.golden_scenarios()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_quality_scaling_is_integer_exact():
    """htslib stores (uint8_t)(0.8 * q); the kernel uses (4*q)/5 -- identical for every byte value."""
    for q in range(256):
        assert int(0.8 * q) == (4 * q) // 5


@pytest.mark.parametrize("path", SCENARIOS, ids=[os.path.basename(p) for p in SCENARIOS])
def test_host_reads_are_the_decoded_records_untouched(path):
    """What crosses the C ABI is the canonical batch: the decoded fields, array by array, plus the batch metadata
    (contig table, longest read) -- nothing derived from a CIGAR, no filter verdict, no mate link."""
    sc = H.load_scenario(path)
    for recs, batch in zip(sc["bams"], H.scenario_batches(sc)):
        hr = device.host_reads(batch)
        assert hr.n_reads == len(recs) == batch.n_placed
        assert hr.contig_off[0] == 0 and hr.contig_off[-1] == len(recs) and len(hr.contig_off) == len(sc["contigs"]) + 1
        assert hr.max_l_qseq == max(len(r["seq"]) for r in recs)
        c = hr.c_struct()
        assert c.n_reads == len(recs) and c.n_cigar_words == len(batch.cigar)
        a = hr.arrays
        for t in range(len(sc["contigs"])):
            lo, hi = int(hr.contig_off[t]), int(hr.contig_off[t + 1])
            assert all(r["tid"] == t for r in recs[lo:hi])
            assert np.all(np.diff(a["pos"][lo:hi].astype(np.int64)) >= 0)
        for i, r in enumerate(recs):
            assert a["pos"][i] == r["pos"] and a["flag"][i] == r["flag"] and a["mapq"][i] == r["mapq"]
            lq = len(r["seq"])
            assert a["l_qseq"][i] == lq
            q0, s0 = int(a["qual_off"][i]), int(a["seq_off"][i])
            assert bytes(a["qual"][q0:q0 + lq]) == bytes(ord(ch) - 33 for ch in r["qual"])
            codes = [(int(a["seq"][s0 + (q >> 1)]) >> (0 if q & 1 else 4)) & 15 for q in range(lq)]
            assert "".join(SEQ_CHARS[x] for x in codes) == r["seq"]
            ops = re.findall(r"(\d+)([MIDNSHP=X])", r["cigar"])
            co, nc = int(a["cigar_off"][i]), int(a["n_cigar"][i])
            assert [(int(w) >> 4, "MIDNSHP=X"[int(w) & 15]) for w in a["cigar"][co:co + nc]] == [(int(n), o) for n, o in ops]


def test_malformed_batches_are_rejected_on_the_host():
    """A placed record with a negative start, reads out of order, unplaced reads in the middle: ValueError before
    anything reaches the device (the kernels check again and flag VFK_ST_BADBATCH: tests/test_gpu_synth.py)."""
    rec = lambda name, tid, pos: dict(qname=name, flag=0, tid=tid, pos=pos, mapq=60, cigar="10M", seq="ACGTACGTAC", qual="I" * 10)
    with pytest.raises(ValueError):
        ReadBatch.from_records([rec("a", 0, -5)], ["c"])
    ok = ReadBatch.from_records([rec("a", 0, 5), rec("b", 0, 9)], ["c"])
    ok.check_sorted()
    bad = ReadBatch.from_records([rec("a", 0, 9), rec("b", 0, 5)], ["c"])
    with pytest.raises(ValueError):
        bad.check_sorted()
    mixed = ReadBatch.from_records([rec("a", 0, 5), rec("u", -1, -1), rec("b", 0, 9)], ["c"])
    with pytest.raises(ValueError):
        mixed.n_placed
    neg = ReadBatch.from_records([rec("a", 0, 5)], ["c"])
    neg.pos[0] = -3
    with pytest.raises(ValueError):
        neg.check_sorted()


def test_units_encoding():
    u = build_units([("c", 10, "A", ["T", "G"]), ("c", 11, "A", ["AT"]), ("c", 12, "ATT", ["A"]),
                     ("c", 13, "AT", ["GC"]), ("c", 14, "a", ["t"]), ("c", 15, "A", ["AnT", "ATT"])], ["c"])
    assert u.n_units == 2 + 1 + 1 + 0 + 1 + 1
    assert u.unit_of[(0, 0)] == 0 and u.unit_of[(0, 1)] == 1 and (3, 0) not in u.unit_of
    assert list(u.pos[:2]) == [9, 9]
    assert (u.meta[0] & 3) == 0 and ((u.meta[0] >> 8) & 0xFF) == SEQ_CHARS.index("A") and ((u.meta[1] >> 16) & 0xFF) == SEQ_CHARS.index("G")
    assert (u.meta[2] & 3) == 1 and u.length[2] == 1 and u.ins_seq[u.seq_off[2]] == SEQ_CHARS.index("T")
    assert (u.meta[3] & 3) == 2 and u.length[3] == 2
    assert ((u.meta[4] >> 8) & 0xFF) == CODE_NONE and ((u.meta[4] >> 16) & 0xFF) == CODE_NONE  # lowercase never matches
    assert u.ins_seq[u.seq_off[5]] == CODE_NONE and u.ins_seq[u.seq_off[5] + 1] == SEQ_CHARS.index("T")
    with pytest.raises(ValueError):
        build_units([("zz", 1, "A", ["T"])], ["c"])


def test_cabi_library_exports_every_declared_symbol():
    """The shared libraries build (nvcc cross-compiles without a GPU), load, and export exactly what
    include/*.h declares.  No compute call is made here."""
    libvfk, libio = build.build_all()
    for header, path in (("vfk.h", libvfk), ("vfk_io.h", libio), ("vfk_synth.h", build.LIBVFKSYNTH)):
        text = open(os.path.join(ROOT, "include", header)).read()
        declared = set(re.findall(r"\b(vfk(?:io)?_[a-z_]+)\s*\(", text))
        lib = C.CDLL(path)
        for name in sorted(declared):
            assert hasattr(lib, name), (header, name)
    lib = C.CDLL(libvfk)
    assert lib.vfk_version() == 200
    assert np.dtype(device.COUNTS_DTYPE).itemsize == 80
    sass = subprocess.run(["cuobjdump", "-lelf", libvfk], capture_output=True, text=True).stdout
    assert "sm_100a" in sass


@pytest.mark.parametrize("by_index", [False, True])
@pytest.mark.parametrize("threads", [1, 5])
def test_fetch_list_host_gather(by_index, threads):
    """The host half of a fetch list (vfk_fetch_gather: what the stream callback of the host entry point runs between two
    kernels) against numpy: both request formats, entries that want nothing, entries that point outside the arrays; lists that
    end inside a block of 64 and lists shorter than the thread count."""
    lib = device.libvfk()
    rng = np.random.default_rng(7)
    n_reads = 5000
    l = rng.integers(30, 200, n_reads)
    qual_off = np.concatenate([[0], np.cumsum(l)[:-1]]).astype(np.int64)
    seq_off = np.concatenate([[0], np.cumsum((l + 1) // 2)[:-1]]).astype(np.int64)
    qual = rng.integers(0, 94, int(l.sum())).astype(np.uint8)
    seq = rng.integers(0, 256, int(((l + 1) // 2).sum())).astype(np.uint8)
    for n in (0, 3, 64, 1000, 12345):
        i = rng.integers(0, n_reads, n)
        q = (rng.random(n) * l[i]).astype(np.int64)
        want = qual[qual_off[i] + q].astype(np.uint16) | (seq[seq_off[i] + q // 2].astype(np.uint16) << 8)
        req = np.empty((n, 2), np.uint64)
        req[:, 0], req[:, 1] = (i, q) if by_index else (qual_off[i] + q, seq_off[i] + q // 2)
        nothing = rng.random(n) < 0.2
        req[nothing, 0] = np.uint64(0xFFFFFFFFFFFFFFFF)
        outside = (rng.random(n) < 0.05) & ~nothing
        req[outside, 0 if by_index else 1] = np.uint64(1 << 40)
        want[nothing | outside] = 0
        out = np.full(n, 0xABCD, np.uint16)
        rc = lib.vfk_fetch_gather(C.c_void_p(req.ctypes.data), C.c_int64(n), C.c_void_p(qual.ctypes.data), C.c_int64(len(qual)),
                                  C.c_void_p(seq.ctypes.data), C.c_int64(len(seq)),
                                  C.c_void_p(qual_off.ctypes.data if by_index else None), C.c_void_p(seq_off.ctypes.data if by_index else None),
                                  C.c_int64(n_reads), C.c_int(threads), C.c_void_p(out.ctypes.data))
        assert rc == 0
        assert np.array_equal(out, want)


def test_decoders_write_contiguous_pools(tmp_path):
    """ReadBatch.contiguous (set by the synthetic generator and the BAM decoder) promises that the three pool offset arrays are the
    prefix sums of the lengths -- the condition under which the host entry point hands them over as NULL (include/vfk.h) and
    slices keep the property."""
    from vafator_b200 import bamio, sharding, synth
    cfg = synth.wgs(contig_len=200_000, n_contigs=2)
    b = synth.reads(cfg)
    p = tmp_path / "c.bam"
    bamio.write_bam_batch(str(p), b)
    decoded = bamio.BamFile(str(p), use_index=False).read_batch()
    for batch in (b, decoded, sharding.slice_reads(b, 1000, 5000)):
        assert batch.contiguous
        n = batch.n_reads
        lq = batch.l_qseq.astype(np.int64)
        assert np.array_equal(batch.cigar_off, np.cumsum(batch.n_cigar.astype(np.int64)) - batch.n_cigar)
        assert np.array_equal(batch.seq_off, np.cumsum((lq + 1) // 2) - (lq + 1) // 2)
        assert np.array_equal(batch.qual_off, np.cumsum(lq) - lq)
        assert len(batch.cigar) == int(batch.n_cigar.sum()) and len(batch.qual) == int(lq.sum()) and n > 0
    hr = device.host_reads(b)
    c = hr.c_struct()
    assert hr.implicit_offsets and not c.cigar_off and not c.seq_off and not c.qual_off  # NULL on the host entry ...
    assert device.host_reads(b, implicit_offsets=False).c_struct().cigar_off           # ... unless asked otherwise
