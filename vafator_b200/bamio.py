"""BAM files <-> columnar read batches.

`BamFile` stands where pysam.AlignmentFile stands in the reference (vafator/annotator.py:161-164):
it decodes a coordinate-sorted BAM with the in-tree BGZF/BAM decoder (csrc/vfkio_bam.cpp, zlib)
into the canonical `ReadBatch`.  `write_bam` is the inverse for tests and synthetic fixtures (pure
Python + zlib, not performance code).  CRAM and SAM are rejected, as SAM is in the reference
(vafator/command_line.py:16-23)."""
from __future__ import annotations

import ctypes as C
import os
import struct
import zlib
from typing import List, Optional, Sequence

import numpy as np

from . import device as _dev
from .batch import ReadBatch, SEQ_CHARS


class _PlanC(C.Structure):
    _fields_ = [("n_reads", C.c_int64), ("n_cigar", C.c_int64), ("n_seq_bytes", C.c_int64), ("n_qual_bytes", C.c_int64)]


def _lib():
    lib = _dev.libio()
    lib.vfkio_bam_contig_name.restype = C.c_char_p
    lib.vfkio_bam_header_text.restype = C.c_char_p
    lib.vfkio_bam_contig_length.restype = C.c_int64
    lib.vfkio_bam_close.restype = None
    lib.vfkio_bam_close.argtypes = [C.c_void_p]
    lib.vfkio_bam_open_regions.argtypes = [C.c_char_p, C.c_char_p, C.c_int32, C.POINTER(C.c_char_p), C.c_void_p, C.c_void_p,
                                           C.POINTER(C.c_void_p)]
    return lib


class BamFile:
    """A BAM file: `.contigs`, `.lengths`, `.read_batch()`.

    With an index next to it (`x.bam.bai` or `x.bai`) only the header is read on open and
    `read_batch(regions)` decodes just the records overlapping the regions -- the reads the reference
    fetches for [first.POS-1, last.POS) of a chromosome group (vafator/pileups.py:71-80, :137-138).
    Without an index (or without regions) the whole file is decoded once and cached."""

    def __init__(self, path: str, mode: str = "r", use_index: bool = True):
        if not os.path.exists(path):
            raise FileNotFoundError("could not open alignment file `%s`: No such file or directory" % path)
        self.path = path
        self._lib = _lib()
        self.index_path: Optional[str] = None
        if use_index:
            for cand in (path + ".bai", os.path.splitext(path)[0] + ".bai"):
                if os.path.exists(cand):
                    self.index_path = cand
                    break
        h = C.c_void_p()
        if self.index_path:
            _dev._io_check(self._lib.vfkio_bam_open_regions(path.encode(), self.index_path.encode(), 0, None, None, None, C.byref(h)))
        else:
            _dev._io_check(self._lib.vfkio_bam_open(path.encode(), C.byref(h)))
        self._h = h
        n = self._lib.vfkio_bam_n_contigs(h)
        self.contigs: List[str] = [self._lib.vfkio_bam_contig_name(h, i).decode() for i in range(n)]
        self.lengths = np.asarray([self._lib.vfkio_bam_contig_length(h, i) for i in range(n)], np.int64)
        self.header_text = self._lib.vfkio_bam_header_text(h).decode(errors="replace")
        self._batch: Optional[ReadBatch] = None

    def _fill(self, h) -> ReadBatch:
        if not self._lib.vfkio_bam_is_sorted(h):
            raise ValueError("%s is not coordinate sorted" % self.path)
        plan = _PlanC()
        _dev._io_check(self._lib.vfkio_bam_plan(h, C.byref(plan)))
        n = plan.n_reads
        arr = dict(tid=np.empty(n, np.int32), pos=np.empty(n, np.int32), flag=np.empty(n, np.uint16),
                   mapq=np.empty(n, np.uint8), l_qseq=np.empty(n, np.int32), n_cigar=np.empty(n, np.int32),
                   cigar_off=np.empty(n, np.int64), seq_off=np.empty(n, np.int64), qual_off=np.empty(n, np.int64),
                   cigar=np.empty(max(plan.n_cigar, 1), np.uint32), seq=np.empty(max(plan.n_seq_bytes, 1), np.uint8),
                   qual=np.empty(max(plan.n_qual_bytes, 1), np.uint8), mtid=np.empty(n, np.int32),
                   mpos=np.empty(n, np.int32), isize=np.empty(n, np.int32), qname_id=np.empty(n, np.uint64),
                   qname_hash=np.empty(n, np.uint32))
        s = _dev._IoReadsC()
        s.n, s.n_contigs = n, len(self.contigs)
        for k, a in arr.items():
            setattr(s, k, a.ctypes.data)
        _dev._io_check(self._lib.vfkio_bam_fill(h, C.byref(s)))
        arr["cigar"], arr["seq"], arr["qual"] = arr["cigar"][:plan.n_cigar], arr["seq"][:plan.n_seq_bytes], arr["qual"][:plan.n_qual_bytes]
        return ReadBatch(contigs=list(self.contigs), contig_lengths=self.lengths, contiguous=True, **arr)  # vfkio_bam_fill: prefix-sum offsets

    def read_batch(self, regions: Optional[Sequence[tuple]] = None) -> ReadBatch:
        """regions: [(contig, start, stop)] 0-based half open, any order; overlapping ones are merged.  A read
        overlapping several regions is delivered once; the batch is in file order.  Ignored (the whole file is
        returned) when the file has no index."""
        if self.index_path and regions is not None:
            order = {c: i for i, c in enumerate(self.contigs)}
            todo = []  # per contig in file order: sorted, overlapping or touching regions merged
            for c, a, b in sorted((r for r in regions if r[0] in order), key=lambda r: (order[r[0]], r[1], r[2])):
                a = max(0, a)
                if todo and todo[-1][0] == c and a <= todo[-1][1][1]:
                    todo[-1][1][1] = max(todo[-1][1][1], b)
                else:
                    todo.append((c, [a, b]))
            n = len(todo)
            names = (C.c_char_p * max(n, 1))(*[c.encode() for c, _ in todo])
            beg = np.asarray([max(0, a) for _, (a, _b) in todo], np.int32)
            end = np.asarray([b for _, (_a, b) in todo], np.int32)
            h = C.c_void_p()
            _dev._io_check(self._lib.vfkio_bam_open_regions(self.path.encode(), self.index_path.encode(), n, names,
                                                            C.c_void_p(beg.ctypes.data), C.c_void_p(end.ctypes.data), C.byref(h)))
            try:
                return self._fill(h)
            finally:
                self._lib.vfkio_bam_close(h)
        if self._batch is None:
            if self.index_path:  # header-only handle: decode the file now
                h = C.c_void_p()
                _dev._io_check(self._lib.vfkio_bam_open(self.path.encode(), C.byref(h)))
                self._lib.vfkio_bam_close(self._h)
                self._h = h
            self._batch = self._fill(self._h)
            self._lib.vfkio_bam_close(self._h)  # the inflated stream has been copied into the batch: let it go
            self._h = None
        return self._batch

    def close(self):
        if self._h:
            self._lib.vfkio_bam_close(self._h)
            self._h = None
            self._batch = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


# ---------------------------------------------------------------------------------------------
# writer (tests / fixtures)
# ---------------------------------------------------------------------------------------------
def _bgzf_block(data: bytes) -> bytes:
    comp = zlib.compressobj(6, zlib.DEFLATED, -15)
    body = comp.compress(data) + comp.flush()
    bsize = len(body) + 25
    return (b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", bsize) + body +
            struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data)))


_BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def _reg2bin(beg: int, end: int) -> int:
    end -= 1
    for shift, base in ((14, 4681), (17, 585), (20, 73), (23, 9), (26, 1)):
        if beg >> shift == end >> shift:
            return base + (beg >> shift)
    return 0


def write_bam(path: str, records: Sequence[dict], contigs: Sequence[str], lengths: Optional[Sequence[int]] = None,
              sort_order: str = "coordinate", index: bool = False) -> None:
    """records: SAM-view dicts (see ReadBatch.from_records), already in the order to be written.
    index=True also writes `path + ".bai"` (bins + linear index, SAMv1 5.2) for coordinate-sorted records."""
    lengths = list(lengths) if lengths is not None else [1 << 28] * len(contigs)
    text = "@HD\tVN:1.6\tSO:%s\n" % sort_order + "".join("@SQ\tSN:%s\tLN:%d\n" % (c, l) for c, l in zip(contigs, lengths))
    out = bytearray(b"BAM\1" + struct.pack("<i", len(text)) + text.encode() + struct.pack("<i", len(contigs)))
    for c, l in zip(contigs, lengths):
        out += struct.pack("<i", len(c) + 1) + c.encode() + b"\0" + struct.pack("<i", l)
    code = {ch: i for i, ch in enumerate(SEQ_CHARS)}
    placed = []  # (tid, pos, end, bin, offset of the record in the uncompressed stream, offset of its end)
    for r in records:
        cig = r["cigar"]
        if isinstance(cig, str):
            ops, num = [], ""
            for ch in cig:
                if ch.isdigit():
                    num += ch
                else:
                    ops.append(("MIDNSHP=X".index(ch), int(num)))
                    num = ""
        else:
            ops = list(cig)
        name = r["qname"].encode() + b"\0"
        seq = r["seq"]
        qual = r["qual"]
        qual = bytes(ord(ch) - 33 for ch in qual) if isinstance(qual, str) else bytes(qual)
        codes = [code[ch] for ch in seq] + ([0] if len(seq) & 1 else [])
        packed = bytes((codes[i] << 4) | codes[i + 1] for i in range(0, len(codes), 2))
        span = sum(l for op, l in ops if op in (0, 2, 3, 7, 8))
        body = struct.pack("<iiBBHHHiiii", r["tid"], r["pos"], len(name), r["mapq"],
                           _reg2bin(r["pos"], r["pos"] + max(span, 1)) if r["tid"] >= 0 else 4680, len(ops), r["flag"],
                           len(seq), r.get("mtid", -1), r.get("mpos", -1), r.get("isize", 0))
        body += name + b"".join(struct.pack("<I", (l << 4) | op) for op, l in ops) + packed + qual
        if r["tid"] >= 0:
            e = r["pos"] + max(span, 1)
            placed.append((r["tid"], r["pos"], e, _reg2bin(r["pos"], e), len(out), len(out) + 4 + len(body)))
        out += struct.pack("<i", len(body)) + body
    block_at = []  # file offset of every BGZF block
    with open(path, "wb") as fh:
        for i in range(0, len(out), 0xFF00):
            block_at.append(fh.tell())
            fh.write(_bgzf_block(bytes(out[i:i + 0xFF00])))
        block_at.append(fh.tell())
        fh.write(_BGZF_EOF)
    if index:
        def voff(u):
            return (block_at[u // 0xFF00] << 16) | (u % 0xFF00)
        bai = bytearray(b"BAI\1" + struct.pack("<i", len(contigs)))
        for t in range(len(contigs)):
            bins, linear = {}, []
            for tid, pos, e, b, u0, u1 in placed:
                if tid != t:
                    continue
                chunks = bins.setdefault(b, [])
                if chunks and chunks[-1][1] == voff(u0):
                    chunks[-1][1] = voff(u1)
                else:
                    chunks.append([voff(u0), voff(u1)])
                for w in range(pos >> 14, ((e - 1) >> 14) + 1):
                    while len(linear) <= w:
                        linear.append(0)
                    if linear[w] == 0:
                        linear[w] = voff(u0)
            for w in range(1, len(linear)):  # htslib fills the windows no read overlaps with the entry before
                if linear[w] == 0:
                    linear[w] = linear[w - 1]
            bai += struct.pack("<i", len(bins))
            for b in sorted(bins):
                bai += struct.pack("<Ii", b, len(bins[b]))
                for c0, c1 in bins[b]:
                    bai += struct.pack("<QQ", c0, c1)
            bai += struct.pack("<i", len(linear))
            for v in linear:
                bai += struct.pack("<Q", v)
        with open(path + ".bai", "wb") as fh:
            fh.write(bytes(bai))


def write_bam_batch(path: str, b: ReadBatch, lengths: Optional[Sequence[int]] = None, level: int = 1) -> int:
    """A coordinate-sorted ReadBatch as BAM + `.bai` (binning index with one chunk per run of records of a bin,
    linear index).  Read names are synthesised from `qname_id` (mates share it).  For fixtures and host-side
    benchmarks of files too large for the record-dict writer above; returns the size of the uncompressed stream."""
    lengths = list(lengths) if lengths is not None else (list(b.contig_lengths) if b.contig_lengths is not None else [1 << 28] * len(b.contigs))
    text = "@HD\tVN:1.6\tSO:coordinate\n" + "".join("@SQ\tSN:%s\tLN:%d\n" % (c, l) for c, l in zip(b.contigs, lengths))
    out = bytearray(b"BAM\1" + struct.pack("<i", len(text)) + text.encode() + struct.pack("<i", len(b.contigs)))
    for c, l in zip(b.contigs, lengths):
        out += struct.pack("<i", len(c) + 1) + c.encode() + b"\0" + struct.pack("<i", int(l))
    cig, seq, qual = b.cigar.tobytes(), b.seq.tobytes(), b.qual.tobytes()
    rec_at = np.zeros(b.n_reads + 1, np.int64)
    for i in range(b.n_reads):
        rec_at[i] = len(out)
        name = b"q%x\0" % int(b.qname_id[i])
        nc, lq = int(b.n_cigar[i]), int(b.l_qseq[i])
        co, so, qo = int(b.cigar_off[i]) * 4, int(b.seq_off[i]), int(b.qual_off[i])
        body = struct.pack("<iiBBHHHiiii", int(b.tid[i]), int(b.pos[i]), len(name), int(b.mapq[i]), 4680, nc, int(b.flag[i]), lq,
                           int(b.mtid[i]), int(b.mpos[i]), int(b.isize[i]))
        body += name + cig[co:co + 4 * nc] + seq[so:so + (lq + 1) // 2] + qual[qo:qo + lq]
        out += struct.pack("<i", len(body)) + body
    rec_at[b.n_reads] = len(out)
    block_at = []
    with open(path, "wb") as fh:
        for i in range(0, len(out), 0xFF00):
            block_at.append(fh.tell())
            data = bytes(out[i:i + 0xFF00])
            comp = zlib.compressobj(level, zlib.DEFLATED, -15)
            z = comp.compress(data) + comp.flush()
            fh.write(b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", len(z) + 25) + z +
                     struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data)))
        block_at.append(fh.tell())
        fh.write(_BGZF_EOF)
    # .bai (SAMv1 5.2): binning index with one chunk per bin + linear index
    block_at = np.asarray(block_at, np.int64)
    voff = (block_at[rec_at // 0xFF00] << 16) | (rec_at % 0xFF00)
    from .sharding import _ref_span
    end = b.pos.astype(np.int64) + np.maximum(_ref_span(b), 1)
    bai = bytearray(b"BAI\1" + struct.pack("<i", len(b.contigs)))
    for t in range(len(b.contigs)):
        idx = np.nonzero(b.tid == t)[0]
        if len(idx) == 0:
            bai += struct.pack("<ii", 0, 0)
            continue
        beg_t, end_t = b.pos[idx].astype(np.int64), end[idx] - 1
        bins = np.zeros(len(idx), np.int64)
        done = np.zeros(len(idx), bool)
        for shift, base in ((14, 4681), (17, 585), (20, 73), (23, 9), (26, 1)):  # reg2bin, SAMv1 5.3
            hit = ~done & ((beg_t >> shift) == (end_t >> shift))
            bins[hit] = base + (beg_t[hit] >> shift)
            done |= hit
        # chunks: runs of consecutive records filed under the same bin
        starts = np.nonzero(np.concatenate([[True], bins[1:] != bins[:-1]]))[0]
        stops = np.concatenate([starts[1:], [len(idx)]])
        order = np.argsort(bins[starts], kind="stable")
        ub, first_of = np.unique(bins[starts][order], return_index=True)
        bai += struct.pack("<i", len(ub))
        for k, bn in enumerate(ub):
            runs = order[first_of[k]:(first_of[k + 1] if k + 1 < len(ub) else len(order))]
            bai += struct.pack("<Ii", int(bn), len(runs))
            for rj in runs:
                bai += struct.pack("<QQ", int(voff[idx[starts[rj]]]), int(voff[idx[stops[rj] - 1] + 1]))
        n_win = int((end[idx].max() - 1) >> 14) + 1
        linear = np.zeros(n_win, np.int64)
        first = np.full(n_win, -1, np.int64)
        w0, w1 = b.pos[idx] >> 14, (end[idx] - 1) >> 14
        for k in range(int((w1 - w0).max()) + 1):  # reads span few windows
            w = np.minimum(w0 + k, w1)
            order = np.argsort(w, kind="stable")
            ww, ii = w[order], idx[order]
            keep = np.concatenate([[True], ww[1:] != ww[:-1]])
            cand = np.full(n_win, np.iinfo(np.int64).max)
            cand[ww[keep]] = ii[keep]
            first = np.where((first < 0) | (cand < first), np.where(cand == np.iinfo(np.int64).max, first, cand), first)
        linear = np.where(first >= 0, voff[np.maximum(first, 0)], 0)
        for w in range(1, n_win):
            if linear[w] == 0:
                linear[w] = linear[w - 1]
        bai += struct.pack("<i", n_win) + linear.astype("<u8").tobytes()
    with open(path + ".bai", "wb") as fh:
        fh.write(bytes(bai))
    return len(out)
