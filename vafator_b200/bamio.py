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
    return lib


class BamFile:
    """A decoded BAM: `.contigs`, `.lengths`, `.read_batch()`."""

    def __init__(self, path: str, mode: str = "r"):
        if not os.path.exists(path):
            raise FileNotFoundError("could not open alignment file `%s`: No such file or directory" % path)
        self.path = path
        self._lib = _lib()
        h = C.c_void_p()
        _dev._io_check(self._lib.vfkio_bam_open(path.encode(), C.byref(h)))
        self._h = h
        n = self._lib.vfkio_bam_n_contigs(h)
        self.contigs: List[str] = [self._lib.vfkio_bam_contig_name(h, i).decode() for i in range(n)]
        self.lengths = np.asarray([self._lib.vfkio_bam_contig_length(h, i) for i in range(n)], np.int64)
        self.header_text = self._lib.vfkio_bam_header_text(h).decode(errors="replace")
        self._batch: Optional[ReadBatch] = None

    def read_batch(self) -> ReadBatch:
        if self._batch is None:
            if not self._lib.vfkio_bam_is_sorted(self._h):
                raise ValueError("%s is not coordinate sorted" % self.path)
            plan = _PlanC()
            _dev._io_check(self._lib.vfkio_bam_plan(self._h, C.byref(plan)))
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
            _dev._io_check(self._lib.vfkio_bam_fill(self._h, C.byref(s)))
            arr["cigar"], arr["seq"], arr["qual"] = arr["cigar"][:plan.n_cigar], arr["seq"][:plan.n_seq_bytes], arr["qual"][:plan.n_qual_bytes]
            self._batch = ReadBatch(contigs=list(self.contigs), contig_lengths=self.lengths, **arr)
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
              sort_order: str = "coordinate") -> None:
    """records: SAM-view dicts (see ReadBatch.from_records), already in the order to be written."""
    lengths = list(lengths) if lengths is not None else [1 << 28] * len(contigs)
    text = "@HD\tVN:1.6\tSO:%s\n" % sort_order + "".join("@SQ\tSN:%s\tLN:%d\n" % (c, l) for c, l in zip(contigs, lengths))
    out = bytearray(b"BAM\1" + struct.pack("<i", len(text)) + text.encode() + struct.pack("<i", len(contigs)))
    for c, l in zip(contigs, lengths):
        out += struct.pack("<i", len(c) + 1) + c.encode() + b"\0" + struct.pack("<i", l)
    code = {ch: i for i, ch in enumerate(SEQ_CHARS)}
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
        out += struct.pack("<i", len(body)) + body
    with open(path, "wb") as fh:
        for i in range(0, len(out), 0xFF00):
            fh.write(_bgzf_block(bytes(out[i:i + 0xFF00])))
        fh.write(_BGZF_EOF)
