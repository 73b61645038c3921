"""Columnar read / variant batches -- the data that crosses the C-ABI.

`ReadBatch` is the canonical, decoder-facing structure-of-arrays form of a set of BAM records
(coordinate sorted, exactly the fields of the BAM record the pileup path reads):

    tid[n] i32, pos[n] i32 (0-based leftmost), flag[n] u16, mapq[n] u8, l_qseq[n] i32,
    n_cigar[n] i32, cigar_off[n] i64 -> cigar[] u32 (BAM encoding len<<4|op),
    seq_off[n] i64 -> seq[] u8 (BAM 4-bit codes, high nibble first, each read byte aligned),
    qual_off[n] i64 -> qual[] u8,
    mtid[n] i32, mpos[n] i32, isize[n] i32,
    qname_id[n] u64 (any 64-bit id that is equal iff the read names are equal),
    qname_hash[n] u32 (khash X31 of the read name; htslib's overlap tie-break hashes it)

It replaces what pysam.AlignmentFile hands to htslib's pileup engine at
vafator/pileups.py:71-80, and it is what crosses the C ABI (include/vfk.h vfk_read_batch): the kernels derive
everything else -- reference spans, the position index, CIGAR shapes, mate pairing -- on the device.

`VariantUnits` is the device-facing form of the VCF records of vafator/pileups.py:133-151:
one *unit* per (record, ALT allele) that needs its own column evaluation.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

SEQ_CHARS = "=ACMGRSVTWYHKDBN"
_CODE_OF = {c: i for i, c in enumerate(SEQ_CHARS)}
CIGAR_CHARS = "MIDNSHP=X"

KIND_SNV, KIND_INS, KIND_DEL = 0, 1, 2
CODE_NONE = 0xFF  # allele text that no decoded read base / slice can ever equal


def x31_hash(name: bytes) -> int:
    """khash.h __ac_X31_hash_string over the NUL-terminated read name."""
    if not name:
        return 0
    h = name[0]
    for c in name[1:]:
        h = ((h << 5) - h + c) & 0xFFFFFFFF
    return h


def fnv1a64(name: bytes) -> int:
    h = 0xCBF29CE484222325
    for c in name:
        h = ((h ^ c) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


_FIELDS = ("tid", "pos", "flag", "mapq", "l_qseq", "n_cigar", "cigar_off", "cigar", "seq_off", "seq",
           "qual_off", "qual", "mtid", "mpos", "isize", "qname_id", "qname_hash")


@dataclass
class ReadBatch:
    contigs: List[str]
    tid: np.ndarray
    pos: np.ndarray
    flag: np.ndarray
    mapq: np.ndarray
    l_qseq: np.ndarray
    n_cigar: np.ndarray
    cigar_off: np.ndarray
    cigar: np.ndarray
    seq_off: np.ndarray
    seq: np.ndarray
    qual_off: np.ndarray
    qual: np.ndarray
    mtid: np.ndarray
    mpos: np.ndarray
    isize: np.ndarray
    qname_id: np.ndarray
    qname_hash: np.ndarray
    contig_lengths: Optional[np.ndarray] = None
    pinned: bool = False  # the arrays already live in page-locked memory (bamio / synth with pinned=True)
    contiguous: bool = False  # the CIGAR / sequence / quality pools are contiguous in read order (a decoder that appends record by
    #                           record produces that): the three offset arrays are prefix sums of the lengths and need not travel

    @property
    def n_reads(self) -> int:
        return int(self.pos.shape[0])

    # ---- batch metadata a decoder has at hand when it finishes a batch (computed once, cached) ----
    def _meta(self):
        m = self.__dict__.get("_meta_cache")
        if m is None:
            tid = self.tid
            n = self.n_reads
            n_placed = int(np.searchsorted(-(tid >= 0).astype(np.int8), 0, side="left")) if n else 0  # placed reads come first
            if n and (np.any(tid[:n_placed] < 0) or np.any(tid[n_placed:] >= 0)):
                raise ValueError("unplaced reads must come last in a coordinate-sorted batch")
            t = tid[:n_placed]
            if n_placed and np.any(t[1:] < t[:-1]):
                raise ValueError("read batch is not coordinate sorted")
            if n_placed and int(t[-1]) >= len(self.contigs):
                raise ValueError("read tid outside the contig table")
            contig_off = np.searchsorted(t, np.arange(len(self.contigs) + 1), side="left").astype(np.int64)
            max_l = int(self.l_qseq[:n_placed].max()) if n_placed else 0
            m = self.__dict__["_meta_cache"] = (n_placed, contig_off, max_l)
        return m

    @property
    def n_placed(self) -> int:
        return self._meta()[0]

    @property
    def contig_off(self) -> np.ndarray:
        """[n_contigs + 1]: the reads of contig t are [off[t], off[t+1])."""
        return self._meta()[1]

    @property
    def max_l_qseq(self) -> int:
        return self._meta()[2]

    def as_dict(self) -> Dict[str, np.ndarray]:
        return {k: getattr(self, k) for k in _FIELDS}

    def check_sorted(self) -> None:
        if self.n_reads and np.any(self.pos[self.tid >= 0] < 0):
            raise ValueError("placed read with a negative start position")
        if self.n_reads < 2:
            return
        key = self.tid.astype(np.int64) * (1 << 32) + self.pos.astype(np.int64)
        mapped = self.tid >= 0
        k = key[mapped]
        if np.any(k[1:] < k[:-1]):
            raise ValueError("read batch is not coordinate sorted")
        if np.any(~mapped[:-1] & mapped[1:]):
            raise ValueError("unplaced reads must come last in a coordinate-sorted batch")
        if np.any(self.pos[mapped] < 0):
            raise ValueError("placed read with a negative start position")

    @staticmethod
    def from_records(records: Sequence[dict], contigs: Sequence[str]) -> "ReadBatch":
        """records: dicts with qname, flag, tid, pos, mapq, cigar ('10M2I..' or [(op,len)]), seq,
        qual (phred+33 text or ints), mtid, mpos, isize -- the SAM view of a record."""
        n = len(records)
        tid = np.zeros(n, np.int32)
        pos = np.zeros(n, np.int32)
        flag = np.zeros(n, np.uint16)
        mapq = np.zeros(n, np.uint8)
        l_qseq = np.zeros(n, np.int32)
        n_cigar = np.zeros(n, np.int32)
        cigar_off = np.zeros(n, np.int64)
        seq_off = np.zeros(n, np.int64)
        qual_off = np.zeros(n, np.int64)
        mtid = np.zeros(n, np.int32)
        mpos = np.zeros(n, np.int32)
        isize = np.zeros(n, np.int32)
        qid = np.zeros(n, np.uint64)
        qh = np.zeros(n, np.uint32)
        cig: List[int] = []
        seq = bytearray()
        qual = bytearray()
        for i, r in enumerate(records):
            tid[i], pos[i], flag[i], mapq[i] = r["tid"], r["pos"], r["flag"], r["mapq"]
            if tid[i] >= 0 and pos[i] < 0:
                raise ValueError("placed read %r with a negative start position" % (r.get("qname"),))
            mtid[i], mpos[i], isize[i] = r.get("mtid", -1), r.get("mpos", -1), r.get("isize", 0)
            c = r["cigar"]
            if isinstance(c, str):
                ops, num = [], ""
                for ch in c:
                    if ch.isdigit():
                        num += ch
                    else:
                        ops.append((CIGAR_CHARS.index(ch), int(num)))
                        num = ""
            else:
                ops = list(c)
            cigar_off[i], n_cigar[i] = len(cig), len(ops)
            cig.extend((ln << 4) | op for op, ln in ops)
            s = r["seq"]
            l_qseq[i] = len(s)
            seq_off[i] = len(seq)
            codes = [_CODE_OF[ch] for ch in s]
            if len(codes) & 1:
                codes.append(0)
            seq.extend((codes[j] << 4) | codes[j + 1] for j in range(0, len(codes), 2))
            q = r["qual"]
            qual_off[i] = len(qual)
            qual.extend(bytes(ord(ch) - 33 for ch in q) if isinstance(q, str) else bytes(q))
            name = r["qname"].encode() if isinstance(r["qname"], str) else bytes(r["qname"])
            qid[i], qh[i] = fnv1a64(name), x31_hash(name)
        return ReadBatch(list(contigs), tid, pos, flag, mapq, l_qseq, n_cigar, cigar_off,
                         np.asarray(cig, np.uint32), seq_off, np.frombuffer(bytes(seq), np.uint8).copy(),
                         qual_off, np.frombuffer(bytes(qual), np.uint8).copy(), mtid, mpos, isize, qid, qh)


def variant_kind(ref: str, alt0: str) -> Optional[int]:
    """vafator/pileup_utils.py:95-119 -- the type is decided on ALT[0] only."""
    if len(ref) == 1 and len(alt0) == 1:
        return KIND_SNV
    if len(ref) == 1 and len(alt0) > 1:
        return KIND_INS
    if len(alt0) == 1 and len(ref) > 1:
        return KIND_DEL
    return None


@dataclass
class VariantUnits:
    """Device-facing variant batch.  unit u evaluates the column of record `rec[u]` for ALT
    allele `alt_index[u]`; `unit_of[(record, alt)]` maps back (absent: no column evaluation is
    needed, the host fills the values the reference would print)."""
    tid: np.ndarray        # i32
    pos: np.ndarray        # i32, 0-based column (POS-1)
    meta: np.ndarray       # u32: kind | ref_code<<8 | alt_code<<16
    length: np.ndarray     # i32: inserted / deleted length
    seq_off: np.ndarray    # u32 -> ins_seq
    ins_seq: np.ndarray    # u8: one BAM 4-bit code per inserted base, CODE_NONE if unmatchable
    rec: np.ndarray        # i32: index of the VCF record
    alt_index: np.ndarray  # i32
    unit_of: Dict[Tuple[int, int], int] = field(default_factory=dict)

    @property
    def n_units(self) -> int:
        return int(self.pos.shape[0])


def _code(allele: str) -> int:
    return _CODE_OF.get(allele, CODE_NONE) if len(allele) == 1 else CODE_NONE


def build_units(records: Sequence[Tuple[str, int, str, Sequence[str]]], contigs: Sequence[str]) -> VariantUnits:
    """records: (CHROM, POS 1-based, REF, [ALT...]).  Raises ValueError for a contig the BAM
    header does not know, as pysam does at vafator/pileups.py:71."""
    tid_of = {c: i for i, c in enumerate(contigs)}
    tid, pos, meta, length, seq_off, rec, alt_index = [], [], [], [], [], [], []
    ins = bytearray()
    unit_of: Dict[Tuple[int, int], int] = {}
    for i, (chrom, p1, ref, alts) in enumerate(records):
        if not alts:
            continue
        kind = variant_kind(ref, alts[0])
        if kind is None:
            continue
        if chrom not in tid_of:
            raise ValueError("invalid contig `%s`" % chrom)
        n_alt = len(alts) if kind == KIND_SNV else 1  # indels are only ever counted for ALT[0]
        for j in range(n_alt):
            unit_of[(i, j)] = len(pos)
            tid.append(tid_of[chrom])
            pos.append(p1 - 1)
            rec.append(i)
            alt_index.append(j)
            if kind == KIND_SNV:
                meta.append(kind | (_code(ref) << 8) | (_code(alts[j]) << 16))
                length.append(0)
                seq_off.append(0)
            elif kind == KIND_INS:
                meta.append(kind | (CODE_NONE << 8) | (CODE_NONE << 16))
                length.append(len(alts[0]) - len(ref))
                seq_off.append(len(ins))
                ins.extend(_CODE_OF.get(ch, CODE_NONE) for ch in alts[0][1:])
            else:
                meta.append(kind | (CODE_NONE << 8) | (CODE_NONE << 16))
                length.append(len(ref) - len(alts[0]))
                seq_off.append(0)
    return VariantUnits(np.asarray(tid, np.int32), np.asarray(pos, np.int32), np.asarray(meta, np.uint32),
                        np.asarray(length, np.int32), np.asarray(seq_off, np.uint32),
                        np.frombuffer(bytes(ins) + b"\0", np.uint8).copy(), np.asarray(rec, np.int32),
                        np.asarray(alt_index, np.int32), unit_of)
