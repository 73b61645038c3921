"""ORACLE (test infrastructure, NOT product code) -- pure-Python restatement of the
pysam 0.21.0 / htslib 1.17 pileup engine as vafator drives it.

PARITY STATUS OF THIS FILE: **parity unpinned**.  pysam/htslib are third-party
code that is absent from /root/reference and not installable here (no network);
the reference's own known-answer tests for this layer
(vafator/tests/test_pileups.py:23-115, vafator/tests/test_annotator.py:184-279)
need a BAM that is missing from the checkout (.MISSING_LARGE_BLOBS:1-4).  What
follows restates the *published* algorithm of

  * pysam 0.21.0  libcalignmentfile.pyx  IteratorColumnRegion / __advance_samtools
  * pysam 0.21.0  libcalignedsegment.pyx PileupColumn.pileups / pileup_base_qual_skip
                  PileupRead, AlignedSegment.query_sequence / query_qualities / query
  * htslib 1.17   sam.c  bam_plp_push / bam_plp64_next / resolve_cigar2 /
                  overlap_push / overlap_remove / tweak_overlap_quality
                  khash.h __ac_X31_hash_string / __ac_Wang_hash

from memory, anchored on the reference's call site
vafator/pileups.py:71-80 (the keyword arguments) and on the attribute accesses
at vafator/pileups.py:199-212, 263-294, 329-351.

The objects here quack like the pysam objects those call sites touch, so the
*unmodified* reference functions in /root/reference/vafator/pileups.py can be
executed on top of this file (oracle/gen_golden.py does exactly that to make
tests/golden/*.json).  It is a streaming state machine (push / next / incremental
CIGAR state) on purpose: oracle/pileup_oracle.c and the CUDA kernels evaluate
every (variant, read) pair directly, so agreement between the two is agreement
between two independent formulations.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import
this module.
"""
from __future__ import annotations

from typing import Iterable, Iterator, List, Optional, Sequence

# BAM CIGAR op codes (SAM spec section 4.2): MIDNSHP=X
M, I, D, N, S, H, P, EQ, X = range(9)
CIGAR_CHARS = "MIDNSHP=X"
SEQ_CHARS = "=ACMGRSVTWYHKDBN"  # 4-bit BAM base codes

FUNMAP, FPAIRED, FPROPER, FMUNMAP = 0x4, 0x1, 0x2, 0x8
FSECONDARY, FQCFAIL, FDUP, FSUPP = 0x100, 0x200, 0x400, 0x800
DEFAULT_FLAG_FILTER = FUNMAP | FSECONDARY | FQCFAIL | FDUP  # pysam IteratorColumn default

_REF_OPS = (M, D, N, EQ, X)
_ALN_OPS = (M, EQ, X)


def x31_hash(name: str) -> int:
    """khash.h __ac_X31_hash_string (uint32)."""
    b = name.encode()
    if not b:
        return 0
    h = b[0]
    for c in b[1:]:
        h = ((h << 5) - h + c) & 0xFFFFFFFF
    return h


def wang_hash(key: int) -> int:
    """khash.h __ac_Wang_hash (uint32)."""
    m = 0xFFFFFFFF
    key = (key + (~(key << 15) & m)) & m
    key ^= key >> 10
    key = (key + (key << 3)) & m
    key ^= key >> 6
    key = (key + (~(key << 11) & m)) & m
    key ^= key >> 16
    return key


def parse_cigar(text: str):
    out, num = [], ""
    for ch in text:
        if ch.isdigit():
            num += ch
        else:
            out.append((CIGAR_CHARS.index(ch), int(num)))
            num = ""
    return out


class Rec:
    """One BAM record (what sam_itr_next hands to the pileup engine)."""

    __slots__ = ("qname", "name_key", "x31", "flag", "tid", "pos", "mapq", "cigar",
                 "seq", "qual", "mtid", "mpos", "isize", "index")

    def __init__(self, qname, flag, tid, pos, mapq, cigar, seq, qual,
                 mtid=-1, mpos=-1, isize=0, name_key=None, x31=None, index=-1):
        self.qname = qname
        self.name_key = qname if name_key is None else name_key
        self.x31 = x31_hash(qname) if x31 is None else x31
        self.flag, self.tid, self.pos, self.mapq = flag, tid, pos, mapq
        self.cigar = parse_cigar(cigar) if isinstance(cigar, str) else list(cigar)
        self.seq = seq
        self.qual = bytes(qual)
        self.mtid, self.mpos, self.isize = mtid, mpos, isize
        self.index = index

    @property
    def l_qseq(self):
        return len(self.seq)

    def rlen(self):
        """bam_cigar2rlen."""
        return sum(l for op, l in self.cigar if op in _REF_OPS)

    def endpos(self):
        """bam_endpos: zero-length alignments count as one base."""
        r = self.rlen()
        return self.pos + (r if r > 0 else 1)


# --------------------------------------------------------------------------
# pysam-facing objects
# --------------------------------------------------------------------------
class AlignedSegment:
    """The bam_dup1 copy held by a pileup node: qualities may have been tweaked."""

    def __init__(self, rec: Rec):
        self.rec = rec
        self.qual = bytearray(rec.qual)  # tweak_overlap_quality mutates this copy

    # attributes touched by vafator/pileups.py
    @property
    def mapping_quality(self):
        return self.rec.mapq

    @property
    def reference_start(self):
        return self.rec.pos

    @property
    def cigartuples(self):
        return list(self.rec.cigar)

    @property
    def query_sequence(self):
        return self.rec.seq

    @property
    def query_qualities(self):
        return self.qual

    @property
    def query(self):
        """pysam `query` == query_alignment_sequence: soft clips trimmed at both ends
        (getQueryStart / getQueryEnd)."""
        cig = self.rec.cigar
        start = 0
        for op, l in cig:
            if op == H:
                continue
            if op == S:
                start += l
            else:
                break
        end = self.rec.l_qseq
        for k in range(len(cig) - 1, 0, -1):
            op, l = cig[k]
            if op == H:
                continue
            if op == S:
                end -= l
            else:
                break
        return self.rec.seq[start:end]


class PileupRead:
    def __init__(self, aln, qpos, indel, is_del, is_refskip):
        self.alignment = aln
        self._qpos = qpos
        self.indel = indel
        self.is_del = is_del
        self.is_refskip = is_refskip

    @property
    def query_position(self):
        return None if (self.is_del or self.is_refskip) else self._qpos

    @property
    def query_position_or_next(self):
        return self._qpos


class PileupColumn:
    def __init__(self, tid, pos, plp, min_base_quality):
        self.reference_id = tid
        self.reference_pos = pos
        self._plp = plp
        self._minq = min_base_quality

    @property
    def pileups(self):
        """PileupColumn.pileups: every bam_pileup1_t except those failing
        pileup_base_qual_skip (quality at qpos, 0 when qpos >= l_qseq)."""
        out = []
        for node, qpos, indel, is_del, is_refskip in self._plp:
            q = node.aln.qual[qpos] if qpos < node.aln.rec.l_qseq else 0
            if q < self._minq:
                continue
            out.append(PileupRead(node.aln, qpos, indel, is_del, is_refskip))
        return out


# --------------------------------------------------------------------------
# htslib pileup engine
# --------------------------------------------------------------------------
class _Node:
    __slots__ = ("aln", "beg", "end", "k", "x", "y")

    def __init__(self, rec: Rec):
        self.aln = AlignedSegment(rec)
        self.beg = rec.pos
        self.end = rec.pos + rec.rlen()  # raw rlen, not bam_endpos
        self.k = -1  # cstate_t: CIGAR op index, its reference start x, its query start y
        self.x = 0
        self.y = 0


def _resolve_cigar2(node: _Node, pos: int):
    """sam.c resolve_cigar2 -> (qpos, indel, is_del, is_refskip)."""
    cig = node.aln.rec.cigar
    n = len(cig)
    if node.k == -1:
        if n == 1:
            if cig[0][0] in _ALN_OPS:
                node.k, node.x, node.y = 0, node.beg, 0
        else:
            node.x, node.y = node.beg, 0
            k = 0
            while k < n:
                op, l = cig[k]
                if op in _REF_OPS:
                    break
                if op in (I, S):
                    node.y += l
                k += 1
            node.k = k
    else:
        op, l = cig[node.k]
        if pos - node.x >= l:
            if op in _ALN_OPS:
                node.y += l
            node.x += l
            k = node.k + 1
            while k < n:
                op2, l2 = cig[k]
                if op2 in _REF_OPS:
                    break
                if op2 in (I, S):
                    node.y += l2
                k += 1
            node.k = k
    op, l = cig[node.k]
    indel = 0
    if node.x + l - 1 == pos and node.k + 1 < n:
        op2, l2 = cig[node.k + 1]
        if op2 == D and op != D:
            indel = -l2
            for k in range(node.k + 2, n):
                o, ll = cig[k]
                if o == D:
                    indel -= ll
                else:
                    break
        elif op2 == I:
            indel = l2
            for k in range(node.k + 2, n):
                o, ll = cig[k]
                if o == I:
                    indel += ll
                elif o != P:
                    break
        elif op2 == P and node.k + 2 < n:
            l3 = 0
            for k in range(node.k + 2, n):
                o, ll = cig[k]
                if o == I:
                    l3 += ll
                elif o in _REF_OPS:
                    break
            if l3 > 0:
                indel = l3
    if op in _ALN_OPS:
        return node.y + (pos - node.x), indel, False, False
    return node.y, indel, True, op == N


def _aligned_pairs(rec: Rec):
    """reference position -> query index for every M/=/X base."""
    out = {}
    r, q = rec.pos, 0
    for op, l in rec.cigar:
        if op in _ALN_OPS:
            for i in range(l):
                out[r + i] = q + i
            r += l
            q += l
        elif op in (D, N):
            r += l
        elif op in (I, S):
            q += l
    return out


def _tweak_overlap_quality(a: AlignedSegment, b: AlignedSegment):
    """sam.c tweak_overlap_quality (htslib >= 1.13 behaviour): for every reference
    position >= b.pos where BOTH mates are inside an M/=/X op."""
    pa, pb = _aligned_pairs(a.rec), _aligned_pairs(b.rec)
    amul = 1 if (wang_hash(a.rec.x31) & 1) else 0
    bmul = 1 - amul
    for r in sorted(pb):
        if r < b.rec.pos or r not in pa:
            continue
        ia, ib = pa[r], pb[r]
        if ia >= a.rec.l_qseq or ib >= b.rec.l_qseq:
            return
        qa, qb = a.qual[ia], b.qual[ib]
        if a.rec.seq[ia] == b.rec.seq[ib]:
            q = min(200, qa + qb)
            a.qual[ia] = amul * q
            b.qual[ib] = bmul * q
        elif qa > qb:
            a.qual[ia] = int(0.8 * qa)
            b.qual[ib] = 0
        elif qa < qb:
            b.qual[ib] = int(0.8 * qb)
            a.qual[ia] = 0
        else:
            a.qual[ia] = int(amul * 0.8 * qa)
            b.qual[ib] = int(bmul * 0.8 * qb)


class _Plp:
    """bam_plp_t with the overlap hash enabled (pysam ignore_overlaps=True)."""

    def __init__(self, maxcnt: int, ignore_overlaps: bool = True):
        self.nodes: List[_Node] = []
        self.tid, self.pos = 0, 0
        self.max_tid, self.max_pos = -1, -1
        self.is_eof = False
        self.maxcnt = maxcnt
        self.overlaps = {} if ignore_overlaps else None

    def _overlap_remove(self, rec: Rec):
        if self.overlaps is not None:
            self.overlaps.pop(rec.name_key, None)

    def _overlap_push(self, node: _Node):
        if self.overlaps is None:
            return
        r = node.aln.rec
        if (r.flag & FMUNMAP) or not (r.flag & FPROPER):
            return
        if (r.mtid >= 0 and r.tid != r.mtid) or (abs(r.isize) >= 2 * r.l_qseq and r.mpos >= node.end):
            return
        other = self.overlaps.get(r.name_key)
        if other is None:
            if r.mpos >= r.pos or ((r.flag & FPAIRED) and r.mpos == -1):
                self.overlaps[r.name_key] = node
        else:
            _tweak_overlap_quality(other.aln, node.aln)
            del self.overlaps[r.name_key]

    def push(self, rec: Optional[Rec]):
        if rec is None:
            self.is_eof = True
            return
        if rec.tid < 0 or (rec.flag & FUNMAP):
            self._overlap_remove(rec)
            return
        if self.tid == rec.tid and self.pos == rec.pos and len(self.nodes) > self.maxcnt:
            self._overlap_remove(rec)
            return
        node = _Node(rec)
        assert not (rec.tid < self.max_tid or (rec.tid == self.max_tid and node.beg < self.max_pos)), "unsorted input"
        self.max_tid, self.max_pos = rec.tid, node.beg
        if node.end > self.pos or rec.tid > self.tid:
            self._overlap_push(node)
            self.nodes.append(node)

    def next(self):
        """bam_plp64_next: returns (tid, pos, [(node,qpos,indel,is_del,is_refskip)...]) or None
        when more input is needed / the stream is exhausted."""
        if self.is_eof and not self.nodes:
            return None
        while self.is_eof or self.max_tid > self.tid or (self.max_tid == self.tid and self.max_pos > self.pos):
            plp, keep = [], []
            for nd in self.nodes:
                t = nd.aln.rec.tid
                if t < self.tid or (t == self.tid and nd.end <= self.pos):
                    self._overlap_remove(nd.aln.rec)
                    continue
                keep.append(nd)
                if t == self.tid and nd.beg <= self.pos:
                    plp.append((nd,) + _resolve_cigar2(nd, self.pos))
            self.nodes = keep
            tid, pos = self.tid, self.pos
            if self.nodes:
                head = self.nodes[0]
                if self.tid < head.aln.rec.tid:
                    self.tid, self.pos = head.aln.rec.tid, head.beg
                elif self.pos < head.beg:
                    self.pos = head.beg
                else:
                    self.pos += 1
            else:
                self.pos += 1
            if plp:
                return tid, pos, plp
            if self.is_eof and not self.nodes:
                break
        return None


class AlignmentFile:
    """Stand-in for pysam.AlignmentFile over an in-memory, coordinate-sorted record list."""

    def __init__(self, recs: Sequence[Rec], contigs: Sequence[str]):
        self.recs = list(recs)
        self.contigs = list(contigs)

    def fetch_region(self, tid: int, start: int, stop: int) -> Iterable[Rec]:
        """sam_itr_queryi: mapped records of `tid` overlapping [start, stop)."""
        for r in self.recs:
            if r.tid == tid and r.pos < stop and r.endpos() > start:
                yield r

    def pileup(self, contig, start, stop, truncate=True, max_depth=8000, min_base_quality=13,
               min_mapping_quality=0, stepper="samtools", flag_filter=DEFAULT_FLAG_FILTER,
               ignore_orphans=True, ignore_overlaps=True) -> Iterator[PileupColumn]:
        assert stepper == "samtools"
        if contig not in self.contigs:
            raise ValueError("invalid contig `%s`" % contig)
        tid = self.contigs.index(contig)
        plp = _Plp(max_depth, ignore_overlaps)

        def advance():  # pysam __advance_samtools (BAQ inactive: no fastafile)
            for rec in self.fetch_region(tid, start, stop):
                if rec.flag & flag_filter:
                    continue
                if rec.mapq < min_mapping_quality:
                    continue
                if ignore_orphans and (rec.flag & FPAIRED) and not (rec.flag & FPROPER):
                    continue
                yield rec

        src = advance()
        while True:
            col = plp.next()
            while col is None and not plp.is_eof:
                plp.push(next(src, None))
                col = plp.next()
            if col is None:
                return
            ctid, cpos, entries = col
            if truncate:
                if start > cpos:
                    continue
                if cpos >= stop:
                    return
            yield PileupColumn(ctid, cpos, entries, min_base_quality)


def recs_from_soa(b, names=None) -> List[Rec]:
    """Canonical SoA read batch (dict of numpy arrays, see vafator_b200/batch.py docstring)
    -> list of Rec.  Kept dependency-free: the oracle never imports the product."""
    out = []
    n = len(b["pos"])
    for i in range(n):
        co, nc = int(b["cigar_off"][i]), int(b["n_cigar"][i])
        cig = [(int(c) & 0xF, int(c) >> 4) for c in b["cigar"][co:co + nc]]
        lq = int(b["l_qseq"][i])
        so = int(b["seq_off"][i])
        packed = b["seq"][so:so + (lq + 1) // 2]
        seq = "".join(SEQ_CHARS[(int(packed[j >> 1]) >> (4 if (j & 1) == 0 else 0)) & 0xF] for j in range(lq))
        qo = int(b["qual_off"][i])
        out.append(Rec(
            qname=None if names is None else names[i],
            flag=int(b["flag"][i]), tid=int(b["tid"][i]), pos=int(b["pos"][i]), mapq=int(b["mapq"][i]),
            cigar=cig, seq=seq, qual=bytes(b["qual"][qo:qo + lq]),
            mtid=int(b["mtid"][i]), mpos=int(b["mpos"][i]), isize=int(b["isize"][i]),
            name_key=(int(b["qname_id"][i]), int(b["qname_hash"][i])), x31=int(b["qname_hash"][i]), index=i))
    return out
