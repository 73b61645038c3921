"""Which reads of an indexed BAM a chromosome group of variants needs.

The reference piles up EVERY column of [first.POS-1, last.POS) of a chromosome group and throws away the
columns that hold no variant (vafator/pileups.py:137-151) -- for a somatic VCF with a few thousand records that
is the whole genome's worth of reads for a few thousand columns.  A column's metrics depend on far fewer reads:

  * the reads overlapping the column (their bases, qualities, CIGARs);
  * their overlapping mates, for htslib's overlap quality adjustment -- a mate that matters shares an aligned
    position with its partner at or right after the column, so it overlaps the window around the column too;
  * the FIRST read the reference pushes after the column: `bam_plp_auto` reads one record past a column before
    emitting it, and if that record is the mate of a read hanging in a deletion / reference skip at the column,
    its arrival has already adjusted the quality the base-quality filter looks at (oracle/htslib_pileup.py).

So the group is fetched as windows [col - pad, col + 1 + pad) merged when they are close, clipped to the span
the reference fetches, and `missing_reads` then proves, column by column, that the first pushed read after the
column is among the fetched ones -- or names the window to extend.  Dense variant sets merge into one stretch
from their first to their last window (for a whole group: nearly the reference's own span).  Not reproduced: three or more alignments sharing one read name (supplementary
alignments) whose extra record lies outside every window while both mates of the pair lie inside -- the same
class of effect DESIGN.md section 7 lists for the fetch-region edge.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

from .batch import ReadBatch
from .sharding import _ref_span

PAD = 1000         # bases fetched on either side of a column
MERGE_GAP = 8192   # windows closer than this are one region: decoding the gap is cheaper than another seek
DENSE = 0.5        # windows covering more than this share of the stretch they span: fetch the stretch as one window
MAX_ROUNDS = 4


def plan_windows(cols: Sequence[int], span_lo: int, span_hi: int, pad: int = PAD, merge_gap: int = MERGE_GAP) -> List[List[int]]:
    """cols: 0-based columns inside [span_lo, span_hi).  Returns sorted disjoint windows [a, b)."""
    out: List[List[int]] = []
    for c in sorted(set(int(c) for c in cols)):
        a, b = max(span_lo, c - pad), min(span_hi, c + 1 + pad)
        if b <= a:
            continue
        if out and a - out[-1][1] < merge_gap:
            out[-1][1] = max(out[-1][1], b)
        else:
            out.append([a, b])
    return out


def missing_reads(batch: ReadBatch, passes: np.ndarray, cols: Sequence[int], windows: List[List[int]], span_hi: int,
                  tid: int) -> List[Tuple[int, int]]:
    """For every column: is the first read the reference pushes after it (the first passing read starting beyond
    the column, before `span_hi`) known from the fetched reads?  It is when a passing read starts inside the rest
    of the column's window, or the window ends where the reference stops reading.  Otherwise it only matters if
    it could be the mate of a read overlapping the column, i.e. if such a read names a mate position at or beyond
    the window end: returns [(window index, end needed)] for those."""
    # (a region fetch of one contig holds reads of that contig only: plain views then, no masked copies)
    all_here = bool(len(batch.tid)) and int(batch.tid[0]) == tid and int(batch.tid[-1]) == tid
    sel = slice(None) if all_here else batch.tid == tid
    pos = batch.pos[sel]
    n = len(pos)
    if n == 0 or not len(cols):
        return []
    ok = passes[sel]
    # position of the first passing read at or after index j (sentinel: beyond everything)
    big = np.int64(1) << 40
    nxt_pos = np.where(ok, pos.astype(np.int64), big)
    nxt_pos = np.minimum.accumulate(nxt_pos[::-1])[::-1]
    nxt_pos = np.concatenate([nxt_pos, [big]])
    cols = np.asarray(sorted(set(int(c) for c in cols)), np.int64)
    w_lo = np.asarray([w[0] for w in windows], np.int64)
    w_hi = np.asarray([w[1] for w in windows], np.int64)
    w_of = np.searchsorted(w_lo, cols, side="right") - 1
    hi = np.searchsorted(pos, cols, side="right")
    known = (nxt_pos[hi] < w_hi[w_of]) | (w_hi[w_of] >= span_hi)
    unknown = np.nonzero(~known)[0].tolist()
    if not unknown:
        return []
    # the rest only for the few columns at a window's end that are left: reference spans, mate fields
    end = pos.astype(np.int64) + _ref_span(batch)[sel]
    pmax = np.maximum.accumulate(end)
    need = {}
    flag = batch.flag[sel]
    mtid = batch.mtid[sel]
    mpos = batch.mpos[sel].astype(np.int64)
    for k in unknown:
        c, w = int(cols[k]), int(w_of[k])
        lo = int(np.searchsorted(pmax, c, side="right"))
        j = slice(lo, int(hi[k]))
        cand = ok[j] & (end[j] > c) & ((flag[j] & 2) != 0) & ((flag[j] & 8) == 0) & (mtid[j] == tid) & (mpos[j] > c) & \
            (mpos[j] >= w_hi[w])
        if cand.any():
            need[w] = max(need.get(w, 0), int(mpos[j][cand].max()) + 1)
    return sorted(need.items())


def fetch_group(bam, chrom: str, first_pos: int, last_pos: int, positions: Sequence[int], params,
                pad: int = PAD, merge_gap: int = MERGE_GAP) -> ReadBatch:
    """The read batch for `positions` -- a chromosome group or a chunk of one -- (1-based POS values; first / last = the
    GROUP's first and last record, whose span the reference fetches: vafator/pileups.py:137-138).  Without an index the
    file's whole batch is returned (bamio.BamFile caches it)."""
    from . import device as _dev
    if not getattr(bam, "index_path", None):
        return bam.read_batch(None)
    span_lo, span_hi = first_pos - 1, last_pos
    span = [(chrom, span_lo, span_hi)]
    cols = [p - 1 for p in positions if first_pos <= p <= last_pos]
    windows = plan_windows(cols, span_lo, span_hi, pad, merge_gap)
    if not windows or chrom not in bam.contigs:
        return bam.read_batch(span)
    tid = bam.contigs.index(chrom)
    for _ in range(MAX_ROUNDS):
        if len(windows) > 1 and sum(b - a for a, b in windows) > DENSE * (windows[-1][1] - windows[0][0]):
            # dense: one stretch from the first to the last window (for a whole group nearly the reference's own span; for a chunk
            # of a group just its share of it), checked and extended like any other window
            windows = [[windows[0][0], windows[-1][1]]]
        batch = bam.read_batch([(chrom, a, b) for a, b in windows])
        ext = missing_reads(batch, _dev.read_passes(batch, params), cols, windows, span_hi, tid)
        if not ext:
            return batch
        for w, e in ext:
            windows[w][1] = min(span_hi, max(windows[w][1], e))
        merged: List[List[int]] = []
        for a, b in windows:
            if merged and a - merged[-1][1] < merge_gap:
                merged[-1][1] = max(merged[-1][1], b)
            else:
                merged.append([a, b])
        windows = merged
    return bam.read_batch(span)
