"""Interval x BAM sharding: every (variant, BAM) unit is independent, so a BAM's units
are cut into contiguous genomic intervals of about equal size and each interval travels with just the
reads that can overlap it.  No data-path collective: ranks compute their shards, the fixed-size
per-unit records are gathered on the host in unit order.

The reference's only parallelism is one CPU process per chromosome (vafator/annotator.py:188-210);
intervals are finer than chromosomes so that 8 GPUs stay busy on one chromosome too."""
from __future__ import annotations

from dataclasses import dataclass, replace
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from .batch import ReadBatch, VariantUnits



@dataclass
class Shard:
    unit_lo: int
    unit_hi: int
    read_lo: int
    read_hi: int


def unit_cuts(n_units: int, n_shards: int) -> List[int]:
    """Boundaries of <= n_shards contiguous runs of units (genome order) of equal size: shard k = [cuts[k], cuts[k+1])."""
    n_shards = max(1, min(n_shards, max(n_units, 1)))
    return [int(round(k * n_units / n_shards)) for k in range(n_shards + 1)]


def plan_shards(batch: ReadBatch, units: VariantUnits, n_shards: int, passes: Optional[np.ndarray] = None,
                stop: Optional[np.ndarray] = None) -> List[Shard]:
    """Cut `units` (in genome order) into <= n_shards contiguous runs and find, for each, the read index
    range that contains every read able to influence it:
      * all reads overlapping a column of the run (found through the running maximum of the read ends),
      * and, after the last column, the reads up to and including the first one the reference would push
        (`passes`): its arrival can still change a base quality seen at that column (mate-overlap
        handling, oracle/htslib_pileup.py header)."""
    n = units.n_units
    if n == 0:
        return []
    order_ok = np.all((units.tid[1:] > units.tid[:-1]) | ((units.tid[1:] == units.tid[:-1]) & (units.pos[1:] >= units.pos[:-1])))
    if not order_ok:
        raise ValueError("units must be in genome order to be sharded by interval")
    cuts = unit_cuts(n, n_shards)
    n_shards = len(cuts) - 1
    placed = batch.tid >= 0
    tid = batch.tid[placed].astype(np.int64)
    pos = batch.pos[placed].astype(np.int64)
    n_reads = len(pos)
    key = tid * (1 << 32) + pos
    span = _ref_span(batch)[placed]
    end_key = tid * (1 << 32) + pos + span
    pmax = np.maximum.accumulate(end_key) if n_reads else end_key
    shards = []
    for k in range(n_shards):
        a, b = cuts[k], cuts[k + 1]
        if a == b:
            continue
        first_key = int(units.tid[a]) * (1 << 32) + int(units.pos[a])
        last_key = int(units.tid[b - 1]) * (1 << 32) + int(units.pos[b - 1])
        lo = int(np.searchsorted(pmax, first_key, side="right"))   # first read whose running-max end exceeds the column
        hi = int(np.searchsorted(key, last_key, side="right"))     # first read starting beyond the last column
        last_tid = int(units.tid[b - 1])
        bound = int(stop[b - 1]) if stop is not None else (1 << 31)
        j = hi
        while j < n_reads and tid[j] == last_tid and pos[j] < bound:
            j += 1
            if passes is None or passes[j - 1]:
                break
        shards.append(Shard(a, b, min(lo, hi), max(j, hi)))
    return shards


def _ref_span(batch: ReadBatch) -> np.ndarray:
    op = batch.cigar & 15
    consumes = (op == 0) | (op == 2) | (op == 3) | (op == 7) | (op == 8)
    length = np.where(consumes, batch.cigar >> 4, 0).astype(np.int64)
    csum = np.concatenate([[0], np.cumsum(length)])
    return csum[batch.cigar_off + batch.n_cigar] - csum[batch.cigar_off]


def slice_reads(batch: ReadBatch, lo: int, hi: int) -> ReadBatch:
    """Reads [lo, hi) as a self-contained canonical batch (pool offsets re-based).  The kernels pair overlapping mates
    among the reads they are given: a partner left outside the slice cannot overlap any column of the shard."""
    sl = slice(lo, hi)
    c0 = int(batch.cigar_off[lo]) if hi > lo else 0
    c1 = int(batch.cigar_off[hi - 1] + batch.n_cigar[hi - 1]) if hi > lo else 0
    s0 = int(batch.seq_off[lo]) if hi > lo else 0
    s1 = int(batch.seq_off[hi - 1] + (batch.l_qseq[hi - 1] + 1) // 2) if hi > lo else 0
    q0 = int(batch.qual_off[lo]) if hi > lo else 0
    q1 = int(batch.qual_off[hi - 1] + batch.l_qseq[hi - 1]) if hi > lo else 0
    return ReadBatch(contigs=batch.contigs, tid=batch.tid[sl], pos=batch.pos[sl], flag=batch.flag[sl], mapq=batch.mapq[sl],
                     l_qseq=batch.l_qseq[sl], n_cigar=batch.n_cigar[sl], cigar_off=batch.cigar_off[sl] - c0,
                     cigar=batch.cigar[c0:c1], seq_off=batch.seq_off[sl] - s0, seq=batch.seq[s0:s1],
                     qual_off=batch.qual_off[sl] - q0, qual=batch.qual[q0:q1], mtid=batch.mtid[sl], mpos=batch.mpos[sl],
                     isize=batch.isize[sl], qname_id=batch.qname_id[sl], qname_hash=batch.qname_hash[sl],
                     contig_lengths=batch.contig_lengths, contiguous=batch.contiguous)


def slice_units(units: VariantUnits, a: int, b: int) -> VariantUnits:
    return replace(units, tid=units.tid[a:b], pos=units.pos[a:b], meta=units.meta[a:b], length=units.length[a:b],
                   seq_off=units.seq_off[a:b], rec=units.rec[a:b], alt_index=units.alt_index[a:b], unit_of={})


def run_sharded(batch: ReadBatch, units: VariantUnits, stop: np.ndarray, passes: np.ndarray,
                n_shards: int, rank: int, world: int,
                compute: Callable[[ReadBatch, VariantUnits, np.ndarray], np.ndarray],
                gather: Callable[[list], list]) -> Optional[np.ndarray]:
    """Shard, compute this rank's shards (shard k -> rank k % world), gather on rank 0 in unit order.
    `compute(sub_batch, sub_units, sub_stop)` returns the per-unit records of one shard;
    `gather(obj)` returns the list of every rank's object on rank 0 (None elsewhere)."""
    shards = plan_shards(batch, units, n_shards, passes, stop)
    mine = []
    for k, sh in enumerate(shards):
        if k % world != rank:
            continue
        sub = slice_reads(batch, sh.read_lo, sh.read_hi)
        res = compute(sub, slice_units(units, sh.unit_lo, sh.unit_hi), stop[sh.unit_lo:sh.unit_hi])
        mine.append((k, res))
    everything = gather(mine)
    if everything is None:
        return None
    parts = dict(x for per_rank in everything for x in per_rank)
    return np.concatenate([parts[k] for k in range(len(shards))]) if shards else None
