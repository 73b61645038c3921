"""ORACLE (test infrastructure, NOT product code) -- CPU restatement of the vafator v3.0.0
layers that sit on top of the pysam pileup column: per-column allele metrics, replicate
merge, rank-sum tests, power model and INFO-string formatting.

PARITY STATUS OF THIS FILE: **pinned**.  tests/test_oracle_pins.py runs it against
  * the reference's own modules imported unmodified from /root/reference (when present):
    vafator/pileups.py _get_metrics_from_column, vafator/annotator.py _add_stats,
    vafator/rank_sum_test.py, vafator/power.py, vafator/ploidies.py
  * tests/golden/*.json produced by oracle/gen_golden.py from those same modules
  * the reference's known answers (vafator/tests/test_rank_sum_test.py:27-35,
    vafator/tests/test_power_calculator.py:11-266, notebooks/power.tsv).
The statistics call scipy exactly as the reference does (scipy is the un-vendored
third-party dependency that holds the arithmetic: setup.cfg:45).

It exists because /root/reference is not present on the GPU box; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg may import it.
"""
from __future__ import annotations

import math
from collections import Counter, OrderedDict
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
import scipy.stats

# vafator/constants.py:5
IUPAC_AMBIGUOUS = ("N", "M", "R", "W", "S", "Y", "K", "V", "H", "D", "B")
_REF_CONSUMING = (0, 2, 3, 7, 8)
_QUERY_CONSUMING = (0, 1, 4, 7, 8)


@dataclass
class Metrics:
    """Same content as vafator/pileup_utils.py:27-49 CoverageMetrics."""
    ac: Counter = field(default_factory=Counter)
    dp: int = 0
    med: Dict[str, Counter] = field(default_factory=lambda: {"bq": Counter(), "mq": Counter(), "pos": Counter()})
    dist: Dict[str, dict] = field(default_factory=lambda: {"bq": {}, "mq": {}, "pos": {}})


def _median(xs) -> float:
    # vafator/pileup_utils.py:82-92
    return float(np.median(xs)) if xs else math.nan


def variant_kind(ref: str, alt0: str) -> Optional[str]:
    # vafator/pileup_utils.py:95-119 -- decided on ALT[0] only
    if len(ref) == 1 and len(alt0) == 1:
        return "snv"
    if len(ref) == 1 and len(alt0) > 1:
        return "ins"
    if len(alt0) == 1 and len(ref) > 1:
        return "del"
    return None


def _group(rows, col):
    out: Dict[str, list] = {}
    for row in rows:
        out.setdefault(row[0], []).append(row[col])
    return out


def snv_metrics(col, include_ambiguous: bool) -> Metrics:
    # vafator/pileups.py:184-238
    rows = []
    for pr in col.pileups:
        if pr.is_refskip:
            continue
        aln = pr.alignment
        if pr.is_del:
            rows.append(("", 0, aln.mapping_quality, pr.query_position_or_next))
        else:
            q = pr.query_position
            rows.append((aln.query_sequence[q].upper(), aln.query_qualities[q], aln.mapping_quality, q))
    m = Metrics()
    for name, c in (("bq", 1), ("mq", 2), ("pos", 3)):
        m.dist[name] = _group(rows, c)
        m.med[name] = Counter({a: _median(v) for a, v in m.dist[name].items()})
    m.ac = Counter(r[0] for r in rows if r[0] != "")
    if include_ambiguous:
        m.dp = len(rows)
    else:
        m.dp = sum(1 for r in rows if r[0] == "" or r[0] not in IUPAC_AMBIGUOUS)
    return m


def _indel_metrics(ref: str, alts: Sequence[str], supports, col) -> Metrics:
    """Shared tail of vafator/pileups.py:241-305 and :308-362.  `supports(pr)` returns how many
    times the read supports ALT[0] when pr.indel has the variant's sign."""
    up0 = alts[0].upper()
    mq = {a.upper(): [] for a in alts}
    ps = {a.upper(): [] for a in alts}
    mq[ref] = []
    ps[ref] = []
    ac = {a.upper(): 0 for a in alts}
    reads = col.pileups
    for pr in reads:
        hits = supports(pr)
        if hits is None:  # pr.indel == 0: counted as reference support (pileups.py:291-294 / :348-351)
            mq[ref].append(pr.alignment.mapping_quality)
            ps[ref].append(pr.query_position_or_next)
        else:
            for _ in range(hits):
                ac[up0] += 1
                mq[up0].append(pr.alignment.mapping_quality)
                ps[up0].append(pr.query_position_or_next)
    m = Metrics()
    m.ac = Counter(ac)
    m.dp = len(reads)
    m.med = {"bq": Counter(), "mq": Counter({k: _median(v) for k, v in mq.items()}),
             "pos": Counter({k: _median(v) for k, v in ps.items()})}
    m.dist = {"bq": Counter(), "mq": mq, "pos": ps}
    return m


def insertion_metrics(pos1: int, ref: str, alts: Sequence[str], col) -> Metrics:
    # vafator/pileups.py:241-305
    length = len(alts[0]) - len(ref)
    inserted = alts[0][1:]

    def supports(pr):
        if pr.indel == 0:
            return None
        if pr.indel < 0:
            return 0
        hits = 0
        refpos = pr.alignment.reference_start
        qoff = 0
        for op, ln in pr.alignment.cigartuples:
            if op in _REF_CONSUMING:
                refpos += ln
                if refpos > pos1:
                    break
            if op in _QUERY_CONSUMING:
                qoff += ln
            if op == 1:
                # NB (reference quirk): qoff already includes this I and any
                # leading soft clip, while `query` has the soft clips trimmed.
                if refpos == pos1 and ln == length and inserted == pr.alignment.query[qoff:qoff + length]:
                    hits += 1
        return hits

    return _indel_metrics(ref, alts, supports, col)


def deletion_metrics(pos1: int, ref: str, alts: Sequence[str], col) -> Metrics:
    # vafator/pileups.py:308-362
    length = len(ref) - len(alts[0])

    def supports(pr):
        if pr.indel == 0:
            return None
        if pr.indel > 0:
            return 0
        refpos = pr.alignment.reference_start
        for op, ln in pr.alignment.cigartuples:
            if op in (0, 3, 7, 8):
                refpos += ln
            elif op == 2:
                if refpos == pos1 and ln == length:
                    return 1
                refpos += ln
            if refpos > pos1:
                break
        return 0

    return _indel_metrics(ref, alts, supports, col)


def column_metrics(pos1: int, ref: str, alts: Sequence[str], col, include_ambiguous=True) -> Optional[Metrics]:
    # vafator/pileups.py:162-181
    kind = variant_kind(ref, alts[0])
    if kind == "snv":
        return snv_metrics(col, include_ambiguous)
    if kind == "ins":
        return insertion_metrics(pos1, ref, alts, col)
    if kind == "del":
        return deletion_metrics(pos1, ref, alts, col)
    return None


def collect_metrics_for_chrom(chrom, variants, bam, min_base_quality, min_mapping_quality,
                              include_ambiguous_bases=False):
    """vafator/pileups.py:106-159.  `variants`: objects with POS/REF/ALT, position sorted."""
    if not variants:
        return {}
    by_pos: Dict[int, list] = {}
    for v in variants:
        by_pos.setdefault(v.POS, []).append(v)
    out = {}
    for col in bam.pileup(contig=chrom, start=variants[0].POS - 1, stop=variants[-1].POS, truncate=True,
                          max_depth=1000000, min_base_quality=min_base_quality,
                          min_mapping_quality=min_mapping_quality, stepper="samtools"):
        p1 = col.reference_pos + 1
        for v in by_pos.get(p1, ()):
            m = column_metrics(p1, v.REF, v.ALT, col, include_ambiguous_bases)
            if m is not None:
                out[(p1, v.REF, v.ALT[0])] = m
    return out


# ---------------------------------------------------------------------------
# statistics
# ---------------------------------------------------------------------------
def rank_sum(alt: list, ref: list):
    # vafator/rank_sum_test.py:6-12
    if not alt or not ref:
        return math.nan, math.nan
    z, p = scipy.stats.ranksums(x=alt, y=ref)
    return round(z, 3), round(p, 5)


def rank_sum_raw(alt: list, ref: list):
    """Unrounded (z, p): what the device epilogue is compared against (1e-6 relative)."""
    if not alt or not ref:
        return math.nan, math.nan
    z, p = scipy.stats.ranksums(x=alt, y=ref)
    return float(z), float(p)


def rank_sum_fields(dist: dict, ref: str, alts: Sequence[str]):
    # vafator/rank_sum_test.py:15-26
    zs, ps = [], []
    for a in alts:
        z, p = rank_sum(dist.get(a, []), dist.get(ref, []))
        if not np.isnan(z) and not np.isnan(p):
            zs.append(str(z))
            ps.append(str(p))
    return ps, zs


class Power:
    """vafator/power.py:14-131 with the ploidy lookup of vafator/ploidies.py:36-50.
    `ploidy`: {sample: float | [(chrom, start, end, copy_number), ...]}"""

    def __init__(self, ploidy=None, purity=None, normal_ploidy=2, fpr=5 * (10 ** -7), error_rate=10 ** -3):
        self.ploidy = ploidy or {}
        self.purity = purity or {}
        self.normal_ploidy = normal_ploidy
        self.fpr = fpr
        self.error_rate = error_rate
        self._k: Dict[int, int] = {}

    def tumor_ploidy(self, sample, chrom=None, pos1=None) -> float:
        src = self.ploidy.get(sample, 2.0)
        if isinstance(src, (int, float)):
            return src
        for c, s, e, cn in src:  # first hit wins, half-open on the 0-based position
            if c == chrom and s <= pos1 - 1 < e:
                return float(cn)
        return 2.0

    def eaf(self, sample, chrom=None, pos1=None) -> float:
        # vafator/power.py:52-82
        pur = self.purity.get(sample, 1.0)
        tp = max(1, self.tumor_ploidy(sample, chrom, pos1))
        return pur / (pur * tp + ((1 - pur) * self.normal_ploidy))

    def pu_raw(self, dp, ac, eaf):
        return scipy.stats.binom.cdf(k=ac, n=dp, p=eaf)

    def pu(self, dp, ac, eaf):
        # vafator/power.py:39-50
        return round(self.pu_raw(dp, ac, eaf), 5)

    def _p(self, m, n):
        # vafator/power.py:84-93
        if m >= 1:
            return 1 - scipy.stats.binom.cdf(k=m - 1, n=n, p=self.error_rate / 3)
        return 1

    def k(self, dp) -> int:
        # vafator/power.py:106-114
        if dp not in self._k:
            k = 1
            while self._p(k, dp) > self.fpr:
                k += 1
            self._k[dp] = k
        return self._k[dp]

    def pw_raw(self, dp, eaf):
        # vafator/power.py:95-104, 116-131
        k = self.k(dp)
        pk, pk1 = self._p(k, dp), self._p(k - 1, dp)
        d = (self.fpr - pk) / (pk1 - pk)
        pw = 1 - scipy.stats.binom.cdf(k=k - 1, n=dp, p=eaf) + d * scipy.stats.binom.pmf(k=k, n=dp, p=eaf)
        return pw, k

    def pw(self, dp, eaf):
        pw, k = self.pw_raw(dp, eaf)
        return round(pw, 5), k


# ---------------------------------------------------------------------------
# INFO formatting (vafator/annotator.py:260-432)
# ---------------------------------------------------------------------------
def _af(ac, dp):
    return round(float(ac) / dp, 5) if dp > 0 else 0.0


def _fields(prefix, suffix, ref, alts, ac, dp, med, dist, power: Power, eaf, with_eaf):
    o = OrderedDict()
    key = lambda name: "{}_{}{}".format(prefix, name, suffix)
    o[key("af")] = ",".join(str(_af(ac[a], dp)) for a in alts)
    o[key("ac")] = ",".join(str(ac[a]) for a in alts)
    o[key("n")] = str(sum(ac.get(b, 0) for b in IUPAC_AMBIGUOUS))
    o[key("dp")] = dp  # assigned as int (annotator.py:387)
    if with_eaf:
        o[key("eaf")] = str(eaf)
    o[key("pu")] = ",".join(str(power.pu(dp, ac[a], eaf)) for a in alts)
    pw, k = power.pw(dp, eaf)
    o[key("pw")] = str(pw)
    o[key("k")] = str(k)
    for name in ("bq", "mq", "pos"):
        o[key(name)] = ",".join([str(med[name][ref])] + [str(med[name][a]) for a in alts])
    for name, tag in (("mq", "rsmq"), ("bq", "rsbq"), ("pos", "rspos")):
        ps, zs = rank_sum_fields(dist[name], ref, alts)
        if zs:
            o[key(tag)] = ",".join(zs)
            o["{}_{}_pv{}".format(prefix, tag, suffix)] = ",".join(ps)
    return o


EMPTY = Metrics()


def info_fields(chrom, pos1, ref, alts, samples: "OrderedDict[str, List[Optional[Metrics]]]", power: Power):
    """Ordered {INFO key: value} for one variant.  `samples`: {sample: [Metrics per BAM]} in
    command-line order; a missing/None entry stands for EMPTY_METRICS.
    Follows vafator/annotator.py:260-302 (_add_stats), :304-353, :355-420."""
    out = OrderedDict()
    for sample, per_bam in samples.items():
        eaf = power.eaf(sample, chrom, pos1)
        g_ac, g_dp = Counter(), 0
        g_med = {"bq": Counter(), "mq": Counter(), "pos": Counter()}
        g_dist = {"bq": {}, "mq": {}, "pos": {}}
        for i, m in enumerate(per_bam):
            m = EMPTY if m is None else m
            if len(per_bam) > 1:
                out.update(_fields(sample, "_%d" % (i + 1), ref, alts, m.ac, m.dp, m.med, m.dist, power, eaf, False))
            g_ac.update(m.ac)
            g_dp += m.dp
            for name in ("bq", "mq", "pos"):
                g_med[name].update(m.med[name])   # medians are SUMMED across replicates
                g_dist[name].update(m.dist[name])  # distributions: last BAM wins per allele key
        out.update(_fields(sample, "", ref, alts, g_ac, g_dp, g_med, g_dist, power, eaf, True))
    return out
