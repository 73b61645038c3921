"""Carrier types of the pileup path (mirror of vafator/pileup_utils.py:9-119)."""
from __future__ import annotations

from collections import Counter
from dataclasses import dataclass, field
from typing import Dict, List, Optional


@dataclass
class VariantRecord:
    """Lightweight variant: the fields of cyvcf2.Variant the path reads (pileup_utils.py:9-24)."""
    CHROM: str
    POS: int
    REF: str
    ALT: List[str]


@dataclass
class RankSum:
    """Rank-sum result for one metric and one ALT allele (raw, unrounded FP64 from the device)."""
    z: float
    p: float


@dataclass
class CoverageMetrics:
    """Pileup metrics of one variant in one BAM (pileup_utils.py:27-49).  Where the reference carries
    the full per-read lists (all_bqs / all_mqs / all_positions) only to feed scipy.stats.ranksums,
    this carries what the device already reduced them to: `ranksums[metric][alt] = RankSum`, and
    `keys[metric]` = the alleles that have a (possibly empty) list in the reference's dictionaries."""
    ac: Counter
    dp: int
    bqs: Counter = field(default_factory=Counter)
    mqs: Counter = field(default_factory=Counter)
    positions: Counter = field(default_factory=Counter)
    ranksums: Dict[str, Dict[str, RankSum]] = field(default_factory=lambda: {"mq": {}, "bq": {}, "pos": {}})
    keys: Dict[str, set] = field(default_factory=lambda: {"mq": set(), "bq": set(), "pos": set()})
    counts: Dict[str, int] = field(default_factory=dict)  # allele -> number of values in its lists


EMPTY_METRICS = CoverageMetrics(ac=Counter(), dp=0)


def is_snp(variant) -> bool:
    return len(variant.REF) == 1 and len(variant.ALT[0]) == 1


def is_insertion(variant) -> bool:
    return len(variant.REF) == 1 and len(variant.ALT[0]) > 1


def is_deletion(variant) -> bool:
    return len(variant.ALT[0]) == 1 and len(variant.REF) > 1
