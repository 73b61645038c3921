"""Carrier types of the pileup path (mirror of vafator/pileup_utils.py:9-119): same names, same fields, same types, so
that the reference's own Annotator._add_stats / _annotate_sample / _annotate_replicate (vafator/annotator.py:260-420) run
unchanged on what vafator_b200.pileups.collect_metrics_for_chrom returns."""
from __future__ import annotations

from collections import Counter
from dataclasses import dataclass
from math import nan
from typing import Dict, List

import numpy as np


@dataclass
class VariantRecord:
    """Lightweight variant: the fields of cyvcf2.Variant the path reads (pileup_utils.py:9-24)."""
    CHROM: str
    POS: int
    REF: str
    ALT: List[str]


@dataclass
class CoverageMetrics:
    """Pileup metrics of one variant in one BAM (vafator/pileup_utils.py:27-49), field for field:
    ac            allele counts per observed base (Counter), including the reference and every IUPAC letter seen
    dp            depth
    bqs / mqs / positions          median per allele (Counter: a missing allele reads 0, an empty list nan)
    all_bqs / all_mqs / all_positions   the per-read values per allele, in read order -- read back from the device
                                        (vfk_pileup_values) by collect_metrics_for_chrom"""
    ac: dict
    dp: int
    bqs: dict = None
    mqs: dict = None
    positions: dict = None
    all_bqs: dict = None
    all_mqs: dict = None
    all_positions: dict = None


EMPTY_METRICS = CoverageMetrics(ac=Counter(), dp=0, bqs=Counter(), mqs=Counter(), positions=Counter(), all_bqs={}, all_mqs={},
                                all_positions={})


def aggregate_list_per_base(bases: List[str], values: list) -> Dict[str, list]:
    """vafator/pileup_utils.py:64-79."""
    out: Dict[str, list] = {}
    for b, v in zip(bases, values):
        out.setdefault(b, []).append(v)
    return out


def safe_median(values: list) -> float:
    """vafator/pileup_utils.py:82-92: the median, nan for an empty list."""
    return float(np.median(values)) if len(values) else nan


def is_snp(variant) -> bool:
    return len(variant.REF) == 1 and len(variant.ALT[0]) == 1


def is_insertion(variant) -> bool:
    return len(variant.REF) == 1 and len(variant.ALT[0]) > 1


def is_deletion(variant) -> bool:
    return len(variant.ALT[0]) == 1 and len(variant.REF) > 1
