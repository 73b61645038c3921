"""Rank-sum tests on explicit value lists (mirror of vafator/rank_sum_test.py:6-26): the mid-rank
sums come from vfk_ranksum_values, z and p from the FP64 epilogue kernel -- scipy.stats.ranksums'
arithmetic (no tie / continuity correction) without scipy on the path."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

_engine = None


def _get_engine():
    global _engine
    if _engine is None:
        from .engine import Engine
        _engine = Engine(0)
    return _engine


def _raw(pairs):
    """pairs: [(alt_list, ref_list)] of non-negative ints < 4096 -> [(z, p)] unrounded."""
    from . import device as _dev
    eng = _get_engine()
    # the metric sits in the read-position bit field of the packed word (12 bits)
    alt = [np.asarray(a, np.uint32) << np.uint32(16) for a, r in pairs]
    ref = [np.asarray(r, np.uint32) << np.uint32(16) for a, r in pairs]
    s2 = eng.ranksum_s2(alt, ref)
    c = np.zeros(len(pairs), _dev.COUNTS_DTYPE)
    c["ac_alt"] = [len(a) for a, r in pairs]
    c["ac_ref"] = [len(r) for a, r in pairs]
    c["dp"] = c["ac_alt"] + c["ac_ref"]
    c["s2"][:, 2] = s2[:, 2]
    st = eng.epilogue(c, np.full(len(pairs), 0.5))
    return [(float(x["z"][2]), float(x["p"][2])) for x in st]


def calculate_rank_sum_test(alternate_dist: List[int], reference_dist: List[int]) -> Tuple[float, float]:
    if not alternate_dist or not reference_dist:
        return np.nan, np.nan
    if min(min(alternate_dist), min(reference_dist)) < 0 or max(max(alternate_dist), max(reference_dist)) > 4095:
        raise ValueError("rank-sum values must be integers in [0, 4095] (qualities and read positions)")
    z, p = _raw([(alternate_dist, reference_dist)])[0]
    return float(np.round(np.float64(z), 3)), float(np.round(np.float64(p), 5))


def get_rank_sum_tests(distributions: dict, variant):
    stats, pvalues = [], []
    for alt in variant.ALT:
        stat, pvalue = calculate_rank_sum_test(distributions.get(alt, []), distributions.get(variant.REF, []))
        if not np.isnan(stat) and not np.isnan(pvalue):
            stats.append(str(stat))
            pvalues.append(str(pvalue))
    return pvalues, stats
