"""collect_metrics_for_chrom -- the first of the three call sites the reference's Annotator makes
(vafator/pileups.py:106-159, called at annotator.py:225-232), backed by the device.  `bam` is a
vafator_b200.bamio.BamFile (or anything with .read_batch()) instead of a pysam.AlignmentFile.

Same signature, same return type: {(POS, REF, ALT[0]): CoverageMetrics} with the reference's fields -- allele counts for every
observed base, medians, and the full per-read lists -- so the reference's own annotator code runs on it unchanged
(tests/test_reference_callsite.py).  It is the compatibility surface: the product's Annotator goes through
vafator_b200.engine.Engine, which keeps everything as arrays and never materialises per-read lists."""
from __future__ import annotations

from collections import Counter
from typing import Dict, List, Sequence, Tuple

import numpy as np

from .batch import KIND_SNV, SEQ_CHARS, VariantUnits, build_units, variant_kind
from .pileup_utils import CoverageMetrics, safe_median

_engines = {}


def _engine(min_base_quality, min_mapping_quality, include_ambiguous_bases):
    from .engine import Engine
    key = (min_base_quality, min_mapping_quality, bool(include_ambiguous_bases))
    if key not in _engines:
        _engines[key] = Engine(0, min_mapq=min_mapping_quality, min_baseq=min_base_quality,
                               include_ambiguous=include_ambiguous_bases)
    return _engines[key]


def stream_variants_by_chrom(vcf):
    """(chrom, [variants]) for consecutive records of one chromosome (vafator/pileups.py:83-103)."""
    chrom, group = None, []
    for v in vcf:
        if v.CHROM != chrom:
            if group:
                yield chrom, group
            chrom, group = v.CHROM, [v]
        else:
            group.append(v)
    if group:
        yield chrom, group


def callsite_records(chrom: str, variants: Sequence) -> List[Tuple[str, int, str, List[str]]]:
    """The records the compatibility call evaluates on the device.  The reference counts EVERY base it observes in an SNV
    column (vafator/pileups.py:214-222): such a record asks for one unit per letter of the BAM alphabet other than REF, so
    that every observed base gets its count, its medians and its lists.  Indels keep ALT[0] (pileups.py:253-261); records
    outside [first.POS, last.POS] get no column (pileups.py:137-138)."""
    first, last = variants[0].POS, variants[-1].POS
    out = []
    for v in variants:
        if not (first <= v.POS <= last) or not v.ALT:
            out.append((chrom, v.POS, "", []))
        elif variant_kind(v.REF, v.ALT[0]) == KIND_SNV:
            out.append((chrom, v.POS, v.REF, [c for c in SEQ_CHARS if c != v.REF]))
        else:
            out.append((chrom, v.POS, v.REF, [v.ALT[0]]))
    return out


def metrics_from_device(variants: Sequence, units: VariantUnits, counts: np.ndarray, values: Sequence[np.ndarray]) -> Dict[Tuple, CoverageMetrics]:
    """Device records -> the reference's CoverageMetrics.  counts[u]: vfk_counts of unit u; values[u]: its packed per-read
    words (vfk_pileup_values: mq | bq << 8 | read position << 16 | in_REF_list << 28 | in_ALT_list << 29), in read order.
    `units` were built from callsite_records(...)."""
    out: Dict[Tuple, CoverageMetrics] = {}
    split = lambda w: ((w & 0xFF).tolist(), ((w >> 8) & 0xFF).tolist(), ((w >> 16) & 0xFFF).tolist())
    for i, v in enumerate(variants):
        u0 = units.unit_of.get((i, 0))
        if u0 is None or counts["n_col"][u0] == 0:
            continue  # no pileup column: the reference has no entry either (-> EMPTY_METRICS at annotator.py:248-250)
        kind = variant_kind(v.REF, v.ALT[0])
        dp = int(counts["dp"][u0])
        if kind == KIND_SNV:
            all_mqs, all_bqs, all_pos = {}, {}, {}
            w0 = np.asarray(values[u0], np.uint32)
            ref_words = w0[(w0 >> 28) & 1 == 1]
            if len(ref_words):
                all_mqs[v.REF], all_bqs[v.REF], all_pos[v.REF] = split(ref_words)
            alleles = [c for c in SEQ_CHARS if c != v.REF]
            for j, a in enumerate(alleles):
                w = np.asarray(values[units.unit_of[(i, j)]], np.uint32)
                w = w[(w >> 29) & 1 == 1]
                if len(w):
                    all_mqs[a], all_bqs[a], all_pos[a] = split(w)
            m = CoverageMetrics(ac=Counter({b: len(l) for b, l in all_mqs.items()}), dp=dp,
                                bqs=Counter({b: safe_median(l) for b, l in all_bqs.items()}),
                                mqs=Counter({b: safe_median(l) for b, l in all_mqs.items()}),
                                positions=Counter({b: safe_median(l) for b, l in all_pos.items()}),
                                all_bqs=all_bqs, all_mqs=all_mqs, all_positions=all_pos)
        else:
            up = v.ALT[0].upper()
            w0 = np.asarray(values[u0], np.uint32)
            ref_mq, _, ref_pos = split(w0[(w0 >> 28) & 1 == 1])
            alt_mq, _, alt_pos = split(w0[(w0 >> 29) & 1 == 1])
            mq = {a.upper(): [] for a in v.ALT}
            pos = {a.upper(): [] for a in v.ALT}
            mq[v.REF], pos[v.REF] = ref_mq, ref_pos
            mq[up], pos[up] = alt_mq, alt_pos
            ac = {a.upper(): 0 for a in v.ALT}
            ac[up] = int(counts["ac_alt"][u0])
            m = CoverageMetrics(ac=Counter(ac), dp=dp, mqs=Counter({k: safe_median(l) for k, l in mq.items()}),
                                positions=Counter({k: safe_median(l) for k, l in pos.items()}), bqs=Counter(),
                                all_mqs=mq, all_positions=pos, all_bqs=Counter())
        out[(v.POS, v.REF, v.ALT[0])] = m
    return out


def collect_metrics_for_chrom(chrom: str, variants: List, bam, min_base_quality: int, min_mapping_quality: int,
                              include_ambiguous_bases: bool = False) -> Dict[Tuple, CoverageMetrics]:
    """{(POS, REF, ALT[0]): CoverageMetrics} for the variants with a pileup column; position-sorted input,
    region [first.POS-1, last.POS) as in the reference."""
    if not variants:
        return {}
    eng = _engine(min_base_quality, min_mapping_quality, include_ambiguous_bases)
    first, last = variants[0].POS, variants[-1].POS
    from .bamio import BamFile
    from .fetch import fetch_group
    # the reads the columns depend on, out of the span the reference fetches (windows around the variants: fetch.py)
    batch = fetch_group(bam, chrom, first, last, [v.POS for v in variants], eng.params) if isinstance(bam, BamFile) else bam.read_batch()
    units = build_units(callsite_records(chrom, variants), batch.contigs)
    stop = np.full(units.n_units, last, np.int32)
    counts = eng.pileup(batch, units, stop)
    values = eng.values(batch, units, stop)
    return metrics_from_device(variants, units, counts, values)
