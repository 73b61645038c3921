"""collect_metrics_for_chrom -- the first of the three call sites the reference's Annotator makes
(vafator/pileups.py:106-159, called at annotator.py:225-232), backed by the device.  `bam` is a
vafator_b200.bamio.BamFile (or anything with .read_batch()) instead of a pysam.AlignmentFile."""
from __future__ import annotations

from collections import Counter
from typing import Dict, List, Tuple

import numpy as np

from .batch import KIND_SNV, build_units, variant_kind
from .pileup_utils import CoverageMetrics, RankSum

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
    records = [(chrom, v.POS, v.REF, list(v.ALT)) if first <= v.POS <= last else (chrom, v.POS, "", []) for v in variants]
    units = build_units(records, batch.contigs)
    counts = eng.pileup(batch, units, np.full(units.n_units, last, np.int32))
    stats = eng.epilogue(counts, np.full(units.n_units, 0.5))
    out: Dict[Tuple, CoverageMetrics] = {}
    for i, v in enumerate(variants):
        u0 = units.unit_of.get((i, 0))
        if u0 is None or counts["n_col"][u0] == 0:
            continue
        kind = variant_kind(v.REF, v.ALT[0])
        c0 = counts[u0]
        m = CoverageMetrics(ac=Counter(), dp=int(c0["dp"]))
        med = lambda x: float(x) / 2.0 if x >= 0 else float("nan")
        names = ("mq", "bq", "pos")
        stores = {"mq": m.mqs, "bq": m.bqs, "pos": m.positions}
        if kind == KIND_SNV:
            if c0["ac_ref"] > 0:
                m.ac[v.REF] = int(c0["ac_ref"])
                m.counts[v.REF] = int(c0["ac_ref"])
                for k, name in enumerate(names):
                    stores[name][v.REF] = med(c0["med2"][k][0])
                    m.keys[name].add(v.REF)
            for j, alt in enumerate(v.ALT):
                u = units.unit_of[(i, j)]
                c = counts[u]
                if c["ac_alt"] > 0:
                    m.ac[alt] = int(c["ac_alt"])
                    m.counts[alt] = int(c["ac_alt"])
                    for k, name in enumerate(names):
                        stores[name][alt] = med(c["med2"][k][1])
                        m.keys[name].add(alt)
                        if c["ac_ref"] > 0:
                            m.ranksums[name][alt] = RankSum(float(stats[u]["z"][k]), float(stats[u]["p"][k]))
            # ambiguous bases are part of `ac` in the reference (Counter over every observed base)
            if c0["n_amb"] > 0:
                m.ac["N*"] = int(c0["n_amb"])  # aggregate of all IUPAC codes; see CoverageMetrics docstring
        else:
            up = v.ALT[0].upper()
            m.ac[up] = int(c0["ac_alt"])
            m.counts[v.REF], m.counts[up] = int(c0["ac_ref"]), int(c0["ac_alt"])
            for k, name in ((0, "mq"), (2, "pos")):
                stores[name][v.REF] = med(c0["med2"][k][0])
                stores[name][up] = med(c0["med2"][k][1])
                m.keys[name].update((v.REF, up))
                if c0["ac_ref"] > 0 and c0["ac_alt"] > 0:
                    m.ranksums[name][up] = RankSum(float(stats[u0]["z"][k]), float(stats[u0]["p"][k]))
            for alt in v.ALT[1:]:
                m.ac.setdefault(alt.upper(), 0)
                for name in ("mq", "pos"):
                    stores[name].setdefault(alt.upper(), float("nan"))
                    m.keys[name].add(alt.upper())
        out[(v.POS, v.REF, v.ALT[0])] = m
    return out
