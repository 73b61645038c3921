"""The call-site swap, with the UNMODIFIED reference annotator on the other side (CPU; needs /root/reference, skipped with a
loud reason elsewhere -- the GPU box has no reference checkout, tests/test_gpu_annotate.py covers the device half there).

vafator/annotator.py is imported as it lies (stub for the two absent third-party modules) and its
Annotator._add_stats / _annotate_sample / _annotate_replicate consume what vafator_b200.pileups builds from DEVICE-SHAPED
records: vfk_counts per unit + the packed per-read words of vfk_pileup_values.  Those records come from test doubles here
(the C oracle for the counts, the reference's own column functions over the pileup restatement for the lists -- exactly what
the goldens were generated from); the product builds them on the GPU (test_collect_metrics_for_chrom_call_site).  The INFO
strings must equal the goldens, replicates included."""
import os
import sys
import types
from collections import OrderedDict

import numpy as np
import pytest

import helpers as H
from oracle import c_oracle, htslib_pileup as hp
from vafator_b200 import device, pileups
from vafator_b200.batch import KIND_SNV, SEQ_CHARS, build_units, variant_kind
from vafator_b200.pileup_utils import EMPTY_METRICS, VariantRecord

REF_ROOT = os.environ.get("VAFATOR_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF_ROOT, "vafator", "annotator.py")),
                                reason="REFERENCE CHECKOUT ABSENT (%s): the unmodified vafator/annotator.py cannot be imported here" % REF_ROOT)


def _import_reference():
    cy = types.ModuleType("cyvcf2")
    cy.Variant = cy.VCF = cy.Writer = object
    sys.modules.setdefault("cyvcf2", cy)
    sys.modules.setdefault("pysam", types.ModuleType("pysam"))
    pl = types.ModuleType("pysam.libcalignmentfile")
    pl.IteratorColumnRegion = pl.AlignmentFile = object
    sys.modules.setdefault("pysam.libcalignmentfile", pl)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import vafator.annotator as ra
    import vafator.pileups as rp
    import vafator.ploidies as rl
    import vafator.power as rw
    return ra, rp, rl, rw


class _Variant:
    def __init__(self, chrom, pos, ref, alts):
        self.CHROM, self.POS, self.REF, self.ALT = chrom, pos, ref, list(alts)
        self.INFO = OrderedDict()


def _device_records(rp, sc, batch_recs, contigs, chrom, variants, prm):
    """(units, counts, values) shaped as the device returns them, from the doubles."""
    recs = pileups.callsite_records(chrom, variants)
    units = build_units(recs, contigs)
    last = variants[-1].POS
    batch = H.ReadBatch.from_records(batch_recs, contigs)
    ou = [(contigs.index(chrom), recs[i][1], recs[i][2], recs[i][3][j], recs[i][3][0], last) for i, j in zip(units.rec, units.alt_index)]
    o = c_oracle.pileup(batch.as_dict(), ou, c_oracle.params(min_mapq=prm["min_mq"], min_baseq=prm["min_bq"]))
    counts = np.zeros(len(ou), device.COUNTS_DTYPE)
    snv = (units.meta & 3) == 0
    counts["dp"] = o["dp"] - (0 if prm["include_ambiguous"] else np.where(snv, o["n_amb"], 0))
    for f in ("n_amb", "ac_ref", "ac_alt", "n_col"):
        counts[f] = o[f]
    # packed per-read words: the reference's own column functions over the restated pileup, re-encoded
    f = hp.AlignmentFile([hp.Rec(d["qname"], d["flag"], d["tid"], d["pos"], d["mapq"], d["cigar"], d["seq"],
                                 bytes(ord(c) - 33 for c in d["qual"]), d["mtid"], d["mpos"], d["isize"], index=k)
                          for k, d in enumerate(batch_recs)], contigs)
    ref_metrics = rp.collect_metrics_for_chrom(chrom=chrom, variants=variants, bam=f, min_base_quality=prm["min_bq"],
                                               min_mapping_quality=prm["min_mq"], include_ambiguous_bases=prm["include_ambiguous"])
    pack = lambda mq, bq, pos, flag: [int(a) | (int(b) << 8) | (int(c) << 16) | flag for a, b, c in zip(mq, bq, pos)]
    values = [np.zeros(0, np.uint32) for _ in range(units.n_units)]
    for i, v in enumerate(variants):
        m = ref_metrics.get((v.POS, v.REF, v.ALT[0]))
        if m is None or (i, 0) not in units.unit_of:
            continue
        if variant_kind(v.REF, v.ALT[0]) == KIND_SNV:
            alleles = [c for c in SEQ_CHARS if c != v.REF]
            refw = pack(m.all_mqs.get(v.REF, []), m.all_bqs.get(v.REF, []), m.all_positions.get(v.REF, []), 1 << 28)
            for j, a in enumerate(alleles):
                w = pack(m.all_mqs.get(a, []), m.all_bqs.get(a, []), m.all_positions.get(a, []), 1 << 29)
                values[units.unit_of[(i, j)]] = np.asarray(refw + w, np.uint32)
        else:
            up = v.ALT[0].upper()
            z = lambda l: [0] * len(l)
            refw = pack(m.all_mqs[v.REF], z(m.all_mqs[v.REF]), m.all_positions[v.REF], 1 << 28)
            altw = pack(m.all_mqs[up], z(m.all_mqs[up]), m.all_positions[up], 1 << 29)
            values[units.unit_of[(i, 0)]] = np.asarray(refw + altw, np.uint32)
    return units, counts, values


@pytest.mark.parametrize("name", ["pileup_handmade.json", "pileup_random_11.json", "pileup_random_13.json"])
def test_unmodified_reference_annotator_runs_on_our_call_site(name):
    ra, rp, rl, rw = _import_reference()
    sc = H.load_scenario(os.path.join(H.GOLDEN, name))
    for run in (sc["runs"][0], sc["runs"][2]):
        prm = run["params"]
        ploidy_mgrs = {}
        for s, p in sc["ploidies"].items():
            if isinstance(p, list):
                import tempfile
                with tempfile.NamedTemporaryFile("w", suffix=".bed", delete=False) as fh:
                    for row in p:
                        fh.write("\t".join(map(str, row)) + "\n")
                ploidy_mgrs[s] = rl.PloidyManager(local_copy_numbers=fh.name)
                os.unlink(fh.name)
            else:
                ploidy_mgrs[s] = rl.PloidyManager(genome_wide_ploidy=p)
        ann = ra.Annotator.__new__(ra.Annotator)  # the reference class, its own methods
        ann.bam_readers = OrderedDict((s, [None] * len(idx)) for s, idx in sc["samples"].items())
        ann.power = rw.PowerCalculator(tumor_ploidies=ploidy_mgrs, purities=dict(sc["purities"]), normal_ploidy=prm.get("normal_ploidy", 2),
                                       fpr=prm.get("fpr", rw.DEFAULT_FPR), error_rate=prm.get("error_rate", rw.DEFAULT_ERROR_RATE))
        by_chrom = OrderedDict()
        for v in sc["variants"]:
            by_chrom.setdefault(v["chrom"], []).append(_Variant(v["chrom"], v["pos"], v["ref"], v["alts"]))
        infos = []
        for chrom, vs in by_chrom.items():
            all_metrics = {}
            for s, idx in sc["samples"].items():
                for j, bi in enumerate(idx):
                    units, counts, values = _device_records(rp, sc, sc["bams"][bi], sc["contigs"], chrom, vs, prm)
                    ours = pileups.metrics_from_device(vs, units, counts, values)  # OUR CoverageMetrics objects
                    gold = run["metrics"][bi]
                    assert {"%s:%d:%s:%s" % ((chrom,) + k) for k in ours} == {k for k in gold if k.startswith(chrom + ":")}
                    all_metrics[(s, j)] = ours
            for v in vs:
                mb = {(s, j): all_metrics[(s, j)].get((v.POS, v.REF, v.ALT[0]), EMPTY_METRICS) for s, idx in sc["samples"].items() for j in range(len(idx))}
                ann._add_stats(v, mb)  # unmodified reference code on our objects
                infos.append([[k, val] for k, val in v.INFO.items()])
        assert len(infos) == len(run["info"])
        for got, want, v in zip(infos, run["info"], sc["variants"]):
            assert [[k, str(x)] for k, x in got] == [[k, str(x)] for k, x in want], (name, prm, v)
