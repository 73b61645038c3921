"""`multiallelics-filter`: keep one SNV per (CHROM, POS, REF) of an annotated VCF -- the one with the highest
tumour allele frequency -- and record the dropped alternatives in INFO/multiallelic.

Host-side mirror of vafator/multiallelic_filter.py:8-130 on top of vafator_b200.vcf (cyvcf2 is not
available here).  Behaviour kept on purpose:
  * only the CURRENT record has to be a SNV (REF and every ALT one base long); it is merged into the pending
    record when CHROM, POS and REF agree                        (multiallelic_filter.py:53-58)
  * `<sample>_af` is read the way cyvcf2 hands it out: an IEEE single widened to a Python float (0.1 comes
    back as 0.10000000149011612), missing -> 0.0               (multiallelic_filter.py:126-127)
  * ties are broken by random.Random(42).sample([current, pending], k=1)  (multiallelic_filter.py:43, :69-76)
  * the annotation is ",".join(old.split(",") + [alt, str(af)]) with old = "" when absent: a first
    annotation therefore starts with a comma                   (multiallelic_filter.py:120-124)
Not reproduced: htslib re-serialises every Float INFO value it parsed (%g, six significant digits); this
writer passes the untouched columns through as text.  An input without records yields a header-only VCF
(the reference crashes writing None)."""
from __future__ import annotations

import datetime
import json
import random
from typing import List, Optional

import numpy as np

from . import VERSION
from .vcf import VcfReader, VcfRecord, VcfWriter


def _info_value(rec: VcfRecord, key: str) -> Optional[str]:
    """Pending annotation first, then the INFO column as read."""
    if key in rec.INFO:
        return str(rec.INFO[key])
    col = rec.fields[7]
    if col in (".", ""):
        return None
    for kv in col.split(";"):
        k, _, v = kv.partition("=")
        if k == key:
            return v
    return None


def _as_cyvcf2_float(text: str):
    """Float INFO value as cyvcf2 returns it: float32 precision; several values come back as a tuple."""
    vals = tuple(float(np.float32(t)) for t in text.split(","))
    return vals[0] if len(vals) == 1 else vals


def _is_snp(rec: VcfRecord) -> bool:
    return len(rec.REF) == 1 and len(rec.ALT) > 0 and all(len(a) == 1 and a in "ACGTN" for a in rec.ALT)


class MultiallelicFilter(object):
    multiallelic_annotation_name = "multiallelic"

    def __init__(self, input_vcf, output_vcf, tumor_sample_name="tumor"):
        self.tumor_sample_name = tumor_sample_name
        self.vcf = VcfReader(input_vcf)
        now = datetime.datetime.now()
        header = {"name": "multiallelic-filter", "version": VERSION, "date": now.ctime(), "timestamp": now.timestamp(),
                  "input_vcf": input_vcf, "output_vcf": output_vcf}
        self.vcf.add_to_header("##vafator_command_line={}".format(json.dumps(header)))
        self.vcf.add_info_to_header({
            "ID": self.multiallelic_annotation_name, "Type": "String", "Number": ".",
            "Description": "Indicates multiallelic variants filtered and their frequencies if any (e.g.: T,0.12)"})
        self.vcf_writer = VcfWriter(output_vcf, self.vcf)
        self.random = random.Random(42)

    def run(self):
        batch: List[VcfRecord] = []
        prev = None
        for variant in self.vcf:
            if prev is None:
                prev = variant
                continue
            if _is_snp(variant) and variant.CHROM == prev.CHROM and variant.POS == prev.POS and variant.REF == prev.REF:
                af1, af2 = self.get_tumor_af(prev), self.get_tumor_af(variant)
                if af1 < af2:
                    self.set_multiallelic_annotation(variant, prev.ALT[0], af1)
                    prev = variant
                elif af1 > af2:
                    self.set_multiallelic_annotation(prev, variant.ALT[0], af2)
                elif af1 == af2:
                    alt1, alt2 = variant.ALT[0], prev.ALT[0]
                    prev = self.random.sample([variant, prev], k=1)[0]
                    self.set_multiallelic_annotation(prev, alt1 if alt1 != prev.ALT[0] else alt2, af1)
                continue
            batch.append(prev)
            prev = variant
            if len(batch) >= 1000:
                self._write_batch(batch)
                batch = []
        if prev is not None:
            batch.append(prev)
        self._write_batch(batch)
        self.vcf_writer.close()
        self.vcf.close()

    def set_multiallelic_annotation(self, variant, alt, af):
        old = _info_value(variant, self.multiallelic_annotation_name)
        variant.INFO[self.multiallelic_annotation_name] = ",".join(("" if old is None else old).split(",") + [alt, str(af)])

    def get_tumor_af(self, variant):
        text = _info_value(variant, "{}_af".format(self.tumor_sample_name))
        return 0.0 if text is None else _as_cyvcf2_float(text)

    def _write_batch(self, batch):
        for v in batch:
            self.vcf_writer.write_record(v)
