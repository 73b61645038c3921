#!/usr/bin/env python
"""ORACLE tooling (test infrastructure) -- extracts the variant set of the reference's integration test
(tests/integration_test_01.sh: vafator/tests/resources/project.NIST.hc.snps.indels.chr1_1000000_2000000.vcf, 368
biallelic records: 310 SNV / 24 INS / 34 DEL, REF up to 18 bases, ALT up to 40) into tests/golden/nist_variants.json:
CHROM / POS / REF / ALT only.  The BAM of that test is absent from the checkout (.MISSING_LARGE_BLOBS), so
tests/test_integration_nist.py pairs the real variant set with simulated reads.

Run only in the build container (needs /root/reference):   python oracle/gen_golden_nist.py
"""
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_ROOT = os.environ.get("VAFATOR_REFERENCE", "/root/reference")
SRC = os.path.join(REF_ROOT, "vafator/tests/resources/project.NIST.hc.snps.indels.chr1_1000000_2000000.vcf")


def main():
    records = []
    with open(SRC) as fh:
        for line in fh:
            if line.startswith("#"):
                continue
            f = line.rstrip("\n").split("\t")
            records.append([f[0], int(f[1]), f[3], f[4].split(",")])
    out = os.path.join(ROOT, "tests", "golden", "nist_variants.json")
    with open(out, "w") as fh:
        json.dump({"source": "vafator/tests/resources/project.NIST.hc.snps.indels.chr1_1000000_2000000.vcf (CHROM, POS, REF, ALT)",
                   "records": records}, fh, separators=(",", ":"))
    print(out, len(records), "records")


if __name__ == "__main__":
    main()
