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
    # The reference's committed (stale, v2.0.0) output for that VCF: normal_ac / normal_dp per record at MQ >= 0, BQ >= 29.
    # Not a golden vector for v3, but consistent with v3's test_nist assertions at all four asserted sites:
    # kept as a 368-row AC / DP table for the day the BAM is available (tests/test_reference_known_answers.py).
    res = os.path.join(REF_ROOT, "vafator/tests/resources/results/project.NIST.hc.snps.indels.chr1_1000000_2000000.vaf.vcf")
    if os.path.exists(res):
        rows = []
        with open(res) as fh:
            for line in fh:
                if line.startswith("#"):
                    continue
                f = line.rstrip("\n").split("\t")
                info = dict(kv.split("=", 1) for kv in f[7].split(";") if "=" in kv)
                if "normal_ac" in info and "normal_dp" in info:
                    rows.append([f[0], int(f[1]), f[3], f[4], [int(x) for x in info["normal_ac"].split(",")], int(info["normal_dp"])])
        out2 = os.path.join(ROOT, "tests", "golden", "nist_v2_ac_dp.json")
        with open(out2, "w") as fh:
            json.dump({"source": "vafator/tests/resources/results/project.NIST.hc.snps.indels.chr1_1000000_2000000.vaf.vcf "
                                 "(vafator 2.0.0 output, MQ>=0 BQ>=29): CHROM, POS, REF, ALT, normal_ac, normal_dp", "rows": rows},
                      fh, separators=(",", ":"))
        print(out2, len(rows), "rows")


if __name__ == "__main__":
    main()
