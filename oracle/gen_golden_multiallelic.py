#!/usr/bin/env python
"""ORACLE tooling (test infrastructure) -- generates tests/golden/multiallelic.json by running the
UNMODIFIED reference vafator/multiallelic_filter.py on hand-made annotated VCFs.

cyvcf2 is not installable here, so the module runs on a stand-in that restates the cyvcf2 behaviour the
filter relies on (parity of the stand-in itself is unpinned):
  * Variant.INFO.get(key, default): Float values are IEEE singles widened to Python floats
    (bcf_get_info_float); String values come back as str; a missing key returns the default
  * Variant.INFO[key] = str: bcf_update_info_string -- replaces the key or appends it
  * Variant.is_snp: REF and every ALT are single bases
  * Writer(path, vcf) copies the template header including add_to_header / add_info_to_header lines
The header line carrying the date is dropped from the golden output.

Run only in the build container (needs /root/reference):   python oracle/gen_golden_multiallelic.py
"""
import json
import os
import struct
import sys
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_ROOT = os.environ.get("VAFATOR_REFERENCE", "/root/reference")

HEADER = """##fileformat=VCFv4.2
##INFO=<ID=tumor_af,Number=A,Type=Float,Description="af">
##INFO=<ID=tumor_dp,Number=1,Type=Integer,Description="dp">
##INFO=<ID=primary_af,Number=A,Type=Float,Description="af">
##contig=<ID=chr1>
##contig=<ID=chr4>
##contig=<ID=chr6>
#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO
"""


def rec(chrom, pos, ref, alt, info):
    return "\t".join([chrom, str(pos), ".", ref, alt, ".", "PASS", info])


CASES = {
    "nothing_to_filter": [rec("chr1", 100, "A", "T", "tumor_af=0.5;tumor_dp=10"), rec("chr1", 101, "A", "T", "tumor_af=0.1"),
                          rec("chr4", 100, "G", "C", ".")],
    "pairs": [rec("chr1", 10, "A", "C", "tumor_af=0.3"),
              rec("chr4", 1235, "A", "T", "tumor_af=0.1;tumor_dp=30"), rec("chr4", 1235, "A", "G", "tumor_af=0.2;tumor_dp=30"),
              rec("chr6", 1234, "C", "T", "tumor_af=0.5"),
              rec("chr6", 1235, "C", "G", "tumor_af=0.2"), rec("chr6", 1235, "C", "A", "tumor_af=0.01")],
    "different_reference_is_kept": [rec("chr1", 50, "A", "T", "tumor_af=0.2"), rec("chr1", 50, "AC", "A", "tumor_af=0.4"),
                                    rec("chr1", 50, "G", "T", "tumor_af=0.4")],
    "three_alleles": [rec("chr1", 7, "A", "C", "tumor_af=0.11111"), rec("chr1", 7, "A", "G", "tumor_af=0.33333"),
                      rec("chr1", 7, "A", "T", "tumor_af=0.22222")],
    "equal_af_random_choice": [rec("chr1", 7, "A", "C", "tumor_af=0.25"), rec("chr1", 7, "A", "G", "tumor_af=0.25"),
                               rec("chr1", 7, "A", "T", "tumor_af=0.25"),
                               rec("chr1", 9, "T", "C", "tumor_af=0.125"), rec("chr1", 9, "T", "G", "tumor_af=0.125")],
    "missing_af_defaults_to_zero": [rec("chr1", 7, "A", "C", "tumor_dp=3"), rec("chr1", 7, "A", "G", "tumor_af=0.00001"),
                                    rec("chr1", 8, "A", "C", "."), rec("chr1", 8, "A", "G", "tumor_dp=1")],
    "indel_after_snv_is_not_merged": [rec("chr1", 7, "A", "C", "tumor_af=0.2"), rec("chr1", 7, "A", "AT", "tumor_af=0.3"),
                                      rec("chr1", 7, "A", "G", "tumor_af=0.1")],
    "other_sample_name": [rec("chr1", 7, "A", "C", "tumor_af=0.9;primary_af=0.1"), rec("chr1", 7, "A", "G", "tumor_af=0.1;primary_af=0.7")],
    "existing_annotation_is_extended": [rec("chr1", 7, "A", "C", "tumor_af=0.9;multiallelic=,T,0.05"), rec("chr1", 7, "A", "G", "tumor_af=0.1")],
}
SAMPLE = {"other_sample_name": "primary"}


def f32(text):
    return struct.unpack("f", struct.pack("f", float(text)))[0]


def install_fake_cyvcf2():
    class Info:
        def __init__(self, col, types_):
            self.items = [] if col in (".", "") else [kv.partition("=")[::2] for kv in col.split(";")]
            self.types = types_

        def get(self, key, default=None):
            for k, v in self.items:
                if k == key:
                    if self.types.get(k) == "Float":
                        vals = tuple(f32(t) for t in v.split(","))
                        return vals[0] if len(vals) == 1 else vals
                    if self.types.get(k) == "Integer":
                        return int(v)
                    return v
            return default

        def __setitem__(self, key, value):
            for i, (k, _) in enumerate(self.items):
                if k == key:
                    self.items[i] = (key, str(value))
                    return
            self.items.append((key, str(value)))

        def text(self):
            return ";".join(k if v == "" and self.types.get(k) == "Flag" else "%s=%s" % (k, v) for k, v in self.items) or "."

    class Variant:
        def __init__(self, line, types_):
            self.f = line.split("\t")
            self.CHROM, self.POS, self.REF = self.f[0], int(self.f[1]), self.f[3]
            self.ALT = [] if self.f[4] == "." else self.f[4].split(",")
            self.INFO = Info(self.f[7], types_)

        @property
        def is_snp(self):
            return len(self.REF) == 1 and len(self.ALT) > 0 and all(len(a) == 1 for a in self.ALT)

        def line(self):
            f = list(self.f)
            f[7] = self.INFO.text()
            return "\t".join(f)

    class VCF:
        def __init__(self, path):
            self.lines = open(path).read().split("\n")
            self.header = [l for l in self.lines if l.startswith("##")]
            self.cols = [l for l in self.lines if l.startswith("#CHROM")][0]
            self.extra = []
            self.types = {}
            for l in self.header:
                if l.startswith("##INFO=<ID="):
                    self.types[l.split("ID=")[1].split(",")[0]] = l.split("Type=")[1].split(",")[0]

        def add_to_header(self, line):
            self.extra.append(line)

        def add_info_to_header(self, d):
            self.types[d["ID"]] = d["Type"]
            self.extra.append('##INFO=<ID={ID},Number={Number},Type={Type},Description="{Description}">'.format(**d))

        def __iter__(self):
            for l in self.lines:
                if l and not l.startswith("#"):
                    yield Variant(l, self.types)

        def close(self):
            pass

    class Writer:
        def __init__(self, path, vcf):
            self.fh = open(path, "w")
            self.fh.write("\n".join(vcf.header + vcf.extra + [vcf.cols]) + "\n")

        def write_record(self, v):
            self.fh.write(v.line() + "\n")

        def close(self):
            self.fh.close()

    cy = types.ModuleType("cyvcf2")
    cy.VCF, cy.Writer, cy.Variant = VCF, Writer, Variant
    sys.modules["cyvcf2"] = cy


def main():
    install_fake_cyvcf2()
    sys.path.insert(0, REF_ROOT)
    from vafator.multiallelic_filter import MultiallelicFilter  # the reference, as it lies
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, records in CASES.items():
            src, dst = os.path.join(tmp, name + ".vcf"), os.path.join(tmp, name + ".out.vcf")
            text = HEADER + "\n".join(records) + "\n"
            open(src, "w").write(text)
            MultiallelicFilter(input_vcf=src, output_vcf=dst, tumor_sample_name=SAMPLE.get(name, "tumor")).run()
            got = [l for l in open(dst).read().split("\n") if l and not l.startswith("##vafator_command_line=")]
            out[name] = {"input": text, "sample": SAMPLE.get(name, "tumor"), "output": got}
    path = os.path.join(ROOT, "tests", "golden", "multiallelic.json")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path, {k: len(v["output"]) for k, v in out.items()})


if __name__ == "__main__":
    main()
