"""File-level end-to-end timing: `Annotator.run()` on synthetic BAM + VCF files -- BAI fetch, BGZF/BAM decode,
packing, the device stages, INFO formatting, VCF writing -- with the stage split the Annotator records.

    python tools/bench_files.py [--mbp 20] [--contigs 2] [--every 1] [--dir /tmp/vfk_files]

--every N keeps every N-th planted variant: 1 = the dense WGS-like set (one record per ~720 bp, the group fetches its
whole span), 100+ = a sparse somatic-like set (windows around the variants).  Needs a GPU; prints one JSON line.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vafator_b200 import bamio, synth  # noqa: E402


def write_inputs(args):
    os.makedirs(args.dir, exist_ok=True)
    length = int(args.mbp * 1e6)
    paths = {}
    for sample, normal in (("tumor", False), ("normal", True)):
        cfg = synth.wgs(n_contigs=args.contigs, contig_len=length, normal_vaf=int(normal), read_salt=7 if normal else 0)
        p = os.path.join(args.dir, "%s_%dx%gMbp.bam" % (sample, args.contigs, args.mbp))
        if not os.path.exists(p + ".bai"):
            batch = synth.reads(cfg)
            bamio.write_bam_batch(p, batch, [length + 1000] * args.contigs, level=args.level)
        paths[sample] = p
    cfg = synth.wgs(n_contigs=args.contigs, contig_len=length)
    recs = synth.variant_records(cfg)[::args.every]
    vcf = os.path.join(args.dir, "in_%d.vcf" % args.every)
    with open(vcf, "w") as fh:
        fh.write("##fileformat=VCFv4.2\n" + "".join("##contig=<ID=%s>\n" % c for c in cfg.contigs))
        fh.write("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n")
        for c, pos, ref, alts in recs:
            fh.write("%s\t%d\t.\t%s\t%s\t.\tPASS\tSOMATIC\n" % (c, pos, ref, ",".join(alts)))
    return paths, vcf, len(recs)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mbp", type=float, default=20.0, help="length of each contig of the 30x BAMs, in Mbp")
    ap.add_argument("--contigs", type=int, default=2)
    ap.add_argument("--every", type=int, default=1)
    ap.add_argument("--level", type=int, default=6, help="deflate level of the BAMs")
    ap.add_argument("--dir", default="/tmp/vfk_files")
    args = ap.parse_args()
    t = time.perf_counter()
    paths, vcf, n_records = write_inputs(args)
    t_inputs = time.perf_counter() - t
    from vafator_b200.annotator import Annotator
    out = os.path.join(args.dir, "out_%d.vcf" % args.every)
    t = time.perf_counter()
    a = Annotator(input_vcf=vcf, output_vcf=out, input_bams={"tumor": [paths["tumor"]], "normal": [paths["normal"]]},
                  mapping_qual_thr=1, base_call_qual_thr=30)
    t_open = time.perf_counter() - t
    a.run()
    tm = a.timings
    units = tm["records"] * 2
    print(json.dumps({
        "what": "file-level end to end: vafator_b200.annotator.Annotator.run() on BAM + VCF files",
        "records": tm["records"], "bams": 2, "reads_decoded": tm["reads"], "bam_bytes": sum(os.path.getsize(p) for p in paths.values()),
        "wall_s": round(tm["wall"], 3), "units_per_s": round(units / tm["wall"], 1),
        "stage_s": {"open": round(t_open, 3), "fetch_background": round(tm["fetch"], 3), "annotate": round(tm["annotate"], 3),
                    "write": round(tm["write"], 3)},
        "inputs_s": round(t_inputs, 1), "every": args.every, "host_cores": os.cpu_count()}))


if __name__ == "__main__":
    main()
