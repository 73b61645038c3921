#!/usr/bin/env python
"""`vafator` command line: same options, defaults and validation errors as
vafator/command_line.py:26-205 (the other three sub-commands of the reference are out of scope)."""
import argparse
import logging
from pathlib import Path

import vafator_b200
from vafator_b200.power import DEFAULT_ERROR_RATE, DEFAULT_FPR

epilog = "B200-native implementation of the VAFator annotation path (reference: TRON-Bioinformatics/vafator)"


def _validate_alignment_file_extension(alignment_file):
    suffixes = [suffix.lower() for suffix in Path(alignment_file).suffixes]
    if ".sam" in suffixes:
        raise ValueError("SAM input is not supported: {}. Please provide BAM or CRAM files.".format(alignment_file))


def build_parser():
    parser = argparse.ArgumentParser(description="vafator v{}".format(vafator_b200.VERSION),
                                     formatter_class=argparse.ArgumentDefaultsHelpFormatter, epilog=epilog)
    parser.add_argument("--input-vcf", dest="input_vcf", action="store", help="The VCF to annotate", required=True)
    parser.add_argument("--output-vcf", dest="output_vcf", action="store", help="The annotated VCF", required=True)
    parser.add_argument("--bam", action="append", nargs=2, metavar=("sample_name", "bam_file"), default=[],
                        help="A sample name and a BAM file. Can be used multiple times to input multiple samples and "
                             "multiple BAM files. The same sample name can be used multiple times with different BAMs, "
                             "this will treated as replicates.")
    parser.add_argument("--mapping-quality", dest="mapping_quality", action="store", type=int, default=1,
                        help="All reads with a mapping quality below this threshold will be filtered out")
    parser.add_argument("--base-call-quality", dest="base_call_quality", action="store", type=int, default=30,
                        help="All bases with a base call quality below this threshold will be filtered out")
    parser.add_argument("--purity", action="append", nargs=2, metavar=("sample_name", "purity"), default=[],
                        help="A sample name and a tumor purity value. Can be used multiple times to input multiple "
                             "samples in combination with --bam. If no purity is provided for a given sample the "
                             "default value is 1.0")
    parser.add_argument("--tumor-ploidy", action="append", nargs=2, metavar=("sample_name", "tumor_ploidy"), default=[],
                        help="A sample name and a tumor ploidy: a genome-wide value (eg: --tumor-ploidy primary 2) or "
                             "local copy numbers in a BED file (eg: --tumor-ploidy primary /path/to/copy_numbers.bed) "
                             "(default: 2)")
    parser.add_argument("--normal-ploidy", dest="normal_ploidy", required=False, default=2, type=int,
                        help="Normal ploidy for the power calculation (default: 2)")
    parser.add_argument("--fpr", dest="fpr", required=False, default=DEFAULT_FPR, type=float,
                        help="False Positive Rate (FPR) to use in the power calculation")
    parser.add_argument("--error-rate", dest="error_rate", required=False, default=DEFAULT_ERROR_RATE, type=float,
                        help="Error rate to use in the power calculation")
    parser.add_argument("--exclude-ambiguous-bases", dest="exclude_ambiguous_bases", action="store_true",
                        help="Flag indicating to include ambiguous bases from the DP calculation")
    parser.add_argument("--num-processes", dest="num_processes", required=False, default=1, type=int,
                        help="Accepted for compatibility with the reference CLI (CPU worker processes there)")
    parser.add_argument("--device", dest="device", required=False, default=0, type=int, help="CUDA device index")
    return parser


def annotator(argv=None):
    args = build_parser().parse_args(argv)
    logging.info("Vafator starting...")
    from vafator_b200.annotator import Annotator
    from vafator_b200.ploidies import PloidyManager

    bams = {}
    for sample_name, bam in args.bam:
        _validate_alignment_file_extension(bam)
        bams.setdefault(sample_name, []).append(bam)

    purities = {}
    for sample_name, purity in args.purity:
        if sample_name in purities:
            raise ValueError("Multiple purity values provided for sample: {}".format(sample_name))
        if sample_name not in bams:
            raise ValueError("Provided a purity value for a sample for which no BAM is provided: {}".format(sample_name))
        purities[sample_name] = float(purity)

    tumor_ploidies = {}
    for sample_name, tumor_ploidy in args.tumor_ploidy:
        if sample_name in tumor_ploidies:
            raise ValueError("Multiple tumor ploidy values provided for sample: {}".format(sample_name))
        if sample_name not in bams:
            raise ValueError(
                "Provided a tumor ploidy value for a sample for which no BAM is provided: {}".format(sample_name))
        try:
            tumor_ploidies[sample_name] = PloidyManager(genome_wide_ploidy=float(tumor_ploidy))
        except ValueError:
            tumor_ploidies[sample_name] = PloidyManager(local_copy_numbers=tumor_ploidy)

    if len(bams) == 0:
        raise ValueError("Please, provide at least one bam file with '--bam sample_name /path/to/file.bam'")

    Annotator(input_vcf=args.input_vcf, output_vcf=args.output_vcf, input_bams=bams,
              mapping_qual_thr=args.mapping_quality, base_call_qual_thr=args.base_call_quality, purities=purities,
              tumor_ploidies=tumor_ploidies, normal_ploidy=int(args.normal_ploidy), fpr=args.fpr,
              error_rate=args.error_rate, include_ambiguous_bases=(not args.exclude_ambiguous_bases),
              num_processes=args.num_processes, device=args.device).run()
    logging.info("Vafator finished!")


def multiallelics_filter(argv=None):
    """`multiallelics-filter` (vafator/command_line.py:208-247): a host-only pass over the annotated VCF."""
    import vafator_b200
    parser = argparse.ArgumentParser(description="vafator v{}".format(vafator_b200.VERSION),
                                     formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument("--input-vcf", dest="input_vcf", action="store", help="The VCF to annotate", required=True)
    parser.add_argument("--output-vcf", dest="output_vcf", action="store", help="The annotated VCF", required=True)
    parser.add_argument("--tumor-sample-name", dest="tumor_sample_name", action="store", default="tumor",
                        help="The tumor sample name (will look for annotation ${SAMPLE_NAME}_af)")
    args = parser.parse_args(argv)
    logging.info("Vafator multiallelic filter starting...")
    from vafator_b200.multiallelic_filter import MultiallelicFilter
    MultiallelicFilter(input_vcf=args.input_vcf, output_vcf=args.output_vcf, tumor_sample_name=args.tumor_sample_name).run()
    logging.info("Vafator multiallelic filter finished!")


if __name__ == "__main__":
    annotator()
