"""Annotator -- same constructor, same `run()`, same output contract as vafator/annotator.py:68-500;
the three arithmetic call sites (collect_metrics_for_chrom, get_rank_sum_tests, PowerCalculator.*)
are served by the CUDA library through vafator_b200.engine.Engine."""
from __future__ import annotations

import datetime
import json
import os
from collections import OrderedDict

import numpy as np

import vafator_b200
from .bamio import BamFile
from .constants import BATCH_SIZE, _HEADER_TEMPLATES, _REPLICATE_HEADER_TEMPLATES
from .engine import Engine
from .fetch import fetch_group
from .ploidies import DEFAULT_PLOIDY
from .power import DEFAULT_ERROR_RATE, DEFAULT_FPR, PowerCalculator
from .vcf import VcfReader, VcfWriter


GROUP_CHUNK = 20000  # records of a chromosome group per engine call: bounds the reads held at once and lets decoding overlap annotation


class Annotator(object):

    def __init__(self, input_vcf: str, output_vcf: str, input_bams: dict, purities: dict = None,
                 mapping_qual_thr: int = 0, base_call_qual_thr: int = 29, tumor_ploidies: dict = None,
                 normal_ploidy: int = 2, fpr: float = DEFAULT_FPR, error_rate: float = DEFAULT_ERROR_RATE,
                 include_ambiguous_bases: bool = True, num_processes: int = 1, device: int = 0):
        """Arguments as vafator/annotator.py:70-98.  `num_processes`: the reference forks that many CPU
        workers, one chromosome each; here it is the number of GPUs (from `device` upwards, as many as the
        box has) the interval x BAM shards are spread over.  `device`: first CUDA device index."""
        purities = {} if purities is None else purities
        tumor_ploidies = {} if tumor_ploidies is None else tumor_ploidies
        self.mapping_quality_threshold = mapping_qual_thr
        self.base_call_quality_threshold = base_call_qual_thr
        self.purities = purities
        self.tumor_ploidies = tumor_ploidies
        self.normal_ploidy = normal_ploidy
        self.include_ambiguous_bases = include_ambiguous_bases
        self.num_processes = num_processes
        self.power = PowerCalculator(normal_ploidy=normal_ploidy, tumor_ploidies=tumor_ploidies, purities=purities,
                                     error_rate=error_rate, fpr=fpr, device_index=device)
        self.vcf = VcfReader(input_vcf)
        # provenance header, key for key as vafator/annotator.py:123-154
        self.header = {
            "name": "vafator",
            "version": vafator_b200.VERSION,
            "date": datetime.datetime.now().ctime(),
            "timestamp": datetime.datetime.now().timestamp(),
            "input_vcf": os.path.abspath(input_vcf),
            "output_vcf": os.path.abspath(output_vcf),
            "bams": ";".join("{}:{}".format(s, ",".join(os.path.abspath(b) for b in bams)) for s, bams in input_bams.items()),
            "mapping_quality_threshold": mapping_qual_thr,
            "base_call_quality_threshold": base_call_qual_thr,
            "purities": ";".join("{}:{}".format(s, p) for s, p in purities.items()),
            "normal_ploidy": normal_ploidy,
            "tumor_ploidy": (";".join("{}:{}".format(s, p.report_value) for s, p in tumor_ploidies.items())
                             if tumor_ploidies else DEFAULT_PLOIDY),
            "include_ambiguous_bases": include_ambiguous_bases,
        }
        self.vcf.add_to_header("##vafator_command_line={}".format(json.dumps(self.header)))
        for a in Annotator._get_headers(input_bams):
            self.vcf.add_info_to_header(a)
        self.vcf_writer = VcfWriter(output_vcf, self.vcf)
        self.bam_paths = input_bams
        self.bam_readers = OrderedDict((s, [BamFile(b, "r") for b in bams]) for s, bams in input_bams.items())
        import torch
        n_dev = max(1, min(int(num_processes), torch.cuda.device_count() - device)) if torch.cuda.is_available() else 1
        self._single_device = n_dev == 1
        # cumulative seconds per stage of run(): fetch (decode, background thread), annotate (device stages +
        # INFO formatting), write; wall = the whole run
        self.timings = {"fetch": 0.0, "annotate": 0.0, "write": 0.0, "wall": 0.0, "records": 0, "reads": 0}
        self.engine = Engine(device, devices=list(range(device, device + n_dev)), min_mapq=mapping_qual_thr, min_baseq=base_call_qual_thr,
                             include_ambiguous=include_ambiguous_bases, fpr=fpr, error_rate=error_rate)

    def run(self) -> None:
        """Annotate every record of the input VCF and write the output VCF.

        The VCF is streamed one chromosome group at a time -- consecutive records of one CHROM, the unit the
        reference piles up (stream_variants_by_chrom, vafator/pileups.py:83-103) -- and a group in chunks of GROUP_CHUNK
        records, so the host never holds more than the reads of two chunks: with an indexed BAM only the reads the chunk's
        columns depend on are decoded (windows around the variants inside [first.POS-1, last.POS), the span the reference
        fetches for the group; fetch.py) and the batch is dropped, on the host and on the device, when the chunk is
        written; the next chunk is decoded in the background meanwhile (a single-chromosome VCF overlaps decoding and
        annotation too).  A BAM without an index is decoded once and stays resident."""
        import time
        from concurrent.futures import ThreadPoolExecutor
        t_run = time.perf_counter()
        with ThreadPoolExecutor(1) as pool:  # decodes chunk k+1 (native code, GIL released) while chunk k is annotated
            pending = None
            for group in self._chromosome_groups():
                span = (group[0].POS, group[-1].POS)  # the span the reference piles up for the whole group
                for at in range(0, len(group), GROUP_CHUNK):
                    chunk = group[at:at + GROUP_CHUNK]
                    fetched = pool.submit(self._fetch_group, chunk, span)
                    if pending is not None:
                        self._annotate_group(pending[0], pending[1].result(), pending[2])
                    pending = (chunk, fetched, span)
            if pending is not None:
                self._annotate_group(pending[0], pending[1].result(), pending[2])
        self.vcf_writer.close()
        self.vcf.close()
        self.engine.release_reads()
        for readers in self.bam_readers.values():
            for bam in readers:
                bam.close()
        self.timings["wall"] += time.perf_counter() - t_run

    def _chromosome_groups(self):
        group = []
        for v in self.vcf:
            if not v.ALT:
                raise ValueError("VCF record %s:%d has no ALT allele" % (v.CHROM, v.POS))
            if group and v.CHROM != group[0].CHROM:
                yield group
                group = []
            group.append(v)
        if group:
            yield group

    def _fetch_group(self, variants: list, span) -> "OrderedDict":
        """The read batches of one chunk of a chromosome group: windows around its variants inside the span the
        reference fetches for the group, [first.POS-1, last.POS) (vafator/pileups.py:137-138); a dense chunk gets the
        stretch from its first to its last window (vafator_b200/fetch.py)."""
        import time
        t0 = time.perf_counter()
        chrom = variants[0].CHROM
        positions = [v.POS for v in variants]
        bams = OrderedDict((s, [fetch_group(r, chrom, span[0], span[1], positions, self.engine.params) for r in readers])
                           for s, readers in self.bam_readers.items())
        self.timings["fetch"] += time.perf_counter() - t0
        self.timings["reads"] += sum(b.n_reads for batches in bams.values() for b in batches)
        return bams

    def _annotate_group(self, variants: list, bams: "OrderedDict", span) -> None:
        """One chunk of a chromosome group: annotate, write, release its reads."""
        records = [(v.CHROM, v.POS, v.REF, v.ALT) for v in variants]
        chrom = records[0][0]
        positions = [r[1] for r in records]
        import time
        t0 = time.perf_counter()
        eaf = {s: self.power.expected_vafs(s, [chrom] * len(records), positions) for s in bams}
        infos = self.engine.annotate(records, bams, eaf, as_text=True, span=span)
        t1 = time.perf_counter()
        batch = []
        for v, info in zip(variants, infos):
            v.INFO = info
            batch.append(v)
            if len(batch) >= BATCH_SIZE:
                self._write_batch(batch)
                batch.clear()
        if batch:
            self._write_batch(batch)
        # region batches are transient (a new object per call); the whole-file batch of an unindexed BAM is
        # cached by its reader and keeps its device copy for the next group
        if any(r.index_path for readers in self.bam_readers.values() for r in readers):
            self.engine.release_reads([b for batches in bams.values() for b in batches])
        self.timings["annotate"] += t1 - t0
        self.timings["write"] += time.perf_counter() - t1
        self.timings["records"] += len(records)

    def _write_batch(self, batch: list) -> None:
        for v in batch:
            self.vcf_writer.write_record(v)

    @staticmethod
    def _get_headers(input_bams: dict) -> list:
        """INFO header entries for all samples and replicates (vafator/annotator.py:443-468)."""
        headers = []
        for s, bams in input_bams.items():
            for suffix, description, typ, number in _HEADER_TEMPLATES:
                headers.append(Annotator._make_header(suffix, description, typ, number, sample=s))
            if len(bams) > 1:
                for i in range(1, len(bams) + 1):
                    for suffix, description, typ, number in _REPLICATE_HEADER_TEMPLATES:
                        headers.append(Annotator._make_header(suffix, description, typ, number, sample=s, index=i))
        return headers

    @staticmethod
    def _make_header(suffix: str, description: str, typ: str, number: str, sample: str, index: int = None) -> dict:
        header_id = "{}_{}".format(sample, suffix)
        if index is not None:
            header_id = "{}_{}".format(header_id, index)
        return {"ID": header_id, "Description": description.format(sample=sample), "Type": typ, "Number": number}
