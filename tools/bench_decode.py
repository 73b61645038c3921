"""Host-side timing of the file path around the kernels (no GPU): BGZF/BAM decode into the canonical columnar batch -- whole
file, one region, a tenth of a contig -- on a synthetic coordinate-sorted BAM written here.  (Everything derived from the
records is the kernels' work now: there is no host packing or mate linking left to time.)

    python tools/bench_decode.py [--mbp 3] [--threads N]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vafator_b200 import bamio, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mbp", type=float, default=3.0, help="contig length of the 30x WGS-like BAM, in Mbp")
    ap.add_argument("--out", default="/tmp/vfk_bench_decode.bam")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    cfg = synth.wgs(n_contigs=1, contig_len=int(args.mbp * 1e6))
    batch = synth.reads(cfg)
    raw = bamio.write_bam_batch(args.out, batch, [int(args.mbp * 1e6) + 1000])
    size = os.path.getsize(args.out)
    print("BAM: %d reads, %.1f MB raw, %.1f MB on disk, %d host threads" % (batch.n_reads, raw / 1e6, size / 1e6, os.cpu_count()))

    def best(f):
        ts = []
        for _ in range(args.reps):
            t = time.perf_counter()
            r = f()
            ts.append(time.perf_counter() - t)
        return min(ts), r

    t, got = best(lambda: bamio.BamFile(args.out, use_index=False).read_batch())
    assert got.n_reads == batch.n_reads and np.array_equal(got.pos, batch.pos) and np.array_equal(got.qual, batch.qual)
    print("decode          %.3f s  %.2f M reads/s  %.0f MB/s inflated" % (t, got.n_reads / t / 1e6, raw / t / 1e6))
    L = int(args.mbp * 1e6)
    t, reg = best(lambda: bamio.BamFile(args.out).read_batch([(batch.contigs[0], 0, L + 1000)]))
    assert reg.n_reads == batch.n_reads and np.array_equal(reg.qual, batch.qual)
    print("decode (.bai, whole contig as one region)  %.3f s  %.2f M reads/s" % (t, reg.n_reads / t / 1e6))
    t, reg = best(lambda: bamio.BamFile(args.out).read_batch([(batch.contigs[0], L // 2, L // 2 + L // 10)]))
    print("decode (.bai, a tenth of the contig)       %.3f s  %.2f M reads/s (%d reads)" % (t, reg.n_reads / t / 1e6, reg.n_reads))


if __name__ == "__main__":
    main()
