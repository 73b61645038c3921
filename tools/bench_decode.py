"""Host-side timing of the file path around the kernels (no GPU): BGZF/BAM decode, packing into the HBM
layout, mate linking, column-store build -- on a synthetic coordinate-sorted BAM written here.

    python tools/bench_decode.py [--mbp 3] [--threads N]
"""
import argparse
import os
import struct
import sys
import time
import zlib

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vafator_b200 import bamio, device, synth  # noqa: E402


def write_bam_from_batch(path, b, lengths):
    """ReadBatch -> BAM (+ .bai through the fixture writer's index code is not needed here)."""
    text = "@HD\tVN:1.6\tSO:coordinate\n" + "".join("@SQ\tSN:%s\tLN:%d\n" % (c, l) for c, l in zip(b.contigs, lengths))
    out = bytearray(b"BAM\1" + struct.pack("<i", len(text)) + text.encode() + struct.pack("<i", len(b.contigs)))
    for c, l in zip(b.contigs, lengths):
        out += struct.pack("<i", len(c) + 1) + c.encode() + b"\0" + struct.pack("<i", int(l))
    cig, seq, qual = b.cigar.tobytes(), b.seq.tobytes(), b.qual.tobytes()
    rec_at = np.zeros(b.n_reads + 1, np.int64)
    for i in range(b.n_reads):
        rec_at[i] = len(out)
        name = b"q%x\0" % int(b.qname_id[i])
        nc, lq = int(b.n_cigar[i]), int(b.l_qseq[i])
        co, so, qo = int(b.cigar_off[i]) * 4, int(b.seq_off[i]), int(b.qual_off[i])
        body = struct.pack("<iiBBHHHiiii", int(b.tid[i]), int(b.pos[i]), len(name), int(b.mapq[i]), 4680, nc, int(b.flag[i]), lq,
                           int(b.mtid[i]), int(b.mpos[i]), int(b.isize[i]))
        body += name + cig[co:co + 4 * nc] + seq[so:so + (lq + 1) // 2] + qual[qo:qo + lq]
        out += struct.pack("<i", len(body)) + body
    rec_at[b.n_reads] = len(out)
    block_at = []
    with open(path, "wb") as fh:
        for i in range(0, len(out), 0xFF00):
            block_at.append(fh.tell())
            data = bytes(out[i:i + 0xFF00])
            comp = zlib.compressobj(1, zlib.DEFLATED, -15)
            z = comp.compress(data) + comp.flush()
            fh.write(b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", len(z) + 25) + z +
                     struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data)))
        block_at.append(fh.tell())
        fh.write(bamio._BGZF_EOF)
    # .bai (SAMv1 5.2): binning index with one chunk per bin + linear index
    block_at = np.asarray(block_at, np.int64)
    voff = (block_at[rec_at // 0xFF00] << 16) | (rec_at % 0xFF00)
    from vafator_b200.sharding import _ref_span
    end = b.pos.astype(np.int64) + np.maximum(_ref_span(b), 1)
    bai = bytearray(b"BAI\1" + struct.pack("<i", len(b.contigs)))
    for t in range(len(b.contigs)):
        idx = np.nonzero(b.tid == t)[0]
        if len(idx) == 0:
            bai += struct.pack("<ii", 0, 0)
            continue
        beg_t, end_t = b.pos[idx].astype(np.int64), end[idx] - 1
        bins = np.zeros(len(idx), np.int64)
        done = np.zeros(len(idx), bool)
        for shift, base in ((14, 4681), (17, 585), (20, 73), (23, 9), (26, 1)):  # reg2bin, SAMv1 5.3
            hit = ~done & ((beg_t >> shift) == (end_t >> shift))
            bins[hit] = base + (beg_t[hit] >> shift)
            done |= hit
        # chunks: runs of consecutive records filed under the same bin
        starts = np.nonzero(np.concatenate([[True], bins[1:] != bins[:-1]]))[0]
        stops = np.concatenate([starts[1:], [len(idx)]])
        order = np.argsort(bins[starts], kind="stable")
        ub, first_of = np.unique(bins[starts][order], return_index=True)
        bai += struct.pack("<i", len(ub))
        for k, bn in enumerate(ub):
            runs = order[first_of[k]:(first_of[k + 1] if k + 1 < len(ub) else len(order))]
            bai += struct.pack("<Ii", int(bn), len(runs))
            for rj in runs:
                bai += struct.pack("<QQ", int(voff[idx[starts[rj]]]), int(voff[idx[stops[rj] - 1] + 1]))
        n_win = int((end[idx].max() - 1) >> 14) + 1
        linear = np.zeros(n_win, np.int64)
        first = np.full(n_win, -1, np.int64)
        w0, w1 = b.pos[idx] >> 14, (end[idx] - 1) >> 14
        for k in range(int((w1 - w0).max()) + 1):  # reads span few windows
            w = np.minimum(w0 + k, w1)
            order = np.argsort(w, kind="stable")
            ww, ii = w[order], idx[order]
            keep = np.concatenate([[True], ww[1:] != ww[:-1]])
            cand = np.full(n_win, np.iinfo(np.int64).max)
            cand[ww[keep]] = ii[keep]
            first = np.where((first < 0) | (cand < first), np.where(cand == np.iinfo(np.int64).max, first, cand), first)
        linear = np.where(first >= 0, voff[np.maximum(first, 0)], 0)
        for w in range(1, n_win):
            if linear[w] == 0:
                linear[w] = linear[w - 1]
        bai += struct.pack("<i", n_win) + linear.astype("<u8").tobytes()
    with open(path + ".bai", "wb") as fh:
        fh.write(bytes(bai))
    return len(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mbp", type=float, default=3.0, help="contig length of the 30x WGS-like BAM, in Mbp")
    ap.add_argument("--out", default="/tmp/vfk_bench_decode.bam")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    cfg = synth.wgs(n_contigs=1, contig_len=int(args.mbp * 1e6))
    batch = synth.reads(cfg)
    raw = write_bam_from_batch(args.out, batch, [int(args.mbp * 1e6) + 1000])
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
    params = device.make_params(min_mapq=1, min_baseq=30)
    t, mate = best(lambda: device.link_mates(got, params))
    print("link_mates      %.3f s  %.2f M reads/s" % (t, got.n_reads / t / 1e6))
    t, _ = best(lambda: device.pack_reads(got, params, mate=mate, columns=False))
    print("pack            %.3f s  %.2f M reads/s" % (t, got.n_reads / t / 1e6))
    t, _ = best(lambda: device.pack_reads(got, params, mate=mate, columns=True))
    print("pack + columns  %.3f s  %.2f M reads/s" % (t, got.n_reads / t / 1e6))


if __name__ == "__main__":
    main()
