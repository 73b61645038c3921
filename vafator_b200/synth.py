"""Seeded synthetic read batches and VCF records -- binding of the generator in
libvfkio.so (csrc/vfkio_synth.cpp).  Used by the benchmark and by the large parity tests; the
reference has no equivalent (its tests ship real BAMs, three of which are missing)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field, replace
from typing import List, Tuple

import numpy as np

from .batch import ReadBatch
from . import device as _dev

BASE_SEED = 20261019  # seed = 20261019 + config number


class SynthCfgC(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("n_contigs", C.c_int32), ("read_len", C.c_int32),
                ("tile_offset", C.c_int32), ("tile_len", C.c_int32), ("tile_stride", C.c_int32),
                ("tiles_per_contig", C.c_int32), ("pairs_per_tile", C.c_int32),
                ("var_lo", C.c_int32), ("var_hi", C.c_int32), ("snv_period", C.c_int32),
                ("indel_period", C.c_int32), ("multiallelic_pct", C.c_int32), ("normal_vaf", C.c_int32),
                ("frag_mean", C.c_float), ("frag_sd", C.c_float),
                ("p_softclip", C.c_float), ("p_del", C.c_float), ("p_ins", C.c_float), ("p_skip", C.c_float),
                ("p_dup", C.c_float), ("p_supp", C.c_float), ("p_secondary", C.c_float), ("p_qcfail", C.c_float),
                ("p_orphan", C.c_float), ("p_mapq60", C.c_float), ("base_error", C.c_float), ("p_n", C.c_float),
                ("read_salt", C.c_uint64), ("g_first", C.c_int64), ("g_count", C.c_int64)]


class _PlanC(C.Structure):
    _fields_ = [("n_reads", C.c_int64), ("n_cigar", C.c_int64), ("n_seq_bytes", C.c_int64), ("n_qual_bytes", C.c_int64)]


@dataclass
class SynthConfig:
    """Python view of vfkio_synth_cfg; defaults are the common mix of SURVEY.md 8(d)."""
    seed: int = BASE_SEED
    n_contigs: int = 1
    read_len: int = 150
    tile_offset: int = 10000
    tile_len: int = 250
    tile_stride: int = 15000
    tiles_per_contig: int = 1000
    pairs_per_tile: int = 36
    var_lo: int = 150
    var_hi: int = 400
    snv_period: int = 500
    indel_period: int = 0
    multiallelic_pct: int = 0
    normal_vaf: int = 0
    frag_mean: float = 300.0
    frag_sd: float = 50.0
    p_softclip: float = 0.03
    p_del: float = 0.025
    p_ins: float = 0.02
    p_skip: float = 0.005
    p_dup: float = 0.02
    p_supp: float = 0.005
    p_secondary: float = 0.003
    p_qcfail: float = 0.002
    p_orphan: float = 0.01
    p_mapq60: float = 0.85
    base_error: float = 0.002
    p_n: float = 0.001
    read_salt: int = 0  # changes the reads, not the genome / planted sites: several samples over one variant set
    # generate only the tiles [g_first, g_first + g_count) of the genome (global tile index = contig * tiles_per_contig +
    # tile; g_count 0 = everything): an interval shard of the SAME input, read for read
    g_first: int = 0
    g_count: int = 0

    def c(self) -> SynthCfgC:
        return SynthCfgC(*[getattr(self, f[0]) for f in SynthCfgC._fields_])

    @property
    def contigs(self) -> List[str]:
        return ["chr%d" % (i + 1) for i in range(self.n_contigs)]

    @property
    def n_tiles_generated(self) -> int:
        return self.g_count if self.g_count > 0 else self.n_contigs * self.tiles_per_contig


def wes(n_tiles=200_000, seed=BASE_SEED + 2, normal=False, **kw) -> SynthConfig:
    """C2: 30x WES-like: `n_tiles` 250-bp targets (200 000 = 50 Mbp), ~1 SNV per 500 target bases."""
    per = min(n_tiles, 10_000)
    n_contigs = (n_tiles + per - 1) // per
    return SynthConfig(seed=seed, n_contigs=n_contigs, tiles_per_contig=per, normal_vaf=int(normal), **kw)


def wgs(n_contigs=1, contig_len=10_000_000, seed=BASE_SEED + 3, **kw) -> SynthConfig:
    """C3: 30x WGS-like contiguous coverage; 1 variant per ~720 bp (90 % SNV, 5 % INS, 5 % DEL)."""
    stride = 4096
    tiles = contig_len // stride
    pairs = int(round(30.0 * stride / (2 * 150)))
    base = dict(seed=seed, n_contigs=n_contigs, tile_offset=0, tile_len=stride, tile_stride=stride, tiles_per_contig=tiles,
                pairs_per_tile=pairs, var_lo=0, var_hi=stride, snv_period=800, indel_period=7200,
                frag_mean=350.0, frag_sd=60.0)
    base.update(kw)
    return SynthConfig(**base)


def multisample(sample: int, n_tiles=40_000, seed=BASE_SEED + 4, **kw) -> SynthConfig:
    """C4: 60x samples over a shared target set (same sites for every sample: site hashes do not
    depend on the read seed), 2 % multiallelic SNV sites."""
    return wes(n_tiles=n_tiles, seed=seed, pairs_per_tile=72, multiallelic_pct=2, indel_period=5000,
               read_salt=0x5A17 * (sample + 1), **kw)


def amplicon(n_tiles=5_000, depth_pairs=900, seed=BASE_SEED + 5, **kw) -> SynthConfig:
    """C5: ~1000x amplicon panel: 250-bp amplicons, ~1 variant per 25 bp."""
    base = dict(seed=seed, n_contigs=1, tiles_per_contig=n_tiles, tile_len=120, tile_stride=2048, pairs_per_tile=depth_pairs,
                var_lo=60, var_hi=310, snv_period=28, indel_period=250, frag_mean=230.0, frag_sd=25.0, tile_offset=5000)
    base.update(kw)
    return SynthConfig(**base)


def _lib():
    lib = _dev.libsynth()
    lib.vfkio_synth_sites.restype = C.c_int64
    return lib


def _check(rc: int):
    if rc != 0:
        raise ValueError("vfkio: %s" % _lib().vfkio_last_error().decode())


def reads(cfg: SynthConfig, pinned: bool = False) -> ReadBatch:
    """Generate the canonical columnar batch (what a BAM decoder would hand over: nothing is derived from the reads).
    `pinned`: the arrays are allocated page-locked, so the host entry point can read them in place."""
    lib = _lib()
    c = cfg.c()
    plan = _PlanC()
    tiles = cfg.n_tiles_generated
    tile_sizes = np.zeros(4 * tiles, np.int64)
    _check(lib.vfkio_synth_plan_reads(C.byref(c), C.byref(plan), C.c_void_p(tile_sizes.ctypes.data)))
    n = plan.n_reads
    E = lambda k, dt: _dev._empty(k, dt, pinned)
    arr = dict(tid=np.empty(n, np.int32), pos=E(n, np.int32), flag=E(n, np.uint16),
               mapq=E(n, np.uint8), l_qseq=E(n, np.int32), n_cigar=E(n, np.int32),
               cigar_off=E(n, np.int64), seq_off=E(n, np.int64), qual_off=E(n, np.int64),
               cigar=E(max(plan.n_cigar, 1), np.uint32), seq=E(max(plan.n_seq_bytes, 1), np.uint8),
               qual=E(max(plan.n_qual_bytes, 1), np.uint8), mtid=E(n, np.int32), mpos=E(n, np.int32),
               isize=E(n, np.int32), qname_id=E(n, np.uint64), qname_hash=E(n, np.uint32))
    s = _dev._IoReadsC()
    s.n, s.n_contigs = n, cfg.n_contigs
    for k, a in arr.items():
        setattr(s, k, a.ctypes.data)
    _check(lib.vfkio_synth_fill_reads(C.byref(c), C.byref(plan), C.c_void_p(tile_sizes.ctypes.data), C.byref(s)))
    arr["cigar"] = arr["cigar"][:plan.n_cigar]
    arr["seq"] = arr["seq"][:plan.n_seq_bytes]
    arr["qual"] = arr["qual"][:plan.n_qual_bytes]
    return ReadBatch(contigs=cfg.contigs, pinned=pinned, contiguous=True, **arr)  # vfkio_synth_fill_reads: prefix-sum offsets


def tile_range(cfg: SynthConfig, tid_a: int, pos_a: int, tid_b: int, pos_b: int, margin: int = 2):
    """(g_first, g_count): the tiles whose reads can reach the columns from (tid_a, pos_a) to (tid_b, pos_b), `margin`
    tiles added on both sides (a read starts at most one stride + its reference span before a column it covers)."""
    def g(tid, pos, d):
        k = (pos - cfg.tile_offset) // cfg.tile_stride + d
        return tid * cfg.tiles_per_contig + min(max(k, 0), cfg.tiles_per_contig - 1)
    a, b = g(tid_a, pos_a, -margin), g(tid_b, pos_b, margin)
    return a, b - a + 1


def sites(cfg: SynthConfig):
    """Planted sites as structured arrays (tid, pos1, kind, length, bases[24]) -- of the whole genome, whatever tile range
    the config generates reads for."""
    lib = _lib()
    c = replace(cfg, g_first=0, g_count=0).c()
    cap = 1 << 16
    while True:
        tid = np.empty(cap, np.int32)
        pos1 = np.empty(cap, np.int32)
        kind = np.empty(cap, np.int32)
        length = np.empty(cap, np.int32)
        bases = np.zeros(cap * 24, np.uint8)
        n = lib.vfkio_synth_sites(C.byref(c), C.c_int64(cap), C.c_void_p(tid.ctypes.data), C.c_void_p(pos1.ctypes.data),
                                  C.c_void_p(kind.ctypes.data), C.c_void_p(length.ctypes.data), C.c_void_p(bases.ctypes.data))
        if n < 0:
            _check(int(n))
        if n <= cap:
            return tid[:n], pos1[:n], kind[:n], length[:n], bases[:n * 24].reshape(n, 24)
        cap = int(n)


def ref_bases(cfg: SynthConfig, tid: int, start: int, length: int) -> str:
    buf = C.create_string_buffer(length + 1)
    _check(_lib().vfkio_synth_ref(C.byref(cfg.c()), tid, start, length, buf))
    return buf.raw[:length].decode()


def variant_records(cfg: SynthConfig, limit: int = None) -> List[Tuple[str, int, str, List[str]]]:
    """(CHROM, POS, REF, [ALT...]) records of the planted sites, in genome order."""
    tid, pos1, kind, length, bases = sites(cfg)
    names = cfg.contigs
    n = len(tid) if limit is None else min(limit, len(tid))
    out = []
    for i in range(n):
        b = bases[i]
        ref = chr(b[0])
        if kind[i] == 0:
            alts = [chr(b[1])] + ([chr(b[2])] if b[2] else [])
            out.append((names[tid[i]], int(pos1[i]), ref, alts))
        elif kind[i] == 1:
            out.append((names[tid[i]], int(pos1[i]), ref, [ref + bytes(b[3:3 + length[i]]).decode()]))
        else:
            out.append((names[tid[i]], int(pos1[i]), ref + ref_bases(cfg, int(tid[i]), int(pos1[i]), int(length[i])), [ref]))
    return out
