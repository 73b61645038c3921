"""ctypes binding of the C ABI (include/vfk.h, include/vfk_io.h).

torch is used for what it is good at here -- device memory, pinned host memory and streams --
and nothing else: every computation happens inside libvfk.so.  There is NO CPU fallback: if the
CUDA library is missing or no device is present the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np

from . import build as _build
from .batch import ReadBatch, VariantUnits

NO_MATE = 0xFFFFFFFF

COUNTS_DTYPE = np.dtype([("dp", "<i4"), ("n_amb", "<i4"), ("ac_ref", "<i4"), ("ac_alt", "<i4"),
                         ("med2", "<i4", (3, 2)), ("status", "<u4"), ("n_cand", "<u4"), ("s2", "<u8", (3,)),
                         ("n_cover", "<u4"), ("n_col", "<u4")])
STATS_DTYPE = np.dtype([("z", "<f8", (3,)), ("p", "<f8", (3,)), ("pu", "<f8"), ("pw", "<f8"), ("k", "<i4"),
                        ("pad", "<i4")])
assert COUNTS_DTYPE.itemsize == 80 and STATS_DTYPE.itemsize == 72

ST_KIND_MASK, ST_DEFERRED, ST_READLEN, ST_BADBATCH = 0x3, 0x10, 0x20, 0x40


class VfkError(RuntimeError):
    pass


_READ_ARRAYS = ("pos", "flag", "mapq", "l_qseq", "n_cigar", "cigar_off", "seq_off", "qual_off", "cigar", "seq", "qual",
                "mtid", "mpos", "isize", "qname_id", "qname_hash")


class _ReadBatchC(C.Structure):
    _fields_ = ([("n_reads", C.c_int64), ("n_contigs", C.c_int32), ("max_l_qseq", C.c_int32), ("contig_off", C.c_void_p)] +
                [(k, C.c_void_p) for k in _READ_ARRAYS] +
                [("n_cigar_words", C.c_int64), ("n_seq_bytes", C.c_int64), ("n_qual_bytes", C.c_int64)])


class _VariantBatchC(C.Structure):
    _fields_ = [("n_units", C.c_int64), ("tid", C.c_void_p), ("pos", C.c_void_p), ("meta", C.c_void_p),
                ("length", C.c_void_p), ("seq_off", C.c_void_p), ("ins_seq", C.c_void_p), ("stop", C.c_void_p),
                ("n_ins_seq", C.c_int64)]


class ParamsC(C.Structure):
    _fields_ = [("min_mapq", C.c_int32), ("min_baseq", C.c_int32), ("flag_filter", C.c_uint32),
                ("ignore_orphans", C.c_int32), ("ignore_overlaps", C.c_int32), ("include_ambiguous", C.c_int32)]


def make_params(min_mapq=0, min_baseq=29, include_ambiguous=True, flag_filter=0x704, ignore_orphans=True,
                ignore_overlaps=True) -> ParamsC:
    """Defaults: Annotator class defaults (vafator/annotator.py:76-77) + pysam 0.21.0 pileup defaults."""
    return ParamsC(int(min_mapq), int(min_baseq), int(flag_filter), int(bool(ignore_orphans)),
                   int(bool(ignore_overlaps)), int(bool(include_ambiguous)))


class _IoReadsC(C.Structure):
    _fields_ = [("n", C.c_int64), ("n_contigs", C.c_int32), ("pad", C.c_int32)] + [(k, C.c_void_p) for k in (
        "tid", "pos", "flag", "mapq", "l_qseq", "n_cigar", "cigar_off", "seq_off", "qual_off", "cigar", "seq", "qual",
        "mtid", "mpos", "isize", "qname_id", "qname_hash")]


_IO_DT = dict(tid=np.int32, pos=np.int32, flag=np.uint16, mapq=np.uint8, l_qseq=np.int32, n_cigar=np.int32,
              cigar_off=np.int64, seq_off=np.int64, qual_off=np.int64, cigar=np.uint32, seq=np.uint8, qual=np.uint8,
              mtid=np.int32, mpos=np.int32, isize=np.int32, qname_id=np.uint64, qname_hash=np.uint32)

_libio = None
_libvfk = None


def libio():
    """libvfkio.so (host companion).  Built in-tree on first use."""
    global _libio
    if _libio is None:
        path = _build.LIBVFKIO if os.path.exists(_build.LIBVFKIO) and not os.environ.get("VFK_REBUILD") else _build.build_host()
        lib = C.CDLL(path)
        lib.vfkio_last_error.restype = C.c_char_p
        _libio = lib
    return _libio


_libsynth = None


def libsynth():
    """libvfksynth.so (synthetic benchmark / test data; not product code).  Built in-tree on first use."""
    global _libsynth
    if _libsynth is None:
        path = _build.LIBVFKSYNTH if os.path.exists(_build.LIBVFKSYNTH) and not os.environ.get("VFK_REBUILD") else _build.build_synth()
        lib = C.CDLL(path)
        lib.vfkio_last_error.restype = C.c_char_p
        _libsynth = lib
    return _libsynth


def libvfk():
    """libvfk.so (CUDA).  Raises if it cannot be loaded -- there is no fallback path."""
    global _libvfk
    if _libvfk is None:
        path = os.environ.get("VFK_LIBVFK") or _build.LIBVFK  # VFK_LIBVFK: an alternative build of the same library (tools/build_variants.py, A/B runs)
        if not os.path.exists(path):
            path = _build.build_cuda()
        try:
            lib = C.CDLL(path)
        except OSError as e:  # pragma: no cover
            raise VfkError("cannot load the CUDA library %s: %s" % (path, e))
        lib.vfk_last_error.restype = C.c_char_p
        lib.vfk_launch_count.restype = C.c_int64
        lib.vfk_launch_count.argtypes = [C.c_void_p]
        lib.vfk_epilogue.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_double, C.c_double,
                                     C.c_void_p, C.c_void_p]
        lib.vfk_pileup.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.vfk_pileup_values.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                          C.c_void_p]
        lib.vfk_ranksum_values.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p]
        lib.vfk_annotate_host_async.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                                                C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.vfk_wait.argtypes = [C.c_void_p, C.c_int]
        lib.vfk_prep_words.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.vfk_set_timing.argtypes = [C.c_void_p, C.c_int]
        lib.vfk_last_timing.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        lib.vfk_transfer_stats.restype = C.c_int
        lib.vfk_transfer_stats.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        lib.vfk_fetch_stats.restype = C.c_int
        lib.vfk_fetch_stats.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        lib.vfk_annotate_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                                          C.c_double, C.c_void_p, C.c_void_p]
        _libvfk = lib
    return _libvfk


def _io_check(rc: int):
    if rc != 0:
        raise ValueError("vfkio: %s" % libio().vfkio_last_error().decode())


def _check(rc: int):
    if rc != 0:
        msg = libvfk().vfk_last_error().decode()
        if rc == -1:
            raise ValueError("vfk: %s" % msg)
        raise VfkError("vfk error %d: %s" % (rc, msg))


def _io_reads(batch: ReadBatch):
    keep = {k: np.ascontiguousarray(getattr(batch, k), dtype=dt) for k, dt in _IO_DT.items()}
    s = _IoReadsC()
    s.n = batch.n_reads
    s.n_contigs = len(batch.contigs)
    for k, a in keep.items():
        setattr(s, k, a.ctypes.data)
    return s, keep


def _empty(n, dtype, pinned: bool):
    """Host array; pinned (page-locked) through torch when asked and a device is present."""
    if pinned:
        import torch
        if torch.cuda.is_available():
            t = torch.empty(max(int(n), 1) * np.dtype(dtype).itemsize, dtype=torch.uint8, pin_memory=True)
            a = t.numpy().view(dtype)[:n]
            return a
    return np.empty(n, dtype)


def pinned_empty(n, dtype):
    """Page-locked host array (inputs read in place / results written directly by the host entry point; pageable arrays work
    too and go through the library's bounce buffers)."""
    return _empty(n, dtype, True)


class HostReads:
    """The canonical columnar batch (include/vfk.h vfk_read_batch) in host memory: the placed reads of a ReadBatch, array
    by array as the decoder wrote them, plus the batch metadata (contig table, longest read).  Nothing is derived from the
    reads here -- spans, position index, CIGAR shapes, mate pairing are the kernels' work.  `pinned`: the arrays are
    copied once into page-locked memory (the decoder / generator can also write there directly: bamio, synth)."""

    def __init__(self, batch: ReadBatch, pinned: bool = False, implicit_offsets: Optional[bool] = None):
        n = batch.n_placed
        self.batch = batch
        # pools contiguous in read order: the offset arrays are handed over as NULL and derived on the device (include/vfk.h)
        self.implicit_offsets = bool(batch.contiguous) if implicit_offsets is None else bool(implicit_offsets)
        self.contigs = list(batch.contigs)
        self.n_reads = n
        self.contig_off = batch.contig_off
        self.max_l_qseq = batch.max_l_qseq
        self.arrays = {}
        for k in _READ_ARRAYS:
            a = getattr(batch, k)
            if k not in ("cigar", "seq", "qual"):
                a = a[:n]
            a = np.ascontiguousarray(a, dtype=_IO_DT[k])
            if pinned and not getattr(batch, "pinned", False):
                p = _empty(a.shape[0], a.dtype, True)
                p[:] = a
                a = p
            self.arrays[k] = a

    @property
    def starts(self):
        return self.arrays["pos"]

    def array_names(self):
        return ("contig_off",) + _READ_ARRAYS

    def nbytes(self) -> int:
        return int(self.contig_off.nbytes + sum(a.nbytes for a in self.arrays.values()))

    def c_struct(self, ptrs: Optional[Dict[str, int]] = None) -> _ReadBatchC:
        ptrs_given = ptrs  # device pointers of a resident copy (DeviceReads): those keep their explicit offsets
        s = _ReadBatchC()
        s.n_reads, s.n_contigs, s.max_l_qseq = self.n_reads, len(self.contigs), self.max_l_qseq
        s.n_cigar_words = int(self.arrays["cigar"].shape[0])
        s.n_seq_bytes = int(self.arrays["seq"].shape[0])
        s.n_qual_bytes = int(self.arrays["qual"].shape[0])
        if ptrs is None:
            ptrs = {k: a.ctypes.data for k, a in self.arrays.items()}
            ptrs["contig_off"] = self.contig_off.ctypes.data
        for k, v in ptrs.items():
            setattr(s, k, v)
        if self.implicit_offsets and ptrs_given is None:  # host entry: derived on the device, nothing of them crosses the bus
            s.cigar_off = s.seq_off = s.qual_off = None
        return s


def host_reads(batch: ReadBatch, pinned: bool = False, implicit_offsets: Optional[bool] = None) -> HostReads:
    return HostReads(batch, pinned=pinned, implicit_offsets=implicit_offsets)


def read_passes(batch: ReadBatch, params: ParamsC) -> np.ndarray:
    """Which reads the reference would push into its pileup under `params` (flag filter, MAPQ, orphans)."""
    f = batch.flag.astype(np.uint32)
    ok = (f & (params.flag_filter | 0x4)) == 0
    ok &= batch.mapq.astype(np.int64) >= params.min_mapq
    if params.ignore_orphans:
        ok &= ~(((f & 1) != 0) & ((f & 2) == 0))
    return ok & (batch.tid >= 0)


def units_c_struct(u: VariantUnits, stop: Optional[np.ndarray], ptrs: Optional[Dict[str, int]] = None) -> _VariantBatchC:
    s = _VariantBatchC()
    s.n_units = u.n_units
    s.n_ins_seq = int(u.ins_seq.shape[0])
    if ptrs is None:
        ptrs = dict(tid=u.tid.ctypes.data, pos=u.pos.ctypes.data, meta=u.meta.ctypes.data,
                    length=u.length.ctypes.data, seq_off=u.seq_off.ctypes.data, ins_seq=u.ins_seq.ctypes.data,
                    stop=stop.ctypes.data if stop is not None else None)
    for k, v in ptrs.items():
        setattr(s, k, v)
    return s


class DeviceReads:
    """A canonical read batch resident in HBM (torch tensors own the memory): the host arrays, copied verbatim."""

    def __init__(self, host: HostReads, device, non_blocking: bool = True):
        import torch
        self.host = host
        self.t = {}
        src = dict(host.arrays)
        src["contig_off"] = host.contig_off
        for k, a in src.items():
            h = torch.from_numpy(a.view(np.uint8).reshape(-1)) if a.size else torch.zeros(16, dtype=torch.uint8)
            self.t[k] = h.to(device, non_blocking=non_blocking) if a.size else h.to(device)
        self.c = host.c_struct({k: v.data_ptr() for k, v in self.t.items()})

    def nbytes(self):
        return self.host.nbytes()


class DeviceUnits:
    def __init__(self, units: VariantUnits, stop: Optional[np.ndarray], device):
        import torch
        self.units = units
        self.t = {}
        src = dict(tid=units.tid, pos=units.pos, meta=units.meta, length=units.length, seq_off=units.seq_off,
                   ins_seq=units.ins_seq)
        if stop is not None:
            src["stop"] = np.ascontiguousarray(stop, np.int32)
        for k, a in src.items():
            h = torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)) if a.size else torch.zeros(16, dtype=torch.uint8)
            self.t[k] = h.to(device)
        ptrs = {k: v.data_ptr() for k, v in self.t.items()}
        ptrs.setdefault("stop", None)
        self.c = units_c_struct(units, stop, ptrs)
        self.n_units = units.n_units


class Device:
    """One vfk_handle on one GPU."""

    def __init__(self, index: int = 0):
        import torch
        if not torch.cuda.is_available():
            raise VfkError("no CUDA device: vafator_b200 has no CPU path")
        self.torch = torch
        self.index = index
        self.device = torch.device("cuda", index)
        self.lib = libvfk()
        h = C.c_void_p()
        _check(self.lib.vfk_create(C.c_int(index), C.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            self.lib.vfk_destroy(self.h)
            self.h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def launch_count(self) -> int:
        return int(self.lib.vfk_launch_count(self.h))

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def upload_reads(self, reads) -> DeviceReads:
        """reads: a HostReads, or a ReadBatch (wrapped first)."""
        if isinstance(reads, ReadBatch):
            reads = HostReads(reads)
        return DeviceReads(reads, self.device)

    def upload_units(self, units: VariantUnits, stop: Optional[np.ndarray] = None) -> DeviceUnits:
        return DeviceUnits(units, stop, self.device)

    def alloc_counts(self, n):
        return self.torch.empty(max(n, 1) * COUNTS_DTYPE.itemsize, dtype=self.torch.uint8, device=self.device)

    def alloc_stats(self, n):
        return self.torch.empty(max(n, 1) * STATS_DTYPE.itemsize, dtype=self.torch.uint8, device=self.device)

    def pileup(self, reads: DeviceReads, units: DeviceUnits, params: ParamsC, out=None):
        """Enqueue the pileup kernel on torch's current stream; returns the device counts buffer."""
        if out is None:
            out = self.alloc_counts(units.n_units)
        _check(self.lib.vfk_pileup(self.h, C.byref(reads.c), C.byref(units.c), C.byref(params),
                                   C.c_void_p(out.data_ptr()), self._stream()))
        return out

    def epilogue(self, n_units: int, counts, eaf, fpr: float, error_rate: float, out=None):
        if out is None:
            out = self.alloc_stats(n_units)
        _check(self.lib.vfk_epilogue(self.h, n_units, C.c_void_p(counts.data_ptr()), C.c_void_p(eaf.data_ptr()),
                                     float(fpr), float(error_rate), C.c_void_p(out.data_ptr()), self._stream()))
        return out

    def pileup_values(self, reads: "DeviceReads", units: "DeviceUnits", params: ParamsC, max_per_unit: int = 128):
        """Per unit, the packed per-read values of the REF / ALT lists (vfk_pileup_values).  Returns
        (values[n_units, max], n_values[n_units]) on the host, growing the buffer until everything fits."""
        t = self.torch
        n = units.n_units
        while True:
            vals = t.empty(max(n, 1) * max_per_unit, dtype=t.int32, device=self.device)
            cnt = t.empty(max(n, 1), dtype=t.int32, device=self.device)
            _check(self.lib.vfk_pileup_values(self.h, C.byref(reads.c), C.byref(units.c), C.byref(params),
                                              C.c_int32(max_per_unit), C.c_void_p(vals.data_ptr()),
                                              C.c_void_p(cnt.data_ptr()), self._stream()))
            counts = cnt.cpu().numpy()[:n]
            if n == 0 or counts.max() <= max_per_unit:
                return vals.cpu().numpy().view(np.uint32).reshape(max(n, 1), max_per_unit)[:n], counts
            max_per_unit = int(counts.max())

    def ranksum_values(self, alt_lists, ref_lists) -> np.ndarray:
        """s2[p][mq,bq,pos] for pairs of packed-value lists (vfk_ranksum_values)."""
        t = self.torch
        n = len(alt_lists)
        if n == 0:
            return np.zeros((0, 3), np.uint64)
        a_off = np.zeros(n + 1, np.int64)
        r_off = np.zeros(n + 1, np.int64)
        a_off[1:] = np.cumsum([len(x) for x in alt_lists])
        r_off[1:] = np.cumsum([len(x) for x in ref_lists])
        a = np.concatenate([np.asarray(x, np.uint32) for x in alt_lists] + [np.zeros(1, np.uint32)])
        r = np.concatenate([np.asarray(x, np.uint32) for x in ref_lists] + [np.zeros(1, np.uint32)])
        dev = lambda x: t.from_numpy(x.view(np.uint8)).to(self.device)
        da, dr, dao, dro = dev(a), dev(r), dev(a_off), dev(r_off)
        out = t.empty(n * 3 * 8, dtype=t.uint8, device=self.device)
        _check(self.lib.vfk_ranksum_values(self.h, C.c_int64(n), C.c_void_p(da.data_ptr()), C.c_void_p(dao.data_ptr()),
                                           C.c_void_p(dr.data_ptr()), C.c_void_p(dro.data_ptr()),
                                           C.c_void_p(out.data_ptr()), self._stream()))
        return out.cpu().numpy().view(np.uint64).reshape(n, 3)

    def annotate_host(self, packed: HostReads, units: VariantUnits, stop, params: ParamsC, eaf: np.ndarray,
                      fpr: float, error_rate: float, counts_out=None, stats_out=None):
        """HOST buffers in, HOST results out (copies inside): the end-to-end entry point."""
        n = units.n_units
        counts = counts_out if counts_out is not None else np.empty(n, COUNTS_DTYPE)
        stats = stats_out if stats_out is not None else np.empty(n, STATS_DTYPE)
        eaf = np.ascontiguousarray(eaf, np.float64)
        rc = packed.c_struct()
        stop = None if stop is None else np.ascontiguousarray(stop, np.int32)
        uc = units_c_struct(units, stop)
        _check(self.lib.vfk_annotate_host(self.h, C.byref(rc), C.byref(uc), C.byref(params),
                                          C.c_void_p(eaf.ctypes.data), float(fpr), float(error_rate),
                                          C.c_void_p(counts.ctypes.data), C.c_void_p(stats.ctypes.data)))
        return counts, stats

    def annotate_host_async(self, packed: HostReads, units: VariantUnits, stop, params: ParamsC, eaf: np.ndarray,
                            fpr: float, error_rate: float, counts_out: np.ndarray, stats_out: np.ndarray):
        """Enqueue one batch on one of the staging slots (`slots` of them); returns a ticket for `wait`.  All host
        arrays (inputs and outputs) must stay alive and untouched until `wait(ticket)` returns."""
        eaf = np.ascontiguousarray(eaf, np.float64)
        stop = None if stop is None else np.ascontiguousarray(stop, np.int32)
        rc = packed.c_struct()
        uc = units_c_struct(units, stop)
        ticket = C.c_int(0)
        _check(self.lib.vfk_annotate_host_async(self.h, C.byref(rc), C.byref(uc), C.byref(params),
                                                C.c_void_p(eaf.ctypes.data), float(fpr), float(error_rate),
                                                C.c_void_p(counts_out.ctypes.data), C.c_void_p(stats_out.ctypes.data),
                                                C.byref(ticket)))
        self.__dict__.setdefault("_inflight", {})[ticket.value] = (packed, units, stop, eaf, counts_out, stats_out)
        return ticket.value

    @property
    def slots(self) -> int:
        """Batches the host entry point keeps in flight (vfk_slot_count)."""
        return int(self.lib.vfk_slot_count())

    def wait(self, ticket: int):
        _check(self.lib.vfk_wait(self.h, C.c_int(ticket)))
        self.__dict__.get("_inflight", {}).pop(ticket, None)

    def transfer_stats(self) -> Dict[str, int]:
        """Host entry point accounting since the handle was created (vfk_transfer_stats)."""
        out = (C.c_int64 * 4)()
        _check(self.lib.vfk_transfer_stats(self.h, out))
        f = (C.c_int64 * 4)()
        _check(self.lib.vfk_fetch_stats(self.h, f))
        return dict(staged_bytes=int(out[0]), zero_copy_batches=int(out[1]), staged_batches=int(out[2]), d2h_bytes=int(out[3]),
                    fetch_pairs=int(f[0]), fetch_batches=int(f[1]), fetch_ns=int(f[2]), fetch_threads=int(f[3]))

    def prep_words(self, reads: "DeviceReads"):
        """(end, shape word, status) the read-preparation kernel derives per read (vfk_prep_words) -- for tests."""
        t = self.torch
        n = reads.host.n_reads
        end = t.empty(max(n, 1), dtype=t.int32, device=self.device)
        shape = t.empty(max(n, 1), dtype=t.int32, device=self.device)
        status = t.zeros(1, dtype=t.int32, device=self.device)
        _check(self.lib.vfk_prep_words(self.h, C.byref(reads.c), C.c_void_p(end.data_ptr()), C.c_void_p(shape.data_ptr()),
                                       C.c_void_p(status.data_ptr()), self._stream()))
        return end.cpu().numpy()[:n], shape.cpu().numpy().view(np.uint32)[:n], int(status.cpu().numpy().view(np.uint32)[0])

    def set_timing(self, on: bool):
        _check(self.lib.vfk_set_timing(self.h, C.c_int(1 if on else 0)))

    def last_timing(self) -> Dict[str, float]:
        """Device milliseconds per phase of the last call made while timing was on (vfk_last_timing)."""
        out = (C.c_double * 5)()
        _check(self.lib.vfk_last_timing(self.h, out))
        return dict(prep=out[0], locate=out[1], pileup=out[2], lists=out[3], epilogue=out[4])

    def to_host(self, buf, dtype, n):
        return buf.cpu().numpy().view(dtype)[:n].copy()

    def sync(self):
        _check(self.lib.vfk_sync(self.h))
