"""ORACLE (test infrastructure, NOT product code) -- ctypes loader for oracle/pileup_oracle.c.

Builds the plain-C restatement with gcc on first use (into oracle/_build/, git-ignored) and
exposes it on numpy arrays.  Only tests/, __graft_entry__ and bench.py's CPU-baseline legs may
import this module; the product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Sequence, Tuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "pileup_oracle.c")
LIB = os.path.join(HERE, "_build", "liboracle.so")


def build(force: bool = False) -> str:
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"])
    return LIB


class _Reads(C.Structure):
    _fields_ = [("n", C.c_int64)] + [(k, C.c_void_p) for k in (
        "tid", "pos", "flag", "mapq", "l_qseq", "n_cigar", "cigar_off", "seq_off", "qual_off", "cigar", "seq",
        "qual", "mtid", "mpos", "isize", "qname_id", "qname_hash")]


class _Params(C.Structure):
    _fields_ = [("min_mapq", C.c_int32), ("min_baseq", C.c_int32), ("flag_filter", C.c_uint32),
                ("ignore_orphans", C.c_int32), ("ignore_overlaps", C.c_int32)]


class _Variants(C.Structure):
    _fields_ = [("n", C.c_int64)] + [(k, C.c_void_p) for k in (
        "tid", "pos1", "stop", "ref_off", "alt_off", "alt0_off", "ref_txt", "alt_txt", "alt0_txt")]


COUNTS_DTYPE = np.dtype([("supported", "<i4"), ("dp", "<i4"), ("n_amb", "<i4"), ("ac_ref", "<i4"), ("ac_alt", "<i4"),
                         ("n_val", "<i4", (3, 2)), ("med2", "<i4", (3, 2)), ("_pad", "<i4"), ("s2", "<u8", (3,)), ("n_col", "<i4"), ("_pad2", "<i4")])

_DT = dict(tid=np.int32, pos=np.int32, flag=np.uint16, mapq=np.uint8, l_qseq=np.int32, n_cigar=np.int32,
           cigar_off=np.int64, seq_off=np.int64, qual_off=np.int64, cigar=np.uint32, seq=np.uint8, qual=np.uint8,
           mtid=np.int32, mpos=np.int32, isize=np.int32, qname_id=np.uint64, qname_hash=np.uint32)

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        assert _lib.orc_sizeof_counts() == COUNTS_DTYPE.itemsize, (_lib.orc_sizeof_counts(), COUNTS_DTYPE.itemsize)
    return _lib


def _reads_struct(b: Dict[str, np.ndarray]):
    keep = {k: np.ascontiguousarray(b[k], dtype=_DT[k]) for k in _DT}
    s = _Reads()
    s.n = len(keep["pos"])
    for k, a in keep.items():
        setattr(s, k, a.ctypes.data)
    return s, keep


def params(min_mapq=0, min_baseq=0, flag_filter=0x704, ignore_orphans=True, ignore_overlaps=True):
    return _Params(min_mapq, min_baseq, flag_filter, int(ignore_orphans), int(ignore_overlaps))


def link_mates(reads: Dict[str, np.ndarray], prm: _Params) -> np.ndarray:
    s, keep = _reads_struct(reads)
    mate = np.empty(s.n, np.int64)
    rc = lib().orc_link_mates(C.byref(s), C.byref(prm), C.c_void_p(mate.ctypes.data))
    assert rc == 0
    return mate


def _pack_text(items: Sequence[str]) -> Tuple[np.ndarray, bytes]:
    off = np.zeros(len(items) + 1, np.int64)
    for i, t in enumerate(items):
        off[i + 1] = off[i] + len(t)
    return off, "".join(items).encode()


def pileup(reads: Dict[str, np.ndarray], units: Sequence[Tuple[int, int, str, str, str]], prm: _Params,
           mate: np.ndarray = None) -> np.ndarray:
    """units: (tid, POS 1-based, REF, ALT_j, ALT_0[, stop]).  `stop` is the last POS of the unit's
    chromosome group (the pysam fetch bound); default: the largest POS given for that tid.
    Returns COUNTS_DTYPE records."""
    s, keep = _reads_struct(reads)
    if mate is None:
        mate = link_mates(reads, prm)
    mate = np.ascontiguousarray(mate, np.int64)
    v = _Variants()
    v.n = len(units)
    tid = np.asarray([u[0] for u in units], np.int32)
    pos1 = np.asarray([u[1] for u in units], np.int32)
    last = {}
    for u in units:
        last[u[0]] = max(last.get(u[0], 0), u[1])
    stop = np.asarray([u[5] if len(u) > 5 else last[u[0]] for u in units], np.int32)
    ro, rt = _pack_text([u[2] for u in units])
    ao, at = _pack_text([u[3] for u in units])
    zo, zt = _pack_text([u[4] for u in units])
    bufs = [C.create_string_buffer(t + b"\0") for t in (rt, at, zt)]
    v.tid, v.pos1, v.stop = tid.ctypes.data, pos1.ctypes.data, stop.ctypes.data
    v.ref_off, v.alt_off, v.alt0_off = ro.ctypes.data, ao.ctypes.data, zo.ctypes.data
    v.ref_txt, v.alt_txt, v.alt0_txt = (C.cast(b, C.c_void_p).value for b in bufs)
    out = np.zeros(len(units), COUNTS_DTYPE)
    rc = lib().orc_pileup(C.byref(s), C.byref(v), C.byref(prm), C.c_void_p(mate.ctypes.data),
                          C.c_void_p(out.ctypes.data))
    assert rc == 0
    return out
