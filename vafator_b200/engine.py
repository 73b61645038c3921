"""Host-side composition of the annotation around the two device stages.

    reads (ReadBatch: the canonical columnar batch, as decoded) + units + expected VAF
        --vfk_annotate_host_async--> integer counts and FP64 z / p / pu / pw / k per unit
    (one call per BAM: read preparation, locate, pileup and epilogue kernels, one copy in, one copy out;
     two batches in flight)

What stays on the host is what the reference also does in Python after its arithmetic
(vafator/annotator.py:260-432): merging replicate BAMs, choosing which values a Counter / dict
lookup would have returned, and turning numbers into INFO strings with the same rounding
functions.  No statistic is computed here.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import device as _dev
from .batch import KIND_SNV, ReadBatch, VariantUnits, build_units, variant_kind
from .constants import AMBIGUOUS_BASES

Record = Tuple[str, int, str, Sequence[str]]  # CHROM, POS, REF, [ALT...]
METRICS = ("mq", "bq", "pos")  # order of the device record


def chrom_groups(records: Sequence[Record]):
    """Consecutive records of one chromosome, as the reference streams them
    (vafator/pileups.py:83-103).  Returns (in_range mask, stop per record): the reference piles up
    [first.POS-1, last.POS) of each group (vafator/pileups.py:137-138), so records outside that span
    get no metrics, and `stop` bounds the reads it ever fetches."""
    n = len(records)
    in_range = np.ones(n, bool)
    stop = np.zeros(n, np.int32)
    i = 0
    while i < n:
        j = i
        while j + 1 < n and records[j + 1][0] == records[i][0]:
            j += 1
        first, last = records[i][1], records[j][1]
        for k in range(i, j + 1):
            in_range[k] = first <= records[k][1] <= last
            stop[k] = last
        i = j + 1
    return in_range, stop


def span_range(records: Sequence[Record], span: Tuple[int, int]):
    """The same for records that are PART of a chromosome group whose first / last record lie at span = (first.POS, last.POS):
    what `chrom_groups` would say about them inside the whole group (a group too large for one call is annotated in chunks)."""
    first, last = int(span[0]), int(span[1])
    pos = np.asarray([r[1] for r in records], np.int64)
    return (pos >= first) & (pos <= last), np.full(len(records), last, np.int32)


class RowTable:
    """One row per (record, ALT allele) -- the granularity of the A-numbered INFO fields."""

    def __init__(self, records: Sequence[Record]):
        self.records = records
        n_alt = np.asarray([len(r[3]) for r in records], np.int64)
        self.first = np.concatenate([[0], np.cumsum(n_alt)]).astype(np.int64)
        self.n_rows = int(self.first[-1])
        self.rec = np.repeat(np.arange(len(records), dtype=np.int64), n_alt)
        self.alt = (np.arange(self.n_rows, dtype=np.int64) - self.first[self.rec]) if self.n_rows else np.zeros(0, np.int64)
        self.kind = np.asarray([(-1 if not r[3] or variant_kind(r[2], r[3][0]) is None else variant_kind(r[2], r[3][0]))
                                for r in records], np.int64)


class BamResult:
    """Per-BAM values on the row table, straight from the device counts."""

    def __init__(self, rows: RowTable, units: VariantUnits, counts: np.ndarray, in_range: np.ndarray):
        records = rows.records
        nrec, nrow = len(records), rows.n_rows
        self.row_unit = np.full(nrow, -1, np.int64)
        for (i, j), u in units.unit_of.items():
            self.row_unit[rows.first[i] + j] = u
        u0 = self.row_unit[np.minimum(rows.first[:-1], max(nrow - 1, 0))] if nrec and nrow else np.full(nrec, -1, np.int64)  # unit of ALT[0]
        u0 = np.where(rows.first[:-1] < rows.first[1:], u0, -1) if nrec else u0
        has_u0 = u0 >= 0
        safe0 = np.where(has_u0, u0, 0)
        z = counts[safe0] if len(counts) else np.zeros(nrec, _dev.COUNTS_DTYPE)
        # the reference has metrics for the record iff a pileup column exists at POS (>= 1 read after
        # the read-level filters) and the record lies inside its chromosome group's span
        self.has = has_u0 & in_range & (z["n_col"] > 0) if nrec else np.zeros(0, bool)
        self.dp = np.where(self.has, z["dp"], 0).astype(np.int64)
        self.n_amb = np.where(self.has, z["n_amb"], 0).astype(np.int64)
        self.ac_ref = np.where(self.has, z["ac_ref"], 0).astype(np.int64)
        self.med2_ref = np.where(self.has[:, None], z["med2"][:, :, 0], -1).astype(np.int64) if nrec else np.zeros((0, 3), np.int64)
        ru = self.row_unit
        has_row = (ru >= 0) & self.has[rows.rec] if nrow else np.zeros(0, bool)
        zr = counts[np.where(ru >= 0, ru, 0)] if len(counts) else np.zeros(nrow, _dev.COUNTS_DTYPE)
        self.ac = np.where(has_row, zr["ac_alt"], 0).astype(np.int64)
        self.med2_alt = np.where(has_row[:, None], zr["med2"][:, :, 1], -1).astype(np.int64) if nrow else np.zeros((0, 3), np.int64)
        self.row_counts = zr.copy()
        self.row_counts["dp"] = self.dp[rows.rec] if nrow else 0
        self.row_counts["ac_alt"] = self.ac
        for f in ("ac_ref", "n_amb"):
            self.row_counts[f] = np.where(has_row, zr[f], 0)
        self.row_counts["s2"] = np.where(has_row[:, None], zr["s2"], 0)
        # the `n` field: ambiguous-base count (SNV); for indels the reference sums ac over the IUPAC
        # letters too, which only bites when an ALT allele is itself one of those letters
        self.n = self.n_amb.copy()
        for i, r in enumerate(records):
            if rows.kind[i] > 0:
                f = rows.first[i]
                if self.has[i] and r[3][0].upper() in AMBIGUOUS_BASES:
                    self.n[i] = int(self.ac[f])
                # indel counts are keyed ALT[0].upper() but looked up with the literal ALT text
                # (vafator/pileups.py:253-261 vs annotator.py:382-383): a lower-case ALT reads 0
                if r[3][0] != r[3][0].upper():
                    self.ac[f] = 0
                    self.row_counts["ac_alt"][f] = 0


def _fmt_float(x: float) -> str:
    return str(float(x))


def _np_round_str(x: float, nd: int) -> str:
    """str(round(np.float64(x), nd)) -- numpy's rounding, as the reference applies to z, p, pu, pw."""
    return str(np.round(np.float64(x), nd))


class RoundedStats:
    """Device statistics of one context, rounded once for the whole array with numpy's rounding (the
    reference rounds np.float64 scalars: same multiply / rint / divide) and held as Python floats, whose
    str() is the same shortest representation numpy prints."""

    def __init__(self, st: np.ndarray):
        self.pu_arr, self.pw_arr, self.k_arr = np.round(st["pu"], 5), np.round(st["pw"], 5), np.asarray(st["k"])
        self.z_arr, self.p_arr = np.round(st["z"], 3).reshape(-1, 3), np.round(st["p"], 5).reshape(-1, 3)
        self.pu = self.pu_arr.tolist()
        self.pw = self.pw_arr.tolist()
        self.k = self.k_arr.tolist()
        self.z = self.z_arr.tolist()
        self.p = self.p_arr.tolist()


class SampleFormatter:
    """Builds the ordered INFO fields of one record (vafator/annotator.py:304-420)."""

    def __init__(self, rows: RowTable):
        self.rows = rows

    def key_state(self, i: int, res: BamResult):
        """Per allele of record i in one BAM: (has_key, n_values, med2 per metric).  `has_key` says
        whether the reference's dictionaries would contain the allele at all (a Counter miss prints
        0, an empty list prints nan)."""
        rows = self.rows
        chrom, pos, ref, alts = rows.records[i]
        kind = rows.kind[i]
        out = []
        if not res.has[i]:
            return [(False, 0, (-1, -1, -1))] * (1 + len(alts))
        if kind == KIND_SNV:
            n = int(res.ac_ref[i])
            out.append((n > 0, n, tuple(int(x) for x in res.med2_ref[i])))
            for j in range(len(alts)):
                r = rows.first[i] + j
                n = int(res.ac[r])
                out.append((n > 0, n, tuple(int(x) for x in res.med2_alt[r])))
        else:
            n = int(res.ac_ref[i])
            out.append((True, n, tuple(int(x) for x in res.med2_ref[i])))  # REF key always present
            for j, a in enumerate(alts):
                r = rows.first[i] + j
                present = a == a.upper()  # lists are keyed ALT.upper(), looked up with the literal
                n = int(res.ac[r]) if (present and j == 0) else 0
                out.append((present, n, tuple(int(x) for x in res.med2_alt[r]) if (present and j == 0) else (-1, -1, -1)))
        return out


def format_fields(prefix: str, suffix: str, with_eaf: bool, alts, kind, ac, dp, n, eaf, pu, pw, k, med, rs) -> "OrderedDict":
    """ac[j], pu[j]: per ALT; med[metric] = list of strings (REF, ALT...); rs[metric] = ([z...], [p...])."""
    o = OrderedDict()
    key = lambda name: "{}_{}{}".format(prefix, name, suffix)
    o[key("af")] = ",".join(str(round(float(a) / dp, 5) if dp > 0 else 0.0) for a in ac)
    o[key("ac")] = ",".join(str(int(a)) for a in ac)
    o[key("n")] = str(int(n))
    o[key("dp")] = int(dp)
    if with_eaf:
        o[key("eaf")] = str(float(eaf))
    o[key("pu")] = ",".join(str(x) for x in pu)  # already rounded (RoundedStats)
    o[key("pw")] = str(pw)
    o[key("k")] = str(int(k))
    for name in ("bq", "mq", "pos"):
        o[key(name)] = ",".join(med[name])
    for name, tag in (("mq", "rsmq"), ("bq", "rsbq"), ("pos", "rspos")):
        zs, ps = rs[name]
        if zs:
            o[key(tag)] = ",".join(zs)
            o["{}_{}_pv{}".format(prefix, tag, suffix)] = ",".join(ps)
    return o


_SINGLE_KEYS = ("af", "ac", "n", "dp", "eaf", "pu", "pw", "k", "bq", "mq", "pos")
_RS_TAGS = (("mq", "rsmq"), ("bq", "rsbq"), ("pos", "rspos"))  # emission order (vafator/annotator.py:350-376)


def format_single_bam(first, kind_rec, alt_idx, present_row, has_rec, dp, n_amb, ac, ac_ref, med2_ref, med2_alt,
                      pu, pw, k, z, p, eaf, templates=None):
    """INFO values of one single-BAM sample for a run of records, from plain arrays (so that a worker process can
    do a chunk: nothing but numpy arrays crosses the process boundary).  Per record: `first[i] .. first[i+1]` are its
    rows (one per ALT).  pu / pw / z / p are already rounded.  Returns (values, masks): values[i] = the strings of
    _SINGLE_KEYS (dp stays an int) followed by the rank-sum pairs present, masks[i] = which of _RS_TAGS those are.
    With `templates` (one "key=%s;key=%s..." string per mask, text_templates) values[i] is the finished INFO text."""
    first = np.asarray(first, np.int64)
    n_rec = len(first) - 1
    n_row = int(first[-1]) if n_rec else 0
    rec_of = np.repeat(np.arange(n_rec, dtype=np.int64), np.diff(first)) if n_rec else np.zeros(0, np.int64)
    first = first.tolist()
    snv_row = kind_rec[rec_of] == KIND_SNV
    has_row = has_rec[rec_of]
    dp_rec = dp.tolist()
    ac_row = ac.tolist()
    dp_row = dp[rec_of].tolist()
    af_row = [str(round(float(a) / d, 5)) if d > 0 else "0.0" for a, d in zip(ac_row, dp_row)]
    acs_row = list(map(str, ac_row))
    pu_row = list(map(str, pu.tolist()))
    pw_row = list(map(str, pw.tolist()))
    k_row = list(map(str, k.tolist()))
    # ---- medians: "0" = the reference's Counter misses the allele, "nan" = key present with an empty list
    ref_ok = has_rec & (ac_ref > 0)
    alt_ok = (ac > 0) & has_row
    indel_rec = has_rec & (kind_rec > KIND_SNV)
    indel_row = indel_rec[rec_of]
    med_ref, med_alt = {}, {}
    for mi, name in enumerate(METRICS):
        vals = list(map(str, (med2_ref[:, mi] / 2.0).tolist())) if n_rec else []
        code = np.where(ref_ok, 2, 0)                                  # SNV: value or "0"
        code = np.where(indel_rec, np.where(ac_ref > 0, 2, 1), code)     # indel: the REF key always exists
        if name == "bq":
            code = np.where(kind_rec != KIND_SNV, 0, code)
        med_ref[name] = [v if c == 2 else ("nan" if c == 1 else "0") for v, c in zip(vals, code.tolist())]
        vals = list(map(str, (med2_alt[:, mi] / 2.0).tolist())) if n_row else []
        code = np.where(alt_ok, 2, 0)
        # an indel ALT list is keyed ALT.upper() and looked up with the literal text (SampleFormatter.key_state)
        code = np.where(indel_row, np.where(present_row, np.where(alt_ok & (alt_idx == 0), 2, 1), 0), code)
        if name == "bq":
            code = np.where(~snv_row, 0, code)
        med_alt[name] = [v if c == 2 else ("nan" if c == 1 else "0") for v, c in zip(vals, code.tolist())]
    # ---- rank sums: reported when both lists hold values and the statistic is a number
    z = np.asarray(z, np.float64).reshape(n_row, 3)
    p = np.asarray(p, np.float64).reshape(n_row, 3)
    rs_ok, rs_z, rs_p = {}, {}, {}
    for mi, name in enumerate(METRICS):
        ok = alt_ok & ref_ok[rec_of] & ~np.isnan(z[:, mi]) & ~np.isnan(p[:, mi])
        if name == "bq":
            ok = ok & snv_row
        rs_ok[name] = ok.tolist()
        rs_z[name] = list(map(str, z[:, mi].tolist()))
        rs_p[name] = list(map(str, p[:, mi].tolist()))
    n_str = list(map(str, n_amb.tolist()))
    eaf_str = list(map(str, np.asarray(eaf, np.float64).tolist()))
    values = []
    masks = []
    for i in range(n_rec):
        r0, r1 = first[i], first[i + 1]
        if r1 - r0 == 1:
            fixed = (af_row[r0], acs_row[r0], n_str[i], dp_rec[i], eaf_str[i], pu_row[r0], pw_row[r0], k_row[r0],
                     med_ref["bq"][i] + "," + med_alt["bq"][r0], med_ref["mq"][i] + "," + med_alt["mq"][r0],
                     med_ref["pos"][i] + "," + med_alt["pos"][r0])
            mask = 0
            extra = ()
            for m, (name, _tag) in enumerate(_RS_TAGS):
                if rs_ok[name][r0]:
                    mask |= 1 << m
                    extra += (rs_z[name][r0], rs_p[name][r0])
        else:
            rr = range(r0, r1)
            fixed = (",".join(af_row[r0:r1]), ",".join(acs_row[r0:r1]), n_str[i], dp_rec[i], eaf_str[i],
                     ",".join(pu_row[r0:r1]), pw_row[r0] if r1 > r0 else "0.0", k_row[r0] if r1 > r0 else "1",
                     ",".join([med_ref["bq"][i]] + med_alt["bq"][r0:r1]), ",".join([med_ref["mq"][i]] + med_alt["mq"][r0:r1]),
                     ",".join([med_ref["pos"][i]] + med_alt["pos"][r0:r1]))
            mask = 0
            extra = ()
            for m, (name, _tag) in enumerate(_RS_TAGS):
                zs = [rs_z[name][r] for r in rr if rs_ok[name][r]]
                if zs:
                    mask |= 1 << m
                    extra += (",".join(zs), ",".join(rs_p[name][r] for r in rr if rs_ok[name][r]))
        values.append(fixed + extra if templates is None else templates[mask] % (fixed + extra))
        masks.append(mask)
    return values, masks


import ctypes as _C


class _FormatIn(_C.Structure):  # include/vfk_io.h vfkio_format_in
    _fields_ = [("sample", _C.c_char_p), ("n_rec", _C.c_int64)] + [(_k, _C.c_void_p) for _k in (
        "first", "kind_rec", "has_rec", "dp", "n", "ac_ref", "med2_ref", "eaf", "alt_idx", "present_row", "ac", "med2_alt", "pu", "pw", "k", "z", "p")]


def format_single_bam_native(sample, first, kind_rec, alt_idx, present_row, has_rec, dp, n_amb, ac, ac_ref, med2_ref, med2_alt,
                             pu, pw, k, z, p, eaf) -> List[str]:
    """The finished INFO text of one single-BAM sample per record -- format_single_bam(..., templates=text_templates(sample))[0],
    printed by libvfkio (vfkio_format_info: all cores, Python's own float formatting restated in C++; tests compare the two
    string for string)."""
    import ctypes as C
    lib = _dev.libio()
    n_rec = len(first) - 1
    if n_rec <= 0:
        return []
    keep = dict(first=np.ascontiguousarray(first, np.int64), kind_rec=np.ascontiguousarray(kind_rec, np.int64),
                has_rec=np.ascontiguousarray(has_rec, np.uint8), dp=np.ascontiguousarray(dp, np.int64), n=np.ascontiguousarray(n_amb, np.int64),
                ac_ref=np.ascontiguousarray(ac_ref, np.int64), med2_ref=np.ascontiguousarray(med2_ref, np.int64).reshape(-1),
                eaf=np.ascontiguousarray(eaf, np.float64), alt_idx=np.ascontiguousarray(alt_idx, np.int64),
                present_row=np.ascontiguousarray(present_row, np.uint8), ac=np.ascontiguousarray(ac, np.int64),
                med2_alt=np.ascontiguousarray(med2_alt, np.int64).reshape(-1), pu=np.ascontiguousarray(pu, np.float64),
                pw=np.ascontiguousarray(pw, np.float64), k=np.ascontiguousarray(k, np.int64), z=np.ascontiguousarray(z, np.float64).reshape(-1),
                p=np.ascontiguousarray(p, np.float64).reshape(-1))
    arg = _FormatIn()
    arg.sample, arg.n_rec = sample.encode(), n_rec
    for name, a in keep.items():
        setattr(arg, name, a.ctypes.data)
    text = C.c_void_p()
    off = np.empty(n_rec + 1, np.int64)
    lib.vfkio_format_info.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.vfkio_free.argtypes = [C.c_void_p]
    _dev._io_check(lib.vfkio_format_info(C.byref(arg), C.byref(text), C.c_void_p(off.ctypes.data)))
    try:
        buf = C.string_at(text, int(off[-1])).decode("utf-8")
    finally:
        lib.vfkio_free(text)
    if len(buf) != int(off[-1]):  # a sample name outside ASCII: byte offsets are not character offsets
        raw = buf.encode("utf-8")
        o = off.tolist()
        return [raw[a:b].decode("utf-8") for a, b in zip(o[:-1], o[1:])]
    o = off.tolist()
    return [buf[a:b] for a, b in zip(o[:-1], o[1:])]


def sample_keys(sample: str):
    """Per mask (which rank-sum pairs are present) the INFO keys of a single-BAM sample, in emission order."""
    base = tuple("{}_{}".format(sample, k) for k in _SINGLE_KEYS)
    out = []
    for mask in range(8):
        ks = base
        for m, (_name, tag) in enumerate(_RS_TAGS):
            if mask >> m & 1:
                ks += ("{}_{}".format(sample, tag), "{}_{}_pv".format(sample, tag))
        out.append(ks)
    return out


def text_templates(sample: str):
    return [";".join(k.replace("%", "%%") + "=%s" for k in ks) for ks in sample_keys(sample)]


def _format_chunk(args):
    return format_single_bam(*args)


_FORMAT_POOL = None


def _format_pool(n_procs: int):
    """Worker processes for the formatting of very large record sets (spawned: they import this module and numpy,
    never torch or the CUDA library)."""
    global _FORMAT_POOL
    if _FORMAT_POOL is None or _FORMAT_POOL[0] != n_procs:
        import atexit
        import multiprocessing
        from concurrent.futures import ProcessPoolExecutor
        if _FORMAT_POOL is not None:
            _FORMAT_POOL[1].shutdown(wait=False)
        pool = ProcessPoolExecutor(n_procs, mp_context=multiprocessing.get_context("spawn"))
        atexit.register(pool.shutdown, wait=False)
        _FORMAT_POOL = (n_procs, pool)
    return _FORMAT_POOL[1]


class SingleBamColumns:
    """The INFO fields of one sample backed by ONE BAM, for all records at once -- the common case, and the same
    strings `Engine._one` builds record by record (which stays the path of replicate BAMs): every value is
    stringified column-wise (format_single_bam), a record then only picks its tuple.  `keys[mask]` / `values[i]`:
    mask bit m set when the rank-sum pair of _RS_TAGS[m] is present.  `processes` > 1: the records are cut into
    chunks formatted by worker processes (worth it from some 10^5 records on)."""

    native_text = True  # as_text: printed by libvfkio (vfkio_format_info); False: by format_single_bam here

    def __init__(self, sample: str, rows: RowTable, res: BamResult, st: RoundedStats, eaf: np.ndarray, processes: int = 1,
                 as_text: bool = False):
        """as_text: `values[i]` is the sample's finished INFO text ("key=value;...") instead of the value tuple."""
        recs = rows.records
        n_rec, n_row = len(recs), rows.n_rows
        present_row = np.asarray([a == a.upper() for r in recs for a in r[3]], bool) if n_row else np.zeros(0, bool)
        first = rows.first
        per_rec = (rows.kind, res.has, res.dp, res.n, res.ac_ref, res.med2_ref, np.asarray(eaf, np.float64))
        per_row = (rows.alt, present_row, res.ac, res.med2_alt, st.pu_arr, st.pw_arr, st.k_arr, st.z_arr, st.p_arr)

        def args(i0, i1):
            r0, r1 = int(first[i0]), int(first[i1])
            kind, has, dp, n, ac_ref, med2_ref, e = (a[i0:i1] for a in per_rec)
            alt, present, ac, med2_alt, pu, pw, k, z, p = (a[r0:r1] for a in per_row)
            return (first[i0:i1 + 1] - r0, kind, alt, present, has, dp, n, ac, ac_ref, med2_ref, med2_alt, pu, pw, k, z, p, e,
                    text_templates(sample) if as_text else None)

        values = masks = None
        if as_text and self.native_text:
            self.values = format_single_bam_native(sample, *args(0, n_rec)[:17])
            self.masks, self.keys = None, sample_keys(sample)
            return
        if processes > 1 and n_rec >= 2 * processes:
            try:
                cuts = [n_rec * c // (4 * processes) for c in range(4 * processes + 1)]
                parts = list(_format_pool(processes).map(_format_chunk, [args(a, b) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]))
                values = [v for part in parts for v in part[0]]
                masks = [m for part in parts for m in part[1]]
            except Exception as e:  # a pool that cannot start (no importable main module, ...) must not cost the run
                import warnings
                warnings.warn("parallel formatting unavailable (%s): formatting in process" % e)
        if values is None:
            values, masks = format_single_bam(*args(0, n_rec))
        self.values, self.masks = values, masks
        self.keys = sample_keys(sample)


class Engine:
    """Drives the device for a set of VCF records against a set of BAM read batches."""
    vectorised_format = True  # single-BAM samples are formatted column-wise (SingleBamColumns); False: record by record
    format_processes = max(1, min(8, (__import__("os").cpu_count() or 1) // 2))  # worker processes for the formatting ...
    format_parallel_from = 200_000                                               # ... of record sets at least this large

    def __init__(self, device_index: int = 0, min_mapq: int = 0, min_baseq: int = 29, include_ambiguous: bool = True,
                 fpr: float = 5e-7, error_rate: float = 1e-3, devices: Optional[Sequence[int]] = None):
        """`devices`: CUDA device indexes to shard the pileup over (interval x BAM shards, one host
        thread per device, no collective); default: just `device_index`."""
        self.dev = _dev.Device(device_index)
        self.extra_devs = [_dev.Device(d) for d in (devices or [])[1:]] if devices and len(devices) > 1 else []
        self.params = _dev.make_params(min_mapq=min_mapq, min_baseq=min_baseq, include_ambiguous=include_ambiguous)
        self.fpr, self.error_rate = fpr, error_rate

    def close(self):
        self.dev.close()
        for d in self.extra_devs:
            d.close()

    # ---- device stages on host arrays -------------------------------------------------------
    def _cache_lock(self):
        import threading
        return self.__dict__.setdefault("_lock", threading.Lock())

    def _host_reads(self, batch: ReadBatch) -> "_dev.HostReads":
        """The batch as the C ABI takes it (no copy: the arrays of the ReadBatch themselves; page-locked when the decoder
        wrote them that way)."""
        with self._cache_lock():
            cache = self.__dict__.setdefault("_hreads", {})
            hit = cache.get(id(batch))
            if hit is None:
                hit = cache[id(batch)] = (batch, _dev.HostReads(batch))
        return hit[1]

    def _device_reads(self, batch: ReadBatch):
        """Device-resident copy (only the list-shaped auxiliary of the replicate merge needs one)."""
        with self._cache_lock():
            cache = self.__dict__.setdefault("_dreads", {})
            hit = cache.get(id(batch))
        if hit is None:
            hit = (batch, self.dev.upload_reads(self._host_reads(batch)))
            with self._cache_lock():
                cache[id(batch)] = hit
        return hit[1]

    def release_reads(self, batches: Optional[Sequence[ReadBatch]] = None):
        """Drop the cached host wrappers / device copies of `batches`, or of every batch."""
        with self._cache_lock():
            for name in ("_dreads", "_hreads"):
                if batches is None:
                    self.__dict__.pop(name, None)
                else:
                    for b in batches:
                        self.__dict__.get(name, {}).pop(id(b), None)

    def submit(self, batch: ReadBatch, units: VariantUnits, stop: np.ndarray, eaf_units: np.ndarray):
        """Enqueue one (BAM, units) batch on the device through the host entry point (vfk_annotate_host_async: one copy in,
        kernels, one copy out; the library keeps two batches in flight).  Returns a job for `finish`."""
        n = units.n_units
        counts = np.empty(n, _dev.COUNTS_DTYPE)
        stats = np.empty(n, _dev.STATS_DTYPE)
        if n == 0:
            return (None, counts, stats)
        if self.extra_devs and n >= 2 * (1 + len(self.extra_devs)):
            in_order = np.all((units.tid[1:] > units.tid[:-1]) | ((units.tid[1:] == units.tid[:-1]) & (units.pos[1:] >= units.pos[:-1])))
            if in_order:
                c, s = self._annotate_sharded(batch, units, stop, eaf_units)
                return (None, c, s)
        ticket = self.dev.annotate_host_async(self._host_reads(batch), units, stop, self.params, eaf_units, self.fpr, self.error_rate,
                                              counts, stats)
        return (ticket, counts, stats)

    def finish(self, job):
        ticket, counts, stats = job
        if ticket is not None:
            self.dev.wait(ticket)
        return counts, stats

    def _annotate_sharded(self, batch: ReadBatch, units: VariantUnits, stop: np.ndarray, eaf_units: np.ndarray):
        """Interval shards of one BAM over several devices (vafator_b200/sharding.py): one host thread per device, each
        shard a self-contained slice of the canonical batch; no collective, the per-unit records are concatenated."""
        from concurrent.futures import ThreadPoolExecutor
        from . import sharding
        devs = [self.dev] + self.extra_devs
        passes = _dev.read_passes(batch, self.params)
        shards = sharding.plan_shards(batch, units, len(devs), passes, stop)

        def work(k):
            sh, dev = shards[k], devs[k % len(devs)]
            sub = sharding.slice_reads(batch, sh.read_lo, sh.read_hi)
            su = sharding.slice_units(units, sh.unit_lo, sh.unit_hi)
            return dev.annotate_host(_dev.HostReads(sub), su, stop[sh.unit_lo:sh.unit_hi], self.params, eaf_units[sh.unit_lo:sh.unit_hi],
                                     self.fpr, self.error_rate)

        with ThreadPoolExecutor(len(devs)) as pool:
            parts = list(pool.map(work, range(len(shards))))
        return np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])

    def pileup_stats(self, batch: ReadBatch, units: VariantUnits, stop: np.ndarray, eaf_units) -> Tuple[np.ndarray, np.ndarray]:
        """(counts, stats) per unit of one BAM: the fused device call."""
        eaf_units = np.broadcast_to(np.asarray(eaf_units, np.float64), (units.n_units,)).copy()
        return self.finish(self.submit(batch, units, stop, eaf_units))

    def pileup(self, batch: ReadBatch, units: VariantUnits, stop: np.ndarray) -> np.ndarray:
        return self.pileup_stats(batch, units, stop, 0.5)[0]

    def epilogue(self, counts: np.ndarray, eaf: np.ndarray) -> np.ndarray:
        """FP64 statistics of explicit count records (merged replicates, rows without a unit of their own)."""
        n = len(counts)
        if n == 0:
            return np.zeros(0, _dev.STATS_DTYPE)
        import torch
        c = torch.from_numpy(np.ascontiguousarray(counts).view(np.uint8).reshape(-1)).to(self.dev.device)
        e = torch.from_numpy(np.ascontiguousarray(eaf, np.float64)).to(self.dev.device)
        buf = self.dev.epilogue(n, c, e, self.fpr, self.error_rate)
        return self.dev.to_host(buf, _dev.STATS_DTYPE, n)

    def _row_stats(self, res: "BamResult", counts_u: np.ndarray, stats_u: np.ndarray, eaf_row: np.ndarray) -> np.ndarray:
        """Statistics on the row table.  A row backed by a unit whose counts the host did not touch takes the unit's
        statistics as the fused call computed them; the others (no unit of their own: second ALT of an indel, MNV, record
        outside its group's span; or a count the reference's key lookup zeroes) are evaluated from the row's counts."""
        rc = res.row_counts
        nrow = len(rc)
        st = np.zeros(nrow, _dev.STATS_DTYPE)
        if nrow == 0:
            return st
        ru = res.row_unit
        have = ru >= 0
        same = have.copy()
        if len(counts_u):
            cu = counts_u[np.where(have, ru, 0)]
            for f in ("dp", "ac_alt", "ac_ref"):
                same &= rc[f] == cu[f]
            same &= np.all(rc["s2"] == cu["s2"], axis=1)
            st[same] = stats_u[ru[same]]
        else:
            same[:] = False
        fix = ~same
        if fix.any():
            st[fix] = self.epilogue(rc[fix], eaf_row[fix])
        return st

    def values(self, batch: ReadBatch, units: VariantUnits, stop: np.ndarray):
        """Packed per-read values of the units' REF / ALT lists: list of uint32 arrays."""
        if units.n_units == 0:
            return []
        vals, cnt = self.dev.pileup_values(self._device_reads(batch), self.dev.upload_units(units, stop), self.params)
        return [vals[u, :cnt[u]] for u in range(units.n_units)]

    def ranksum_s2(self, alt_lists, ref_lists) -> np.ndarray:
        return self.dev.ranksum_values(alt_lists, ref_lists)

    def _cross_bam(self, records, in_range, stop_rec, batches, needs, eaf_sample):
        """Rank sums whose ALT values come from one replicate BAM and REF values from another
        (vafator/annotator.py:285-287 dict.update semantics).  needs: [(i, j, b_ref, b_alt)]."""
        out = {}
        if not needs:
            return out
        lists = {}
        for b in sorted({n[2] for n in needs} | {n[3] for n in needs}):
            idx = sorted({(i, j) for i, j, br, ba in needs if b in (br, ba)})
            sub = [(records[i][0], records[i][1], records[i][2], [records[i][3][j]]) for i, j in idx]
            units = build_units(sub, batches[b].contigs)
            vals = self.values(batches[b], units, np.asarray([stop_rec[i] for i, j in idx], np.int32)[units.rec])
            for k, (i, j) in enumerate(idx):
                lists[(b, i, j)] = vals[units.unit_of[(k, 0)]]
        alt_lists = [lists[(ba, i, j)][(lists[(ba, i, j)] >> 29) & 1 == 1] for i, j, br, ba in needs]
        ref_lists = [lists[(br, i, j)][(lists[(br, i, j)] >> 28) & 1 == 1] for i, j, br, ba in needs]
        s2 = self.ranksum_s2(alt_lists, ref_lists)
        c = np.zeros(len(needs), _dev.COUNTS_DTYPE)
        c["ac_alt"] = [len(x) for x in alt_lists]
        c["ac_ref"] = [len(x) for x in ref_lists]
        c["s2"] = s2
        st = RoundedStats(self.epilogue(c, np.asarray([eaf_sample[i] for i, j, br, ba in needs], np.float64)))
        for q, n in enumerate(needs):
            out[(n[0], n[1])] = (st.z[q], st.p[q])
        return out

    # ---- the annotation -----------------------------------------------------------------------
    def annotate(self, records: Sequence[Record], bams: "OrderedDict[str, List[ReadBatch]]",
                 eaf: Dict[str, np.ndarray], as_text: bool = False, span: Optional[Tuple[int, int]] = None) -> List["OrderedDict"]:
        """INFO fields (ordered) for every record.  `eaf[sample][i]` = expected VAF of record i.
        as_text: the same fields as one "key=value;key=value" string per record (what a VCF writer appends).
        span: the records are a chunk of ONE chromosome group whose first / last record lie at (first.POS, last.POS) -- the span
        the reference piles up (vafator/pileups.py:137-138) and the bound on the reads it ever fetches."""
        rows = RowTable(records)
        in_range, stop_rec = chrom_groups(records) if span is None else span_range(records, span)
        fmt = SampleFormatter(rows)
        per_sample = OrderedDict()
        units_of_header = {}
        # every (sample, BAM) is one fused device call; they are submitted ahead of being collected so that the copies of one
        # batch overlap the host gather and the kernels of the previous ones (the device's staging slots)
        tasks = []
        for sample, batches in bams.items():
            for batch in batches:
                contigs_key = tuple(batch.contigs)
                if contigs_key not in units_of_header:  # BAMs of one run nearly always share their reference
                    recs_in = records if in_range.all() else [r if ok else (r[0], r[1], "", []) for r, ok in zip(records, in_range)]
                    units_of_header[contigs_key] = build_units(recs_in, batch.contigs)
                tasks.append((sample, batch, units_of_header[contigs_key]))
        jobs, done = [], {}
        n_slots = getattr(getattr(self, "dev", None), "slots", 2)  # batches the host entry point keeps in flight
        for t, (sample, batch, units) in enumerate(tasks):
            if t >= n_slots:
                done[t - n_slots] = self.finish(jobs[t - n_slots])
            stop_u = stop_rec[units.rec] if units.n_units else np.zeros(0, np.int32)
            eaf_u = np.asarray(eaf[sample], np.float64)[units.rec] if units.n_units else np.zeros(0)
            jobs.append(self.submit(batch, units, stop_u, eaf_u))
        for t in range(len(tasks)):
            if t not in done:
                done[t] = self.finish(jobs[t])
        t = 0
        for sample, batches in bams.items():
            eaf_row = np.asarray(eaf[sample], np.float64)[rows.rec] if rows.n_rows else np.zeros(0)
            results, stats = [], []
            for batch in batches:
                units = tasks[t][2]
                counts, stats_u = done[t]
                t += 1
                bad = counts["status"] & _dev.ST_READLEN if len(counts) else np.zeros(0, np.uint32)
                if len(counts) and bad.any():
                    raise _dev.VfkError("a read position exceeded the histogram range (reads longer than supported)")
                res = BamResult(rows, units, counts, in_range)
                results.append(res)
                stats.append(RoundedStats(self._row_stats(res, counts, stats_u, eaf_row)))
            merged_stats = None
            if len(batches) > 1:
                mc = np.zeros(rows.n_rows, _dev.COUNTS_DTYPE)
                mc["dp"] = sum(r.dp for r in results)[rows.rec] if rows.n_rows else 0
                mc["ac_alt"] = sum(r.ac for r in results)
                merged_stats = RoundedStats(self.epilogue(mc, eaf_row))
            cross = {}
            if len(batches) > 1:
                needs = []
                for i in range(len(records)):
                    if rows.kind[i] != KIND_SNV:
                        continue  # indel lists exist in every BAM that has a column: the last one wins for both
                    states = [fmt.key_state(i, res) for res in results]
                    b_ref = max([b for b in range(len(results)) if states[b][0][0]], default=-1)
                    for j in range(len(records[i][3])):
                        b_alt = max([b for b in range(len(results)) if states[b][1 + j][0]], default=-1)
                        if b_ref >= 0 and b_alt >= 0 and b_ref != b_alt:
                            needs.append((i, j, b_ref, b_alt))
                cross = self._cross_bam(records, in_range, stop_rec, batches, needs, np.asarray(eaf[sample], np.float64))
            per_sample[sample] = (results, stats, merged_stats, cross)
        columns = {}
        if self.vectorised_format:
            procs = self.format_processes if len(records) >= self.format_parallel_from else 1
            all_single = all(len(results) == 1 for results, _s, _m, _c in per_sample.values())
            if as_text and all_single:  # the text of every sample, concatenated per record
                texts = [SingleBamColumns(sample, rows, results[0], stats[0], eaf[sample], processes=procs, as_text=True).values
                         for sample, (results, stats, _m, _c) in per_sample.items()]
                return list(texts[0]) if len(texts) == 1 else [";".join(t) for t in zip(*texts)]
            columns = {sample: SingleBamColumns(sample, rows, results[0], stats[0], eaf[sample], processes=procs)
                       for sample, (results, stats, _m, _c) in per_sample.items() if len(results) == 1}
        as_wanted = (lambda infos: [";".join(["%s=%s" % kv for kv in info.items()]) for info in infos]) if as_text else (lambda infos: infos)
        out = []
        if len(columns) == len(per_sample):  # every sample has a single BAM: nothing is left to do record by record
            cols = list(columns.values())
            for i in range(len(records)):
                info = OrderedDict()
                for col in cols:
                    info.update(zip(col.keys[col.masks[i]], col.values[i]))
                out.append(info)
            return as_wanted(out)
        first, kinds = rows.first.tolist(), rows.kind.tolist()
        for i, (chrom, pos, ref, alts) in enumerate(records):
            info = OrderedDict()
            r0, r1 = first[i], first[i + 1]
            kind = kinds[i]
            for sample, (results, stats, merged_stats, cross) in per_sample.items():
                col = columns.get(sample)
                if col is not None:
                    info.update(zip(col.keys[col.masks[i]], col.values[i]))
                    continue
                e = float(eaf[sample][i])
                states = [fmt.key_state(i, res) for res in results]
                nb = len(results)
                if nb > 1:
                    for b, (res, st) in enumerate(zip(results, stats)):
                        info.update(self._one(sample, "_%d" % (b + 1), False, alts, kind, res, st, states[b], i, r0, r1, e))
                    info.update(self._merged(sample, alts, kind, results, stats, merged_stats, states, i, r0, r1, e, cross))
                else:
                    info.update(self._one(sample, "", True, alts, kind, results[0], stats[0], states[0], i, r0, r1, e))
            out.append(info)
        return as_wanted(out)

    @staticmethod
    def _median_strings(kind, state_ref, state_alts):
        med = {}
        for mi, name in enumerate(METRICS):
            vals = []
            for has_key, n, m2 in [state_ref] + list(state_alts):
                if name == "bq" and kind != KIND_SNV:
                    vals.append("0")  # indel metrics carry no base qualities (pileups.py:301,304)
                elif not has_key:
                    vals.append("0")  # Counter miss
                elif n == 0:
                    vals.append("nan")  # key present, empty list
                else:
                    vals.append(_fmt_float(m2[mi] / 2.0))
            med[name] = vals
        return med

    @staticmethod
    def _ranksum_strings(kind, state_ref, state_alts, st, r0):
        rs = {}
        for mi, name in enumerate(METRICS):
            zs, ps = [], []
            if not (name == "bq" and kind != KIND_SNV):
                for j, (has_key, n, _) in enumerate(state_alts):
                    if has_key and n > 0 and state_ref[0] and state_ref[1] > 0:
                        z, p = st.z[r0 + j][mi], st.p[r0 + j][mi]
                        if not (z != z or p != p):
                            zs.append(str(z))
                            ps.append(str(p))
            rs[name] = (zs, ps)
        return rs

    def _one(self, sample, suffix, with_eaf, alts, kind, res, st, state, i, r0, r1, e):
        ac = [int(res.ac[r]) for r in range(r0, r1)]
        dp = int(res.dp[i])
        pu = st.pu[r0:r1]
        pw, k = (st.pw[r0], st.k[r0]) if r1 > r0 else (0.0, 1)
        med = self._median_strings(kind, state[0], state[1:])
        rs = self._ranksum_strings(kind, state[0], state[1:], st, r0)
        return format_fields(sample, suffix, with_eaf, alts, kind, ac, dp, int(res.n[i]), e, pu, pw, k, med, rs)

    def _merged(self, sample, alts, kind, results, stats, merged_stats, states, i, r0, r1, e, cross):
        """Sample-level fields over replicate BAMs (vafator/annotator.py:277-302): counts and depth add,
        MEDIANS ADD (Counter.update), distributions are replaced allele by allele by the last BAM that
        has the allele (dict.update), so the rank sums describe that BAM only."""
        nb = len(results)
        ac = [int(sum(res.ac[r] for res in results)) for r in range(r0, r1)]
        dp = int(sum(res.dp[i] for res in results))
        n = int(sum(res.n[i] for res in results))
        pu = merged_stats.pu[r0:r1]
        pw, k = (merged_stats.pw[r0], merged_stats.k[r0]) if r1 > r0 else (0.0, 1)
        med = {}
        for mi, name in enumerate(METRICS):
            vals = []
            for a in range(1 + len(alts)):
                total, seen = 0.0, False
                for b in range(nb):
                    has_key, cnt, m2 = states[b][a]
                    if name == "bq" and kind != KIND_SNV:
                        has_key = False
                    if has_key:
                        seen = True
                        total = total + (m2[mi] / 2.0 if cnt > 0 else math.nan)
                vals.append(_fmt_float(total) if seen else "0")
            med[name] = vals
        rs = {}
        for mi, name in enumerate(METRICS):
            zs, ps = [], []
            if not (name == "bq" and kind != KIND_SNV):
                def last_with(a):
                    for b in range(nb - 1, -1, -1):
                        if states[b][a][0]:
                            return b
                    return -1
                b_ref = last_with(0)
                for j in range(len(alts)):
                    b_alt = last_with(1 + j)
                    if b_ref < 0 or b_alt < 0:
                        continue
                    if states[b_ref][0][1] == 0 or states[b_alt][1 + j][1] == 0:
                        continue
                    zz, pp = cross[(i, j)] if b_ref != b_alt else (stats[b_alt].z[r0 + j], stats[b_alt].p[r0 + j])
                    z, p = zz[mi], pp[mi]
                    if not (z != z or p != p):
                        zs.append(str(z))
                        ps.append(str(p))
            rs[name] = (zs, ps)
        return format_fields(sample, "", True, alts, kind, ac, dp, n, e, pu, pw, k, med, rs)
