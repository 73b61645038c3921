"""CPU-side checks of the product's host code: HBM packing, mate linking, C-ABI surface."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import helpers as H
from oracle import c_oracle
from vafator_b200 import build, device
from vafator_b200.batch import ReadBatch, build_units, SEQ_CHARS, CODE_NONE

SCENARIOS = H.golden_scenarios()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_quality_scaling_is_integer_exact():
    """htslib stores (uint8_t)(0.8 * q); the kernel uses (4*q)/5 -- identical for every byte value."""
    for q in range(256):
        assert int(0.8 * q) == (4 * q) // 5


def one_event_word(ops, l_qseq):
    """include/vfk.h VFK_SHAPE_ONE_EVENT: [H][S] M [ I|D|N M ] [S][H] -> lead | m1 << 12 | event << 24 | len << 26."""
    m = re.fullmatch(r"(?:\d+H)*(?:(\d+)S)?(\d+)[M=X](?:(\d+)([IDN])(\d+)[M=X])?(?:\d+S)?(?:\d+H)*", "".join(n + o for n, o in ops))
    if not m:
        return None
    lead, m1 = int(m.group(1) or 0), int(m.group(2))
    ev, ln = ("IDN".index(m.group(4)) + 1, int(m.group(3))) if m.group(4) else (0, 0)
    if m1 == 0 or m1 >= 4096 or lead >= 4096 or ln >= 64 or l_qseq > 4095 or (m.group(4) and (ln == 0 or int(m.group(5)) == 0)):
        return None
    return lead | (m1 << 12) | (ev << 24) | (ln << 26)


@pytest.mark.parametrize("path", SCENARIOS, ids=[os.path.basename(p) for p in SCENARIOS])
def test_pack_layout_round_trips(path):
    sc = H.load_scenario(path)
    for recs, batch in zip(sc["bams"], H.scenario_batches(sc)):
        pk = device.pack_reads(batch, device.make_params())
        assert pk.n_reads == len(recs)
        assert pk.contig_off[0] == 0 and pk.contig_off[-1] == len(recs)
        assert np.all(np.diff(pk.starts.astype(np.int64))[batch.tid[1:] == batch.tid[:-1]] >= 0)
        for i, r in enumerate(recs):
            assert pk.hot["pos"][i] == r["pos"]
            assert (pk.hot["fmk"][i] & 0xFFFF) == r["flag"] and ((pk.hot["fmk"][i] >> 16) & 0xFF) == r["mapq"]
            lq = len(r["seq"])
            base = int(pk.hot["pay"][i]) * 32
            gran = pk.payload[base:base + 32 * ((lq + 19) // 20)].view(np.uint64)
            for q in range(lq):
                g, rr = divmod(q, 5)
                w = (int(gran[g]) >> (12 * rr)) & 0xFFF
                assert (w & 0xFF) == ord(r["qual"][q]) - 33
                assert SEQ_CHARS[w >> 8] == r["seq"][q]
            ops = re.findall(r"(\d+)([MIDNSHP=X])", r["cigar"])
            span = sum(int(n) for n, o in ops if o in "MDN=X")
            assert pk.hot["span"][i] == span
            simple = len(ops) == 1 and ops[0][1] in "M=X"
            event = None if simple else one_event_word(ops, lq)
            assert (pk.hot["fmk"][i] >> 24) == (0 if simple else 1 if event is None else 2)
            assert pk.cold["event"][i] == (event or 0)
            if not simple:
                co, nc = int(pk.cold["cigar_off"][i]), int(pk.cold["n_cigar"][i])
                assert [(int(w) >> 4, "MIDNSHP=X"[int(w) & 15]) for w in pk.cigar[co:co + nc]] == [(int(n), o) for n, o in ops]
            assert pk.cold["l_qseq"][i] == lq
        # position index: back pointers
        for t in range(len(sc["contigs"])):
            a, b = int(pk.contig_off[t]), int(pk.contig_off[t + 1])
            ends = pk.hot["pos"][a:b].astype(np.int64) + pk.hot["span"][a:b]
            pmax = np.maximum.accumulate(ends) if b > a else ends
            for i in range(a, b):
                want = a + int(np.searchsorted(pmax[:i - a], pk.starts[i], side="right")) if i > a else a
                # first j <= i with pmax[j] > start[i]; pmax is non-decreasing so searchsorted applies
                assert pk.sb[i, 1] == min(want, i), (i, pk.sb[i, 1], want)


@pytest.mark.parametrize("path", SCENARIOS, ids=[os.path.basename(p) for p in SCENARIOS])
def test_link_mates_matches_oracle(path):
    sc = H.load_scenario(path)
    for batch in H.scenario_batches(sc):
        for mq, orphans, overlaps in ((0, True, True), (1, True, True), (30, False, True), (0, True, False)):
            prm = device.make_params(min_mapq=mq, ignore_orphans=orphans, ignore_overlaps=overlaps)
            got = device.link_mates(batch, prm)
            want = c_oracle.link_mates(batch.as_dict(), c_oracle.params(min_mapq=mq, ignore_orphans=orphans,
                                                                        ignore_overlaps=overlaps))
            partner = np.where(got == device.NO_MATE, -1, (got & 0x7FFFFFFF).astype(np.int64))
            assert np.array_equal(partner, want)
            linked = np.nonzero(partner >= 0)[0]
            assert np.array_equal(partner[partner[linked]], linked)  # symmetric
            if overlaps:
                assert len(linked) > 0 or sc["name"] == "none"


def test_units_encoding():
    u = build_units([("c", 10, "A", ["T", "G"]), ("c", 11, "A", ["AT"]), ("c", 12, "ATT", ["A"]),
                     ("c", 13, "AT", ["GC"]), ("c", 14, "a", ["t"]), ("c", 15, "A", ["AnT", "ATT"])], ["c"])
    assert u.n_units == 2 + 1 + 1 + 0 + 1 + 1
    assert u.unit_of[(0, 0)] == 0 and u.unit_of[(0, 1)] == 1 and (3, 0) not in u.unit_of
    assert list(u.pos[:2]) == [9, 9]
    assert (u.meta[0] & 3) == 0 and ((u.meta[0] >> 8) & 0xFF) == SEQ_CHARS.index("A") and ((u.meta[1] >> 16) & 0xFF) == SEQ_CHARS.index("G")
    assert (u.meta[2] & 3) == 1 and u.length[2] == 1 and u.ins_seq[u.seq_off[2]] == SEQ_CHARS.index("T")
    assert (u.meta[3] & 3) == 2 and u.length[3] == 2
    assert ((u.meta[4] >> 8) & 0xFF) == CODE_NONE and ((u.meta[4] >> 16) & 0xFF) == CODE_NONE  # lowercase never matches
    assert u.ins_seq[u.seq_off[5]] == CODE_NONE and u.ins_seq[u.seq_off[5] + 1] == SEQ_CHARS.index("T")
    with pytest.raises(ValueError):
        build_units([("zz", 1, "A", ["T"])], ["c"])


def test_cabi_library_exports_every_declared_symbol():
    """The shared libraries build (nvcc cross-compiles without a GPU), load, and export exactly what
    include/*.h declares.  No compute call is made here."""
    libvfk, libio = build.build_all()
    for header, path in (("vfk.h", libvfk), ("vfk_io.h", libio)):
        text = open(os.path.join(ROOT, "include", header)).read()
        declared = set(re.findall(r"\b(vfk(?:io)?_[a-z_]+)\s*\(", text))
        lib = C.CDLL(path)
        for name in sorted(declared):
            assert hasattr(lib, name), (header, name)
    lib = C.CDLL(libvfk)
    assert lib.vfk_version() == 100
    assert np.dtype(device.COUNTS_DTYPE).itemsize == 80
    sass = subprocess.run(["cuobjdump", "-lelf", libvfk], capture_output=True, text=True).stdout
    assert "sm_100a" in sass


# ---- column store (include/vfk.h): every sector against an independent CIGAR walk ----
def _walk(read):
    """[(ref_pos, state, base_char, qual, qpos)] for every reference position the read covers; states as in vfk.h."""
    ops = re.findall(r"(\d+)([MIDNSHP=X])", read["cigar"])
    x, y, out = read["pos"], 0, []
    for n, o in ops:
        n = int(n)
        if o in "M=X":
            for _ in range(n):
                out.append((x, 1, read["seq"][y], ord(read["qual"][y]) - 33, y))
                x += 1
                y += 1
        elif o in "DN":
            for _ in range(n):
                out.append((x, 2 if o == "D" else 3, None, 0, y))
                x += 1
        elif o in "IS":
            y += n
    return out


@pytest.mark.parametrize("path", SCENARIOS, ids=[os.path.basename(p) for p in SCENARIOS])
def test_column_store_matches_cigar_walk(path):
    sc = H.load_scenario(path)
    W = 12
    for recs, batch in zip(sc["bams"], H.scenario_batches(sc)):
        for mq, orphans in ((0, True), (20, False)):
            params = device.make_params(min_mapq=mq, ignore_orphans=orphans)
            pk = device.pack_reads(batch, params, columns=True)
            passes = device.read_passes(batch, params)
            sectors = pk.col_sectors.reshape(-1, 32)
            entries = sectors[:, :24].copy().view("<u2").reshape(-1, 12)
            h0, h1 = sectors[:, 24:28].copy().view("<u4").ravel(), sectors[:, 28:32].copy().view("<u4").ravel()
            walks = [_walk(r) for r in recs]
            n_seen = 0
            for t in range(len(sc["contigs"])):
                n_win = int(pk.col_wbase[t + 1] - pk.col_wbase[t])
                members = [[] for _ in range(n_win)]
                for i, (r, wk) in enumerate(zip(recs, walks)):
                    if r["tid"] == t and wk:
                        for w in range(wk[0][0] // W, wk[-1][0] // W + 1):
                            members[w].append(i)
                for w in range(n_win):
                    a, b = int(pk.col_off[pk.col_wbase[t] + w]), int(pk.col_off[pk.col_wbase[t] + w + 1])
                    assert b - a == len(members[w])
                    for s, i in zip(range(a, b), members[w]):
                        n_seen += 1
                        by_pos = {x: (st, ch, q, y) for x, st, ch, q, y in walks[i]}
                        complex_ = bool(h0[s] >> 31)
                        assert bool((h0[s] >> 28) & 1) == bool(passes[i]) and ((h0[s] >> 12) & 0xFF) == recs[i]["mapq"]
                        mask, has_ins = int(h1[s]) & 0xFFF, bool((h1[s] >> 28) & 1)
                        ins_at, ins_len = (int(h1[s]) >> 12) & 15, (int(h1[s]) >> 16) & 0xFFF
                        for p in range(W):
                            x = w * W + p
                            e = int(entries[s, p])
                            st = (e >> 12) & 3
                            if x not in by_pos:
                                assert e == 0
                                continue
                            want_st, ch, q, y = by_pos[x]
                            assert st == want_st
                            if st == 1:
                                assert (e & 0xFF) == q and SEQ_CHARS[(e >> 8) & 15] == ch and (mask >> p) & 1
                                if not complex_:
                                    qpos = (int(h0[s]) & 0xFFF) + bin(mask & ((1 << p) - 1)).count("1") + (ins_len if has_ins and p > ins_at else 0)
                                    assert qpos == y, (i, x, recs[i]["cigar"])
                        # overlap partner: the sector `off` away belongs to the linked read and points back
                        off = (int(h0[s]) >> 20) & 0xFF
                        off = off - 256 if off >= 128 else off
                        linked = pk.mate[i] != 0xFFFFFFFF and (recs[i]["flag"] & 2) and not (recs[i]["flag"] & 8)
                        if off:
                            m = int(pk.mate[i] & 0x7FFFFFFF)
                            assert linked and members[w][s - a + off] == m
                            assert bool((h0[s] >> 30) & 1) == (i < m) and bool((h0[s] >> 29) & 1) == bool(pk.mate[i] >> 31)
                        elif linked and not complex_:
                            assert int(pk.mate[i] & 0x7FFFFFFF) not in members[w]
            assert n_seen == sectors.shape[0]


def test_column_store_on_noisy_synthetic_reads():
    """Same check on generated reads with many clips / indels / skips / dense pairing: the sectors the builder flags
    unusable are rare, everything else decodes to the CIGAR walk, and partner offsets point at the linked read."""
    from vafator_b200 import synth
    cfg = synth.wes(n_tiles=40, indel_period=40, snv_period=30, p_softclip=0.3, p_del=0.2, p_ins=0.2, p_skip=0.08,
                    frag_mean=180.0, frag_sd=30.0, pairs_per_tile=60)
    batch = synth.reads(cfg)
    params = device.make_params(min_mapq=1)
    pk = device.pack_reads(batch, params, columns=True)
    W = 12
    sectors = pk.col_sectors.reshape(-1, 32)
    entries = sectors[:, :24].copy().view("<u2").reshape(-1, 12)
    h0, h1 = sectors[:, 24:28].copy().view("<u4").ravel(), sectors[:, 28:32].copy().view("<u4").ravel()
    n = pk.n_reads
    spans, cov = np.zeros(n, np.int64), []
    for i in range(n):
        c = batch.cigar[batch.cigar_off[i]:batch.cigar_off[i] + batch.n_cigar[i]]
        x, y, states = int(batch.pos[i]), 0, {}
        for w_ in c:
            op, l = int(w_) & 15, int(w_) >> 4
            if op in (0, 7, 8):
                for _ in range(l):
                    states[x] = (1, y)
                    x += 1
                    y += 1
            elif op in (2, 3):
                for _ in range(l):
                    states[x] = (2 if op == 2 else 3, y)
                    x += 1
            elif op in (1, 4):
                y += l
        cov.append(states)
        spans[i] = x - int(batch.pos[i])
    members = {}
    for i in range(n):
        if spans[i] > 0:
            t = int(batch.tid[i])
            for w_ in range(int(batch.pos[i]) // W, (int(batch.pos[i]) + int(spans[i]) - 1) // W + 1):
                members.setdefault((t, w_), []).append(i)
    assert sum(len(v) for v in members.values()) == sectors.shape[0]
    n_complex = 0
    for (t, w_), reads in members.items():
        a = int(pk.col_off[pk.col_wbase[t] + w_])
        assert int(pk.col_off[pk.col_wbase[t] + w_ + 1]) - a == len(reads)
        for k, i in enumerate(reads):
            s = a + k
            complex_ = bool(h0[s] >> 31)
            n_complex += complex_
            mask, has_ins = int(h1[s]) & 0xFFF, bool((h1[s] >> 28) & 1)
            ins_at, ins_len = (int(h1[s]) >> 12) & 15, (int(h1[s]) >> 16) & 0xFFF
            for p in range(W):
                st = (int(entries[s, p]) >> 12) & 3
                want = cov[i].get(w_ * W + p)
                assert st == (want[0] if want else 0)
                if st == 1 and not complex_:
                    q = (int(h0[s]) & 0xFFF) + bin(mask & ((1 << p) - 1)).count("1") + (ins_len if has_ins and p > ins_at else 0)
                    assert q == want[1]
                    assert (int(entries[s, p]) & 0xFF) == int(batch.qual[batch.qual_off[i] + want[1]])
            off = (int(h0[s]) >> 20) & 0xFF
            off = off - 256 if off >= 128 else off
            if off:
                assert reads[k + off] == int(pk.mate[i] & 0x7FFFFFFF)
    assert n_complex < 0.02 * sectors.shape[0]


def test_column_store_window_transposed_form():
    """vfkio_columns_transpose: header pairs first, then one row of entries per window position -- checked against
    a numpy restatement, window by window, on noisy synthetic reads."""
    import ctypes as C
    from vafator_b200 import synth
    batch = synth.reads(synth.wgs(n_contigs=2, contig_len=60_000, p_del=0.05, p_ins=0.05, p_skip=0.01))
    params = device.make_params(min_mapq=1, min_baseq=30)
    pk = device.pack_reads(batch, params, columns=True)
    woff = np.ascontiguousarray(pk.col_off, np.uint32)
    src = np.ascontiguousarray(pk.col_sectors, np.uint8)
    out = np.full_like(src, 0xEE)
    lib = device.libio()
    lib.vfkio_columns_transpose.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    assert lib.vfkio_columns_transpose(woff.ctypes.data, len(woff) - 1, src.ctypes.data, out.ctypes.data) == 0
    assert lib.vfkio_columns_transpose(woff.ctypes.data, len(woff) - 1, src.ctypes.data, src.ctypes.data) != 0  # not in place
    sec = src.reshape(-1, 32)
    entries = sec[:, :24].copy().view(np.uint16).reshape(-1, 12)
    headers = sec[:, 24:].copy().view(np.uint32).reshape(-1, 2)
    checked = 0
    for w in np.random.default_rng(3).choice(len(woff) - 1, 400, replace=False):
        a, b = int(woff[w]), int(woff[w + 1])
        n = b - a
        if n == 0:
            continue
        run = out[a * 32:b * 32]
        assert np.array_equal(run[:8 * n].view(np.uint32).reshape(n, 2), headers[a:b])
        assert np.array_equal(run[8 * n:].view(np.uint16).reshape(12, n), entries[a:b].T)
        checked += n
    assert checked > 1000


def test_column_store_is_skipped_for_splice_dominated_batches():
    """RNA-like reads (most of them spanning kilobases of reference skip) would blow the column store up to many
    times the size of the reads: pack_reads keeps such a batch read-major (with a warning), ordinary batches --
    including skip-heavy test mixes -- get the store."""
    import warnings
    from vafator_b200.batch import ReadBatch
    params = device.make_params(min_mapq=1, min_baseq=30)
    recs = [dict(qname="r%d" % i, flag=0, tid=0, pos=100 + 7 * i, mapq=60, cigar="20M60000N30M", seq="ACGTA" * 10, qual=[35] * 50)
            for i in range(40)]
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        pk = device.pack_reads(ReadBatch.from_records(recs, ["c1"]), params, columns=True)
    assert not pk.has_columns and any("reference skips dominate" in str(x.message) for x in w)
    from vafator_b200 import synth
    noisy = synth.reads(synth.wes(n_tiles=300, p_skip=0.25, p_del=0.15))
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        assert device.pack_reads(noisy, params, columns=True).has_columns and not w


def test_transposed_store_addressing_matches_the_sector_form(monkeypatch):
    """VFK_COL_TRANSPOSED=1 (experimental): pack_reads hands out the window-transposed store; the addressing the
    kernel's load_col_read<true> uses (header pair k of the run, entry k of row p) returns what the sector form
    returns for sector off + k, position p."""
    from vafator_b200 import synth
    batch = synth.reads(synth.wgs(n_contigs=1, contig_len=40_000, p_del=0.05, p_ins=0.05))
    params = device.make_params(min_mapq=1, min_baseq=30)
    monkeypatch.delenv("VFK_COL_TRANSPOSED", raising=False)
    a = device.pack_reads(batch, params, columns=True)
    monkeypatch.setenv("VFK_COL_TRANSPOSED", "1")
    t = device.pack_reads(batch, params, columns=True)
    assert np.array_equal(a.col_off, t.col_off) and a.col_sectors.shape == t.col_sectors.shape
    sa, st = np.asarray(a.col_sectors), np.asarray(t.col_sectors)
    rng = np.random.default_rng(0)
    n = 0
    for w in rng.choice(len(a.col_off) - 1, 1500):
        off, cnt = int(a.col_off[w]), int(a.col_off[w + 1] - a.col_off[w])
        if cnt == 0:
            continue
        k, p = int(rng.integers(0, cnt)), int(rng.integers(0, 12))
        sec = sa[(off + k) * 32:(off + k + 1) * 32]
        run = st[off * 32:(off + cnt) * 32]
        assert int(sec[:24].view(np.uint16)[p]) == int(run[8 * cnt:].view(np.uint16)[p * cnt + k])
        assert np.array_equal(sec[24:].view(np.uint32), run[:8 * cnt].view(np.uint32).reshape(cnt, 2)[k])
        n += 1
    assert n > 1000


def test_select_compare_emulation_matches_the_definition():
    """The staged MQ ranking (vfk_pileup.cu select_compare: one round per distinct value -- warp minimum, ballot of
    its holders) restated lane by lane: for every contributing lane, lt = the contributing lanes holding a smaller
    value and eq = those holding an equal one -- what radix_compare delivers and radix_finish consumes."""
    rng = np.random.default_rng(9)
    for _ in range(300):
        contributing = int(rng.integers(0, 1 << 32)) & int(rng.integers(0, 1 << 32)) if rng.random() < 0.7 else int(rng.integers(0, 1 << 32))
        v = rng.choice([0, 1, 20, 29, 60, 60, 60, 255], 32).astype(np.uint32)
        lanes = [k for k in range(32) if contributing >> k & 1]
        lt = [0] * 32
        eq = [contributing] * 32
        rem, done = contributing, 0
        rounds = 0
        while rem:
            waiting = [(rem >> k) & 1 for k in range(32)]
            m = min(int(v[k]) if waiting[k] else 0xFFFFFFFF for k in range(32))       # __reduce_min_sync
            g = sum(1 << k for k in range(32) if waiting[k] and int(v[k]) == m)        # __ballot_sync
            for k in range(32):
                if g >> k & 1:
                    lt[k], eq[k] = done, g
            done |= g
            rem &= ~g
            rounds += 1
        assert rounds == len({int(v[k]) for k in lanes})
        for k in lanes:
            assert lt[k] == sum(1 << j for j in lanes if v[j] < v[k])
            assert eq[k] == sum(1 << j for j in lanes if v[j] == v[k])


@pytest.mark.parametrize("transposed", [0, 1], ids=["sector_form", "transposed_form"])
def test_host_gather_of_column_runs(monkeypatch, transposed):
    """vfk_host_gather_columns (host half of the staged VFK_HOST_GATHER path, no CUDA call): for every unit the compact
    run holds exactly what load_col_read<.., CP> will address -- header pair k at 8 k, entry k of the unit's own
    position behind the header pairs -- and equals the sector form's sector off + k; units the column kernel does not
    serve are left without a run."""
    import ctypes as C
    from vafator_b200 import synth
    from vafator_b200.batch import build_units
    cfg = synth.wes(n_tiles=400, indel_period=300, multiallelic_pct=10)
    batch = synth.reads(cfg)
    units = build_units(synth.variant_records(cfg), cfg.contigs)
    params = device.make_params(min_mapq=1, min_baseq=30)
    monkeypatch.delenv("VFK_COL_TRANSPOSED", raising=False)
    ref = device.pack_reads(batch, params, columns=True)
    if transposed:
        monkeypatch.setenv("VFK_COL_TRANSPOSED", "1")
    pk = device.pack_reads(batch, params, columns=True)
    w = pk.col_wbase[units.tid] + units.pos // 12
    cu = np.zeros((units.n_units, 4), np.uint32)
    cu[:, 0] = pk.col_off[w]
    cu[:, 1] = pk.col_off[w + 1] - pk.col_off[w]
    cu[:, 2] = units.pos % 12
    cu[:, 3] = units.meta
    cu[5, 1] = 70  # a column too deep for the kernel
    lib = device.libvfk()
    lib.vfk_host_gather_columns.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    cu2 = np.zeros_like(cu)
    need = C.c_int64(0)
    sectors = np.ascontiguousarray(pk.col_sectors)
    assert lib.vfk_host_gather_columns(cu.ctypes.data, len(cu), sectors.ctypes.data, transposed, cu2.ctypes.data, None, 0, C.byref(need)) == 0
    out = np.full(need.value, 0xCD, np.uint8)
    assert lib.vfk_host_gather_columns(cu.ctypes.data, len(cu), sectors.ctypes.data, transposed, cu2.ctypes.data, out.ctypes.data,
                                       need.value, C.byref(need)) == 0
    assert lib.vfk_host_gather_columns(cu.ctypes.data, len(cu), sectors.ctypes.data, transposed, cu2.ctypes.data, out.ctypes.data,
                                       need.value - 32, C.byref(need)) != 0  # too small a buffer is refused
    sa = np.asarray(ref.col_sectors)
    kind = cu[:, 3] & 3
    n_checked = 0
    for u in range(len(cu)):
        off, cnt, p = int(cu[u, 0]), int(cu[u, 1]), int(cu[u, 2])
        assert cu2[u, 2] == p and cu2[u, 3] == cu[u, 3]
        if kind[u] != 0:
            assert cu2[u, 1] == 0
            continue
        if cnt > 64:
            assert cu2[u, 1] == 65 and cu2[u, 0] == 0
            continue
        assert cu2[u, 1] == cnt
        run = out[int(cu2[u, 0]) * 32:]
        for k in range(cnt):
            sec = sa[(off + k) * 32:(off + k + 1) * 32]
            assert np.array_equal(run[8 * k:8 * k + 8], sec[24:32])                                  # header pair k
            assert np.array_equal(run[8 * cnt + 2 * k:8 * cnt + 2 * k + 2], sec[2 * p:2 * p + 2])   # entry k of position p
            n_checked += 1
    assert n_checked > 3000 and need.value < 0.45 * 32 * cu[kind == 0, 1].sum()  # ~10 of every 32 bytes
