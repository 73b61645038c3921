"""GPU vs the C oracle on seeded synthetic batches (sizes the oracle finishes in seconds), plus
size-independent properties of the path: idempotence, unit-order invariance, shard invariance."""
import numpy as np
import pytest

from oracle import c_oracle
from vafator_b200 import device, synth
from vafator_b200.batch import build_units

pytestmark = pytest.mark.gpu

INT_FIELDS = ("dp", "n_amb", "ac_ref", "ac_alt")


@pytest.fixture(scope="module")
def dev():
    d = device.Device(0)
    yield d
    d.close()


def _setup(cfg):
    batch = synth.reads(cfg)
    recs = synth.variant_records(cfg)
    units = build_units(recs, cfg.contigs)
    last = {}
    for c, p, r, a in recs:
        last[c] = p
    stop = np.asarray([last[recs[i][0]] for i in units.rec], np.int32)
    tid_of = {c: i for i, c in enumerate(cfg.contigs)}
    ounits = [(tid_of[recs[i][0]], recs[i][1], recs[i][2], recs[i][3][j], recs[i][3][0], last[recs[i][0]])
              for i, j in zip(units.rec, units.alt_index)]
    return batch, recs, units, stop, ounits


def _gpu(dev, batch, units, stop, mq, bq, amb=True):
    params = device.make_params(min_mapq=mq, min_baseq=bq, include_ambiguous=amb)
    dreads = dev.upload_reads(batch)
    buf = dev.pileup(dreads, dev.upload_units(units, stop), params)
    dev.sync()
    return dev.to_host(buf, device.COUNTS_DTYPE, units.n_units)


def _compare(got, want, amb=True):
    assert np.all(want["supported"] == 1)
    assert int((got["status"] & (device.ST_DEFERRED | device.ST_READLEN)).max()) == 0
    dp = want["dp"] if amb else want["dp"] - np.where((got["status"] & 3) == 0, want["n_amb"], 0)
    assert np.array_equal(got["dp"], dp)
    for f in ("n_amb", "ac_ref", "ac_alt"):
        assert np.array_equal(got[f], want[f]), f
    assert np.array_equal(got["med2"], want["med2"])
    both = (want["ac_ref"] > 0) & (want["ac_alt"] > 0)
    snv = (got["status"] & 3) == 0
    assert np.array_equal(got["s2"][both][:, [0, 2]], want["s2"][both][:, [0, 2]])
    assert np.array_equal(got["s2"][both & snv][:, 1], want["s2"][both & snv][:, 1])


CASES = {
    "wes": (lambda: synth.wes(n_tiles=3000), [(1, 30), (0, 0), (20, 13)]),
    "wgs_indels": (lambda: synth.wgs(contig_len=1_500_000, n_contigs=2, multiallelic_pct=5), [(1, 30), (0, 0)]),
    "amplicon_1000x": (lambda: synth.amplicon(n_tiles=24, depth_pairs=900), [(1, 30), (0, 20)]),       # tile_kernel (<= 2048 reads per tile)
    "amplicon_5000x": (lambda: synth.amplicon(n_tiles=14, depth_pairs=2500), [(1, 30)]),               # too deep for a tile: tail_kernel
    "multisample_60x": (lambda: synth.multisample(1, n_tiles=1500), [(1, 30), (0, 0)]),               # 65..192 candidates: medium_kernel
    "noisy_cigars": (lambda: synth.wes(n_tiles=1500, indel_period=40, snv_period=30, multiallelic_pct=20, p_softclip=0.2,
                                       p_del=0.15, p_ins=0.15, p_skip=0.05, p_supp=0.05, p_orphan=0.05, p_n=0.02,
                                       frag_mean=200.0, frag_sd=40.0), [(1, 30), (0, 0), (0, 41)]),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_gpu_matches_c_oracle(dev, name):
    make, thresholds = CASES[name]
    batch, recs, units, stop, ounits = _setup(make())
    assert units.n_units > 100
    for mq, bq in thresholds:
        want = c_oracle.pileup(batch.as_dict(), ounits, c_oracle.params(min_mapq=mq, min_baseq=bq))
        _compare(_gpu(dev, batch, units, stop, mq, bq), want)
    want = c_oracle.pileup(batch.as_dict(), ounits, c_oracle.params(min_mapq=1, min_baseq=30))
    got = _gpu(dev, batch, units, stop, 1, 30, amb=False)
    _compare(got, want, amb=False)
    assert want["ac_alt"].sum() > 0 and want["n_amb"].sum() >= 0
    # the case really takes the path it is named for (n_cand = candidate reads of the position index)
    n_cand = got["n_cand"]
    if name == "multisample_60x":
        assert np.mean((n_cand > 64) & (n_cand <= 192)) > 0.1
    elif name == "amplicon_1000x":
        assert np.mean((n_cand > 192) & (n_cand <= 2048)) > 0.5
    elif name == "amplicon_5000x":
        assert np.mean(n_cand > 2048) > 0.5


def test_deep_column_block_path(dev):
    """Columns deeper than the warp path's u16 bins (> 32767 candidate reads) take the CTA-per-unit kernel."""
    cfg = synth.amplicon(n_tiles=2, depth_pairs=30000, snv_period=60, indel_period=120)
    batch, recs, units, stop, ounits = _setup(cfg)
    got = _gpu(dev, batch, units, stop, 1, 30)
    assert int(got["n_cand"].max()) > 32767
    want = c_oracle.pileup(batch.as_dict(), ounits, c_oracle.params(min_mapq=1, min_baseq=30))
    _compare(got, want)


def test_properties_at_scale(dev):
    """A full-size-like batch (1.4 M reads): the result does not depend on the order or grouping of units,
    and a second launch reproduces it bit for bit."""
    cfg = synth.wes(n_tiles=20_000, indel_period=3000)
    batch = synth.reads(cfg)
    recs = synth.variant_records(cfg)
    units = build_units(recs, cfg.contigs)
    params = device.make_params(min_mapq=1, min_baseq=30)
    dreads = dev.upload_reads(batch)
    def run(u):
        buf = dev.pileup(dreads, dev.upload_units(u, None), params)
        dev.sync()
        return dev.to_host(buf, device.COUNTS_DTYPE, u.n_units)
    a, b = run(units), run(units)
    assert a.tobytes() == b.tobytes()
    rng = np.random.default_rng(7)
    perm = rng.permutation(units.n_units)
    from dataclasses import replace
    pu = replace(units, tid=units.tid[perm], pos=units.pos[perm], meta=units.meta[perm], length=units.length[perm],
                 seq_off=units.seq_off[perm], rec=units.rec[perm], alt_index=units.alt_index[perm])
    c = run(pu)
    assert c.tobytes() == a[perm].tobytes()
    half = units.n_units // 2
    first = replace(units, tid=units.tid[:half], pos=units.pos[:half], meta=units.meta[:half], length=units.length[:half],
                    seq_off=units.seq_off[:half], rec=units.rec[:half], alt_index=units.alt_index[:half])
    assert run(first).tobytes() == a[:half].tobytes()
    # depth sanity: dp <= reads covering the column <= candidates inspected
    assert np.all(a["dp"] <= a["n_cover"]) and np.all(a["n_cover"] <= a["n_cand"])
    assert np.all(a["ac_ref"] + a["ac_alt"] <= a["dp"] + (a["status"] & 3 != 0) * a["dp"])


def test_error_paths(dev):
    cfg = synth.wes(n_tiles=1100)
    batch = synth.reads(cfg)
    with pytest.raises(ValueError):
        build_units([("nope", 5, "A", ["T"])], cfg.contigs)
    # a batch that is not coordinate sorted: refused on the host by check_sorted, flagged by the kernels (VFK_ST_BADBATCH on
    # every record through the device-pointer call, ValueError from the host entry point)
    bad = synth.reads(cfg)
    bad.pos[:] = bad.pos[::-1].copy()
    with pytest.raises(ValueError):
        bad.check_sorted()
    units1 = build_units(synth.variant_records(cfg)[:50], cfg.contigs)
    params = device.make_params()
    buf = dev.pileup(dev.upload_reads(bad), dev.upload_units(units1, None), params)
    dev.sync()
    got = dev.to_host(buf, device.COUNTS_DTYPE, units1.n_units)
    assert np.all(got["status"] & device.ST_BADBATCH) and got["dp"].max() == 0
    with pytest.raises(ValueError):
        dev.annotate_host(device.host_reads(bad), units1, None, params, np.full(units1.n_units, 0.5), 5e-7, 1e-3)
    # the handle stays usable
    good = dev.annotate_host(device.host_reads(batch), units1, None, params, np.full(units1.n_units, 0.5), 5e-7, 1e-3)[0]
    assert int((good["status"] & device.ST_BADBATCH).max()) == 0 and good["dp"].max() > 0
    # empty unit list and empty read batch are fine
    units = build_units([], cfg.contigs)
    params = device.make_params()
    dreads = dev.upload_reads(batch)
    dev.pileup(dreads, dev.upload_units(units, None), params)
    dev.sync()


def test_interval_sharding_on_device():
    """Two handles (here on the same GPU) process interval shards of one BAM concurrently; the gathered
    records equal the single-launch result bit for bit (n_cand differs by construction: it counts the
    candidates inspected inside the shard's own read slice)."""
    from vafator_b200 import engine
    cfg = synth.wgs(contig_len=600_000, n_contigs=2, seed=5, indel_period=1500)
    batch, recs, units, stop, ounits = _setup(cfg)
    one = engine.Engine(0, min_mapq=1, min_baseq=30)
    two = engine.Engine(0, min_mapq=1, min_baseq=30, devices=[0, 0])
    a = one.pileup(batch, units, stop)
    b = two.pileup(batch, units, stop)
    for f in ("dp", "n_amb", "ac_ref", "ac_alt", "med2", "s2", "n_cover", "n_col", "status"):
        assert np.array_equal(a[f], b[f]), f
    one.close()
    two.close()


@pytest.mark.parametrize("read_len", [300, 600, 1500])
def test_longer_reads_use_wider_position_range(dev, read_len):
    """Read positions beyond 255 select the 512 / 1024 / 4096-bin instantiations (and 9 / 10 / 12-bit ranking)."""
    cfg = synth.wes(n_tiles=1100, read_len=read_len, frag_mean=2.2 * read_len, frag_sd=0.2 * read_len, tile_len=2 * read_len,
                    var_lo=read_len, var_hi=3 * read_len, snv_period=200, indel_period=1500, pairs_per_tile=40)
    batch, recs, units, stop, ounits = _setup(cfg)
    assert batch.l_qseq.max() == read_len and units.n_units > 200
    want = c_oracle.pileup(batch.as_dict(), ounits, c_oracle.params(min_mapq=1, min_baseq=30))
    got = _gpu(dev, batch, units, stop, 1, 30)
    _compare(got, want)
    assert int(want["med2"][:, 2, :].max()) > 2 * 255  # medians of read positions beyond one byte


@pytest.mark.parametrize("pairs", [160, 500, 4000])
def test_longer_reads_through_the_medium_tile_and_tail_paths(dev, pairs):
    """300-bp reads (the 512-bin instantiations) at depths that take the medium kernel (65..192 candidates), the tile kernel
    (<= 2048 reads per tile) and the tail kernel (deeper than a tile)."""
    cfg = synth.wes(n_tiles=40 if pairs > 1000 else 160, read_len=300, frag_mean=660.0, frag_sd=60.0, tile_len=600, var_lo=300, var_hi=900,
                    snv_period=60, indel_period=400, pairs_per_tile=pairs)
    batch, recs, units, stop, ounits = _setup(cfg)
    assert batch.l_qseq.max() == 300 and units.n_units > 100
    want = c_oracle.pileup(batch.as_dict(), ounits, c_oracle.params(min_mapq=1, min_baseq=30))
    got = _gpu(dev, batch, units, stop, 1, 30)
    _compare(got, want)
    lo, hi = {160: (64, 192), 500: (192, 2048), 4000: (2048, 1 << 30)}[pairs]
    assert np.mean((got["n_cand"] > lo) & (got["n_cand"] <= hi)) > 0.3, np.percentile(got["n_cand"], [5, 50, 95])


def test_degenerate_inputs(dev):
    """Empty read batch, contig without reads, column 0, every read filtered out."""
    from vafator_b200.batch import ReadBatch
    params = device.make_params(min_mapq=1, min_baseq=30)
    empty = ReadBatch.from_records([], ["c1", "c2"])
    units = build_units([("c1", 1, "A", ["T"]), ("c2", 500, "AT", ["A"])], empty.contigs)
    buf = dev.pileup(dev.upload_reads(empty), dev.upload_units(units, None), params)
    dev.sync()
    got = dev.to_host(buf, device.COUNTS_DTYPE, units.n_units)
    assert got["dp"].tolist() == [0, 0] and got["n_col"].tolist() == [0, 0] and got["med2"].max() == -1
    recs = [dict(qname="a", flag=1024, tid=0, pos=0, mapq=60, cigar="10M", seq="ACGTACGTAC", qual=[40] * 10),
            dict(qname="b", flag=0, tid=0, pos=0, mapq=0, cigar="10M", seq="ACGTACGTAC", qual=[40] * 10),
            dict(qname="c", flag=0, tid=0, pos=0, mapq=60, cigar="10M", seq="ACGTACGTAC", qual=[40] * 10)]
    b = ReadBatch.from_records(recs, ["c1", "c2"])
    units = build_units([("c1", 1, "A", ["T"]), ("c1", 10, "C", ["G"]), ("c1", 11, "A", ["T"]), ("c2", 1, "A", ["T"])], b.contigs)
    buf = dev.pileup(dev.upload_reads(b), dev.upload_units(units, None), params)
    dev.sync()
    got = dev.to_host(buf, device.COUNTS_DTYPE, units.n_units)
    assert got["dp"].tolist() == [1, 1, 0, 0] and got["n_cover"].tolist() == [3, 3, 0, 0] and got["ac_ref"].tolist() == [1, 1, 0, 0]


@pytest.mark.parametrize("mode", ["pageable", "pinned", "pinned_stage_all", "pinned_no_fetch", "pinned_short_list", "pinned_explicit_offsets"])
def test_host_entry_point_transfer_strategies(mode, monkeypatch):
    """vfk_annotate_host moves a batch in one of two ways -- page-locked arrays: the position-index inputs are copied, the
    rest is read in place over PCIe, except the base / quality bytes at the columns, which the device asks for on a fetch
    list and host threads gather (VFK_HOST_FETCH=0: read in place too; a list too short for the call: the units that do not
    fit read in place; a batch that hands over its pool offsets: requests by read index); pageable arrays (or
    VFK_HOST_STAGE_ALL=1 when the handle is created): everything is staged -- and returns exactly what the device-pointer
    entry points return, for VCF-ordered and for shuffled units."""
    from dataclasses import replace
    if mode == "pinned_stage_all":
        monkeypatch.setenv("VFK_HOST_STAGE_ALL", "1")
    if mode == "pinned_no_fetch":
        monkeypatch.setenv("VFK_HOST_FETCH", "0")
    if mode == "pinned_short_list":
        monkeypatch.setenv("VFK_HOST_FETCH_SLACK", "0.4")
    if mode in ("pinned", "pinned_short_list", "pinned_explicit_offsets"):
        monkeypatch.setenv("VFK_HOST_FETCH", "1")  # (the default asks for six host threads)
    in_place = mode in ("pinned", "pinned_no_fetch", "pinned_short_list", "pinned_explicit_offsets")
    dev = device.Device(0)
    cfg = synth.wes(n_tiles=3000, indel_period=700, multiallelic_pct=5)
    batch, recs, units, stop, _ = _setup(cfg)
    rng = np.random.default_rng(11)
    perm = rng.permutation(units.n_units)
    shuffled = replace(units, tid=units.tid[perm], pos=units.pos[perm], meta=units.meta[perm], length=units.length[perm],
                       seq_off=units.seq_off[perm], rec=units.rec[perm], alt_index=units.alt_index[perm])
    params = device.make_params(min_mapq=1, min_baseq=30)
    host = device.host_reads(batch, pinned=mode != "pageable", implicit_offsets=False if mode == "pinned_explicit_offsets" else None)
    dreads = dev.upload_reads(host)
    eaf = np.full(units.n_units, 0.4)
    for u, s in ((units, stop), (shuffled, stop[perm])):
        counts_dev = dev.pileup(dreads, dev.upload_units(u, s), params)
        stats_dev = dev.epilogue(u.n_units, counts_dev, dev.torch.as_tensor(eaf, device=dev.device), 5e-7, 1e-3)
        dev.sync()
        want_c = dev.to_host(counts_dev, device.COUNTS_DTYPE, u.n_units)
        want_s = dev.to_host(stats_dev, device.STATS_DTYPE, u.n_units)
        before = dev.transfer_stats()
        got_c, got_s = dev.annotate_host(host, u, s, params, eaf, 5e-7, 1e-3)
        after = dev.transfer_stats()
        assert got_c.tobytes() == want_c.tobytes()
        assert got_s.tobytes() == want_s.tobytes()
        assert (after["zero_copy_batches"] - before["zero_copy_batches"]) == (1 if in_place else 0)
        assert (after["staged_batches"] - before["staged_batches"]) == (0 if in_place else 1)
        pairs = after["fetch_pairs"] - before["fetch_pairs"]
        if in_place and mode != "pinned_no_fetch":
            n_cand = int(want_c["n_cand"].astype(np.int64)[want_c["n_cand"] <= 192].sum())  # (register and medium kernel units)
            assert after["fetch_batches"] - before["fetch_batches"] == 1
            # every candidate of every register-kernel unit is on the list -- or, with a list cut short, as many as it holds
            assert pairs == n_cand if mode != "pinned_short_list" else 0 < pairs < n_cand
        else:
            assert pairs == 0
        # two batches in flight through the double-buffered entry
        c2 = [np.empty(u.n_units, device.COUNTS_DTYPE) for _ in range(3)]
        s2 = [np.empty(u.n_units, device.STATS_DTYPE) for _ in range(3)]
        c2[1], s2[2] = device.pinned_empty(u.n_units, device.COUNTS_DTYPE), device.pinned_empty(u.n_units, device.STATS_DTYPE)  # written directly
        tickets = []
        for k in range(3):
            if k >= 2:
                dev.wait(tickets[k - 2])
            tickets.append(dev.annotate_host_async(host, u, s, params, eaf, 5e-7, 1e-3, c2[k], s2[k]))
        for t in tickets[1:]:
            dev.wait(t)
        for k in range(3):
            assert c2[k].tobytes() == want_c.tobytes() and s2[k].tobytes() == want_s.tobytes()
    dev.close()


def test_fetch_lists_cover_medium_columns(monkeypatch):
    """60x-like columns (65..192 candidate reads: medium kernel) through the host entry point with the arrays in place: their
    base / quality bytes come from the fetch list too, and the results equal the resident call byte for byte."""
    monkeypatch.setenv("VFK_HOST_FETCH", "1")
    dev = device.Device(0)
    cfg = synth.wes(n_tiles=600, indel_period=700, multiallelic_pct=5, pairs_per_tile=130)
    batch, recs, units, stop, _ = _setup(cfg)
    params = device.make_params(min_mapq=1, min_baseq=30)
    host = device.host_reads(batch, pinned=True)
    eaf = np.full(units.n_units, 0.5)
    counts_dev = dev.pileup(dev.upload_reads(host), dev.upload_units(units, stop), params)
    stats_dev = dev.epilogue(units.n_units, counts_dev, dev.torch.as_tensor(eaf, device=dev.device), 5e-7, 1e-3)
    dev.sync()
    want_c = dev.to_host(counts_dev, device.COUNTS_DTYPE, units.n_units)
    want_s = dev.to_host(stats_dev, device.STATS_DTYPE, units.n_units)
    assert np.mean((want_c["n_cand"] > 64) & (want_c["n_cand"] <= 192)) > 0.3, np.percentile(want_c["n_cand"], [5, 50, 95])
    before = dev.transfer_stats()
    got_c, got_s = dev.annotate_host(host, units, stop, params, eaf, 5e-7, 1e-3)
    after = dev.transfer_stats()
    assert after["zero_copy_batches"] - before["zero_copy_batches"] == 1 and after["fetch_batches"] - before["fetch_batches"] == 1
    assert after["fetch_pairs"] - before["fetch_pairs"] == int(want_c["n_cand"].astype(np.int64)[want_c["n_cand"] <= 192].sum())
    assert got_c.tobytes() == want_c.tobytes() and got_s.tobytes() == want_s.tobytes()
    dev.close()


def test_short_parity_campaign():
    """tests/fuzz_parity.py for twenty seconds: random read / CIGAR / pairing / variant mixes x thresholds."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, os.path.join(here, "fuzz_parity.py"), "--seconds", "20", "--seed0", "5000"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "fuzz ok" in r.stdout


@pytest.mark.parametrize("name", ["wes", "wgs_indels", "amplicon_1000x", "noisy_cigars"])
def test_register_pairing_matches_the_link_kernels(dev, name, monkeypatch):
    """The register kernel pairs the overlapping mates of a column in registers (MATCH.ANY over the read-name ids of its
    candidate reads); the list path replays htslib's name hash through the link kernels.  VFK_DEBUG_LINK_ALL=1 (read when a
    handle is created) sends every unit's pairing through the link kernels: both must return the same records."""
    make, thresholds = CASES[name]
    batch, recs, units, stop, ounits = _setup(make())
    monkeypatch.setenv("VFK_DEBUG_LINK_ALL", "1")
    dev2 = device.Device(0)
    monkeypatch.delenv("VFK_DEBUG_LINK_ALL")
    for mq, bq in thresholds[:2]:
        a = _gpu(dev, batch, units, stop, mq, bq)
        b = _gpu(dev2, batch, units, stop, mq, bq)
        both = (a["ac_ref"] > 0) & (a["ac_alt"] > 0)  # the rank sums exist only then
        for f in a.dtype.names:
            if f == "s2":
                assert np.array_equal(a[f][both], b[f][both]), f
            else:
                assert np.array_equal(a[f], b[f]), f
        _compare(b, c_oracle.pileup(batch.as_dict(), ounits, c_oracle.params(min_mapq=mq, min_baseq=bq)))
    dev2.close()


def test_listed_units_without_candidate_reads(dev):
    """A unit that lies where no read is (between two targets) must not disturb the units processed after it by the same
    warp, on the register path and on the list path alike (regression of round 1)."""
    cfg = synth.wes(n_tiles=1200, indel_period=60, snv_period=200)
    batch, recs, units, stop, ounits = _setup(cfg)
    contig = cfg.contigs[0]
    tile0 = min(p for c, p, r, a in recs if c == contig)
    # uncovered positions: far before the first target of the contig and in the gaps between targets
    gaps = [(contig, tile0 + 900 + 2048 * k, "A", ["ATT"]) for k in range(0, 400, 7)] + \
           [(contig, tile0 + 901 + 2048 * k, "ACG", ["A"]) for k in range(3, 400, 11)]
    recs2 = sorted(recs + gaps, key=lambda r: (cfg.contigs.index(r[0]), r[1]))
    units2 = build_units(recs2, cfg.contigs)
    last = {}
    for c, p, r, a in recs2:
        last[c] = p
    stop2 = np.asarray([last[recs2[i][0]] for i in units2.rec], np.int32)
    tid_of = {c: i for i, c in enumerate(cfg.contigs)}
    ou2 = [(tid_of[recs2[i][0]], recs2[i][1], recs2[i][2], recs2[i][3][j], recs2[i][3][0], last[recs2[i][0]])
           for i, j in zip(units2.rec, units2.alt_index)]
    want = c_oracle.pileup(batch.as_dict(), ou2, c_oracle.params(min_mapq=1, min_baseq=30))
    for rep in range(3):
        got = _gpu(dev, batch, units2, stop2, 1, 30)
        _compare(got, want)
    empty = got["n_cover"] == 0
    assert int(empty.sum()) >= 50 and int((got["status"] & 3 != 0).sum()) > 100


def test_three_alignments_of_one_name_take_the_general_protocol(dev):
    """A pair plus a supplementary record of the same name inside one column's window: htslib's hash pairs the first two
    arrivals and leaves the third alone.  The register kernel hands such a column to the list path (link kernels)."""
    from vafator_b200.batch import ReadBatch
    q = lambda n, v=30: "".join(chr(v + 33) for _ in range(n))
    R = []
    for k in range(40):  # filler pairs so that the column is shallow but not trivial
        R.append(dict(qname="f%d" % k, flag=99, tid=0, pos=100 + k % 7, mapq=60, cigar="30M", seq="ACGTACGTACGTACGTACGTACGTACGTAC", qual=q(30),
                      mtid=0, mpos=115 + k % 7, isize=45))
        R.append(dict(qname="f%d" % k, flag=147, tid=0, pos=115 + k % 7, mapq=60, cigar="30M", seq="ACGTACGTACGTACGTACGTACGTACGTAC", qual=q(30),
                      mtid=0, mpos=100 + k % 7, isize=-45))
    R.append(dict(qname="tri", flag=99, tid=0, pos=108, mapq=60, cigar="30M", seq="CACGTACGTACGCACGTACGTACGCACGTA", qual=q(30), mtid=0, mpos=112, isize=34))
    R.append(dict(qname="tri", flag=2048 | 99, tid=0, pos=110, mapq=60, cigar="4S8M18S", seq="GGGGCGTACGTAGGGGGGGGGGGGGGGGGG", qual=q(30), mtid=0, mpos=112, isize=34))
    R.append(dict(qname="tri", flag=147, tid=0, pos=112, mapq=60, cigar="30M", seq="TACGTACGGGGGTACGTACGGGGGTACGTA", qual=q(30), mtid=0, mpos=108, isize=-34))
    R.sort(key=lambda d: d["pos"])
    batch = ReadBatch.from_records(R, ["c"])
    recs = [("c", p, "A", ["C"]) for p in range(109, 140)]
    units = build_units(recs, ["c"])
    stop = np.full(units.n_units, 139, np.int32)
    ou = [(0, r[1], r[2], r[3][0], r[3][0], 139) for r in recs]
    for mq, bq in ((0, 0), (1, 30), (0, 41)):
        want = c_oracle.pileup(batch.as_dict(), ou, c_oracle.params(min_mapq=mq, min_baseq=bq))
        _compare(_gpu(dev, batch, units, stop, mq, bq), want)


def _one_event_word(ops, l_qseq):
    """include/vfk.h vfk_prep_words: [H*][S] M [ I|D|N M ] [S][H*] -> lead | m1 << 12 | event << 24 | len << 26, else None."""
    import re
    m = re.fullmatch(r"(?:\d+H)*(?:(\d+)S)?(\d+)[M=X](?:(\d+)([IDN])(\d+)[M=X])?(?:\d+S)?(?:\d+H)*", "".join("%d%s" % (n, o) for n, o in ops))
    if not m:
        return None
    lead, m1 = int(m.group(1) or 0), int(m.group(2))
    ev, ln = ("IDN".index(m.group(4)) + 1, int(m.group(3))) if m.group(4) else (0, 0)
    if m1 == 0 or m1 >= 4096 or lead >= 4096 or ln >= 64 or l_qseq > 4095 or (m.group(4) and (ln == 0 or int(m.group(5)) == 0)):
        return None
    return lead | (m1 << 12) | (ev << 24) | (ln << 26)


def test_read_preparation_words(dev):
    """prep_kernel against an independent reading of every CIGAR: reference end (htslib bam_endpos) and the shape word, on
    synthetic reads with heavy CIGAR noise and on handmade shapes (hard clips, pads, consecutive events, zero lengths)."""
    import random
    from vafator_b200.batch import ReadBatch, CIGAR_CHARS
    cfg = synth.wes(n_tiles=1500, p_softclip=0.3, p_del=0.2, p_ins=0.2, p_skip=0.1)
    batch = synth.reads(cfg)
    rng = random.Random(5)
    recs = []
    shapes = ["20M", "5S15M", "3H5S12M3S2H", "10M2D10M", "10M2I8M", "4S6M100N10M", "20S", "10M2D", "5M0D15M", "5M1P1I14M", "5M2I2I11M",
              "2H18M2H", "5M3D5M2D10M", "10M70D10M", "10M63I10M", "18M2S", "2S3S15M", "5H5H20M", "10M2D0M", "1M", "20="]
    for k in range(400):
        if k < len(shapes):
            cg = shapes[k]
        else:
            parts = []
            for _ in range(rng.randint(1, 6)):
                parts.append("%d%s" % (rng.choice([0, 1, 2, 5, 17, 63, 64, 120]), rng.choice("MIDNSHP=X")))
            cg = "".join(parts)
        import re
        ops = [(int(n), o) for n, o in re.findall(r"(\d+)([MIDNSHP=X])", cg)]
        lq = sum(n for n, o in ops if o in "MIS=X")
        recs.append(dict(qname="r%d" % k, flag=0, tid=0, pos=1000 + k, mapq=60, cigar=cg, seq="A" * lq, qual=[30] * lq))
    hand = ReadBatch.from_records(recs, ["c"])
    for b in (batch, hand):
        end, shape, status = dev.prep_words(dev.upload_reads(b))
        assert status == 0
        n = b.n_placed
        for i in range(n):
            ops = [(int(w) >> 4, CIGAR_CHARS[int(w) & 15]) for w in b.cigar[int(b.cigar_off[i]):int(b.cigar_off[i]) + int(b.n_cigar[i])]]
            span = sum(l for l, o in ops if o in "MDN=X")
            assert end[i] == b.pos[i] + span, (i, ops)
            lq = int(b.l_qseq[i])
            simple = len(ops) == 1 and ops[0][1] in "M=X" and ops[0][0] == lq
            word = 0 if simple else _one_event_word(ops, lq)
            assert shape[i] == (0xFFFFFFFF if word is None else word), (i, ops, hex(int(shape[i])), word)


def test_implicit_pool_offsets(dev):
    """A batch with contiguous pools may leave cigar_off / seq_off / qual_off NULL (include/vfk.h): the device derives them.  Same
    records out as with the explicit arrays, through the host entry (pageable and page-locked) and the device-pointer entry."""
    cfg = synth.wgs(contig_len=700_000, n_contigs=2, multiallelic_pct=5)
    batch, recs, units, stop, ounits = _setup(cfg)
    params = device.make_params(min_mapq=1, min_baseq=30)
    eaf = np.full(units.n_units, 0.4)
    ref = dev.annotate_host(device.host_reads(batch, implicit_offsets=False), units, stop, params, eaf, 5e-7, 1e-3)
    for pinned in (False, True):
        got = dev.annotate_host(device.host_reads(batch, pinned=pinned, implicit_offsets=True), units, stop, params, eaf, 5e-7, 1e-3)
        assert np.array_equal(got[0].view(np.uint8), ref[0].view(np.uint8))
        assert np.array_equal(got[1].view(np.uint8), ref[1].view(np.uint8), equal_nan=False) or np.array_equal(got[1]["k"], ref[1]["k"])
    # device-pointer entry: explicit offsets resident vs NULL
    dreads = dev.upload_reads(device.host_reads(batch, implicit_offsets=False))
    want = dev.to_host(dev.pileup(dreads, dev.upload_units(units, stop), params), device.COUNTS_DTYPE, units.n_units)
    dreads.c.cigar_off = dreads.c.seq_off = dreads.c.qual_off = None
    got = dev.to_host(dev.pileup(dreads, dev.upload_units(units, stop), params), device.COUNTS_DTYPE, units.n_units)
    assert np.array_equal(got.view(np.uint8), want.view(np.uint8))
    # one of the three missing: refused
    dreads.c.cigar_off = dreads.t["cigar_off"].data_ptr()
    with pytest.raises(ValueError):
        dev.pileup(dreads, dev.upload_units(units, stop), params)
