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


def _gpu(dev, batch, units, stop, mq, bq, amb=True, columns=False):
    params = device.make_params(min_mapq=mq, min_baseq=bq, include_ambiguous=amb)
    dreads = dev.upload_reads(device.pack_reads(batch, params, columns=columns))
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
    "amplicon_1000x": (lambda: synth.amplicon(n_tiles=24, depth_pairs=900), [(1, 30), (0, 20)]),
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
    _compare(_gpu(dev, batch, units, stop, 1, 30, amb=False), want, amb=False)
    assert want["ac_alt"].sum() > 0 and want["n_amb"].sum() >= 0


@pytest.mark.parametrize("columns", [False, True], ids=["read_major", "columns"])
def test_deep_column_block_path(dev, columns):
    """Columns deeper than the warp path's u16 bins (> 32767 candidate reads) take the CTA-per-unit kernel -- also when
    the units come off the column kernel's list and are located afterwards."""
    cfg = synth.amplicon(n_tiles=2, depth_pairs=30000, snv_period=60, indel_period=120)
    batch, recs, units, stop, ounits = _setup(cfg)
    got = _gpu(dev, batch, units, stop, 1, 30, columns=columns)
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
    dreads = dev.upload_reads(device.pack_reads(batch, params))
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
    bad = synth.reads(cfg)
    bad.pos = bad.pos[::-1].copy()
    with pytest.raises(ValueError):
        device.pack_reads(bad, device.make_params())
    # empty unit list and empty read batch are fine
    units = build_units([], cfg.contigs)
    params = device.make_params()
    dreads = dev.upload_reads(device.pack_reads(batch, params))
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


def test_degenerate_inputs(dev):
    """Empty read batch, contig without reads, column 0, every read filtered out."""
    from vafator_b200.batch import ReadBatch
    params = device.make_params(min_mapq=1, min_baseq=30)
    empty = ReadBatch.from_records([], ["c1", "c2"])
    units = build_units([("c1", 1, "A", ["T"]), ("c2", 500, "AT", ["A"])], empty.contigs)
    buf = dev.pileup(dev.upload_reads(device.pack_reads(empty, params)), dev.upload_units(units, None), params)
    dev.sync()
    got = dev.to_host(buf, device.COUNTS_DTYPE, units.n_units)
    assert got["dp"].tolist() == [0, 0] and got["n_col"].tolist() == [0, 0] and got["med2"].max() == -1
    recs = [dict(qname="a", flag=1024, tid=0, pos=0, mapq=60, cigar="10M", seq="ACGTACGTAC", qual=[40] * 10),
            dict(qname="b", flag=0, tid=0, pos=0, mapq=0, cigar="10M", seq="ACGTACGTAC", qual=[40] * 10),
            dict(qname="c", flag=0, tid=0, pos=0, mapq=60, cigar="10M", seq="ACGTACGTAC", qual=[40] * 10)]
    b = ReadBatch.from_records(recs, ["c1", "c2"])
    units = build_units([("c1", 1, "A", ["T"]), ("c1", 10, "C", ["G"]), ("c1", 11, "A", ["T"]), ("c2", 1, "A", ["T"])], b.contigs)
    buf = dev.pileup(dev.upload_reads(device.pack_reads(b, params)), dev.upload_units(units, None), params)
    dev.sync()
    got = dev.to_host(buf, device.COUNTS_DTYPE, units.n_units)
    assert got["dp"].tolist() == [1, 1, 0, 0] and got["n_cover"].tolist() == [3, 3, 0, 0] and got["ac_ref"].tolist() == [1, 1, 0, 0]


@pytest.mark.parametrize("pinned", [False, True], ids=["pageable", "pinned"])
def test_host_entry_point_transfer_strategies(dev, pinned, monkeypatch):
    """vfk_annotate_host moves a batch in one of several ways (units located on the host or by the locate
    kernel; headers / payload staged or gathered from page-locked memory): every combination returns
    exactly what the device-pointer entry points return, for VCF-ordered and for shuffled units."""
    from dataclasses import replace
    cfg = synth.wes(n_tiles=3000, indel_period=700, multiallelic_pct=5)
    batch, recs, units, stop, _ = _setup(cfg)
    rng = np.random.default_rng(11)
    perm = rng.permutation(units.n_units)
    shuffled = replace(units, tid=units.tid[perm], pos=units.pos[perm], meta=units.meta[perm], length=units.length[perm],
                       seq_off=units.seq_off[perm], rec=units.rec[perm], alt_index=units.alt_index[perm])
    params = device.make_params(min_mapq=1, min_baseq=30)
    packed = device.pack_reads(batch, params, pinned=pinned)
    dreads = dev.upload_reads(packed)
    eaf = np.full(units.n_units, 0.4)
    for u, s in ((units, stop), (shuffled, stop[perm])):
        counts_dev = dev.pileup(dreads, dev.upload_units(u, s), params)
        stats_dev = dev.epilogue(u.n_units, counts_dev, dev.torch.as_tensor(eaf, device=dev.device), 5e-7, 1e-3)
        dev.sync()
        want_c = dev.to_host(counts_dev, device.COUNTS_DTYPE, u.n_units)
        want_s = dev.to_host(stats_dev, device.STATS_DTYPE, u.n_units)
        modes = [dict(VFK_HOST_LOCATE=a, VFK_ZERO_COPY_PAYLOAD=b, VFK_ZERO_COPY_HOT=c)
                 for a in "01" for b in "01" for c in "01"] + [dict()]
        for env in modes:
            for k in ("VFK_HOST_LOCATE", "VFK_ZERO_COPY_PAYLOAD", "VFK_ZERO_COPY_HOT"):
                monkeypatch.delenv(k, raising=False)
            for k, v in env.items():
                monkeypatch.setenv(k, v)
            before = dev.zero_copy_batches()
            got_c, got_s = dev.annotate_host(packed, u, s, params, eaf, 5e-7, 1e-3)
            assert got_c.tobytes() == want_c.tobytes(), env
            assert got_s.tobytes() == want_s.tobytes(), env
            if env.get("VFK_ZERO_COPY_PAYLOAD") == "1":
                assert (dev.zero_copy_batches() > before) == pinned


@pytest.mark.parametrize("layout", ["read_major", "columns"])
def test_short_parity_campaign(layout):
    """tests/fuzz_parity.py for ten seconds per layout: random read / CIGAR / pairing / variant mixes x thresholds."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, os.path.join(here, "fuzz_parity.py"), "--seconds", "10", "--seed0", "5000"]
    r = subprocess.run(cmd + (["--columns"] if layout == "columns" else []), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "fuzz ok" in r.stdout


@pytest.mark.parametrize("name", ["wes", "wgs_indels", "amplicon_1000x", "noisy_cigars"])
def test_column_store_kernel_matches_read_major_kernels(dev, name):
    """With a column store in the batch SNV units are evaluated from contiguous sector runs (pileup_column_kernel);
    every field but the diagnostic n_cand equals what the read-major kernels return, and both equal the oracle."""
    make, thresholds = CASES[name]
    batch, recs, units, stop, ounits = _setup(make())
    for mq, bq in thresholds[:2]:
        params = device.make_params(min_mapq=mq, min_baseq=bq)
        got = {}
        for columns in (False, True):
            dreads = dev.upload_reads(device.pack_reads(batch, params, columns=columns))
            buf = dev.pileup(dreads, dev.upload_units(units, stop), params)
            dev.sync()
            got[columns] = dev.to_host(buf, device.COUNTS_DTYPE, units.n_units)
        both = (got[False]["ac_ref"] > 0) & (got[False]["ac_alt"] > 0)  # the rank sums exist only then
        for f in got[True].dtype.names:
            if f == "s2":
                assert np.array_equal(got[True][f][both], got[False][f][both]), f
            elif f != "n_cand":
                assert np.array_equal(got[True][f], got[False][f]), f
        _compare(got[True], c_oracle.pileup(batch.as_dict(), ounits, c_oracle.params(min_mapq=mq, min_baseq=bq)))


def test_host_entry_point_reads_the_column_store_in_place(dev, monkeypatch):
    """Page-locked batch with a column store: vfk_annotate_host gathers the sector runs straight from host memory and
    returns exactly what the device-resident call returns; VFK_HOST_COLUMNS=0 falls back to the read-major arrays."""
    for k in ("VFK_HOST_LOCATE", "VFK_ZERO_COPY_PAYLOAD", "VFK_ZERO_COPY_HOT", "VFK_HOST_COLUMNS"):
        monkeypatch.delenv(k, raising=False)
    cfg = synth.wes(n_tiles=3000, indel_period=700, multiallelic_pct=5)
    batch, recs, units, stop, _ = _setup(cfg)
    params = device.make_params(min_mapq=1, min_baseq=30)
    packed = device.pack_reads(batch, params, pinned=True, columns=True)
    dreads = dev.upload_reads(packed)
    eaf = np.full(units.n_units, 0.4)
    counts_dev = dev.pileup(dreads, dev.upload_units(units, stop), params)
    stats_dev = dev.epilogue(units.n_units, counts_dev, dev.torch.as_tensor(eaf, device=dev.device), 5e-7, 1e-3)
    dev.sync()
    want_c = dev.to_host(counts_dev, device.COUNTS_DTYPE, units.n_units)
    want_s = dev.to_host(stats_dev, device.STATS_DTYPE, units.n_units)
    before = dev.column_batches()
    got_c, got_s = dev.annotate_host(packed, units, stop, params, eaf, 5e-7, 1e-3)
    assert dev.column_batches() == before + 1
    assert got_c.tobytes() == want_c.tobytes() and got_s.tobytes() == want_s.tobytes()
    monkeypatch.setenv("VFK_HOST_COLUMNS", "0")
    alt_c, alt_s = dev.annotate_host(packed, units, stop, params, eaf, 5e-7, 1e-3)
    assert dev.column_batches() == before + 1
    both = (want_c["ac_ref"] > 0) & (want_c["ac_alt"] > 0)
    for f in alt_c.dtype.names:
        if f == "s2":
            assert np.array_equal(alt_c[f][both], want_c[f][both]), f
        elif f != "n_cand":
            assert np.array_equal(alt_c[f], want_c[f]), f
    assert alt_s.tobytes() == want_s.tobytes()


def test_listed_units_without_candidate_reads(dev):
    """Indel units always reach the histogram kernel when the batch has a column store; one that lies where no read
    is (between two targets) must not disturb the units processed after it by the same warp (regression: it used to
    arm the TMA staging ring without consuming it)."""
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
    params = device.make_params(min_mapq=1, min_baseq=30)
    got = {}
    for columns in (False, True):
        for rep in range(3):  # the staging ring state carries over between launches of a persistent warp only within one
            dreads = dev.upload_reads(device.pack_reads(batch, params, columns=columns))
            buf = dev.pileup(dreads, dev.upload_units(units2, stop2), params)
            dev.sync()
            got[columns] = dev.to_host(buf, device.COUNTS_DTYPE, units2.n_units)
    empty = got[False]["n_cover"] == 0
    assert int(empty.sum()) >= 50 and int((got[False]["status"] & 3 != 0).sum()) > 100
    both = (got[False]["ac_ref"] > 0) & (got[False]["ac_alt"] > 0)
    for f in got[True].dtype.names:
        if f == "s2":
            assert np.array_equal(got[True][f][both], got[False][f][both])
        elif f != "n_cand":
            assert np.array_equal(got[True][f], got[False][f]), f


STAGED_FLAGS = ("VFK_COL_TRANSPOSED", "VFK_LIST_SHALLOW", "VFK_MQ_SELECT")


@pytest.mark.skipif(not __import__("os").environ.get("VFK_TEST_STAGED"), reason="staged kernel variants: not yet run on a GPU "
                    "(set VFK_TEST_STAGED=1; DESIGN.md section 6)")
@pytest.mark.parametrize("name", ["wes", "wgs_indels", "amplicon_1000x", "noisy_cigars"])
@pytest.mark.parametrize("combo", [c for r in (1, 2, 3) for c in __import__("itertools").combinations(STAGED_FLAGS, r)],
                         ids=lambda c: "+".join(x[4:] for x in c))
def test_staged_variants_match_the_oracle(dev, monkeypatch, name, combo):
    """Every combination of the staged switches (read per call by libvfk.so and by pack_reads) on the column-store
    path: every field of every unit equals the C oracle."""
    make, thresholds = CASES[name]
    batch, recs, units, stop, ounits = _setup(make())
    for f in STAGED_FLAGS:
        monkeypatch.setenv(f, "1" if f in combo else "0")
    for mq, bq in thresholds[:2]:
        got = _gpu(dev, batch, units, stop, mq, bq, columns=True)
        _compare(got, c_oracle.pileup(batch.as_dict(), ounits, c_oracle.params(min_mapq=mq, min_baseq=bq)))


@pytest.mark.skipif(not __import__("os").environ.get("VFK_TEST_STAGED"), reason="staged host gather: not yet run on a GPU "
                    "(set VFK_TEST_STAGED=1; DESIGN.md section 6)")
@pytest.mark.parametrize("transposed", ["0", "1"], ids=["sector_form", "transposed_form"])
def test_staged_host_gather_matches_the_resident_call(dev, monkeypatch, transposed):
    """VFK_HOST_GATHER=1: the host entry point gathers every unit's column run (10 bytes per read) into one bulk copy;
    results equal the device-resident call on the same batch, field for field (the diagnostic n_cand aside)."""
    for k in ("VFK_HOST_LOCATE", "VFK_ZERO_COPY_PAYLOAD", "VFK_ZERO_COPY_HOT", "VFK_HOST_COLUMNS", "VFK_LIST_SHALLOW", "VFK_MQ_SELECT"):
        monkeypatch.delenv(k, raising=False)
    monkeypatch.setenv("VFK_COL_TRANSPOSED", transposed)
    cfg = synth.wes(n_tiles=3000, indel_period=700, multiallelic_pct=5)
    batch, recs, units, stop, ounits = _setup(cfg)
    params = device.make_params(min_mapq=1, min_baseq=30)
    packed = device.pack_reads(batch, params, pinned=True, columns=True)
    eaf = np.full(units.n_units, 0.4)
    monkeypatch.setenv("VFK_HOST_GATHER", "0")
    want_c, want_s = dev.annotate_host(packed, units, stop, params, eaf, 5e-7, 1e-3)
    monkeypatch.setenv("VFK_HOST_GATHER", "1")
    got_c, got_s = dev.annotate_host(packed, units, stop, params, eaf, 5e-7, 1e-3)
    for f in got_c.dtype.names:
        if f != "n_cand":
            assert np.array_equal(got_c[f], want_c[f]), f
    assert got_s.tobytes() == want_s.tobytes()
    _compare(got_c, c_oracle.pileup(batch.as_dict(), ounits, c_oracle.params(min_mapq=1, min_baseq=30)))
