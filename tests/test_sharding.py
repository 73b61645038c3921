"""Interval x BAM sharding on CPU: world_size-2 gloo run (two processes) + single-process checks.  The
per-shard compute is the oracle here (no GPU in this container); what is under test is the product's
planning / slicing / gathering logic: sharded result == unsharded result, bit for bit."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers as H
from oracle import c_oracle
from vafator_b200 import device, sharding, synth
from vafator_b200.batch import build_units

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _case():
    cfg = synth.wgs(contig_len=400_000, n_contigs=2, seed=77, indel_period=900, snv_period=300, p_del=0.1, p_skip=0.02)
    batch = synth.reads(cfg)
    recs = synth.variant_records(cfg)
    units = build_units(recs, cfg.contigs)
    last = {}
    for c, p, r, a in recs:
        last[c] = p
    stop = np.asarray([last[recs[i][0]] for i in units.rec], np.int32)
    return cfg, batch, recs, units, stop


def _oracle_compute(recs, contigs):
    def compute(sub, sub_units, sub_stop):
        # every shard is a self-contained canonical batch: the mates are paired among the reads of the slice, as the kernels do
        ou = [(int(t), int(p) + 1, recs[i][2], recs[i][3][j], recs[i][3][0], int(s))
              for t, p, i, j, s in zip(sub_units.tid, sub_units.pos, sub_units.rec, sub_units.alt_index, sub_stop)]
        return c_oracle.pileup(sub.as_dict(), ou, c_oracle.params(min_mapq=1, min_baseq=30))
    return compute


def test_sharded_equals_unsharded_single_process():
    cfg, batch, recs, units, stop = _case()
    prm = device.make_params(min_mapq=1, min_baseq=30)
    passes = device.read_passes(batch, prm)
    compute = _oracle_compute(recs, cfg.contigs)
    whole = compute(batch, units, stop)
    for n_shards in (1, 2, 7, 64):
        got = sharding.run_sharded(batch, units, stop, passes, n_shards, 0, 1, compute, lambda x: [x])
        assert got.tobytes() == whole.tobytes(), n_shards
    shards = sharding.plan_shards(batch, units, 8, passes, stop)
    assert shards[0].unit_lo == 0 and shards[-1].unit_hi == units.n_units
    assert all(a.unit_hi == b.unit_lo for a, b in zip(shards, shards[1:]))
    assert sum(s.read_hi - s.read_lo for s in shards) < 1.5 * batch.n_reads  # little duplication at the cuts


def _worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg, batch, recs, units, stop = _case()
    prm = device.make_params(min_mapq=1, min_baseq=30)
    passes = device.read_passes(batch, prm)

    def gather(obj):
        box = [None] * world if rank == 0 else None
        dist.gather_object(obj, box, dst=0)
        return box

    got = sharding.run_sharded(batch, units, stop, passes, 6, rank, world, _oracle_compute(recs, cfg.contigs), gather)
    if rank == 0:
        whole = _oracle_compute(recs, cfg.contigs)(batch, units, stop)
        np.save(out_path, np.asarray([got.tobytes() == whole.tobytes(), len(got)], np.int64))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_two_ranks_gloo(tmp_path):
    out = str(tmp_path / "ok.npy")
    mp.spawn(_worker, args=(2, 29533, out), nprocs=2, join=True)
    ok, n = np.load(out)
    assert ok == 1 and n > 500
