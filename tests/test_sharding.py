"""N > 1 host logic on CPU (gloo, world size 2): block partition of particle ids and the all-reduce of the
observable series.  The per-shard partial sums come from the oracle (test infrastructure) stepping each shard
with the Philox streams keyed by GLOBAL particle id -- the same contract the GPU kernels implement
(tests/test_bulk_gpu.py::test_sharding_invariance_and_determinism is the GPU counterpart)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pyoracle as po
from scenarios import build_si
from viennaemc_b200 import sharding


def test_shard_range_partitions_every_id_exactly_once():
    for n, world in ((10, 3), (100000000, 8), (7, 8), (0, 2), (12500, 1)):
        seen = []
        for r in range(world):
            a, b = sharding.shard_range(n, r, world)
            assert 0 <= a <= b <= n
            seen += [(a, b)]
        assert seen[0][0] == 0 and seen[-1][1] == n
        assert all(seen[i][1] == seen[i + 1][0] for i in range(world - 1))
        sizes = [b - a for a, b in seen]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


STEPS, DT, SEED = 40, 5e-16, 77
BOX = [2e-7] * 3


def _ensemble():
    m = build_si()
    ens, _ = m.generate_initial(BOX, [2, 2, 2], 1e23, po.mt_state(3))
    return m, ens


def _run_shard(m, ens, first, last):
    sub = ens.copy()
    for f in po.Ensemble.F64 + po.Ensemble.I32:
        setattr(sub, f, getattr(ens, f)[first:last].copy())
    sub.n = last - first
    res = m.bulk_steps(sub, BOX, [-1, 0, 0], 1e6, DT, STEPS, po.rng_philox(SEED, first), first_step=1)
    return res["obs"]


def _worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m, ens = _ensemble()
    first, last = sharding.shard_range(ens.n, rank, world)
    part = torch.from_numpy(np.ascontiguousarray(_run_shard(m, ens, first, last)))
    sharding.allreduce_observables(part)
    if rank == 0:
        np.save(out_path, part.numpy())
    dist.destroy_process_group()


def test_two_rank_allreduce_of_shard_observables_equals_the_single_process_run(tmp_path):
    out = str(tmp_path / "reduced.npy")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    reduced = np.load(out)
    m, ens = _ensemble()
    whole = _run_shard(m, ens, 0, ens.n)
    assert np.array_equal(reduced[:, :, 2], whole[:, :, 2])  # counts: exact
    assert np.allclose(reduced, whole, rtol=1e-12, atol=0)   # sums: order of summation differs
    e, v, occ = sharding.finalize_observables(reduced, ens.n)
    assert np.all(occ == 1.0) and np.all(e > 0)


# ---- device runs: how the ranks split the contact handling -----------------------------------------------------------
def test_inject_shares_add_up_and_rotate():
    for world in (1, 2, 3, 8):
        for missing in range(0, 20):
            for cell in (0, 5, 41):
                for step in (1, 2, 77):
                    shares = [sharding.inject_share_of_rank(missing, r, world, cell, step) for r in range(world)]
                    assert sum(shares) == missing and max(shares) - min(shares) <= 1
    # the remainder moves from rank to rank with the step: no rank collects the injected particles
    firsts = {max(range(4), key=lambda r: sharding.inject_share_of_rank(1, r, 4, 3, s)) for s in range(8)}
    assert firsts == {0, 1, 2, 3}


def _reservoir_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)  # the same on every rank: the "global" ensemble
    cells, nr_carriers = 12, 1.0
    expected = np.array([0, 2.5, 5, 5, 5, 2.5, 0, 0, 3, 1, 4.5, 0.25])
    cell_of = rng.integers(0, cells, size=200)
    a, b = sharding.shard_range(len(cell_of), rank, world)
    mine = np.bincount(cell_of[a:b], minlength=cells)
    share = torch.zeros((world, cells), dtype=torch.float64)
    share[rank] = torch.from_numpy(mine.astype(np.float64))
    dist.all_reduce(share)  # what emcgpu_device_set_sharding's callback does with the share table
    kept, dropped, injected = sharding.reservoir_decisions(share.numpy(), expected, nr_carriers, rank, step=7)
    t = torch.from_numpy(np.stack([kept, dropped, injected]).astype(np.float64))
    dist.all_reduce(t)
    out[rank] = (t.numpy(), np.bincount(cell_of, minlength=cells), expected)
    dist.destroy_process_group()


def test_sharded_contact_handling_equals_the_single_rank_decision():
    """world 2 on gloo: kept / deleted / injected per reservoir cell, summed over the ranks, are what one rank holding the
    whole ensemble decides (the reference's rule); the first particles in GLOBAL index order are the ones kept"""
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_reservoir_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    total, counts, expected = out[0]
    assert np.array_equal(out[1][0], total)
    single = sharding.reservoir_decisions(counts[None, :], expected, 1.0, 0, step=7)
    for got, want in zip(total, single):
        assert np.array_equal(got, want)
    slots = np.where(expected > 0, np.ceil(expected), 0)
    assert np.array_equal(total[0], np.minimum(counts, slots)) and np.array_equal(total[0] + total[2] >= np.floor(expected), np.ones(12, bool))
