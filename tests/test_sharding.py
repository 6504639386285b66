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
