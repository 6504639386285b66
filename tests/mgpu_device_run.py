"""Sharded device run on N GPUs (torchrun, one rank per GPU): SURVEY.md 8e.  The 'device_bar' scenario (tests/golden) is
split in blocks of particles over the ranks; every rank runs emcgpu_device_run with the all-reduce callback of
viennaemc_b200.sharding.DeviceRunSharding.  Checks (rank 0 prints one JSON line, exit code 1 on failure):
  * potential and concentration are bitwise identical on all ranks (replicated Poisson solve on identical inputs),
  * the carriers-per-grid-point grid is the sum over the ranks and adds up to the global ensemble size,
  * every reservoir cell holds its expected population GLOBALLY after each step's contact handling,
  * particle bookkeeping: sum of ensemble sizes = initial - left through contacts + injected - deleted,
  * the run agrees with the same run on ONE rank within the ensemble statistics (ensemble size, mean energy).
    torchrun --nproc-per-node 2 tests/mgpu_device_run.py [--steps 400]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

from helpers import download_ensemble, load_golden, upload_ensemble, upload_model  # noqa: E402
from oracle import pyoracle as po  # noqa: E402  (test infrastructure: scenario + initial ensemble)
from scenarios import DEVICE_CASES, build_device  # noqa: E402
from test_device_gpu import configure  # noqa: E402
from test_oracle_device import ens_from  # noqa: E402
from viennaemc_b200 import capi, sharding  # noqa: E402


def run(ctx_rank, world, group, steps, case="device_bar", replicate=8):
    g = load_golden(case)
    a = DEVICE_CASES[case]
    m, dev = build_device(case)
    ens = ens_from(g, "init_")
    first, last = sharding.shard_range(ens.n, ctx_rank, world)
    sub = po.Ensemble(last - first)
    for f in po.Ensemble.F64 + po.Ensemble.I32:
        getattr(sub, f)[:] = getattr(ens, f)[first:last]
    sub.n = last - first
    ctx = capi.Context(torch.cuda.current_device())
    upload_model(ctx, m)
    configure(ctx, dev, math_mode=capi.MATH_FAST)
    ctx.device_set_grid(capi.GRID_POTENTIAL, g["pot_eq"])
    ctx.device_set_grid(capi.GRID_CONCENTRATION, g["conc_eq"])
    upload_ensemble(ctx, sub, particle_id_base=ctx_rank << 40)
    ctx.device_reserve(2 * sub.n + 4096)
    ctx.rng_philox(4242)
    ctx.set_step_index(1)
    shard = sharding.DeviceRunSharding(ctx, group) if world > 1 else None
    counters, sweeps = ctx.device_run(a["dt"], steps, 1e-4, 1.8, True)
    e = download_ensemble(ctx)
    return dict(ctx=ctx, shard=shard, counters=counters, sweeps=sweeps, ens=e, n0=ens.n, dev=dev, g=g)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=400)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    dist.init_process_group("nccl")
    ok, notes = True, {}

    def check(name, cond):
        nonlocal ok
        notes[name] = bool(cond)
        ok = ok and bool(cond)

    r = run(rank, world, None, args.steps)
    ctx, dev, g = r["ctx"], r["dev"], r["g"]
    pot = torch.from_numpy(ctx.device_get_grid(capi.GRID_POTENTIAL)).cuda()
    conc = torch.from_numpy(ctx.device_get_grid(capi.GRID_CONCENTRATION)).cuda()
    count = ctx.device_get_grid(capi.GRID_COUNT)
    for name, t in (("potential", pot), ("concentration", conc)):
        lo, hi = t.clone(), t.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        check(name + "_identical_on_all_ranks", torch.equal(lo, hi))
    sizes = torch.tensor([ctx.size], dtype=torch.int64).cuda()
    dist.all_reduce(sizes)
    n_total = int(sizes.item())
    check("count_grid_sums_to_global_ensemble", count.sum() == n_total)
    expected = g["expected_at_contact"].ravel()
    res = g["is_reservoir"].ravel().astype(bool)
    check("reservoir_cells_at_expected_population",
          np.all(count[res] >= np.floor(expected[res])) and np.all(count[res] <= np.ceil(expected[res])))
    total = r["shard"].sum_counters(r["counters"]) if world > 1 else r["counters"].astype(np.int64)
    left, net = int(total[:, 0, :].sum()), int(total[:, 1, :].sum())
    check("bookkeeping", n_total == r["n0"] - left + net)
    check("all_reduce_called_twice_per_step", world == 1 or r["shard"].calls == 2 * args.steps)
    e = r["ens"]
    esum = torch.tensor([float(e.energy.sum()), float(e.n)], dtype=torch.float64).cuda()
    dist.all_reduce(esum)
    mean_energy = float(esum[0] / esum[1])
    # the same run on one rank (every rank does it on its own GPU; rank 0 compares)
    solo = run(0, 1, None, args.steps)
    solo_n, solo_e = solo["ctx"].size, float(solo["ens"].energy.mean())
    check("ensemble_size_close_to_single_gpu_run", abs(n_total - solo_n) <= 6 * np.sqrt(left + abs(net) + 1) + 0.01 * solo_n)
    check("mean_energy_close_to_single_gpu_run", abs(mean_energy / solo_e - 1) < 6 / np.sqrt(solo_n))
    flag = torch.tensor([1 if ok else 0]).cuda()
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps(dict(world=world, steps=args.steps, n_total=n_total, n_single=solo_n, left=left, net=net,
                              mean_energy=mean_energy, mean_energy_single=solo_e, sweeps_mean=float(r["sweeps"].mean()),
                              checks=notes, ok=bool(flag.item()))))
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
