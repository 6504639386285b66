"""A few self-consistent steps of a device run (NEC-VWD scheme, gate, red-black cluster solver) through the C ABI: small enough for
compute-sanitizer's racecheck (tools/gpu_sanitize_*.sh).  Prints the bookkeeping; exits non-zero when it does not add up.

    compute-sanitizer --tool racecheck python tools/short_device_run.py [steps]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import numpy as np  # noqa: E402

from helpers import load_golden, upload_ensemble, upload_model  # noqa: E402
from scenarios import build_device  # noqa: E402
from test_device_gpu import configure  # noqa: E402
from test_oracle_device import ens_from  # noqa: E402
from viennaemc_b200 import capi  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
for case in ("device_vwd", "device_bar"):
    g = load_golden(case)
    m, dev = build_device(case)
    ctx = capi.Context(0)
    upload_model(ctx, m)
    configure(ctx, dev, math_mode=capi.MATH_FAST)
    ctx.set_option("sor_order", 1)
    ctx.device_set_grid(capi.GRID_POTENTIAL, g["pot_eq"])
    ctx.device_set_grid(capi.GRID_CONCENTRATION, g["conc_eq"])
    upload_ensemble(ctx, ens_from(g, "init_"))
    ctx.device_reserve(4096)
    ctx.rng_philox(11)
    ctx.set_step_index(1)
    n0 = ctx.size
    counters, sweeps = ctx.device_run(2e-15, steps, 1e-4, 1.8, True, n_average=steps // 2)
    left, net = int(counters[:, 0, :].sum()), int(counters[:, 1, :].sum())
    count = ctx.device_get_grid(capi.GRID_COUNT)
    print(f"{case}: {n0} -> {ctx.size} particles, {left} left, {net} net injected, sweeps {sweeps.tolist()}, count sum {count.sum()}")
    assert ctx.size == n0 - left + net and count.sum() == ctx.size and np.all(sweeps >= 1)
    ctx.close()
