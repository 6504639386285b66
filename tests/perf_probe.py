"""Developer probe (TEST INFRASTRUCTURE, uses the oracle only to build the Si tables):
times the bulk kernel for a few ensemble sizes / steps-per-launch.  Not the benchmark."""
import argparse
import ctypes
import sys
import os
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from helpers import upload_model  # noqa: E402
from scenarios import build_si  # noqa: E402
from viennaemc_b200 import capi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, nargs="+", default=[100000, 10_000_000, 100_000_000])
    ap.add_argument("--spl", type=int, nargs="+", default=[1, 4, 16])
    ap.add_argument("--steps", type=int, default=32)
    ap.add_argument("--math", type=int, default=1)
    ap.add_argument("--dt", type=float, default=1e-16)
    ap.add_argument("--vec", type=int, nargs="+", default=[2])
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--opt", nargs="*", default=[], help="name=value context options")
    ap.add_argument("--settle", type=int, default=0, help="untimed steps (fused) to leave the initial transient")
    args = ap.parse_args()
    import torch
    m = build_si()
    for n in args.n:
        ctx = capi.Context(0)
        upload_model(ctx, m)
        for o in args.opt:
            k, v = o.split("=")
            ctx.set_option(k, int(v))
        box = [(n / 1e23) ** (1 / 3)] * 3
        ctx.generate_bulk_ensemble(n, box, 300.0, 0, seed=1)
        ctx.rng_philox(5)
        ctx.bulk_configure(box, [-1, 0, 0], 1e6, math_mode=args.math)
        obs = torch.zeros(args.steps * 3, dtype=torch.float64, device="cuda")
        if args.settle:
            obs0 = torch.zeros(args.settle * 3, dtype=torch.float64, device="cuda")
            ctx.bulk_step_device(args.dt, args.settle, 16, obs0.data_ptr())
        for spl, vec in [(s, v) for s in args.spl for v in (args.vec if s == 1 else [0])]:
            if vec:
                ctx.set_option("vec", vec)
            ctx.bulk_step_device(args.dt, args.steps, spl, obs.data_ptr())  # warm-up
            ctx.synchronize()
            ms = 1e30
            for _ in range(args.reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ctx.bulk_step_device(args.dt, args.steps, spl, obs.data_ptr())
                e1.record()
                torch.cuda.synchronize()
                ms = min(ms, e0.elapsed_time(e1))
            rate = n * args.steps / (ms * 1e-3)
            launches = (args.steps + spl - 1) // spl
            gbs = 136.0 * n * launches / (ms * 1e-3) / 1e9
            print(f"n={n:>10} spl={spl:>3} vec={vec} steps={args.steps} {ms:9.3f} ms  {rate:.3e} p-steps/s  "
                  f"state traffic {gbs:8.1f} GB/s ({gbs/6555.8*100:5.1f}% of measured HBM peak)", flush=True)
        ctx.close()


if __name__ == "__main__":
    main()
