"""Times the several-steps-per-launch bulk kernels against each other on one GPU (development tool).

    python tools/kernel_sweep.py [--particles 1e8] [--steps 192]
prints one JSON line per configuration (kernel, steps per launch, particles per lane)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    from viennaemc_b200 import capi, hostapi

    ap = argparse.ArgumentParser()
    ap.add_argument("--particles", type=float, default=1e8)
    ap.add_argument("--steps", type=int, default=192)
    ap.add_argument("--settle", type=int, default=2000)
    ap.add_argument("--grain-rate", type=float, default=0.0, help="> 0: a grain mechanism with this event rate [1/s]")
    ap.add_argument("--configs", default="2:8:4,3:8:4,3:16:4,3:24:4,3:16:2,3:8:2")
    a = ap.parse_args()
    n = int(a.particles)
    box = [(n / 1e23) ** (1.0 / 3.0)] * 3
    ctx = capi.Context(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    hostapi.si_upload(ctx, hostapi.si_spec(box=box, spacing=[b / 5 for b in box], doping=1e23))
    ctx.generate_bulk_ensemble(n, box, 300.0, 0, seed=12345)
    ctx.rng_philox(12345)
    if a.grain_rate > 0:
        import numpy as np
        ctx.set_grain(0.5, a.grain_rate)
        ctx.set_grain_clock(np.random.default_rng(1).exponential(1.0 / a.grain_rate, n))
    ctx.bulk_configure(box, [-1, 0, 0], 1e6, math_mode=capi.MATH_FAST)
    ctx.set_step_index(1)
    obs = torch.zeros(max(a.settle, a.steps) * 3, dtype=torch.float64, device="cuda")
    if a.settle:
        ctx.bulk_step_device(1e-16, a.settle, 16, obs.data_ptr())
    for cfg in a.configs.split(","):
        mk, spl, ppl = (int(x) for x in cfg.split(":")[:3])
        ctx.set_option("multi_kernel", mk)
        ctx.set_option("split_ppl", ppl)
        ctx.bulk_step_device(1e-16, 2 * spl, spl, obs.data_ptr())
        ctx.set_option("kernel_timing", 1)
        ctx.kernel_times(reset=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        l0 = ctx.launch_count
        e0.record()
        ctx.bulk_step_device(1e-16, a.steps, spl, obs.data_ptr())
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        kms, kn = ctx.kernel_times(reset=True)
        o = obs[: a.steps * 3].view(a.steps, 3).cpu().numpy()
        print(json.dumps({"multi_kernel": mk, "spl": spl, "ppl": ppl, "ms_per_step": ms / a.steps,
                          "particle_steps_per_s": n * a.steps / (ms * 1e-3), "launches": ctx.launch_count - l0, "flight_ms": kms[0] / max(1, kn[0]), "event_ms": kms[1] / max(1, kn[1]),
                          "other_ms": kms[2] / max(1, kn[0]),
                          "mean_E": float((o[:, 0] / o[:, 2]).mean()), "mean_v": float((o[:, 1] / o[:, 2]).mean())}),
              flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
