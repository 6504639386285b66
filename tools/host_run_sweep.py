"""Slice-count sweep of emcgpu_bulk_run_host on the bench workload (development aid; prints ms per slice count)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from viennaemc_b200 import capi, hostapi

N, K, SPL, DT, DOPING = 100_000_000, int(os.environ.get("SWEEP_K", "1000")), 24, 1e-16, 1e23
box = [(N / DOPING) ** (1.0 / 3.0)] * 3
ctx = capi.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
hostapi.si_upload(ctx, hostapi.si_spec(box=box, spacing=[b / 5 for b in box], doping=DOPING))
ctx.generate_bulk_ensemble(N, box, 300.0, 0, seed=1, particle_id_base=0)
ctx.rng_philox(1)
ctx.bulk_configure(box, [-1, 0, 0], 1e6, math_mode=capi.MATH_FAST)
host = [torch.empty(N, dtype=torch.float64, pin_memory=True) for _ in range(capi.N_STREAMS)]
hp = torch.empty(N, dtype=torch.int32, pin_memory=True)
streams = [h.numpy() for h in host]
packed = hp.numpy().view(np.uint32)
ctx.get_ensemble_into(streams, packed)
out = {}
quantum = 148 * 16 * 64
for slices in (32, 24, 16, 12, 8, 6, 4):
    sl = ((N // slices + quantum - 1) // quantum) * quantum
    ctx.bulk_run_host(streams, packed, DT, SPL, SPL, sl, want_obs=False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    ctx.bulk_run_host(streams, packed, DT, K, SPL, sl)
    e1.record(); torch.cuda.synchronize()
    out[slices] = e0.elapsed_time(e1)
    print(slices, sl, out[slices], flush=True)
json.dump(out, open("gpurun_out/host_run_sweep_k%d.json" % K, "w"))
