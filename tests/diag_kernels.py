"""Developer diagnostic: compare the three step kernels (TMA pipeline, plain streaming, fused) bit for bit."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from helpers import upload_model, download_ensemble
from scenarios import build_si
from viennaemc_b200 import capi
from oracle import pyoracle as po

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 16
dt = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-15
m = build_si()
box = [1e-6] * 3
res = {}
for name, kernel, spl in (("tma", 0, 1), ("stream", 1, 1), ("fused", 0, 8)):
    ctx = capi.Context(0)
    upload_model(ctx, m)
    ctx.generate_bulk_ensemble(n, box, 300.0, 0, seed=3)
    ctx.rng_philox(11)
    ctx.set_option("kernel", kernel)
    ctx.bulk_configure(box, [-1, 0, 0], 1e6, math_mode=capi.MATH_FAST)
    ctx.set_step_index(1)
    obs = ctx.bulk_step(dt, steps, spl)
    res[name] = (download_ensemble(ctx), obs)
    ctx.close()
ref = res["fused"][0]
for name in ("tma", "stream"):
    e = res[name][0]
    for f in po.Ensemble.F64[:5] + po.Ensemble.F64[6:] + po.Ensemble.I32:
        a, b = getattr(e, f), getattr(ref, f)
        bad = np.nonzero(a != b)[0]
        if len(bad):
            rel = np.abs(a[bad].astype(float) - b[bad]) / np.maximum(np.abs(b[bad]), 1e-300)
            print(f"{name}/{f}: {len(bad)} differ, first idx {bad[:5]}, max rel {rel.max():.3e}, median rel {np.median(rel):.3e}")
        else:
            print(f"{name}/{f}: identical")
    print(name, "obs max rel diff", np.max(np.abs(res[name][1] - res["fused"][1]) / np.abs(res["fused"][1]).max(axis=0)))
