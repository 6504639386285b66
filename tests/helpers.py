"""Glue between the oracle model (test infrastructure) and the GPU C ABI (product)."""
from __future__ import annotations

import os

import numpy as np

from oracle import pyoracle as po
from viennaemc_b200 import capi

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# tolerance of BASELINE.json north_star: fp64 particle state within 1e-12 relative
STATE_RTOL = 1e-12


def load_golden(case):
    g = np.load(os.path.join(GOLDEN_DIR, case + ".npz"))
    if "draws" in g.files or "draws_count" not in g.files:
        return g
    # large fixture without the raw draws: they are the mt19937_64 stream of the case's seed -- regenerate them and hold them
    # against the digest of what the reference's recorder saw
    import hashlib
    from scenarios import GOLDEN_CASES
    class _Golden(dict):  # like NpzFile: the names of the arrays in .files
        @property
        def files(self):
            return list(self.keys())

    d = _Golden((k, g[k]) for k in g.files)
    from scenarios import DEVICE_LONG_CASES, MHP_CASES, MOS2_CASES
    seed = (GOLDEN_CASES[case]["args"]["seed"] if case in GOLDEN_CASES else
            MOS2_CASES[case]["seed"] if case in MOS2_CASES else
            DEVICE_LONG_CASES[case]["seed"] if case in DEVICE_LONG_CASES else MHP_CASES[case]["seed"])
    draws = po.mt_fill(int(seed), int(d["draws_count"][0]))
    assert hashlib.sha256(draws.tobytes()).digest() == d["draws_sha256"].tobytes(), "regenerated draws differ from the recording"
    d["draws"] = draws
    if "events" in d:
        d["events"] = d["events"].astype(np.int64)
    return d


def golden_ensemble(g, prefix):
    return po.Ensemble.from_arrays(g[prefix + "k"], g[prefix + "pos"], g[prefix + "energy"], g[prefix + "tau"],
                                   g[prefix + "grainTau"], g[prefix + "idx"])


def upload_baths(ctx: capi.Context, baths):
    """oracle phonon baths -> emcgpu_set_phonon_baths (prefix sums included: q-resolved angles may need them)"""
    if baths:
        ctx.set_phonon_baths(len(baths), baths[0].n_bins, baths[0].dq, np.stack([b.cum_w for b in baths]),
                             np.stack([b.cum_wn for b in baths]))


def upload_model(ctx: capi.Context, model: po.Model, valleys_too=True):
    """oracle model -> emcgpu_set_valleys / emcgpu_set_tables"""
    valleys = []
    for v in model.valleys():
        rot = np.array([list(v.rot[s]) for s in range(po.MAX_SUB)])
        valleys.append(capi.make_valley(v.kind, v.deg, v.mCond, v.mDos, v.alpha, v.eBottom, list(v.vogt), rot))
    if valleys_too:
        ctx.set_valleys(valleys)
    sets = []
    for ts in model.tablesets():
        mechs = []
        for m in ts["mech"]:
            fs = np.array([[m.finalSub[s][f] for f in range(max(1, m.nFinal))] for s in range(po.MAX_SUB)])
            mechs.append(capi.make_mech(m.sampler, name=f"mech{m.globalId}", mech_id=m.globalId,
                                        final_valley=m.finalValley, final_sub=fs if m.nFinal > 0 else None,
                                        params=[m.p[0], m.p[1], m.p[2], m.p[3]]))
        sets.append(dict(valley=ts["valley"], region=ts["region"], tau=ts["tau"], cum=ts["cum"], mech=mechs))
    ctx.set_tables(sets, model.n_levels, model.max_energy)
    grain = getattr(model, "grain", None)
    if grain:
        ctx.set_grain(*grain)


def upload_ensemble(ctx: capi.Context, ens: po.Ensemble, particle_id_base=0):
    ctx.set_ensemble([ens.kx, ens.ky, ens.kz, ens.energy, ens.tau, ens.x, ens.y, ens.z], ens.packed(),
                     particle_id_base)
    if getattr(ctx, "grain_on", False):
        ctx.set_grain_clock(ens.grainTau[: ens.n])


def download_ensemble(ctx: capi.Context) -> po.Ensemble:
    streams, packed = ctx.get_ensemble()
    e = po.Ensemble(len(packed))
    e.n = len(packed)
    e.kx, e.ky, e.kz, e.energy, e.tau, e.x, e.y, e.z = streams
    e.valley = (packed & 0xFF).astype(np.int32)
    e.sub = ((packed >> 8) & 0xFF).astype(np.int32)
    e.region = (packed >> 16).astype(np.int32)
    if getattr(ctx, "grain_on", False):
        e.grainTau = ctx.get_grain_clock()
    return e


def rel_err(a, b, scale=None):
    """max |a-b| / max(|b|, scale); scale defaults to the RMS magnitude of b."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if scale is None:
        scale = float(np.sqrt(np.mean(b * b))) if b.size else 1.0
    den = np.maximum(np.abs(b), scale if scale > 0 else 1.0)
    return float(np.max(np.abs(a - b) / den)) if b.size else 0.0


def assert_state_close(got: po.Ensemble, want: po.Ensemble, box, rtol=STATE_RTOL, what=""):
    assert got.n == want.n
    n = want.n
    # indices: bit-exact
    for f in ("valley", "sub", "region"):
        assert np.array_equal(getattr(got, f)[:n], getattr(want, f)[:n]), f"{what}: {f} differs"
    kmag = float(np.sqrt(np.mean(want.kx[:n] ** 2 + want.ky[:n] ** 2 + want.kz[:n] ** 2)))
    errs = {}
    for f in ("kx", "ky", "kz"):
        errs[f] = rel_err(getattr(got, f)[:n], getattr(want, f)[:n], kmag)
    errs["energy"] = rel_err(got.energy[:n], want.energy[:n])
    errs["tau"] = rel_err(got.tau[:n], want.tau[:n])
    for f, b in zip(("x", "y", "z"), box):
        errs[f] = rel_err(getattr(got, f)[:n], getattr(want, f)[:n], b)
    bad = {k: v for k, v in errs.items() if not v <= rtol}
    assert not bad, f"{what}: state differs beyond {rtol}: {bad}"
    return errs


def field_dir_of(args):
    return [float(x) for x in args["fdir"].split(",")]
