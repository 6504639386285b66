"""GPU parity tests of the bulk time-step kernel, through the C ABI (include/emcgpu.h).

Correctness contract (BASELINE.json north_star):
  * replay mode fed the reference's random draws: mechanism / valley / sub-valley indices
    bit-exact, fp64 particle state within 1e-12 relative after N steps;
  * Philox mode against the oracle consuming the identical counter-based streams: same bar.
"""
import numpy as np
import pytest

from helpers import (STATE_RTOL, assert_state_close, download_ensemble, field_dir_of, golden_ensemble, load_golden,
                     upload_ensemble, upload_model)
from oracle import pyoracle as po
from scenarios import GOLDEN_CASES, build_model, build_si
from viennaemc_b200 import capi

pytestmark = pytest.mark.gpu

CASES = list(GOLDEN_CASES)


def _reference_replay_streams(case):
    """Per-particle replay streams of the REFERENCE's recorded draws.  The attribution of each raw
    draw to a particle comes from the oracle run that reproduces the reference bit for bit
    (tests/test_oracle_golden.py proves the equality)."""
    g = load_golden(case)
    a = GOLDEN_CASES[case]["args"]
    m = build_model(case)
    st = po.mt_state(int(a["seed"]))
    ens, used = m.generate_initial([a["box"]] * 3, [a["cells"]] * 3, a["doping"], st)
    res = m.bulk_steps(ens.copy(), [a["box"]] * 3, field_dir_of(a), a["field"], a["dt"], a["steps"], po.rng_mt(st),
                       first_step=1, record=True)
    draws, offsets = po.streams_from_record(g["draws"][used:], res["rec_pid"], ens.n)
    return g, a, m, draws, offsets


@pytest.mark.parametrize("math_mode", [capi.MATH_EXACT, capi.MATH_FAST], ids=["exact", "fast"])
@pytest.mark.parametrize("steps_per_launch,multi_kernel", [(1, 0), (7, 1), (7, 2), (16, 2), (7, 3), (24, 3)],
                         ids=["spl1", "spl7-inplace", "spl7-deferred", "spl16-deferred", "spl7-split", "spl24-split"])
@pytest.mark.parametrize("case", CASES)
def test_replay_of_reference_draws(gpu_ctx_factory, case, steps_per_launch, multi_kernel, math_mode):
    g, a, m, draws, offsets = _reference_replay_streams(case)
    ctx = gpu_ctx_factory()
    # 1: K1b (events in place), 2: K1c (deferred events), 3: K1d (flight kernel + event kernel; FAST arithmetic only,
    # EXACT runs K1c)
    ctx.set_option("multi_kernel", multi_kernel)
    upload_model(ctx, m)
    upload_ensemble(ctx, golden_ensemble(g, "init_"))
    ctx.rng_replay(draws, offsets)
    box = [a["box"]] * 3
    ctx.bulk_configure(box, field_dir_of(a), a["field"], math_mode=math_mode)
    ctx.set_step_index(1)
    ctx.event_log_enable(1 << 20)
    obs = ctx.bulk_step(a["dt"], a["steps"], steps_per_launch)
    got = download_ensemble(ctx)
    want = golden_ensemble(g, "final_")
    # state: indices exact, fp64 within 1e-12 relative -- against the REFERENCE's final ensemble
    assert_state_close(got, want, box, STATE_RTOL, f"{case}/replay")
    if getattr(m, "grain", None):  # the grain clocks (a new one is drawn at every grain event) agree as well
        assert np.allclose(got.grainTau, want.grainTau[: want.n], rtol=1e-11, atol=1e-27)
    # every scatter event: same step, same particle, same mechanism as the reference logged
    ev, n_ev = ctx.event_log_read(1 << 20)
    assert n_ev == len(ev)
    real = ev[ev[:, 2] >= 0][:, [0, 1, 3]]
    real = real[np.lexsort((real[:, 2], real[:, 1], real[:, 0]))]  # device log order is arbitrary
    ref_ev = g["events"]
    ref_ev = ref_ev[np.lexsort((ref_ev[:, 2], ref_ev[:, 1], ref_ev[:, 0]))]
    assert np.array_equal(real, ref_ev)
    # per-step observables against the reference's getAvgEnergy / getAvgDriftVelocity / occupation
    cnt = obs[:, :, 2]
    assert np.all(cnt.sum(axis=1) == want.n)
    with np.errstate(invalid="ignore", divide="ignore"):
        avg_e = np.where(cnt > 0, obs[:, :, 0] / cnt, 0.0)
        avg_v = np.where(cnt > 0, obs[:, :, 1] / cnt, 0.0)
    assert np.array_equal(cnt / want.n, g["obs"][1:, 2, :])
    assert np.allclose(avg_e, g["obs"][1:, 0, :], rtol=1e-11, atol=0)
    vscale = np.abs(g["obs"][1:, 1, :]).max()
    assert np.max(np.abs(avg_v - g["obs"][1:, 1, :])) <= 1e-11 * vscale


@pytest.mark.parametrize("math_mode", [capi.MATH_EXACT, capi.MATH_FAST], ids=["exact", "fast"])
@pytest.mark.parametrize("multi_kernel", [1, 2, 3], ids=["inplace", "deferred", "split"])
@pytest.mark.parametrize("case", ["si_bulk", "mixed", "si_grain"])
def test_philox_against_oracle(gpu_ctx_factory, case, math_mode, multi_kernel):
    """Independent (counter-based) RNG: GPU and oracle consume identical Philox streams."""
    a = dict(GOLDEN_CASES[case]["args"])
    m = build_model(case)
    box = [4e-7] * 3
    st = po.mt_state(99)
    ens, _ = m.generate_initial(box, [4, 4, 4], 1e23, st)  # 6400 particles
    n_steps, dt, seed, base = 150, 4 * a["dt"], 0xC0FFEE1234, 1000
    ctx = gpu_ctx_factory()
    ctx.set_option("multi_kernel", multi_kernel)
    upload_model(ctx, m)
    upload_ensemble(ctx, ens, particle_id_base=base)
    ctx.rng_philox(seed)
    ctx.bulk_configure(box, field_dir_of(a), a["field"], math_mode=math_mode)
    ctx.set_step_index(1)
    ctx.event_log_enable(1 << 22)
    obs = ctx.bulk_step(dt, n_steps, 5)
    got = download_ensemble(ctx)
    ref = ens.copy()
    res = m.bulk_steps(ref, box, field_dir_of(a), a["field"], dt, n_steps, po.rng_philox(seed, base), first_step=1,
                       log_events=True)
    assert_state_close(got, ref, box, STATE_RTOL, f"{case}/philox")
    ev, n_ev = ctx.event_log_read(1 << 22)
    assert n_ev == len(res["events"]) and n_ev > 1000
    dev = ev[np.lexsort((ev[:, 3], ev[:, 2], ev[:, 1], ev[:, 0]))]
    dev[:, 1] -= base
    cpu = res["events"]
    cpu = cpu[np.lexsort((cpu[:, 3], cpu[:, 2], cpu[:, 1], cpu[:, 0]))]
    assert np.array_equal(dev, cpu)
    assert np.allclose(obs[:, :, 2], res["obs"][:, :, 2], rtol=0, atol=0)
    assert np.allclose(obs[:, :, 0], res["obs"][:, :, 0], rtol=1e-11)
    assert np.max(np.abs(obs[:, :, 1] - res["obs"][:, :, 1])) <= 1e-11 * np.abs(res["obs"][:, :, 1]).max()


@pytest.mark.parametrize("multi_kernel", [1, 2, 3], ids=["inplace", "deferred", "split"])
def test_sharding_invariance_and_determinism(gpu_ctx_factory, multi_kernel):
    """Philox key = global particle id: a shard [lo,hi) evolves exactly as inside the full ensemble,
    and two runs give bit-identical particle state."""
    m = build_si()
    box = [5e-7] * 3
    st = po.mt_state(5)
    ens, _ = m.generate_initial(box, [5, 5, 5], 1e23, st)  # 12500 particles = the shipped example
    assert abs(ens.n - 12500) <= 2

    def run(sub: po.Ensemble, base):
        ctx = gpu_ctx_factory()
        ctx.set_option("multi_kernel", multi_kernel)
        upload_model(ctx, m)
        upload_ensemble(ctx, sub, particle_id_base=base)
        ctx.rng_philox(42)
        ctx.bulk_configure(box, [-1, 0, 0], 1e6, math_mode=capi.MATH_FAST)
        ctx.set_step_index(1)
        obs = ctx.bulk_step(1e-15, 64, 8)
        return download_ensemble(ctx), obs

    full, obs_full = run(ens, 0)
    again, _ = run(ens, 0)
    for f in po.Ensemble.F64[:5] + po.Ensemble.F64[6:] + po.Ensemble.I32:
        assert np.array_equal(getattr(full, f), getattr(again, f)), f
    cut = 5000
    lo, hi = ens.copy(), ens.copy()
    for f in po.Ensemble.F64 + po.Ensemble.I32:
        setattr(lo, f, getattr(ens, f)[:cut].copy())
        setattr(hi, f, getattr(ens, f)[cut:].copy())
    lo.n, hi.n = cut, ens.n - cut
    a, obs_a = run(lo, 0)
    b, obs_b = run(hi, cut)
    for f in po.Ensemble.F64[:5] + po.Ensemble.F64[6:] + po.Ensemble.I32:
        assert np.array_equal(np.concatenate([getattr(a, f), getattr(b, f)]), getattr(full, f)), f
    # the "allreduce" of the shard observables equals the single-GPU observables
    assert np.array_equal((obs_a + obs_b)[:, :, 2], obs_full[:, :, 2])
    assert np.allclose(obs_a + obs_b, obs_full, rtol=1e-12)


def test_observables_only_and_empty_steps(gpu_ctx_factory):
    g = load_golden("mixed")
    a = GOLDEN_CASES["mixed"]["args"]
    m = build_model("mixed")
    ens = golden_ensemble(g, "init_")
    ctx = gpu_ctx_factory()
    upload_model(ctx, m)
    upload_ensemble(ctx, ens)
    ctx.bulk_configure([a["box"]] * 3, field_dir_of(a), a["field"], math_mode=capi.MATH_EXACT)
    obs = ctx.bulk_observables()
    cnt = obs[:, 2]
    with np.errstate(invalid="ignore", divide="ignore"):
        assert np.allclose(np.where(cnt > 0, obs[:, 0] / cnt, 0), g["obs"][0, 0, :], rtol=1e-12)
        ref_v = g["obs"][0, 1, :]
        assert np.max(np.abs(np.where(cnt > 0, obs[:, 1] / cnt, 0) - ref_v)) <= 1e-11 * np.abs(ref_v).max()
    assert np.array_equal(cnt / ens.n, g["obs"][0, 2, :])


def test_unsupported_mechanism_is_rejected_with_its_name(gpu_ctx_factory):
    m = build_si()
    ctx = gpu_ctx_factory()
    valleys = []
    for v in m.valleys():
        rot = np.array([list(v.rot[s]) for s in range(po.MAX_SUB)])
        valleys.append(capi.make_valley(v.kind, v.deg, v.mCond, v.mDos, v.alpha, v.eBottom, list(v.vogt), rot))
    ctx.set_valleys(valleys)
    ts = m.tablesets()[0]
    mechs = [capi.make_mech(capi.SAMPLER_ISOTROPIC_ELASTIC, "Acoustic")] + \
            [capi.make_mech(capi.SAMPLER_NONE, "RemoteSurfaceOpticalPhonon")] * (ts["cum"].shape[0] - 1)
    with pytest.raises(capi.EmcGpuError) as ei:
        ctx.set_tables([dict(valley=0, region=0, tau=ts["tau"], cum=ts["cum"], mech=mechs)], m.n_levels, m.max_energy)
    assert ei.value.code == capi.E_UNSUPPORTED_MECHANISM
    assert "RemoteSurfaceOpticalPhonon" in str(ei.value) and "no CPU fallback" in str(ei.value)


def test_call_order_errors(gpu_ctx_factory):
    ctx = gpu_ctx_factory()
    with pytest.raises(capi.EmcGpuError) as ei:
        ctx.n_valleys = 1
        ctx.bulk_step(1e-16, 1, 1)
    assert ei.value.code == capi.E_INVALID


def test_replay_exhaustion_is_reported(gpu_ctx_factory):
    g, a, m, draws, offsets = _reference_replay_streams("si_coulomb_bigdt")
    ctx = gpu_ctx_factory()
    upload_model(ctx, m)
    upload_ensemble(ctx, golden_ensemble(g, "init_"))
    ctx.rng_replay(draws, offsets)
    ctx.bulk_configure([a["box"]] * 3, field_dir_of(a), a["field"])
    with pytest.raises(capi.EmcGpuError) as ei:
        ctx.bulk_step(a["dt"], a["steps"] + 20, 4)  # more steps than were recorded
    assert ei.value.code == capi.E_REPLAY_EXHAUSTED


def test_device_generated_ensemble_statistics(gpu_ctx_factory):
    """emcgpu_generate_bulk_ensemble draws from the reference's initial distributions
    (emcElectron.hpp:75-90, emcParticleInitialization.hpp:36-51)."""
    m = build_si()
    ctx = gpu_ctx_factory()
    upload_model(ctx, m)
    n = 1 << 20
    box = [1e-6, 2e-6, 3e-6]
    ctx.generate_bulk_ensemble(n, box, 300.0, 0, seed=7)
    e = download_ensemble(ctx)
    vt = 1.38066e-23 / 1.60219e-19 * 300.0
    se = 1.5 * vt / np.sqrt(n)
    assert abs(e.energy.mean() - 1.5 * vt) < 6 * se  # E = -1.5 Vt ln U  -> mean 1.5 Vt (truncated at U=1e-6)
    assert np.all(e.valley == 0) and set(np.unique(e.sub)) == {0, 1, 2}
    for arr, b in zip((e.x, e.y, e.z), box):
        assert arr.min() >= 0 and arr.max() <= b and abs(arr.mean() / b - 0.5) < 6 / np.sqrt(12 * n)
    assert abs(e.tau.mean() / m.tau(0, 0) - 1.0) < 6 / np.sqrt(n)
    # |k| consistent with E through the valley dispersion
    import ctypes as C
    v = m.valley(0)
    for i in range(0, 1000, 37):
        k = np.array([e.kx[i], e.ky[i], e.kz[i]])
        assert abs(po.lib().orc_energy(C.byref(v), po._dp(k)) / e.energy[i] - 1) < 1e-12
    # isotropy
    kn = np.sqrt(e.kx ** 2 + e.ky ** 2 + e.kz ** 2)
    assert abs((e.kz / kn).mean()) < 6 / np.sqrt(3 * n)


def test_large_ensemble_properties(gpu_ctx_factory):
    """Full-size shard (BASELINE configs[1] scale is 1e8; here 2^24 to bound test time): particle count
    conserved every step, positions stay inside the periodic box, energies positive, drift velocity
    anti-parallel to the field for electrons, fused and single-step launches agree bit for bit."""
    m = build_si()
    n = 1 << 24
    box = [1e-6] * 3
    res = []
    # one step per launch (K1a), deferred events (K1c), in place (K1b), flight + event kernels (K1d: forced, and as the
    # automatic choice for an ensemble of this size)
    for spl, mk in ((1, 0), (8, 2), (8, 1), (16, 3), (8, 0)):
        ctx = gpu_ctx_factory()
        ctx.set_option("multi_kernel", mk)
        upload_model(ctx, m)
        ctx.generate_bulk_ensemble(n, box, 300.0, 0, seed=3)
        ctx.rng_philox(11)
        ctx.bulk_configure(box, [-1, 0, 0], 1e6, math_mode=capi.MATH_FAST)
        ctx.set_step_index(1)
        obs = ctx.bulk_step(1e-15, 16, spl)
        e = download_ensemble(ctx)
        res.append((e, obs))
        ctx.close()
    (e1, o1) = res[0]
    for e8, o8 in res[1:]:
        for f in po.Ensemble.F64[:5] + po.Ensemble.F64[6:] + po.Ensemble.I32:
            assert np.array_equal(getattr(e1, f), getattr(e8, f)), f
        assert np.array_equal(o1[:, :, 2], o8[:, :, 2])
        assert np.allclose(o1, o8, rtol=1e-12)
    assert np.all(o1[:, :, 2].sum(axis=1) == n)
    for arr in (e1.x, e1.y, e1.z):
        assert arr.min() >= 0 and arr.max() <= 1e-6
    assert e1.energy.min() > 0 and np.isfinite(e1.energy).all()
    vd = o1[:, 0, 1] / n
    assert vd[-1] < 0 and abs(vd[-1]) > abs(vd[0])  # electrons accelerate against the field direction (-1,0,0)... sign per reference
