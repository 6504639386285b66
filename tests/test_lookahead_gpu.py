"""GPU tests of the look-ahead that lets the reference's one-step-per-call driver loop
(examples/bulkSimulation/bulkSimulation.cpp:150-157: moveParticles(dt), getAvgEnergy, getAvgDriftVelocity,
getValleyOccupationProbability per time step) run on the several-steps-per-launch kernels:

  * C ABI: emcgpu_bulk_step_ahead advances like emcgpu_bulk_step and keeps the ensemble it started from;
    emcgpu_bulk_rewind + stepping the first k steps again gives bit for bit the state of a run that only ever did k steps
    (flight + event kernels: the first flight launch writes a second set of streams; other kernels copy first);
  * drop-in handler: the driver's results do not depend on the look-ahead depth, also when the driver asks for the
    ensemble in the middle of a look-ahead window (handler.print) -- the handler rewinds and repeats the served steps.
"""
import os
import subprocess

import numpy as np
import pytest

from helpers import download_ensemble, upload_model
from oracle import pyoracle as po
from scenarios import build_si
from viennaemc_b200 import capi

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "viennaemc_b200", "bin")
FIELDS = po.Ensemble.F64[:5] + po.Ensemble.F64[6:] + po.Ensemble.I32  # everything but the never-read grain clock


def _ctx(gpu_ctx_factory, m, n, box, multi_kernel):
    ctx = gpu_ctx_factory()
    ctx.set_option("multi_kernel", multi_kernel)
    upload_model(ctx, m)
    ctx.generate_bulk_ensemble(n, box, 300.0, 0, seed=21)
    ctx.rng_philox(77)
    ctx.bulk_configure(box, [-1, 0, 0], 1e6, math_mode=capi.MATH_FAST)
    ctx.set_step_index(1)
    return ctx


@pytest.mark.parametrize("multi_kernel,n", [(3, 200_003), (1, 30_011), (2, 70_001), (0, 3_000_017)],
                         ids=["split", "inplace", "deferred", "auto-large"])
def test_step_ahead_then_rewind_reproduces_the_earlier_state(gpu_ctx_factory, multi_kernel, n):
    m = build_si()
    box = [1e-6] * 3
    dt, ahead, served = 1e-15, 24, 9
    a = _ctx(gpu_ctx_factory, m, n, box, multi_kernel)
    obs_ahead = a.bulk_step_ahead(dt, ahead, ahead)
    assert a.L.emcgpu_get_step_index(a.h) == 1 + ahead
    full = download_ensemble(a)  # the ensemble after all the steps the device ran ahead
    a.bulk_rewind()
    assert a.L.emcgpu_get_step_index(a.h) == 1
    obs_again = a.bulk_step(dt, served, served)
    part = download_ensemble(a)
    # the same steps without any look-ahead
    b = _ctx(gpu_ctx_factory, m, n, box, multi_kernel)
    obs_b = b.bulk_step(dt, ahead, ahead)
    ref_full = download_ensemble(b)
    c = _ctx(gpu_ctx_factory, m, n, box, multi_kernel)
    obs_c = c.bulk_step(dt, served, served)
    ref_part = download_ensemble(c)
    for f in FIELDS:
        assert np.array_equal(getattr(full, f), getattr(ref_full, f)), f
        assert np.array_equal(getattr(part, f), getattr(ref_part, f)), f
    assert np.array_equal(obs_ahead[:, :, 2], obs_b[:, :, 2])
    assert np.allclose(obs_ahead, obs_b, rtol=1e-12) and np.allclose(obs_again, obs_c, rtol=1e-12)
    assert np.allclose(obs_ahead[:served], obs_again, rtol=1e-12)  # what was handed out ahead is what the steps give
    # a second rewind has nothing to go back to
    with pytest.raises(capi.EmcGpuError):
        a.bulk_rewind()
    for x in (a, b, c):
        x.close()


@pytest.mark.parametrize("multi_kernel", [3, 1], ids=["split", "inplace"])
def test_step_ahead_carries_the_grain_clocks(gpu_ctx_factory, multi_kernel):
    """a grain mechanism (second free-flight clock per particle): the look-ahead copy has clocks of its own, a rewind brings
    back the clocks of the state before the launch"""
    from helpers import upload_ensemble
    m = build_si()
    m.set_grain(0.4, 2e13)
    m.grain = (0.4, 2e13)
    box = [8e-7] * 3
    ens, _ = m.generate_initial(box, [8, 8, 8], 1e23, po.mt_state(3))  # 51200 particles
    dt, ahead, served = 1e-15, 24, 9

    def fresh():
        ctx = gpu_ctx_factory()
        ctx.set_option("multi_kernel", multi_kernel)
        upload_model(ctx, m)
        upload_ensemble(ctx, ens)
        ctx.rng_philox(77)
        ctx.bulk_configure(box, [-1, 0, 0], 1e6, math_mode=capi.MATH_FAST)
        ctx.set_step_index(1)
        return ctx

    a = fresh()
    a.bulk_step_ahead(dt, ahead, ahead)
    full = download_ensemble(a)
    a.bulk_rewind()
    assert np.array_equal(download_ensemble(a).grainTau, ens.grainTau[: ens.n])
    a.bulk_step(dt, served, served)
    part = download_ensemble(a)
    b = fresh()
    b.bulk_step(dt, ahead, ahead)
    ref_full = download_ensemble(b)
    c = fresh()
    c.bulk_step(dt, served, served)
    ref_part = download_ensemble(c)
    assert not np.array_equal(ref_full.grainTau, ens.grainTau[: ens.n])
    for f in FIELDS + ("grainTau",):
        assert np.array_equal(getattr(full, f), getattr(ref_full, f)), f
        assert np.array_equal(getattr(part, f), getattr(ref_part, f)), f


def _run_driver(tmp, prefix, *extra):
    r = subprocess.run([os.path.join(BIN, "bulkSimulation"), "--seed", "11", "--particles", "20000", "--steps", "150",
                        "--dt", "1e-15", "--prefix", prefix, *extra], cwd=tmp, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    out = {k: np.loadtxt(os.path.join(tmp, prefix + k + ".txt")) for k in ("AvgEnergy", "AvgDriftVelocity",
                                                                          "valleyOccupation")}
    return out


def test_driver_results_do_not_depend_on_the_lookahead(tmp_path):
    tmp = str(tmp_path)
    base = _run_driver(tmp, "la1", "--lookahead", "1", "--print-at", "37")
    for la in ("16", "7", "24"):
        got = _run_driver(tmp, "la" + la, "--lookahead", la, "--print-at", "37")
        for k, ref in base.items():
            assert got[k].shape == ref.shape == (151, 2)
            assert np.allclose(got[k], ref, rtol=1e-5 if k != "valleyOccupation" else 0, atol=0), (la, k)  # 6 digits in the files
        # the ensemble the driver printed in the middle of a look-ahead window is the ensemble of THAT step
        with open(os.path.join(tmp, "la1Electrons37.txt")) as f, open(os.path.join(tmp, f"la{la}Electrons37.txt")) as g:
            assert f.read() == g.read(), la


@pytest.mark.parametrize("components", [1, 3])
def test_recorded_velocities_against_the_oracle(gpu_ctx_factory, components):
    """emcgpu_bulk_record_velocities: the per-particle velocities of every step (printDriftVelocities / printVelocities,
    basicBulkParticleHandler.hpp:251-285) as the step kernel streams them to the host, against the oracle stepping the same
    Philox streams one step at a time and evaluating getVelocity on its ensemble."""
    import ctypes as C
    m = build_si()
    box = [4e-7] * 3
    ens, _ = m.generate_initial(box, [4, 4, 4], 1e23, po.mt_state(5))
    n_steps, dt, seed = 23, 4e-16, 99
    ctx = gpu_ctx_factory()
    upload_model(ctx, m)
    from helpers import upload_ensemble
    upload_ensemble(ctx, ens)
    ctx.rng_philox(seed)
    ctx.bulk_configure(box, [-1, 0, 0], 1e6, math_mode=capi.MATH_FAST)
    ctx.set_step_index(1)
    vel = ctx.record_velocities(components, n_steps, ens.n)
    ctx.bulk_step(dt, n_steps, 8)
    ctx.record_velocities(0)
    ref = ens.copy()
    v = m.valley(0)
    out = np.zeros(3)
    scale = None
    for s in range(n_steps):
        m.bulk_steps(ref, box, [-1, 0, 0], 1e6, dt, 1, po.rng_philox(seed), first_step=1 + s)
        want = np.zeros((ens.n, 3))
        for i in range(ens.n):
            k = np.array([ref.kx[i], ref.ky[i], ref.kz[i]])
            po.lib().orc_velocity(C.byref(v), po._dp(k), float(ref.energy[i]), int(ref.sub[i]), po._dp(out))
            want[i] = out
        scale = scale or float(np.sqrt((want ** 2).sum(axis=1).mean()))
        if components == 1:
            assert np.max(np.abs(vel[s, :, 0] + want[:, 0])) <= 1e-12 * scale, s  # field direction (-1, 0, 0)
        else:
            assert np.max(np.abs(vel[s] - want)) <= 1e-12 * scale, s


def test_driver_velocity_lines_do_not_depend_on_the_lookahead(tmp_path):
    tmp = str(tmp_path)
    for comps in ("1", "3"):
        _run_driver(tmp, "v1_" + comps, "--lookahead", "1", "--velocities", comps)
        ref = np.loadtxt(os.path.join(tmp, "v1_" + comps + "Velocities.txt"))
        assert ref.shape[0] == 150 and ref.shape[1] % int(comps) == 0
        for la in ("16", "7"):
            _run_driver(tmp, f"v{la}_{comps}", "--lookahead", la, "--velocities", comps)
            got = np.loadtxt(os.path.join(tmp, f"v{la}_{comps}Velocities.txt"))
            assert got.shape == ref.shape
            assert np.allclose(got, ref, rtol=2e-5, atol=1e-3), (comps, la)  # 6 significant digits in the file
