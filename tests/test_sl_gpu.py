"""GPU parity of the single-layer (2-D material) path -- SURVEY.md 8 f4, examples/singleLayerMoS2 with its default parameter
set: the single-layer valley classes (K isotropic, Q anisotropic with in-plane frames) and the final-state samplers of
emcAcousticSingleLayerMechanism / emcZeroOrderSingleLayerInterValley*ScatterMechanism behind the same mechanism-ID table.

  * replay mode fed the REFERENCE's recorded draws (tests/golden/mos2_*.npz, oracle/_ref/ref_bulk_driver --material mos2):
    every scatter event (step, particle, mechanism) and the valley / sub-valley indices exact, fp64 state within 1e-12 of the
    reference's final ensemble, per-step observables of both valleys within 1e-11;
  * Philox mode against the oracle consuming the identical counter-based streams: same bars, larger ensemble."""
import numpy as np
import pytest

from helpers import STATE_RTOL, assert_state_close, download_ensemble, field_dir_of, golden_ensemble, load_golden, upload_ensemble, upload_model
from oracle import pyoracle as po
from scenarios import MOS2_CASES, MOS2_LZ, build_mos2, build_mos2_pilotto
from viennaemc_b200 import capi

pytestmark = pytest.mark.gpu
CASES = list(MOS2_CASES)
KERNELS = [(1, 0), (7, 1), (7, 2), (16, 2)]
KERNEL_IDS = ["spl1", "spl7-inplace", "spl7-deferred", "spl16-deferred"]


def box_of(a):
    return [a["box"], a["box"], MOS2_LZ]


@pytest.mark.parametrize("math_mode", [capi.MATH_EXACT, capi.MATH_FAST], ids=["exact", "fast"])
@pytest.mark.parametrize("steps_per_launch,multi_kernel", KERNELS, ids=KERNEL_IDS)
@pytest.mark.parametrize("case", CASES)
def test_replay_of_reference_draws(gpu_ctx_factory, case, steps_per_launch, multi_kernel, math_mode):
    g = load_golden(case)
    a = MOS2_CASES[case]
    m = build_mos2(case)
    box = box_of(a)
    # which particle consumed which of the reference's draws: from the oracle run that reproduces the reference bit for bit
    # (tests/test_oracle_sl.py)
    st = po.mt_state(int(a["seed"]))
    ens, used = m.generate_initial(box, [a["cells"], a["cells"], 1], 1.0, st, capacity=4096)
    res = m.bulk_steps(ens.copy(), box, field_dir_of(a), a["field"], a["dt"], a["steps"], po.rng_mt(st), first_step=1, record=True)
    draws, offsets = po.streams_from_record(g["draws"][used:], res["rec_pid"], ens.n)
    ctx = gpu_ctx_factory()
    ctx.set_option("multi_kernel", multi_kernel)
    upload_model(ctx, m)
    upload_ensemble(ctx, golden_ensemble(g, "init_"))
    ctx.rng_replay(draws, offsets)
    ctx.bulk_configure(box, field_dir_of(a), a["field"], math_mode=math_mode)
    ctx.set_step_index(1)
    ctx.event_log_enable(1 << 20)
    obs = ctx.bulk_step(a["dt"], a["steps"], steps_per_launch)
    got = download_ensemble(ctx)
    want = golden_ensemble(g, "final_")
    assert_state_close(got, want, box, STATE_RTOL, f"{case}/replay")
    assert np.all(got.kz == 0) and np.array_equal(got.z, want.z[: want.n])  # nothing leaves the plane
    ev, n_ev = ctx.event_log_read(1 << 20)
    assert n_ev == len(ev)
    real = ev[ev[:, 2] >= 0][:, [0, 1, 3]]
    real = real[np.lexsort((real[:, 2], real[:, 1], real[:, 0]))]
    ref_ev = g["events"].astype(np.int64)
    ref_ev = ref_ev[np.lexsort((ref_ev[:, 2], ref_ev[:, 1], ref_ev[:, 0]))]
    assert np.array_equal(real, ref_ev)
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
@pytest.mark.parametrize("multi_kernel", [1, 2], ids=["inplace", "deferred"])
def test_philox_against_oracle(gpu_ctx_factory, math_mode, multi_kernel):
    m = build_mos2_pilotto()
    box = [4e-7, 4e-7, MOS2_LZ]
    ens, _ = m.generate_initial(box, [40, 40, 1], 1.0, po.mt_state(99), capacity=20000)  # 41 x 41 x 2 x 4 = 13448 electrons
    assert ens.n == 13448
    n_steps, dt, seed, base, field, fdir = 120, 5e-16, 0xC0FFEE77, 500, 6e6, [1.0, 0.3, 0.0]
    ctx = gpu_ctx_factory()
    ctx.set_option("multi_kernel", multi_kernel)
    upload_model(ctx, m)
    upload_ensemble(ctx, ens, particle_id_base=base)
    ctx.rng_philox(seed)
    ctx.bulk_configure(box, fdir, field, math_mode=math_mode)
    ctx.set_step_index(1)
    ctx.event_log_enable(1 << 22)
    obs = ctx.bulk_step(dt, n_steps, 6)
    got = download_ensemble(ctx)
    ref = ens.copy()
    res = m.bulk_steps(ref, box, fdir, field, dt, n_steps, po.rng_philox(seed, base), first_step=1, log_events=True)
    assert_state_close(got, ref, box, STATE_RTOL, "mos2/philox")
    ev, n_ev = ctx.event_log_read(1 << 22)
    assert n_ev == len(res["events"]) and n_ev > 10000
    dev = ev[np.lexsort((ev[:, 3], ev[:, 2], ev[:, 1], ev[:, 0]))]
    dev[:, 1] -= base
    cpu = res["events"]
    cpu = cpu[np.lexsort((cpu[:, 3], cpu[:, 2], cpu[:, 1], cpu[:, 0]))]
    assert np.array_equal(dev, cpu)
    assert (ref.valley == 1).sum() > 100  # the Q valleys fill up at this field
    assert np.array_equal(obs[:, :, 2], res["obs"][:, :, 2])
    assert np.allclose(obs[:, :, 0], res["obs"][:, :, 0], rtol=1e-11)
    assert np.max(np.abs(obs[:, :, 1] - res["obs"][:, :, 1])) <= 1e-11 * np.abs(res["obs"][:, :, 1]).max()
