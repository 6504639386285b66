"""GPU parity of config 5 (hot-phonon Ga2O3 bulk): the polar-optical samplers (EMCGPU_SAMPLER_FROEHLICH /
_SCREENED_FROEHLICH incl. the q-resolved angle) and the phonon-bath event counters behind the C ABI, fed the
REFERENCE's random draws (tests/golden/ga2o3_*.npz, recorded from the unmodified reference) and driven through the
reference's own loop: step -> counters -> bath update -> table rebuild -> next step.

Bars: event counters per |q| bin exact in every step (hence occupations and rebuilt tables identical to the
reference's), valley / index words exact, fp64 state within 1e-12 of the REFERENCE's final ensemble."""
import numpy as np
import pytest

from helpers import assert_state_close, download_ensemble, load_golden, upload_baths, upload_ensemble, upload_model
from oracle import pyoracle as po
from scenarios import GA2O3, GA2O3_CASES, build_ga2o3
from test_oracle_ga2o3 import KB, Q, ens3

pytestmark = pytest.mark.gpu
CASES = list(GA2O3_CASES)


@pytest.mark.parametrize("math_mode", ["exact", "fast"])
@pytest.mark.parametrize("case", CASES)
def test_hot_phonon_loop_replays_the_reference(gpu_ctx_factory, case, math_mode):
    from viennaemc_b200 import capi
    g = load_golden(case)
    m, baths, a = build_ga2o3(case)
    hot = len(baths) > 0
    box = [a["box"]] * 3
    ctx = gpu_ctx_factory()
    upload_baths(ctx, baths)
    upload_model(ctx, m)
    init = ens3(g, "init_")
    upload_ensemble(ctx, init)
    ctx.bulk_configure(box, [-1, 0, 0], a["field"], math_mode=capi.MATH_EXACT if math_mode == "exact" else capi.MATH_FAST)
    # a second, CPU copy of the run only tells which particle consumed which of the reference's draws
    shadow = init.copy()
    mt = po.mt_state(a["seed"])
    used = int(g["draws_init_count"][0])
    for _ in range(used):
        po.lib().orc_mt_next(mt)
    after = g["draw_count_after_step"]
    n_events = 0
    for s in range(a["steps"]):
        res = m.bulk_steps(shadow, box, [-1, 0, 0], a["field"], a["dt"], 1, po.rng_mt(mt), first_step=s + 1, record=True)
        step_draws = g["draws"][used:int(after[s])]
        assert len(res["rec_pid"]) == len(step_draws)
        used = int(after[s])
        sd, offsets = po.streams_from_record(step_draws, res["rec_pid"], shadow.n)
        ctx.rng_replay(sd, offsets)
        ctx.set_step_index(s + 1)
        obs = ctx.bulk_step(a["dt"], 1, 1)
        e_mean = obs[0, 0, 0] / obs[0, 0, 2]
        assert abs(e_mean / g["obs"][s, 0] - 1) < 1e-11 and abs(obs[0, 0, 1] / obs[0, 0, 2] / g["obs"][s, 1] - 1) < 1e-9
        stale = False
        if a["screening"]:
            # the reference's own mean energy: the device sum differs from it in the last bits (summation order)
            qs2 = po.plasmon_qs2(a["doping"], 2.0 * g["obs"][s, 0] * Q / (3.0 * KB), GA2O3["eps_lo"])
            assert qs2 == g["qs2"][s]
            m.set_qs2(qs2)
            for b in baths:
                b.set_qs2(qs2)
            stale = True
        if hot:
            em, ab = ctx.get_phonon_counts(reset=True)
            for i, b in enumerate(baths):
                assert np.array_equal(em[i], g["bath_counts"][s, i, 0]), f"step {s}: emission counters of bath {i}"
                assert np.array_equal(ab[i], g["bath_counts"][s, i, 1]), f"step {s}: absorption counters of bath {i}"
                n_events += int(em[i].sum() + ab[i].sum())
                # the shadow run has counted the same events into the oracle bath already
                b.update(a["dt"])
                assert b.mean_nq() == g["mean_nq"][s, i]
            stale = True
        if stale and (s + 1) % a["reinit_every"] == 0:
            m.build_tables()
            upload_baths(ctx, baths)
            upload_model(ctx, m, valleys_too=False)
    got = download_ensemble(ctx)
    assert_state_close(got, ens3(g, "final_"), box, 1e-12, case)
    if hot:
        assert n_events > 100
