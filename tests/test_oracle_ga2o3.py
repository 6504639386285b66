"""CPU: the polar-optical (Froehlich) family and the phonon bath of the oracle (oracle/emc_oracle.c) against the
UNMODIFIED reference (tests/golden/ga2o3_*.npz, recorded by oracle/_ref/ref_ga2o3_driver from
examples/hotPhononGa2O3/Ga2O3Functions.hpp driven like hotPhononGa2O3.cpp:152-365): rate tables, every step of the
moveParticles -> observables -> screening update -> bath update -> table rebuild loop, event counters per |q| bin,
occupations and the final ensemble -- bit for bit, consuming the reference's own draw sequence."""
import numpy as np
import pytest

from helpers import load_golden
from oracle import pyoracle as po
from scenarios import GA2O3, GA2O3_CASES, build_ga2o3

CASES = list(GA2O3_CASES)
KB, Q = 1.38066e-23, 1.60219e-19


def ens3(g, p):
    e = po.Ensemble.from_arrays(g[p + "k"], g[p + "pos"], g[p + "energy"], g[p + "tau"], g[p + "graintau"], g[p + "idx"])
    return e


def run_reference_loop(g, m, baths, a, rng, on_step=None):
    """the loop of runOneField(): returns the ensemble after a['steps'] steps"""
    ens = ens3(g, "init_")
    box = [a["box"]] * 3
    for s in range(a["steps"]):
        res = m.bulk_steps(ens, box, [-1, 0, 0], a["field"], a["dt"], 1, rng, first_step=s + 1)
        e_mean = res["obs"][0, 0, 0] / res["obs"][0, 0, 2]
        stale = False
        if a["screening"]:
            te = 2.0 * e_mean * Q / (3.0 * KB)
            qs2 = po.plasmon_qs2(a["doping"], te, GA2O3["eps_lo"])
            m.set_qs2(qs2)
            for b in baths:
                b.set_qs2(qs2)
            stale = True
        counts = [(b.n_em, b.n_abs) for b in baths]
        for b in baths:
            b.update(a["dt"])
            stale = True
        if stale and (s + 1) % a["reinit_every"] == 0:
            m.build_tables()
        if on_step:
            on_step(s, res, counts)
    return ens


@pytest.mark.parametrize("case", CASES)
def test_rate_tables_equal_the_reference(case):
    g = load_golden(case)
    m, baths, a = build_ga2o3(case)
    ts = m.tablesets()
    assert len(ts) == 1
    assert np.array_equal(ts[0]["cum"], g["init_cum_v0_r0"]) and ts[0]["tau"] == g["init_tau_v0_r0"][0]


@pytest.mark.parametrize("case", CASES)
def test_hot_phonon_loop_bit_for_bit(case):
    g = load_golden(case)
    m, baths, a = build_ga2o3(case)
    mt = po.mt_state(a["seed"])
    for _ in range(int(g["draws_init_count"][0])):
        po.lib().orc_mt_next(mt)
    hot = len(baths) > 0

    def check(s, res, counts):
        e_mean = res["obs"][0, 0, 0] / res["obs"][0, 0, 2]
        v_mean = res["obs"][0, 0, 1] / res["obs"][0, 0, 2]
        assert (e_mean, v_mean) == (g["obs"][s, 0], g["obs"][s, 1]), f"step {s}: observables"
        assert m.tau(0, 0) == g["tau_series"][s], f"step {s}: tau after the table rebuild"
        if hot:
            for i, b in enumerate(baths):
                assert np.array_equal(counts[i][0], g["bath_counts"][s, i, 0]), f"step {s}: emission counters"
                assert np.array_equal(counts[i][1], g["bath_counts"][s, i, 1]), f"step {s}: absorption counters"
                assert b.mean_nq() == g["mean_nq"][s, i], f"step {s}: <N_q>"
            assert baths[0].acoustic_temp() == g["t_acoustic"][s]

    ens = run_reference_loop(g, m, baths, a, po.rng_mt(mt), check)
    ref = ens3(g, "final_")
    assert ens.n == ref.n
    for f in ("kx", "ky", "kz", "energy", "tau", "x", "y", "z", "valley", "sub", "region"):
        assert np.array_equal(getattr(ens, f)[: ens.n], getattr(ref, f)[: ref.n]), f
    ts = m.tablesets()
    assert np.array_equal(ts[0]["cum"], g["final_cum_v0_r0"])
    if hot:
        for i, b in enumerate(baths):
            assert np.array_equal(b.nq, g["final_nq"][i])
            assert np.array_equal(b.cum_w, g["final_cumw"][i]) and np.array_equal(b.cum_wn, g["final_cumwn"][i])
        assert g["bath_counts"].sum() > 100  # the counters were exercised
    # the oracle consumed exactly as many draws as the reference
    nxt = po.lib().orc_mt_next(mt)
    ref_mt = po.mt_state(a["seed"])
    for _ in range(len(g["draws"])):
        po.lib().orc_mt_next(ref_mt)
    assert nxt == po.lib().orc_mt_next(ref_mt)
