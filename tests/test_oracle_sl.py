"""CPU: the single-layer (2-D material) path -- SURVEY.md 8 f4 -- of the oracle against the UNMODIFIED reference classes as
examples/singleLayerMoS2/singleLayerMoS2.cpp sets them up with its default parameter set (parameterPilotto.hpp): electron2D,
emcNonParabolicIsotropSingleLayerValley (K), emcNonParabolicAnisotropSingleLayerValley (Q, six in-plane frames),
emcAcousticSingleLayerMechanism and the 36 emcZeroOrderSingleLayerInterValley{Absorption,Emission}ScatterMechanism objects.
Recorded by oracle/_ref/ref_bulk_driver --material mos2 (tests/golden/mos2_*.npz, oracle/make_golden.py mos2).  Valley
constants, rate tables, initial ensemble, every scatter event, per-step observables and the final ensemble: bit for bit,
consuming the reference's own mt19937_64 stream."""
import numpy as np
import pytest

from helpers import field_dir_of, golden_ensemble, load_golden
from oracle import pyoracle as po
from scenarios import MOS2_CASES, MOS2_LZ, build_mos2

CASES = list(MOS2_CASES)


def box_of(a):
    return [a["box"], a["box"], MOS2_LZ]


@pytest.mark.parametrize("case", ["mos2_pilotto", "mos2_kaasbjerg", "mos2_kaasbjerg_screened", "mos2_kaasbjerg_supported",
                                  "mos2_pilotto_screened"])
def test_valley_constants_and_rate_tables_equal_the_reference(case):
    g = load_golden(case)
    m = build_mos2(case)
    for v, val in enumerate(m.valleys()):
        assert np.array_equal(np.array([val.mCond, val.mDos, val.alpha, val.eBottom, *val.vogt]), g["valley_consts"][v])
        assert val.deg == g["valley_deg"][v]
        rot = np.array([list(val.rot[s]) for s in range(val.deg)])
        if v == 1:  # the in-plane frames of the Q valleys (the isotropic classes have none)
            assert np.array_equal(rot, g["valley_rot"][v][: val.deg])
    assert np.array_equal(m.raw_rates()[:, ::25], g["raw_rates"])
    sets = m.tablesets()
    for ts in sets:
        key = f"_v{ts['valley']}_r{ts['region']}"
        assert np.array_equal(ts["cum"], g["cum" + key])
        assert ts["tau"] == g["tau" + key][0]
        assert [x.globalId for x in ts["mech"]] == list(g["mech" + key])
    assert [len(ts["mech"]) for ts in sets] == {"mos2": [15, 23], "mos2ps": [15, 23], "mos2kf": [18], "mos2kx": [24]}[MOS2_CASES[case]["material"]]


@pytest.mark.parametrize("case", CASES)
def test_initial_ensemble_and_full_run_bit_for_bit(case):
    g = load_golden(case)
    a = MOS2_CASES[case]
    m = build_mos2(case)
    st = po.mt_state(int(a["seed"]))
    ens, used = m.generate_initial(box_of(a), [a["cells"], a["cells"], 1], 1.0, st, capacity=4096)
    assert used == int(g["draws_init_count"][0])
    ref0 = golden_ensemble(g, "init_")
    assert ens.n == ref0.n == 4 * 2 * (a["cells"] + 1) ** 2
    for f in po.Ensemble.F64 + po.Ensemble.I32:
        assert np.array_equal(getattr(ens, f), getattr(ref0, f)), f
    assert np.all(ens.kz == 0) and np.all(ens.valley == 0)
    res = m.bulk_steps(ens, box_of(a), field_dir_of(a), a["field"], a["dt"], a["steps"], po.rng_mt(st), first_step=1,
                       record=True, log_events=True)
    assert used + res["n_draws"] == int(g["draws_count"][0])
    ref1 = golden_ensemble(g, "final_")
    for f in po.Ensemble.F64 + po.Ensemble.I32:
        assert np.array_equal(getattr(ens, f), getattr(ref1, f)), f
    ev = res["events"]
    real = ev[ev[:, 2] >= 0]
    assert np.array_equal(real[:, [0, 1, 3]], g["events"])
    if a["material"] in ("mos2", "mos2ps"):
        assert len(set(real[:, 3])) > 15 and (ens.valley == 1).sum() > 0  # many of the 38 mechanisms fired, Q valleys populated
        if a["material"] == "mos2ps":  # ids 2, 3: the screened K -> K Gamma-phonon pair
            assert (real[:, 3] == 2).sum() >= 20 and (real[:, 3] == 3).sum() >= 20
    else:
        fired = set(real[:, 3])  # one-valley model: ids 6-13 first order, 14-15 Froehlich, 16-17 piezoelectric
        assert len(fired) >= 12 and len(fired & set(range(6, 14))) >= 6 and len(real) > 1000
        if a["material"] == "mos2kx":  # 18 charged impurities, 19 interface roughness, 20-23 remote surface-optical phonons
            assert (real[:, 3] == 18).sum() >= 100 and (real[:, 3] == 19).sum() >= 20 and {20, 21, 22, 23} <= fired
        if a["material"] == "mos2kf":  # (screening suppresses the long-range piezoelectric terms)
            assert {14, 15} <= fired and len(fired & {16, 17}) >= (1 if "sheet-density" in a else 2)
    obs = res["obs"]
    cnt = obs[:, :, 2]
    with np.errstate(invalid="ignore", divide="ignore"):
        avg_e = np.where(cnt > 0, obs[:, :, 0] / cnt, 0.0)
        avg_v = np.where(cnt > 0, obs[:, :, 1] / cnt, 0.0)
    assert np.array_equal(avg_e, g["obs"][1:, 0, :])
    assert np.array_equal(avg_v, g["obs"][1:, 1, :])
    assert np.array_equal(cnt / ens.n, g["obs"][1:, 2, :])
