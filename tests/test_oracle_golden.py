"""CPU tests: the oracle (CPU restatement) against the reference's own outputs and
known-answer tests.  The golden archives were produced by the UNMODIFIED reference
(oracle/make_golden.py -> oracle/_ref/ref_bulk_driver); every comparison here is
bit-for-bit, which is what pins the oracle."""
import ctypes as C

import numpy as np
import pytest

from helpers import field_dir_of, golden_ensemble, load_golden
from oracle import pyoracle as po
from scenarios import GOLDEN_CASES, build_model, build_si

CASES = list(GOLDEN_CASES)


def test_mt19937_64_known_answer():
    # C++ standard [rand.predef]: 10000th invocation of a default-constructed mt19937_64
    out = po.mt_fill(5489, 10000)
    assert int(out[-1]) == 9981545732273789042


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    L = po.lib()
    L.orc_philox4x32.argtypes = [C.POINTER(C.c_uint32)] * 3
    kats = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
            ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
            ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
             (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kats:
        c = (C.c_uint32 * 4)(*ctr)
        k = (C.c_uint32 * 2)(*key)
        o = (C.c_uint32 * 4)()
        L.orc_philox4x32(c, k, o)
        assert tuple(o) == want


def test_uniform_mapping_matches_libstdcxx():
    # the first recorded draws of the reference feed initParticlePos: u*0.5*h etc.
    assert po.lib().orc_uniform(0, 0.0, 1.0) == 0.0
    assert po.lib().orc_uniform(2 ** 64 - 1, 0.0, 1.0) == np.nextafter(1.0, 0.0)
    assert po.lib().orc_uniform(2 ** 63, 1e-6, 1.0) == 0.5 * (1.0 - 1e-6) + 1e-6


@pytest.mark.parametrize("case", CASES)
def test_valley_constants_bitwise(case):
    g = load_golden(case)
    m = build_model(case)
    for v, val in enumerate(m.valleys()):
        mine = np.array([val.mCond, val.mDos, val.alpha, val.eBottom, *val.vogt])
        assert np.array_equal(mine, g["valley_consts"][v])
        assert val.deg == g["valley_deg"][v]
        rot = np.array([list(val.rot[s]) for s in range(val.deg)])
        assert np.array_equal(rot, g["valley_rot"][v][: val.deg])


def test_si_known_constants():
    # m_c ~ 0.26559 me, m_DOS ~ 0.32769 me, vogt ~ (0.53846, 1.16406, 1.16406); tau = 8.73807e-15 s
    m = build_si()
    v = m.valley(0)
    assert abs(v.mCond / po.ME - 0.26559) < 1e-5
    assert abs(v.mDos / po.ME - 0.32769) < 1e-5
    assert np.allclose(list(v.vogt), [0.53846, 1.16406, 1.16406], atol=1e-5)
    assert f"{m.tau(0, 0):.5e}" == "8.73807e-15"
    # un-normalised acoustic rates printed by the reference (6 significant digits)
    raw = m.raw_rates()
    assert f"{raw[0, 0]:.5e}" == "3.40429e+11"
    assert f"{raw[0, 99]:.4e}" == "3.8324e+12"
    assert f"{raw[0, 999]:.5e}" == "2.63366e+13"


@pytest.mark.parametrize("case", CASES)
def test_rates_and_tables_bitwise(case):
    g = load_golden(case)
    m = build_model(case)
    assert np.array_equal(m.raw_rates(), g["raw_rates"])
    sets = m.tablesets()
    assert len(sets) == sum(1 for k in g.files if k.startswith("cum_"))
    for ts in sets:
        key = f"_v{ts['valley']}_r{ts['region']}"
        assert np.array_equal(ts["cum"], g["cum" + key])
        assert ts["tau"] == g["tau" + key][0]
        assert [x.globalId for x in ts["mech"]] == list(g["mech" + key])


@pytest.mark.parametrize("case", CASES)
def test_initial_ensemble_and_full_run_bitwise(case):
    """generateInitialParticles + N x (moveParticles; observables), one mt19937_64 stream:
    initial state, every raw draw, every scatter event, the final state and every
    per-step observable equal the reference's bit for bit."""
    g = load_golden(case)
    a = GOLDEN_CASES[case]["args"]
    m = build_model(case)
    st = po.mt_state(int(a["seed"]))
    ens, used = m.generate_initial([a["box"]] * 3, [a["cells"]] * 3, a["doping"], st)
    assert used == int(g["draws_init_count"][0])
    ref0 = golden_ensemble(g, "init_")
    for f in po.Ensemble.F64 + po.Ensemble.I32:
        assert np.array_equal(getattr(ens, f), getattr(ref0, f)), f
    assert np.array_equal(po.mt_fill(int(a["seed"]), len(g["draws"])), g["draws"])

    res = m.bulk_steps(ens, [a["box"]] * 3, field_dir_of(a), a["field"], a["dt"], a["steps"], po.rng_mt(st),
                       first_step=1, record=True, log_events=True)
    assert used + res["n_draws"] == len(g["draws"])
    ref1 = golden_ensemble(g, "final_")
    for f in po.Ensemble.F64 + po.Ensemble.I32:
        assert np.array_equal(getattr(ens, f), getattr(ref1, f)), f
    ev = res["events"]
    real = ev[ev[:, 2] >= 0]
    assert np.array_equal(real[:, [0, 1, 3]], g["events"])
    obs = res["obs"]
    cnt = obs[:, :, 2]
    with np.errstate(invalid="ignore", divide="ignore"):
        avg_e = np.where(cnt > 0, obs[:, :, 0] / cnt, 0.0)
        avg_v = np.where(cnt > 0, obs[:, :, 1] / cnt, 0.0)
    assert np.array_equal(avg_e, g["obs"][1:, 0, :])
    assert np.array_equal(avg_v, g["obs"][1:, 1, :])
    assert np.array_equal(cnt / ens.n, g["obs"][1:, 2, :])


@pytest.mark.parametrize("case", CASES)
def test_replay_streams_reproduce_global_stream(case):
    """Splitting the reference's draw log into per-particle streams (the form the GPU
    replay mode consumes) reproduces the same run."""
    g = load_golden(case)
    a = GOLDEN_CASES[case]["args"]
    m = build_model(case)
    st = po.mt_state(int(a["seed"]))
    ens, used = m.generate_initial([a["box"]] * 3, [a["cells"]] * 3, a["doping"], st)
    e1 = ens.copy()
    res = m.bulk_steps(e1, [a["box"]] * 3, field_dir_of(a), a["field"], a["dt"], a["steps"], po.rng_mt(st),
                       first_step=1, record=True)
    draws, offsets = po.streams_from_record(g["draws"][used:], res["rec_pid"], ens.n)
    e2 = ens.copy()
    cursor = np.zeros(ens.n, dtype=np.int64)
    m.bulk_steps(e2, [a["box"]] * 3, field_dir_of(a), a["field"], a["dt"], a["steps"],
                 po.rng_streams(draws, offsets, cursor), first_step=1)
    for f in po.Ensemble.F64[:5] + po.Ensemble.F64[6:] + po.Ensemble.I32:
        assert np.array_equal(getattr(e1, f), getattr(e2, f)), f
    assert np.array_equal(cursor, np.diff(offsets))


# ---- the reference's own known-answer tests, restated against the oracle ----------------------

HBAR = 1.05459e-34


def _expected(force, eff_mass, dT):
    """tests/testParticleMovement/testParticleMovement.cpp:46-64 (calcExpectedValues).
    NB: the reference accumulates 1/m with an *int* initial value 0 (std::accumulate(..., 0, ...)),
    i.e. it truncates every partial sum; reproduced here because the expected values depend on it."""
    acc = 0
    for mm in eff_mass:
        acc = int(acc + 1.0 / mm)
    mass_cond = 3.0 / acc
    exp_k = [f * dT * np.sqrt(mass_cond / mm) / HBAR for f, mm in zip(force, eff_mass)]
    exp_pos = [k / 2.0 * dT / np.sqrt(mm * mass_cond) * HBAR for k, mm in zip(exp_k, eff_mass)]
    return exp_k, exp_pos


def _to_cs(dirs, vec):
    return [float(np.dot(d, vec) / np.linalg.norm(d)) for d in dirs]


def _from_cs(dirs, vec):
    return [sum(dirs[i][c] * vec[i] / np.linalg.norm(dirs[i]) for i in range(3)) for c in range(3)]


KAT_DIRS = [[[1, 0, 0], [0, 1, 0], [0, 0, 1]], [[0, 1, 0], [1, 0, 0], [0, 0, 1]], [[0, 0, 1], [0, 1, 0], [1, 0, 0]],
            [[-1, 1, 1], [1, 1, 0], [1, -1, 2]], [[1, 1, 1], [-1, 1, 0], [-1, -1, 2]],
            [[-1, -1, 1], [1, 0, 1], [-1, 2, 1]]]


def test_reference_kat_particle_movement_isotropic():
    """tests/testParticleMovement/testParticleMovement.cpp:109-143"""
    dT = 1e-12
    m = po.Model()
    m.add_valley(po.VALLEY_NONPARABOLIC_ISO, 1.0, 6, 0.5)
    v = m.valley(0)
    force = np.array([1e-19, 5e-19, 1e-20])
    exp_k, exp_pos = _expected(force, [1, 1, 1], dT)
    for sub in range(6):
        k = np.zeros(3)
        pos = np.zeros(3)
        e = C.c_double(0)
        po.lib().orc_drift(C.byref(v), dT, po._dp(k), C.byref(e), sub, po._dp(pos), 3, po._dp(force))
        assert np.all(np.abs(pos - exp_pos) < 1e-9)
        assert np.all(np.abs(k - exp_k) < 1e-9)


def test_reference_kat_particle_movement_anisotropic():
    """tests/testParticleMovement/testParticleMovement.cpp:150-217"""
    dT = 1e-12
    eff = [0.1, 0.5, 1]
    m = po.Model()
    m.add_valley(po.VALLEY_NONPARABOLIC_ANISO, eff, 6, 0.5, 0.0, KAT_DIRS)
    v = m.valley(0)
    force = np.array([1e-17, 1e-17, 1e-17])
    for sub in range(6):
        dirs = np.array(KAT_DIRS[sub], dtype=float)
        k = np.zeros(3)
        pos = np.full(3, 1e-7)
        e = C.c_double(0)
        po.lib().orc_drift(C.byref(v), dT, po._dp(k), C.byref(e), sub, po._dp(pos), 3, po._dp(force))
        # the reference's expectation uses the harmonic mass with its int-truncating accumulate;
        # with masses (0.1, 0.5, 1) 1/m sums to exactly 13 so no truncation happens
        exp_k, exp_pos = _expected(_to_cs(dirs, force), eff, dT)
        exp_k = _from_cs(dirs, exp_k)
        exp_pos = np.array(_from_cs(dirs, exp_pos)) + 1e-7
        kk = np.array(exp_k)
        exp_e = po.lib().orc_energy(C.byref(v), po._dp(kk))
        assert np.all(np.abs(k - exp_k) < 1e-5)
        assert np.all(np.abs(pos - exp_pos) < 1e-10)
        assert abs(e.value - exp_e) < 1e-10


def test_reference_kat_valley_coordinate_transformation():
    """tests/testValleyCoordinateTransformation/testValleyCoordinateTransformation.cpp:10-84"""
    x_dirs = [[[1, 0, 0], [0, 1, 0], [0, 0, 1]], [[-1, 0, 0], [0, 1, 0], [0, 0, 1]], [[0, 1, 0], [1, 0, 0], [0, 0, 1]],
              [[0, -1, 0], [1, 0, 0], [0, 0, 1]], [[0, 0, 1], [0, 1, 0], [1, 0, 0]], [[0, 0, -1], [0, 1, 0], [1, 0, 0]]]
    l_dirs = [[[1, 1, 1], [-1, 1, 0], [-1, -1, 2]], [[-1, -1, -1], [-1, 1, 0], [-1, -1, 2]],
              [[-1, 1, 1], [1, 1, 0], [1, -1, 2]], [[1, -1, 1], [1, 1, 0], [-1, 1, 2]],
              [[1, 1, -1], [1, 0, 1], [-1, 2, 1]], [[-1, -1, 1], [1, 0, 1], [-1, 2, 1]],
              [[-1, 1, -1], [1, 1, 0], [-1, 1, 2]], [[-1, -1, 1], [1, 0, 1], [-1, 2, 1]]]
    for masses, alpha, dirs in (([0.9, 0.1, 0.1], 0.5, x_dirs), ([0.8, 0.2, 0.2], 0.3, l_dirs)):
        m = po.Model()
        m.add_valley(po.VALLEY_NONPARABOLIC_ANISO, masses, len(dirs), alpha, 0.0, dirs)
        v = m.valley(0)
        vec = np.array([1.0, 2.0, 3.0])
        for s, d in enumerate(dirs):
            ve = np.zeros(3)
            vd = np.zeros(3)
            po.lib().orc_to_ellipse(C.byref(v), s, po._dp(vec), po._dp(ve))
            po.lib().orc_to_device(C.byref(v), s, po._dp(ve), po._dp(vd))
            for i in range(3):
                dn = np.array(d[i], dtype=float) / np.linalg.norm(d[i])
                assert abs(ve[i] - float(np.dot(vec, dn))) < 1e-9
            assert np.all(np.abs(vd - vec) < 1e-9)
