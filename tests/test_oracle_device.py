"""CPU: the device-run part of the oracle (oracle/emc_oracle.c, orc_device_* / orc_sor / orc_efield / ...)
against the UNMODIFIED reference (tests/golden/device_*.npz, recorded by oracle/_ref/ref_device_driver from
emcSORSolver, calcEFieldAtGridPts, emcNGPScheme, emcSimulationResults, emcBasicParticleHandler).
Everything is compared bit for bit, step by step."""
import numpy as np
import pytest

from helpers import load_golden
from oracle import pyoracle as po
from scenarios import DEVICE_CASES, DEVICE_LONG_CASES, build_device

CASES = list(DEVICE_CASES)


def ens_from(g, p):
    pos = g[p + "pos"]
    if pos.shape[1] == 2:
        pos = np.concatenate([pos, np.zeros((len(g[p + "energy"]), 1))], axis=1)
    return po.Ensemble.from_arrays(g[p + "k"], pos, g[p + "energy"], g[p + "tau"], g[p + "label"], g[p + "idx"])


def assert_same_ensemble(a, b, what):
    assert a.n == b.n, what
    for f in ("kx", "ky", "kz", "energy", "tau", "x", "y", "z", "valley", "sub", "region"):
        assert np.array_equal(getattr(a, f)[: a.n], getattr(b, f)[: b.n]), f"{what}: {f}"


@pytest.mark.parametrize("case", CASES)
def test_flattened_device_description_equals_the_reference(case):
    g = load_golden(case)
    m, dev = build_device(case)
    assert dev.extent == list(g["region"].shape[::-1])
    assert np.array_equal(dev.region, g["region"].ravel())
    assert np.array_equal(dev.doping / dev.ni, g["doping_norm"].ravel())
    assert np.array_equal(dev.face_contact, g["face_contact"].reshape(-1, 2 * dev.dim))
    vt, debye, vol, ni = g["device_consts"][:4]
    assert (dev.vt, dev.debye, dev.cell_volume, dev.ni) == (vt, debye, vol, ni)
    d = dev.c()
    is_ohmic = np.array([dev.L.orc_dev_is_ohmic(po.C.byref(d), i) for i in range(dev.cells)])
    is_res = np.array([dev.L.orc_dev_is_reservoir(po.C.byref(d), i) for i in range(dev.cells)])
    cidx = np.array([dev.L.orc_dev_contact_idx(po.C.byref(d), i) for i in range(dev.cells)])
    assert np.array_equal(is_ohmic, g["is_ohmic"].ravel())
    assert np.array_equal(is_res, g["is_reservoir"].ravel())
    assert np.array_equal(cidx, g["contact_idx"].ravel())
    assert np.array_equal(dev.expected_at_contact(), g["expected_at_contact"].ravel())
    for ts in m.tablesets():
        key = f"_v{ts['valley']}_r{ts['region']}"
        assert np.array_equal(ts["cum"], g["cum" + key]) and ts["tau"] == g["tau" + key][0]


@pytest.mark.parametrize("case", CASES)
def test_equilibrium_chain_bit_for_bit(case):
    """calcEquilibriumCharacteristics (emcSimulation.hpp:139-146): potential guess, SOR, E field, initial
    particles, NGP assignment, concentration."""
    g = load_golden(case)
    a = DEVICE_CASES[case]
    m, dev = build_device(case)
    pot = dev.initial_potential()
    assert np.array_equal(pot, g["pot_guess"].ravel())
    dev.sor(pot, None, 1e-4, 1.8, True)
    assert np.array_equal(pot, g["pot_eq"].ravel())
    e = dev.efield(pot)
    assert np.array_equal(e[0], g["ex_eq"].ravel()) and np.array_equal(e[1], g["ey_eq"].ravel())
    assert dev.dim == 2 or np.array_equal(e[2], g["ez_eq"].ravel())
    mt = po.mt_state(a["seed"])
    # electronVWD always starts from the equilibrium potential; emcElectron(usePotentialForInit = false) from the doping
    ens = dev.generate_initial(m, mt, pot=pot if a.get("electron") == "vwd" else None)
    ref = ens_from(g, "init_")
    assert_same_ensemble(ens, ref, "initial ensemble")
    count = dev.assign(ens)
    assert np.array_equal(count, g["count_eq"].ravel()) and abs(count.sum() - ens.n) <= 1e-9 * ens.n
    assert np.array_equal(dev.concentration(count), g["conc_eq"].ravel())


@pytest.mark.parametrize("case", CASES)
def test_emc_steps_bit_for_bit(case):
    """performEMCStep (emcSimulation.hpp:177-192) step by step, consuming the reference's own draw sequence."""
    g = load_golden(case)
    a = DEVICE_CASES[case]
    m, dev = build_device(case)
    draws = g["draws"]
    marks = g["draw_marks"].reshape(-1, 3)
    expected = dev.expected_at_contact()
    pot = g["pot_eq"].ravel().copy()
    conc = g["conc_eq"].ravel().copy()
    # one global mt19937_64 stream like the reference's single-thread rngs[0]: fast-forward past the creation draws
    mt = po.mt_state(a["seed"])
    for _ in range(int(g["draws_init_count"][0])):
        po.lib().orc_mt_next(mt)
    ens = ens_from(g, "init_")
    cap = ens.n + 2000
    big = po.Ensemble(cap)
    for f in po.Ensemble.F64 + po.Ensemble.I32:
        getattr(big, f)[: ens.n] = getattr(ens, f)
    big.n = ens.n
    ens = big
    for s in range(a["steps"]):
        p = f"s{s}_"
        dev.sor(pot, conc, 1e-4, 1.8, s == 0)
        assert np.array_equal(pot, g[p + "pot"].ravel()), f"step {s}: potential"
        e = dev.efield(pot)
        assert np.array_equal(e[0], g[p + "ex"].ravel()) and np.array_equal(e[1], g[p + "ey"].ravel())
        assert dev.dim == 2 or np.array_equal(e[2], g[p + "ez"].ravel())
        assert_same_ensemble(ens, ens_from(g, p + "pre_"), f"step {s}: pre")
        assert int(marks[s, 0]) == int(g["draws_init_count"][0]) + sum(
            int(marks[j, 2] - marks[j, 0]) for j in range(s))
        res = dev.step(m, ens, e, a["dt"], po.rng_mt(mt), step_index=s + 1)
        assert np.array_equal(res["removed_per_contact"], g[p + "removed_per_contact"]), f"step {s}: removed"
        dev.compact(ens, res["removed"])
        assert_same_ensemble(ens, ens_from(g, p + "drift_"), f"step {s}: after drift")
        net = dev.contacts(m, ens, expected, mt)
        assert np.array_equal(net, g[p + "net_injected_per_contact"]), f"step {s}: contacts"
        assert_same_ensemble(ens, ens_from(g, p + "post_"), f"step {s}: after contacts")
        count = dev.assign(ens)
        assert np.array_equal(count, g[p + "count"].ravel())
        conc = dev.concentration(count)
        assert np.array_equal(conc, g[p + "conc"].ravel())
    # the oracle consumed exactly as many draws as the reference
    nxt = po.lib().orc_mt_next(mt)
    ref_mt = po.mt_state(a["seed"])
    for _ in range(len(draws)):
        po.lib().orc_mt_next(ref_mt)
    assert nxt == po.lib().orc_mt_next(ref_mt)


@pytest.mark.parametrize("case", list(DEVICE_LONG_CASES))
def test_long_chained_run_bit_for_bit(case):
    """250 chained time steps (Poisson -> field -> drift / scatter -> contacts -> assignment -> concentration, each step from
    the state the step before left): the per-contact counters and the ensemble size of EVERY step, and the grids and ensembles
    of every 25th step and of the last one, consuming the reference's own draw sequence."""
    g = load_golden(case)
    a = DEVICE_LONG_CASES[case]
    m, dev = build_device(case)
    expected = dev.expected_at_contact()
    pot = g["pot_eq"].ravel().copy()
    conc = g["conc_eq"].ravel().copy()
    mt = po.mt_state(a["seed"])
    for _ in range(int(g["draws_init_count"][0])):
        po.lib().orc_mt_next(mt)
    ens = ens_from(g, "init_")
    big = po.Ensemble(ens.n + 4000)
    for f in po.Ensemble.F64 + po.Ensemble.I32:
        getattr(big, f)[: ens.n] = getattr(ens, f)
    big.n = ens.n
    ens = big
    n_snaps = 0
    for s in range(a["steps"]):
        p = f"s{s}_"
        snap = (p + "pot") in g
        dev.sor(pot, conc, 1e-4, 1.8, s == 0)
        e = dev.efield(pot)
        if snap:
            assert np.array_equal(pot, g[p + "pot"].ravel()), f"step {s}: potential"
            assert np.array_equal(e[0], g[p + "ex"].ravel()) and np.array_equal(e[1], g[p + "ey"].ravel())
        res = dev.step(m, ens, e, a["dt"], po.rng_mt(mt), step_index=s + 1)
        assert np.array_equal(res["removed_per_contact"], g["removed_all"][s]), f"step {s}: removed"
        dev.compact(ens, res["removed"])
        if snap:
            assert_same_ensemble(ens, ens_from(g, p + "drift_"), f"step {s}: after drift")
        net = dev.contacts(m, ens, expected, mt)
        assert np.array_equal(net, g["net_injected_all"][s]), f"step {s}: contacts"
        assert ens.n == int(g["size_all"][s])
        count = dev.assign(ens)
        conc = dev.concentration(count)
        if snap:
            assert_same_ensemble(ens, ens_from(g, p + "post_"), f"step {s}: after contacts")
            assert np.array_equal(count, g[p + "count"].ravel()) and np.array_equal(conc, g[p + "conc"].ravel())
            n_snaps += 1
    assert n_snaps == 11 and g["removed_all"].sum() > 10 and np.abs(g["net_injected_all"]).sum() > 10
    nxt = po.lib().orc_mt_next(mt)
    ref_mt = po.mt_state(a["seed"])
    for _ in range(int(g["draws_count"][0])):
        po.lib().orc_mt_next(ref_mt)
    assert nxt == po.lib().orc_mt_next(ref_mt)
