"""CPU: TWO species on shared phonon baths (examples/hotCarrierMHP): the oracle against the UNMODIFIED
reference classes as hotCarrierMHP.cpp drives them (tests/golden/mhp_*.npz, recorded by oracle/_ref/ref_mhp_driver) --
emcElectron + emcHole with a mono-energetic start, screened / unscreened hot-phonon Froehlich mechanisms feeding ONE bath,
Debye screening from the live density and temperature of both species, table rebuild of both species every step.  Rate
tables, per-step observables of both species, event counters per |q| bin, occupations and both final ensembles -- bit for
bit, consuming the reference's own draw sequence."""
import numpy as np
import pytest

from helpers import load_golden
from oracle import pyoracle as po
from scenarios import MHP, MHP_CASES, build_mhp

CASES = list(MHP_CASES)
KB, Q, EPS0 = 1.38066e-23, 1.60219e-19, 8.85419e-12


def ens_of(g, p):
    return po.Ensemble.from_arrays(g[p + "k"], g[p + "pos"], g[p + "energy"], g[p + "tau"], g[p + "graintau"], g[p + "idx"])


def screening_qs2(a, sizes, mean_energies):
    """hotCarrierMHP.cpp:563-582: q_s^2 = sum over the species of n q^2 / (eps eps0 kB T), T = 2 <E> q / (3 kB)"""
    v_sim = a["box"] * a["box"] * a["box"]
    qs2 = 0.0
    for n, e in zip(sizes, mean_energies):
        if n == 0:
            continue
        n_s = float(n) / v_sim
        t_s = 2.0 * e * Q / (3.0 * KB)
        if t_s <= 0.0:
            continue
        qs2 += n_s * Q * Q / (MHP["eps_hi"] * EPS0 * KB * t_s)
    return qs2


def mean_energy(model, ens, box):
    obs = model.bulk_observables(ens, [0.0, 0.0, 0.0])
    return obs[0, 0] / obs[0, 2]


def apply_screening(models, baths, qs2):
    for m in models:
        m.set_qs2(qs2)
    for b in baths:
        b.set_qs2(qs2)


def run_reference_loop(g, models, baths, a, rng, on_step=None, perturb=None):
    """the loop of hotCarrierMHP.cpp:655-705 without the pairwise / host-side steps"""
    ens = [ens_of(g, "init_e_"), ens_of(g, "init_h_")]
    if perturb:
        for e in ens:
            perturb(e)
    box = [a["box"]] * 3
    hot = len(baths) > 0
    if a["screening"]:
        apply_screening(models, baths, screening_qs2(a, [e.n for e in ens], [mean_energy(m, e, box) for m, e in zip(models, ens)]))
        for m in models:
            m.build_tables()
    for s in range(a["steps"]):
        res = [m.bulk_steps(e, box, [0.0, 0.0, 0.0], 0.0, a["dt"], 1, rng, first_step=s + 1, charge=c)
               for m, e, c in zip(models, ens, (-Q, +Q))]
        counts = [(b.n_em, b.n_abs) for b in baths]
        for b in baths:
            b.update(a["dt"])
        if a["screening"]:
            means = [r["obs"][0, 0, 0] / r["obs"][0, 0, 2] for r in res]
            apply_screening(models, baths, screening_qs2(a, [e.n for e in ens], means))
        if hot or a["screening"]:
            for m in models:
                m.build_tables()
        if on_step:
            on_step(s, res, counts)
    return ens


@pytest.mark.parametrize("case", CASES)
def test_rate_tables_of_both_species_equal_the_reference(case):
    g = load_golden(case)
    models, baths, a = build_mhp(case)
    for m, p in zip(models, ("init_e_", "init_h_")):
        ts = m.tablesets()
        assert len(ts) == 1
        assert np.array_equal(ts[0]["cum"], g[p + "tab_cum"]) and ts[0]["tau"] == g[p + "tab_tau"][0]


@pytest.mark.parametrize("case", CASES)
def test_two_species_loop_bit_for_bit(case):
    g = load_golden(case)
    models, baths, a = build_mhp(case)
    mt = po.mt_state(a["seed"])
    # both initial ensembles from the reference's draw sequence: electrons first, then holes (map order)
    used_total = 0
    for m, p in zip(models, ("init_e_", "init_h_")):
        ens, used = m.generate_initial([a["box"]] * 3, [10, 10, 10], a["density"], mt)
        ref = ens_of(g, p)
        assert ens.n == ref.n
        for f in ("kx", "ky", "kz", "energy", "tau", "x", "y", "z", "valley", "sub", "region"):
            assert np.array_equal(getattr(ens, f)[: ens.n], getattr(ref, f)[: ref.n]), (p, f)
        used_total += used
    assert used_total == int(g["draws_init_count"][0])
    hot = len(baths) > 0

    def check(s, res, counts):
        for p in range(2):
            o = res[p]["obs"][0, 0]
            assert (o[0] / o[2], o[1] / o[2]) == (g["obs"][s, p, 0], g["obs"][s, p, 1]), f"step {s}, species {p}: observables"
            assert models[p].tau(0, 0) == g["tau_series"][s, p], f"step {s}, species {p}: tau after the table rebuild"
        if hot:
            assert np.array_equal(counts[0][0], g["bath_counts"][s, 0, 0]), f"step {s}: emission counters (both species)"
            assert np.array_equal(counts[0][1], g["bath_counts"][s, 0, 1]), f"step {s}: absorption counters"
            assert baths[0].mean_nq() == g["mean_nq"][s, 0], f"step {s}: <N_q>"

    ens = run_reference_loop(g, models, baths, a, po.rng_mt(mt), check)
    for e, p, m in zip(ens, ("final_e_", "final_h_"), models):
        ref = ens_of(g, p)
        assert e.n == ref.n
        for f in ("kx", "ky", "kz", "energy", "tau", "x", "y", "z", "valley", "sub", "region"):
            assert np.array_equal(getattr(e, f)[: e.n], getattr(ref, f)[: ref.n]), (p, f)
        assert np.array_equal(m.tablesets()[0]["cum"], g[p + "tab_cum"])
    if hot:
        assert np.array_equal(baths[0].nq, g["final_nq"][0])
        assert g["bath_counts"].sum() > 1000  # the shared counters were exercised by both species
    # the oracle consumed exactly as many draws as the reference
    nxt = po.lib().orc_mt_next(mt)
    ref_mt = po.mt_state(a["seed"])
    for _ in range(int(g["draws_count"][0])):
        po.lib().orc_mt_next(ref_mt)
    assert nxt == po.lib().orc_mt_next(ref_mt)


def test_limit_angle_samples_are_made_of_rounding_in_the_reference_algorithm():
    """Why tests/test_mhp_gpu.py compares a FEW particles of the q-resolved case at 1e-6 instead of 1e-12.  When the |q| sample
    of emcPhononBath::sampleQ lies on a kinematic limit q = |kI - kF| (forward scattering), cos(theta) = (kI^2 + kF^2 - q^2) /
    (2 kI kF) is 1 - O(eps): sin(theta) = O(sqrt(eps)) ~ 1e-8 is made of the rounding of kI and kF.  Shown on the reference's
    algorithm itself: the same loop, same draws, started from a state ONE ULP away in k_x.  Particles that never met a limit
    sample stay within 1e-12 of the unperturbed run; among the ones that did, the direction moves by > 1e-10 -- while |k| and
    the energy of every particle stay within 1e-12."""
    case = "mhp_qres"
    g = load_golden(case)
    a = dict(build_mhp(case)[2], steps=12)
    finals = []
    for nudge in (False, True):
        models, baths, _ = build_mhp(case)
        mt = po.mt_state(a["seed"])
        for _ in range(int(g["draws_init_count"][0])):
            po.lib().orc_mt_next(mt)

        def perturb(e):
            e.kx[: e.n] = np.nextafter(e.kx[: e.n], np.inf)

        flags = [po.set_limit_flags(4096), None]  # the electron loop and the hole loop index their own ensembles: mark either
        try:
            ens = run_reference_loop(g, models, baths, a, po.rng_mt(mt), perturb=perturb if nudge else None)
        finally:
            marked = flags[0].copy()
            po.set_limit_flags(0)
        finals.append((ens, marked))
    (base, marked), (moved, marked2) = finals
    assert np.array_equal(marked, marked2) and marked.sum() > 0
    worst_marked = 0.0
    for b, m in zip(base, moved):
        n = b.n
        k_b, k_m = np.stack([b.kx[:n], b.ky[:n], b.kz[:n]]), np.stack([m.kx[:n], m.ky[:n], m.kz[:n]])
        norm_b, norm_m = np.sqrt((k_b * k_b).sum(0)), np.sqrt((k_m * k_m).sum(0))
        assert np.max(np.abs(norm_m / norm_b - 1)) < 1e-12 and np.max(np.abs(m.energy[:n] / b.energy[:n] - 1)) < 1e-12
        dev = np.sqrt(((k_m - k_b) ** 2).sum(0)) / norm_b
        clean = marked[:n] == 0
        assert dev[clean].max() < 1e-12, "a particle that met no limit sample moved"
        worst_marked = max(worst_marked, float(dev[~clean].max()) if (~clean).any() else 0.0)
        assert dev.max() < 1e-6
    assert worst_marked > 1e-10, worst_marked
