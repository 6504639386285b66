"""GPU parity tests of the device-run path through the C ABI (emcgpu_device_*),
against the golden fixtures recorded from the UNMODIFIED reference (tests/golden/device_*.npz) and against the
oracle on seeded inputs.

Bars (BASELINE.json north_star): integer / index results bit-exact (NGP counts, contact bookkeeping, removed and
injected particles, scatter events, valley / region indices); fp64 within 1e-12 relative (particle state,
potential, field, concentration) -- the device libm (exp, log, sincos, asinh) differs from glibc in the last bit.
"""
import numpy as np
import pytest

from helpers import STATE_RTOL, download_ensemble, load_golden, rel_err, upload_ensemble, upload_model
from oracle import pyoracle as po
from scenarios import DEVICE_CASES, build_device
from test_oracle_device import ens_from
from viennaemc_b200 import capi

pytestmark = pytest.mark.gpu
CASES = list(DEVICE_CASES)


def configure(ctx, dev: po.Device, expected=None, math_mode=capi.MATH_EXACT, nr_carriers=1.0):
    if expected is None and dev.electron_kind == po.ELECTRON_VWD:
        expected = dev.expected_at_contact()  # the library's default is emcElectron's rule (no rounding)
    ctx.device_configure(dev.dim, dev.extent, dev.spacing, dev.max_pos, dev.vt, dev.debye, dev.ni, dev.cell_volume,
                         dev.eps_r, dev.contact_type, dev.contact_voltage, dev.gate_eps, dev.gate_thick,
                         dev.gate_barrier, dev.region, dev.face_contact, dev.doping, expected=expected,
                         math_mode=math_mode, pm_scheme=dev.pm_scheme, nr_carriers=nr_carriers)
    for face in range(2 * dev.dim):
        if dev.surface_kind[face]:
            ctx.device_set_surface(face, dev.surface_kind[face], dev.surface_param[face])
    ctx.device_set_particle_kind(dev.electron_kind)


def assert_grid_close(got, want, what, rtol=STATE_RTOL, scale=None):
    want = np.asarray(want).ravel()
    scale = scale or float(np.abs(want).max()) or 1.0
    err = float(np.max(np.abs(got - want))) / scale
    assert err <= rtol, f"{what}: {err:.3e}"


def assert_counts(dev, got, want, what):
    """carriers per grid point: exact for NGP / NEC (every deposit is a multiple of 1/8, order-independent sums),
    1e-12 for CIC (fractional weights, the order of the atomic adds shows in the last bits)"""
    if dev.pm_scheme == po.PM_CIC:
        assert_grid_close(got, want, what, 1e-12)
    else:
        assert np.array_equal(got, np.asarray(want).ravel()), what


def assert_ensemble_close(got: po.Ensemble, want: po.Ensemble, dev, what, check_grain=False):
    assert got.n == want.n, what
    n = want.n
    for f in ("valley", "sub", "region"):
        assert np.array_equal(getattr(got, f)[:n], getattr(want, f)[:n]), f"{what}: {f}"
    kmag = float(np.sqrt(np.mean(want.kx[:n] ** 2 + want.ky[:n] ** 2 + want.kz[:n] ** 2)))
    errs = {f: rel_err(getattr(got, f)[:n], getattr(want, f)[:n], kmag) for f in ("kx", "ky", "kz")}
    errs["energy"] = rel_err(got.energy[:n], want.energy[:n])
    errs["tau"] = rel_err(got.tau[:n], want.tau[:n])
    errs["x"] = rel_err(got.x[:n], want.x[:n], dev.max_pos[0])
    errs["y"] = rel_err(got.y[:n], want.y[:n], dev.max_pos[1])
    if dev.dim > 2:
        errs["z"] = rel_err(got.z[:n], want.z[:n], dev.max_pos[2])
    if check_grain:
        errs["grainTau"] = rel_err(got.grainTau[:n], want.grainTau[:n])
    bad = {k: v for k, v in errs.items() if not v <= STATE_RTOL}
    assert not bad, f"{what}: {bad}"


@pytest.mark.parametrize("case", CASES)
def test_grid_chain_against_the_reference(gpu_ctx_factory, case):
    """initial guess -> equilibrium SOR -> E field; NGP assignment -> concentration; non-equilibrium SOR of step 0."""
    g = load_golden(case)
    m, dev = build_device(case)
    ctx = gpu_ctx_factory()
    upload_model(ctx, m)
    configure(ctx, dev)
    assert np.array_equal(ctx.device_get_grid(capi.GRID_EXPECTED), g["expected_at_contact"].ravel())
    assert_grid_close(ctx.device_get_grid(capi.GRID_POTENTIAL), g["pot_guess"], "initial guess")
    # the oracle's sweep count is the reference's (bit-identical iterates, tests/test_oracle_device.py)
    pot = dev.initial_potential()
    ref_sweeps = dev.sor(pot, None, 1e-4, 1.8, True)
    sweeps = ctx.device_poisson(True, 1e-4, 1.8, True)
    assert sweeps == ref_sweeps
    assert_grid_close(ctx.device_get_grid(capi.GRID_POTENTIAL), g["pot_eq"], "equilibrium potential")
    ctx.device_efield()
    assert_grid_close(ctx.device_get_grid(capi.GRID_EFIELD_X), g["ex_eq"], "Ex")
    assert_grid_close(ctx.device_get_grid(capi.GRID_EFIELD_Y), g["ey_eq"], "Ey")
    if dev.dim > 2:
        # the box is uniform along z: Ez is a difference of (almost) equal potentials, compared on the scale of the field
        assert_grid_close(ctx.device_get_grid(capi.GRID_EFIELD_Z), g["ez_eq"], "Ez", scale=float(np.abs(g["ex_eq"]).max()))
    # particles of the reference -> counts (exact) -> concentration
    upload_ensemble(ctx, ens_from(g, "init_"))
    ctx.device_assign()
    assert_counts(dev, ctx.device_get_grid(capi.GRID_COUNT), g["count_eq"], "counts")
    ctx.device_concentration()
    assert_grid_close(ctx.device_get_grid(capi.GRID_CONCENTRATION), g["conc_eq"], "concentration",
                      1e-12 if dev.pm_scheme == po.PM_CIC else 1e-15)
    # non-equilibrium solve of step 0 (Dirichlet values reset), from the reference's own inputs
    ctx.device_set_grid(capi.GRID_POTENTIAL, g["pot_eq"])
    ctx.device_set_grid(capi.GRID_CONCENTRATION, g["conc_eq"])
    pot = g["pot_eq"].ravel().copy()
    ref_sweeps = dev.sor(pot, g["conc_eq"].ravel().copy(), 1e-4, 1.8, True)
    assert ctx.device_poisson(False, 1e-4, 1.8, True) == ref_sweeps
    assert_grid_close(ctx.device_get_grid(capi.GRID_POTENTIAL), g["s0_pot"], "non-equilibrium potential")


def _step_streams(g, a, m, dev, s):
    """Per-particle replay streams of the reference's drift/scatter phase of step s.  The attribution of each raw draw
    comes from the oracle run that reproduces the reference bit for bit (tests/test_oracle_device.py)."""
    p = f"s{s}_"
    marks = g["draw_marks"].reshape(-1, 3)
    ens = ens_from(g, p + "pre_")
    e = np.stack([g[p + "e" + ax].ravel() for ax in "xyz"[: dev.dim]])
    draws = g["draws"][int(marks[s, 0]):int(marks[s, 1])]
    mt = np.zeros(0)
    # replay through the oracle with a flat stream to learn who consumed what
    flat = po.rng_streams(np.ascontiguousarray(draws), np.zeros(ens.n + 1, dtype=np.int64), np.zeros(ens.n, dtype=np.int64))
    # (a flat stream shared by all particles = offsets all zero is not what STREAMS means; use the recorder instead)
    st = po.mt_state(a["seed"])
    for _ in range(int(marks[s, 0])):
        po.lib().orc_mt_next(st)
    work = ens.copy()
    res = dev.step(m, work, e, a["dt"], po.rng_mt(st), step_index=s + 1, record=True, log_events=True)
    assert len(res["rec_pid"]) == len(draws)
    sd, offsets = po.streams_from_record(draws, res["rec_pid"], ens.n)
    return ens, e, sd, offsets, work, res


@pytest.mark.parametrize("math_mode", [capi.MATH_EXACT, capi.MATH_FAST], ids=["exact", "fast"])
@pytest.mark.parametrize("case", CASES)
def test_particle_step_replays_the_reference(gpu_ctx_factory, case, math_mode):
    """driftScatterParticles of every recorded step, fed the reference's own draws: removed particles, per-contact
    counts, scatter events and indices exact; state within 1e-12 of the REFERENCE's ensemble."""
    g = load_golden(case)
    a = DEVICE_CASES[case]
    m, dev = build_device(case)
    ctx = gpu_ctx_factory()
    upload_model(ctx, m)
    configure(ctx, dev, math_mode=math_mode)
    n_events = 0
    for s in range(a["steps"]):
        p = f"s{s}_"
        ens, e, sd, offsets, oracle_after, res = _step_streams(g, a, m, dev, s)
        upload_ensemble(ctx, ens)
        ctx.rng_replay(sd, offsets)
        ctx.device_set_grid(capi.GRID_EFIELD_X, e[0])
        ctx.device_set_grid(capi.GRID_EFIELD_Y, e[1])
        if dev.dim > 2:
            ctx.device_set_grid(capi.GRID_EFIELD_Z, e[2])
        ctx.set_step_index(s + 1)
        ctx.event_log_enable(1 << 16)
        removed = ctx.device_step(a["dt"])
        assert np.array_equal(removed, g[p + "removed_per_contact"]), f"step {s}"
        got = download_ensemble(ctx)
        assert_ensemble_close(got, ens_from(g, p + "drift_"), dev, f"{case} step {s}", check_grain="grain-rate" in a)
        ev, n_ev = ctx.event_log_read(1 << 16)
        dev_ev = ev[np.lexsort((ev[:, 3], ev[:, 2], ev[:, 1], ev[:, 0]))]
        cpu_ev = res["events"][np.lexsort((res["events"][:, 3], res["events"][:, 2], res["events"][:, 1], res["events"][:, 0]))]
        assert np.array_equal(dev_ev, cpu_ev), f"step {s}: scatter events"
        n_events += n_ev
    assert n_events > 100


@pytest.mark.parametrize("case", CASES)
def test_contacts_replay_the_reference(gpu_ctx_factory, case):
    """handleOhmicContacts of every recorded step: which particles are deleted, how many are injected per contact
    (exact), and the injected particles themselves from the reference's draws."""
    g = load_golden(case)
    a = DEVICE_CASES[case]
    m, dev = build_device(case)
    marks = g["draw_marks"].reshape(-1, 3)
    ctx = gpu_ctx_factory()
    upload_model(ctx, m)
    configure(ctx, dev, expected=g["expected_at_contact"].ravel())
    ctx.rng_philox(1)
    injected_total = 0
    for s in range(a["steps"]):
        p = f"s{s}_"
        before = ens_from(g, p + "drift_")
        upload_ensemble(ctx, before)
        ctx.set_step_index(s + 2)
        draws = g["draws"][int(marks[s, 1]):int(marks[s, 2])]
        net = ctx.device_contacts(replay_draws=draws)
        assert np.array_equal(net, g[p + "net_injected_per_contact"]), f"step {s}"
        assert_ensemble_close(download_ensemble(ctx), ens_from(g, p + "post_"), dev, f"{case} contacts {s}",
                              check_grain="grain-rate" in a)
        injected_total += len(draws) // (dev.dim + 7)
        ctx.device_assign()
        assert_counts(dev, ctx.device_get_grid(capi.GRID_COUNT), g[p + "count"], f"counts {s}")
        ctx.device_concentration()
        assert_grid_close(ctx.device_get_grid(capi.GRID_CONCENTRATION), g[p + "conc"], "concentration",
                          1e-12 if dev.pm_scheme == po.PM_CIC else 1e-15)
    assert injected_total > 0


@pytest.mark.parametrize("case", ["device_nec", "device_vwd"])
@pytest.mark.parametrize("nr_carriers", [1.0, 3.0, 2.5])
def test_nec_deposit_by_integer_hits(gpu_ctx_factory, case, nr_carriers):
    """NEC / NEC-VWD charge assignment (2-D schemes): the default kernel counts particles per mesh cell with integer atomics
    and lets every node collect its four cells (integer-valued carriers per particle only; 2.5 takes the fp64 atomics); option
    assign_fp64 forces one fp64 atomic per corner.  Both equal the oracle's deposit bit for bit, also with particles on the
    lower and just inside the upper faces of the device."""
    g = load_golden(case)
    m, dev = build_device(case)
    ens = ens_from(g, "init_")
    rng = np.random.default_rng(7)
    if ens.n < 1000:  # some fixtures hold few particles: crowd the box instead
        ens = po.Ensemble(5000)
        ens.n = 5000
        ens.x[:], ens.y[:] = rng.uniform(0, dev.max_pos[0], ens.n), rng.uniform(0, dev.max_pos[1], ens.n)
        if dev.dim > 2:
            ens.z[:] = rng.uniform(0, dev.max_pos[2], ens.n)
    k = min(64, ens.n)  # some particles on the lower faces and just inside the upper faces and corners of the box
    ens.x[:k] = rng.choice([0.0, np.nextafter(dev.max_pos[0], 0.0)], k)
    ens.y[k // 2:k] = rng.choice([0.0, np.nextafter(dev.max_pos[1], 0.0)], k - k // 2)
    if dev.dim > 2:
        ens.z[k // 4:k // 2] = np.nextafter(dev.max_pos[2], 0.0)
    want = dev.assign(ens, nr_carriers)
    got = {}
    for fp64 in (0, 1):
        ctx = gpu_ctx_factory()
        upload_model(ctx, m)
        configure(ctx, dev, nr_carriers=nr_carriers)
        ctx.set_option("assign_fp64", fp64)
        upload_ensemble(ctx, ens)
        ctx.device_assign()
        got[fp64] = ctx.device_get_grid(capi.GRID_COUNT)
        assert got[fp64].sum() == nr_carriers * ens.n
        assert np.array_equal(got[fp64], np.asarray(want).ravel()), f"assign_fp64={fp64}"


def test_self_consistent_run_conserves_bookkeeping_and_stays_physical(gpu_ctx_factory):
    """emcgpu_device_run (Philox): particle count changes exactly by the per-contact counters, counts grid sums to
    the ensemble size, the potential keeps its Dirichlet values, reservoir cells stay at their expected population,
    and the run is deterministic."""
    case = "device_bar"
    g = load_golden(case)
    a = DEVICE_CASES[case]
    m, dev = build_device(case)
    finals = []
    for _ in range(2):
        ctx = gpu_ctx_factory()
        upload_model(ctx, m)
        configure(ctx, dev, math_mode=capi.MATH_FAST)
        ctx.device_set_grid(capi.GRID_POTENTIAL, g["pot_eq"])
        ctx.device_set_grid(capi.GRID_CONCENTRATION, g["conc_eq"])
        upload_ensemble(ctx, ens_from(g, "init_"))
        ctx.device_reserve(4096)
        ctx.rng_philox(2026)
        ctx.set_step_index(1)
        n0 = ctx.size
        counters, sweeps = ctx.device_run(a["dt"], 200, 1e-4, 1.8, True, n_average=50)
        assert np.all(sweeps >= 1)
        avg_pot = ctx.device_get_grid(capi.GRID_SUM_POTENTIAL) / 50
        assert np.abs(avg_pot - ctx.device_get_grid(capi.GRID_POTENTIAL)).max() < 0.5  # Vt units: same solution, noise only
        left, net = counters[:, 0, :].sum(), counters[:, 1, :].sum()
        assert ctx.size == n0 - left + net
        count = ctx.device_get_grid(capi.GRID_COUNT)
        assert count.sum() == ctx.size
        pot = ctx.device_get_grid(capi.GRID_POTENTIAL)
        ohmic = g["is_ohmic"].ravel().astype(bool)
        assert np.allclose(pot[ohmic], g["s0_pot"].ravel()[ohmic], rtol=1e-13)
        expected = g["expected_at_contact"].ravel()
        res_cells = g["is_reservoir"].ravel().astype(bool)
        assert np.all(count[res_cells] >= np.floor(expected[res_cells])) and np.all(count[res_cells] <= np.ceil(expected[res_cells]))
        e = download_ensemble(ctx)
        assert e.x.min() >= 0 and e.x.max() <= dev.max_pos[0] and e.y.min() >= 0 and e.y.max() <= dev.max_pos[1]
        assert np.isfinite(e.energy).all() and e.energy.min() > 0
        assert counters[:, 0, :].sum() > 0  # particles did leave through the contacts and were replaced
        finals.append(e)
    for f in ("kx", "energy", "x", "y", "tau"):
        assert np.array_equal(getattr(finals[0], f), getattr(finals[1], f)), f


@pytest.mark.parametrize("case", CASES)
def test_sor_variants_agree(gpu_ctx_factory, case):
    """The two kernels of the lexicographic solver (thread pair per row / hyperplane loop) give bit-identical potentials
    and sweep counts; the opt-in red-black ordering converges to the same potential within the accuracy of the solver
    (and to 1e-8 V when both are driven to 1e-10 V)."""
    g = load_golden(case)
    m, dev = build_device(case)
    results = {}
    for name, opts in (("rows", {}), ("planes", {"sor_kernel": 1}), ("redblack", {"sor_order": 1}),
                       ("redblack_fast_8_ctas", {"sor_order": 1, "sor_kernel": 3}),
                       ("redblack_general_cluster", {"sor_order": 1, "sor_kernel": 2}),
                       ("redblack_one_cta", {"sor_order": 1, "sor_kernel": 1})):
        ctx = gpu_ctx_factory()
        upload_model(ctx, m)
        configure(ctx, dev)
        for k, v in opts.items():
            ctx.set_option(k, v)
        out = []
        for acc in (1e-4, 1e-10):
            ctx.device_set_grid(capi.GRID_POTENTIAL, g["pot_eq"])
            ctx.device_set_grid(capi.GRID_CONCENTRATION, g["conc_eq"])
            sweeps = ctx.device_poisson(False, acc, 1.8, True)
            out.append((sweeps, ctx.device_get_grid(capi.GRID_POTENTIAL)))
        eq_sweeps = ctx.device_poisson(True, 1e-4, 1.8, True)
        out.append((eq_sweeps, ctx.device_get_grid(capi.GRID_POTENTIAL)))
        results[name] = out
    for (s_r, p_r), (s_p, p_p) in zip(results["rows"], results["planes"]):
        assert s_r == s_p and np.array_equal(p_r, p_p)
    # red-black on a cluster (distributed shared memory) -- the fast 2-D form on 16 CTAs of 512 threads where the device
    # places such a cluster (the default) and on the portable 8 CTAs of 1024, the general cluster kernel -- and on one CTA:
    # the same iterates
    for other in ("redblack_fast_8_ctas", "redblack_general_cluster", "redblack_one_cta"):
        for (s_c, p_c), (s_1, p_1) in zip(results["redblack"], results[other]):
            assert s_c == s_1 and np.array_equal(p_c, p_1), other
    vt = dev.vt
    for k, tol_volt in ((0, 5e-4), (1, 1e-8), (2, 5e-4)):
        diff = float(np.abs(results["redblack"][k][1] - results["rows"][k][1]).max()) * vt
        assert diff <= tol_volt, (k, diff)
        assert results["redblack"][k][0] >= 1


def test_large_grid_poisson_and_empty_ensemble(gpu_ctx_factory):
    """A grid that does not fit into the shared memory of a CTA (201 x 201 points, gate strip, two ohmic contacts): the
    lexicographic solver falls back to the hyperplane kernel on global memory, red-black to the single-CTA form; both against
    the oracle's sequential sweep.  Then a device run with NO particles: contacts fill the reservoir cells from nothing."""
    lx = 2e-7
    dev = po.Device([lx, lx], [1e-9, 1e-9], device_width=1e-6)
    dev.add_doping_region([0, 0], [lx, lx], 2e23)
    dev.add_doping_region([lx / 2, 0], [lx, lx], -1e23)
    dev.add_contact(1, po.CONTACT_OHMIC, 0.0, [0.0], [lx])
    dev.add_contact(0, po.CONTACT_OHMIC, 0.2, [0.0], [lx])
    dev.add_contact(2, po.CONTACT_GATE, 0.4, [lx / 3], [2 * lx / 3], 3.9, 1.2e-9, 1.15 / 2)
    m = build_device("device_bar")[0]
    pot_ref = dev.initial_potential()
    sweeps_ref = dev.sor(pot_ref, None, 1e-4, 1.8, True)
    ctx = gpu_ctx_factory()
    upload_model(ctx, m)
    configure(ctx, dev)
    assert ctx.device_poisson(True, 1e-4, 1.8, True) == sweeps_ref
    assert_grid_close(ctx.device_get_grid(capi.GRID_POTENTIAL), pot_ref, "equilibrium potential, 201 x 201")
    conc = np.exp(pot_ref) * (1 + 0.05 * np.sin(np.arange(dev.cells)))  # some non-equilibrium density
    pot2 = pot_ref.copy()
    sweeps2 = dev.sor(pot2, conc.copy(), 1e-4, 1.8, True)
    ctx.device_set_grid(capi.GRID_CONCENTRATION, conc)
    assert ctx.device_poisson(False, 1e-4, 1.8, True) == sweeps2
    assert_grid_close(ctx.device_get_grid(capi.GRID_POTENTIAL), pot2, "non-equilibrium potential, 201 x 201")
    # red-black converges to the same potential; on a grid this large the stopping rule (max |delta| of a sweep) leaves both
    # orders well short of the fixed point at 1e-4 V, so compare them where they have converged
    pot3 = pot_ref.copy()
    dev.sor(pot3, conc.copy(), 1e-7, 1.8, True)
    ctx.set_option("sor_order", 1)
    ctx.device_set_grid(capi.GRID_POTENTIAL, pot_ref)
    assert ctx.device_poisson(False, 1e-7, 1.8, True) >= 1
    assert float(np.abs(ctx.device_get_grid(capi.GRID_POTENTIAL) - pot3).max()) * dev.vt <= 1e-4
    # empty ensemble
    small = build_device("device_bar")[1]
    ctx2 = gpu_ctx_factory()
    upload_model(ctx2, m)
    configure(ctx2, small, math_mode=capi.MATH_FAST)
    empty = po.Ensemble(0)
    empty.n = 0
    upload_ensemble(ctx2, empty)
    ctx2.device_reserve(4096)
    ctx2.rng_philox(1)
    counters, sweeps = ctx2.device_run(1e-15, 3, 1e-4, 1.8, True)
    expected = small.expected_at_contact()
    assert counters[0, 1, :].sum() == int(np.ceil(expected[expected > 0]).sum()) and counters[:, 0, :].sum() >= 0
    assert ctx2.size == counters[:, 1, :].sum() - counters[:, 0, :].sum()
    assert ctx2.device_get_grid(capi.GRID_COUNT).sum() == ctx2.size


@pytest.mark.parametrize("math_mode", [capi.MATH_EXACT, capi.MATH_FAST], ids=["exact", "fast"])
def test_long_chained_run_replays_the_reference(gpu_ctx_factory, math_mode):
    """250 CHAINED time steps on the device -- Poisson, field, drift / scatter, contacts, assignment, concentration, every step
    from the state the device's own previous step left, nothing re-uploaded -- fed the reference's draws (per-particle replay
    streams for the drift / scatter phase, the contact phase's draws in order).  Every step: removed and injected particles per
    contact and the ensemble size exact.  Every 25th step and the last: particle state within 1e-12 of the REFERENCE's
    ensemble (indices exact), potential within 1e-12 of the bias, counts exact."""
    case = "device_bar_long"
    from scenarios import DEVICE_LONG_CASES
    g = load_golden(case)
    a = DEVICE_LONG_CASES[case]
    m, dev = build_device(case)
    marks = g["draw_marks"].reshape(-1, 3)
    expected = dev.expected_at_contact()
    ctx = gpu_ctx_factory()
    upload_model(ctx, m)
    configure(ctx, dev, expected=g["expected_at_contact"].ravel(), math_mode=math_mode)
    ctx.device_reserve(2048)
    ctx.rng_philox(1)  # (not consumed: every draw of the run is replayed)
    ctx.device_set_grid(capi.GRID_POTENTIAL, g["pot_eq"])
    ctx.device_set_grid(capi.GRID_CONCENTRATION, g["conc_eq"])
    upload_ensemble(ctx, ens_from(g, "init_"))
    # the oracle runs the same chain bit for bit with the reference (tests/test_oracle_device.py): it tells which particle
    # consumed which draw of the drift / scatter phase
    mt = po.mt_state(a["seed"])
    for _ in range(int(g["draws_init_count"][0])):
        po.lib().orc_mt_next(mt)
    shadow = po.Ensemble(4096)
    ens0 = ens_from(g, "init_")
    for f in po.Ensemble.F64 + po.Ensemble.I32:
        getattr(shadow, f)[: ens0.n] = getattr(ens0, f)
    shadow.n = ens0.n
    pot, conc = g["pot_eq"].ravel().copy(), g["conc_eq"].ravel().copy()
    n_snaps = 0
    for s in range(a["steps"]):
        p = f"s{s}_"
        snap = (p + "pot") in g
        # ---- oracle shadow: fields and the attribution of the draws
        dev.sor(pot, conc, 1e-4, 1.8, s == 0)
        e = dev.efield(pot)
        res = dev.step(m, shadow, e, a["dt"], po.rng_mt(mt), step_index=s + 1, record=True)
        draws = g["draws"][int(marks[s, 0]):int(marks[s, 1])]
        assert len(res["rec_pid"]) == len(draws)
        sd, offsets = po.streams_from_record(draws, res["rec_pid"], shadow.n)
        dev.compact(shadow, res["removed"])
        dev.contacts(m, shadow, expected, mt)
        conc = dev.concentration(dev.assign(shadow))
        # ---- device: its own chain
        ctx.device_poisson(False, 1e-4, 1.8, s == 0)
        ctx.device_efield()
        if snap:
            assert_grid_close(ctx.device_get_grid(capi.GRID_POTENTIAL), g[p + "pot"], f"step {s}: potential", 1e-12, scale=a["voltage"])
        ctx.rng_replay(sd, offsets)
        ctx.set_step_index(s + 1)
        removed = ctx.device_step(a["dt"])
        assert np.array_equal(removed, g["removed_all"][s]), f"step {s}: removed per contact"
        if snap:
            assert_ensemble_close(download_ensemble(ctx), ens_from(g, p + "drift_"), dev, f"step {s}: after drift")
        net = ctx.device_contacts(replay_draws=g["draws"][int(marks[s, 1]):int(marks[s, 2])])
        assert np.array_equal(net, g["net_injected_all"][s]), f"step {s}: contacts"
        ctx.device_assign()
        ctx.device_concentration()
        if snap:
            got = download_ensemble(ctx)
            assert got.n == int(g["size_all"][s])
            assert_ensemble_close(got, ens_from(g, p + "post_"), dev, f"step {s}: after contacts")
            assert_counts(dev, ctx.device_get_grid(capi.GRID_COUNT), g[p + "count"], f"step {s}: counts")
            n_snaps += 1
    assert n_snaps == 11
