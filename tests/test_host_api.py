"""CPU tests of the drop-in C++ host API (viennaemc_b200/host/include) through libemchost.so:
the host-built rate tables and the initial ensemble must equal the reference's, bit for bit.

Pins: the oracle tables were proven identical to the unmodified reference's by
tests/test_oracle_golden.py; the golden init_* arrays were recorded from the unmodified reference."""
import ctypes
import glob
import os
import subprocess

import numpy as np
import pytest

from helpers import load_golden
from scenarios import GOLDEN_CASES, build_si
from viennaemc_b200 import hostapi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST_INC = os.path.join(ROOT, "viennaemc_b200", "host", "include")
REFERENCE = "/root/reference"


def test_libemchost_exports_every_declared_symbol():
    lib = ctypes.CDLL(hostapi.LIB_PATH)
    header = open(os.path.join(ROOT, "include", "emchost.h")).read()
    for sym in hostapi.EXPORTED_SYMBOLS:
        assert sym + "(" in header
        assert hasattr(lib, sym)


@pytest.mark.parametrize("kw,oracle_kw", [
    (dict(), dict()),
    (dict(n_levels=500, max_energy=4.0, mechanisms=15, coulomb_second=True),
     dict(mechs=("acoustic", "zero", "first", "coulomb"), n_levels=500, max_energy=4.0)),
    (dict(n_levels=200, max_energy=2.0, temperature=77.0, mechanisms=hostapi.ACOUSTIC | hostapi.ZERO_ORDER),
     dict(mechs=("acoustic", "zero"), n_levels=200, max_energy=2.0, temperature=77.0)),
])
def test_host_built_tables_equal_the_oracle_bit_for_bit(kw, oracle_kw):
    cum, tau = hostapi.si_tables(hostapi.si_spec(**kw))
    ts = build_si(**oracle_kw).tablesets()[0]
    assert cum.shape == ts["cum"].shape
    assert np.array_equal(cum, ts["cum"])
    assert tau == ts["tau"]


def test_tau_known_answers_of_the_reference():
    """tau = 1/Gamma_max as printed by the reference at table build:
    bulk Si 8.73807e-15 s; resistor set (10 mechanisms incl. Coulomb at 1e22, 4 eV) 4.29583e-16 s."""
    _, tau = hostapi.si_tables(hostapi.si_spec())
    assert f"{tau:.6g}" == "8.73807e-15"
    _, tau = hostapi.si_tables(hostapi.si_spec(n_levels=1000, max_energy=4.0, doping=1e22, mechanisms=15))
    assert f"{tau:.6g}" == "4.29583e-16"


@pytest.mark.parametrize("case", ["si_bulk", "si_coulomb_bigdt"])
def test_initial_ensemble_equals_the_reference_bit_for_bit(case):
    g = load_golden(case)
    a = GOLDEN_CASES[case]["args"]
    mech = {"acoustic": 1, "zero": 2, "first": 4, "coulomb": 8}
    mask = sum(mech[m] for m in a["mechs"].split(","))
    spec = hostapi.si_spec(n_levels=a["levels"], max_energy=a["emax"], doping=a["doping"], box=(a["box"],) * 3,
                           spacing=(a["box"] / a["cells"],) * 3, mechanisms=mask, coulomb_second=True)
    streams, packed, grain = hostapi.si_initial_ensemble(spec, a["seed"])
    assert len(packed) == len(g["init_energy"])
    assert np.array_equal(np.stack(streams[0:3], 1), g["init_k"])
    assert np.array_equal(np.stack(streams[5:8], 1), g["init_pos"])
    assert np.array_equal(streams[3], g["init_energy"])
    assert np.array_equal(streams[4], g["init_tau"])
    assert np.array_equal(grain, g["init_grainTau"])
    idx = g["init_idx"]
    assert np.array_equal(packed & 0xFF, idx[:, 0]) and np.array_equal((packed >> 8) & 0xFF, idx[:, 1])


def test_every_public_header_is_self_contained():
    headers = [h for h in glob.glob(os.path.join(HOST_INC, "**", "*.hpp"), recursive=True)]
    assert len(headers) >= 20
    for h in headers:
        r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-I", HOST_INC, "-I", os.path.join(ROOT, "include"),
                            "-x", "c++", h], capture_output=True, text=True)
        assert r.returncode == 0, f"{h}:\n{r.stderr[:2000]}"


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference tree not mounted")
def test_unmodified_reference_example_compiles_against_the_dropin_headers():
    """The reference's own bulk example main() and its SiliconFunctions.hpp build against our headers
    (our basicBulkParticleHandler.hpp is pre-included; its include guard is the reference's)."""
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-I", HOST_INC, "-I", os.path.join(ROOT, "include"),
                        "-include", os.path.join(HOST_INC, "basicBulkParticleHandler.hpp"),
                        os.path.join(REFERENCE, "examples", "bulkSimulation", "bulkSimulation.cpp")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[:4000]


# ---- config 5: host side of the hot-phonon loop (emcPhononBath, emcPlasmonScreening, Froehlich rate classes) ----------
@pytest.mark.parametrize("case", ["ga2o3_hot", "ga2o3_qres", "ga2o3_plain_hot", "ga2o3_eq", "ga2o3_screened_eq"])
def test_ga2o3_host_loop_reproduces_the_reference_bit_for_bit(case):
    """The drop-in host classes driven with the event counts and mean energies the UNMODIFIED reference recorded
    (tests/golden/ga2o3_*.npz): rate tables before the first and after the last step, tau after every rebuild, <N_q> after
    every bath update and the final occupations are the reference's, bit for bit."""
    from helpers import load_golden
    from scenarios import ga2o3_args
    g = load_golden(case)
    a = ga2o3_args(case)
    spec = hostapi.ga2o3_spec(**a)
    n_baths = g["mean_nq"].shape[1] if "mean_nq" in g.files else 0
    counts = g["bath_counts"] if n_baths else None
    out = hostapi.ga2o3_host_loop(spec, a["dt"], counts, g["obs"][:, 0], n_baths)
    assert np.array_equal(out["cum_initial"], g["init_cum_v0_r0"])
    assert np.array_equal(out["tau"], g["tau_series"])
    assert np.array_equal(out["cum_final"], g["final_cum_v0_r0"])
    if n_baths:
        assert np.array_equal(out["mean_nq"], g["mean_nq"])
        assert np.array_equal(out["final_nq"], g["final_nq"])
