"""Scenario definitions shared by the parity tests (TEST INFRASTRUCTURE).

`si`    : Si X valleys + the shipped bulkSimulation mechanism set
          (reference: examples/SiliconFunctions.hpp:21-145, examples/bulkSimulation/bulkSimulation.cpp:100-103)
`mixed` : synthetic four-valley material, one valley of every 3-D valley class, built identically in
          oracle/ref_bulk_driver.cpp::buildMixed against the reference headers.
"""
from __future__ import annotations

import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from oracle import pyoracle as po  # noqa: E402

SI_G = [[0], [1], [2]]
SI_F = [[1, 1, 2, 2], [0, 0, 2, 2], [0, 0, 1, 1]]
SI_DIRS = [[[1, 0, 0], [0, 1, 0], [0, 0, 1]],
           [[0, 1, 0], [1, 0, 0], [0, 0, 1]],
           [[0, 0, 1], [0, 1, 0], [1, 0, 0]]]


def build_si(mechs=("acoustic", "zero", "first"), n_levels=1000, max_energy=1.0, temperature=300.0,
             doping=1e23, regions=(0,)):
    m = po.Model(n_levels, max_energy, temperature, 2329.0, 9040.0)
    m.add_valley(po.VALLEY_NONPARABOLIC_ANISO, [0.916, 0.196, 0.196], 3, 0.5, 0.0, SI_DIRS)
    for reg in regions:
        if "acoustic" in mechs:
            m.add_acoustic(0, reg, 9.0)
        if "coulomb" in mechs:
            m.add_coulomb(0, reg, 11.8, doping)
        if "zero" in mechs:
            m.add_intervalley(0, False, 0, 0, reg, 5.23e10, 0.06, SI_F)
            m.add_intervalley(0, True, 0, 0, reg, 5.23e10, 0.06, SI_F)
            m.add_intervalley(0, False, 0, 0, reg, 5.23e10, 0.06, SI_G)
            m.add_intervalley(0, True, 0, 0, reg, 5.23e10, 0.06, SI_G)
        if "first" in mechs:
            m.add_intervalley(1, False, 0, 0, reg, 2.5, 0.023, SI_F)
            m.add_intervalley(1, True, 0, 0, reg, 2.5, 0.023, SI_F)
            m.add_intervalley(1, False, 0, 0, reg, 4.0, 0.018, SI_G)
            m.add_intervalley(1, True, 0, 0, reg, 4.0, 0.018, SI_G)
    m.build_tables()
    return m


MIXED_DEG = [1, 4, 2, 3]
MIXED_L_DIRS = [[[1, 1, 1], [-1, 1, 0], [-1, -1, 2]],
                [[-1, 1, 1], [1, 1, 0], [1, -1, 2]],
                [[1, -1, 1], [1, 1, 0], [-1, 1, 2]],
                [[1, 1, -1], [1, 0, 1], [-1, 2, 1]]]


def build_mixed(mechs=("acoustic", "zero", "first", "coulomb"), n_levels=250, max_energy=2.0,
                temperature=300.0, doping=1e23):
    m = po.Model(n_levels, max_energy, temperature, 2329.0, 9040.0)
    m.add_valley(po.VALLEY_NONPARABOLIC_ISO, 0.067, 1, 0.61, 0.0)
    m.add_valley(po.VALLEY_NONPARABOLIC_ANISO, [1.9, 0.075, 0.11], 4, 0.46, 0.05, MIXED_L_DIRS)
    m.add_valley(po.VALLEY_PARABOLIC_ISO, 0.3, 2, 0.0, 0.03)
    m.add_valley(po.VALLEY_PARABOLIC_ANISO, [0.9, 0.2, 0.3], 3, 0.0, 0.08, SI_DIRS)
    for vi in range(4):
        if "acoustic" in mechs:
            m.add_acoustic(vi, 0, 7.0 + vi)
        if "coulomb" in mechs:
            m.add_coulomb(vi, 0, 11.8, doping)
        for vf in range(4):
            if vf == vi:
                continue
            sm = [[sf for sf in range(MIXED_DEG[vf])] for _ in range(MIXED_DEG[vi])]
            if "zero" in mechs:
                m.add_intervalley(0, False, vi, vf, 0, 6e10, 0.03, sm)
                m.add_intervalley(0, True, vi, vf, 0, 6e10, 0.03, sm)
            if "first" in mechs and (vi + vf) % 2 == 1:
                m.add_intervalley(1, False, vi, vf, 0, 3.0, 0.02, sm)
                m.add_intervalley(1, True, vi, vf, 0, 3.0, 0.02, sm)
    m.build_tables()
    return m


# ---- single-layer MoS2: examples/singleLayerMoS2 with its default parameter set (parameterPilotto.hpp) ----
MOS2 = dict(rho=3.1e-6, v_sound=6.6e3, e_qk=0.16,
            ac_k=(23.1 + 29.1) / 2000., ac_m=(19.2 + 29.2) / 2000., ac_q=(17.9 + 23.6) / 2000.,
            op_gamma=(48.6 + 48.9 + 50.9) / 3000., op_k=(46.4 + 42.2 + 51.9) / 3000., op_m=(48.2 + 44.3 + 50.1) / 3000.,
            op_q=(48.0 + 44.2 + 52.2) / 3000.)
MOS2_SUB = dict(same=[[0], [1], [2], [3], [4], [5]], next=[[1], [2], [3], [4], [5], [0]],
                nbr=[[1, 5], [0, 2], [1, 3], [2, 4], [3, 5], [4, 0]], nbrnbr=[[2, 4], [3, 5], [4, 0], [5, 1], [0, 2], [1, 3]],
                opposite=[[3], [4], [5], [0], [1], [2]],
                slash=[[1, 3, 5], [0, 2, 4], [1, 3, 5], [0, 2, 4], [1, 3, 5], [0, 2, 4]],
                # (the last row is the reference's: parameterPilotto.hpp:61-62)
                samekind=[[0, 2, 4], [1, 3, 5], [0, 2, 4], [1, 3, 5], [0, 2, 4], [0, 2, 4]])


def build_mos2_pilotto(temperature=300.0, sheet_density=0.0):
    """parameterPilotto.hpp:64-176 on electron2D (5000 energy levels up to 0.5 eV): K valleys (isotropic), Q valleys
    (anisotropic, six in-plane frames 60 degrees apart), acoustic + zero-order intervalley mechanisms, in the order the
    example adds them"""
    import math
    P, S = MOS2, MOS2_SUB
    m = po.Model(5000, 0.5, temperature, 1.0, 1.0)
    m.set_electron2d(4)
    m.add_valley(po.VALLEY_NONPARABOLIC_ISO_SL, 0.47, 6, 0.94)
    a60 = math.pi / 3.
    m.add_valley(po.VALLEY_NONPARABOLIC_ANISO_SL, [0.54, 1.14, 0.0], 6, 1.16, P["e_qk"],
                 angles=[0, a60, 2 * a60, math.pi, 4 * a60, 5 * a60])
    m.add_acoustic_sl(0, 0, 4.5, P["rho"], P["v_sound"])
    m.add_acoustic_sl(1, 0, 2.8, P["rho"], P["v_sound"])

    def pair(vi, vf, sigma, ph, sub):
        m.add_intervalley_sl(False, vi, vf, 0, sigma, P["rho"], ph, S[sub])
        m.add_intervalley_sl(True, vi, vf, 0, sigma, P["rho"], ph, S[sub])

    if sheet_density > 0:  # parameterPilotto.hpp:119-131: the K -> K Gamma-phonon pair screened by the 2-D carrier gas
        qs = twod_screening_wavevector(sheet_density, temperature, 1.0, 0.47 * 9.11e-31, 4.0)
        m.add_screened_optical_sl(False, 0, 0, 5.8e10, P["rho"], P["op_gamma"], qs)
        m.add_screened_optical_sl(True, 0, 0, 5.8e10, P["rho"], P["op_gamma"], qs)
    else:
        pair(0, 0, 5.8e10, P["op_gamma"], "same")
    pair(0, 0, 1.4e10, P["ac_k"], "next")
    pair(0, 0, 2.0e10, P["op_k"], "next")
    pair(0, 1, 0.93e9, P["ac_q"], "samekind")
    pair(0, 1, 1.9e10, P["op_q"], "samekind")
    pair(0, 1, 4.4e10, P["ac_m"], "slash")
    pair(0, 1, 5.6e10, P["op_m"], "slash")
    pair(1, 1, 7.1e10, P["op_gamma"], "same")
    pair(1, 1, 2.1e10, P["ac_q"], "nbr")
    pair(1, 1, 4.8e10, P["op_q"], "nbr")
    pair(1, 1, 2.0e10, P["ac_m"], "nbrnbr")
    pair(1, 1, 4.0e10, P["op_m"], "nbrnbr")
    pair(1, 1, 4.8e10, P["ac_k"], "opposite")
    pair(1, 1, 6.5e10, P["op_k"], "opposite")
    pair(1, 0, 1.5e10, P["ac_q"], "same")
    pair(1, 0, 2.4e10, P["op_q"], "same")
    pair(1, 0, 4.4e10, P["ac_m"], "next")
    pair(1, 0, 6.6e10, P["op_m"], "next")
    m.build_tables()
    return m


def twod_screening_wavevector(sheet_density, temperature, env_permittivity, dos_mass, degeneracy=4.0):
    """emc2DScreening.hpp:49-62 in the reference's operation order"""
    import math
    Q, KB, HBAR, EPS0 = 1.60219e-19, 1.38066e-23, 1.05459e-34, 8.85419e-12
    if sheet_density <= 0 or temperature <= 0 or dos_mass <= 0:
        return 0.0
    d0 = degeneracy * dos_mass / (2 * 3.14159265358979323846 * HBAR * HBAR)
    kbt = KB * temperature
    dndmu = d0 * (1 - math.exp(-sheet_density / (d0 * kbt)))
    return Q * Q * dndmu / (2 * EPS0 * env_permittivity)


def build_mos2_kaasbjerg_subset(temperature=300.0, full=False, sheet_density=0.0, supported=False):
    """the part of parameterKaasbjerg.hpp whose mechanisms have device samplers (oracle/ref_bulk_driver.cpp:
    buildMoS2KaasbjergSubset): ONE parabolic single-layer valley with one sub-valley, acoustic TA / LA, zero-order LO / homopolar
    and the four first-order pairs through the constructors without a sub-valley map (emission added before absorption,
    :143-201)"""
    dp_cal, rho = 1.60, 3.1e-6
    m = po.Model(5000, 0.5, temperature, 1.0, 1.0)
    m.set_electron2d(4)
    m.add_valley(po.VALLEY_PARABOLIC_ISO_SL, 0.48, 1)
    m.add_acoustic_sl(0, 0, dp_cal * 1.6, rho, 4.2e3)
    m.add_acoustic_sl(0, 0, dp_cal * 2.8, rho, 6.7e3)
    for sigma, ph in ((dp_cal * 2.6e10, 0.041), (dp_cal * 4.1e10, 0.05)):
        m.add_intervalley_sl(True, 0, 0, 0, sigma, rho, ph, None)
        m.add_intervalley_sl(False, 0, 0, 0, sigma, rho, ph, None)
    # first order (:162-201): TO at K, TO at Gamma, TA, LA -- deformation potentials in eV
    for sigma, ph in ((dp_cal * 1.9, 0.048), (dp_cal * 4.0, 0.048), (dp_cal * 5.9, 0.023), (dp_cal * 3.9, 0.029)):
        m.add_intervalley_sl(True, 0, 0, 0, sigma, rho, ph, None, order=1)
        m.add_intervalley_sl(False, 0, 0, 0, sigma, rho, ph, None, order=1)
    if full:  # the whole set (:211-259): Froehlich absorption / emission at the LO phonon, piezoelectric TA / LA
        qs = twod_screening_wavevector(sheet_density, temperature, 1.0, 0.48 * 9.11e-31, 4.0)
        cc, width = dp_cal * (0.286 * 1e-10), 5.41e-10
        m.add_froehlich_sl(False, 0, 0, 0.048, cc, width, qs)
        m.add_froehlich_sl(True, 0, 0, 0.048, cc, width, qs)
        m.add_piezo_sl(0, 0, 3.0e-11, width, rho, 4.2e3, qs)
        m.add_piezo_sl(0, 0, 3.0e-11, width, rho, 6.7e3, qs)
    if supported:  # the three optional extrinsic helpers (:272-351) as oracle/ref_bulk_driver.cpp calls them (material mos2kx)
        Q, EPS0 = 1.60219e-19, 8.85419e-12
        qs4 = twod_screening_wavevector(sheet_density, temperature, 4.0, 0.48 * 9.11e-31, 4.0)
        m.add_charged_impurity_sl(0, 0, 1e16, 4.0, qs4, 4.0e-9, 0.0, 1.0)
        m.add_surface_roughness_sl(0, 0, Q * sheet_density / (2 * EPS0 * 4.0) + 0.0, 3.0e-10, 1.5e-9, qs4)
        qs1 = twod_screening_wavevector(sheet_density, temperature, 1.0, 0.48 * 9.11e-31, 4.0)
        d_per_mode = 0.5 * (1. / (5.03 + 1.0) - 1. / (23.0 + 1.0))
        for w_so in (0.0124, 0.0484):
            m.add_remote_so_sl(False, 0, 0, w_so, d_per_mode, 5.0e-10, qs1)
            m.add_remote_so_sl(True, 0, 0, w_so, d_per_mode, 5.0e-10, qs1)
    m.build_tables()
    return m


def build_mos2(case):
    a = MOS2_CASES[case]
    if a["material"] in ("mos2", "mos2ps"):
        return build_mos2_pilotto(sheet_density=a.get("sheet-density", 0.0) if a["material"] == "mos2ps" else 0.0)
    return build_mos2_kaasbjerg_subset(full=a["material"] != "mos2k", sheet_density=a.get("sheet-density", 0.0),
                                       supported=a["material"] == "mos2kx")


# recorder cases of the single-layer path (oracle/_ref/ref_bulk_driver --material mos2): box = (box, box, 0.65 nm), one cell in z
MOS2_CASES = {
    # the example's own time step and a high field (valley transfer K -> Q sets in)
    "mos2_pilotto": dict(material="mos2", cells=6, box=6e-8, field=4e6, fdir="1,0,0", dt=1e-16, steps=1500, seed=17),
    # large time step (several events per step), field off the axes
    "mos2_pilotto_bigdt": dict(material="mos2", cells=5, box=5e-8, field=1e7, fdir="0.6,-1,0", dt=2e-15, steps=120, seed=23),
    # one parabolic single-layer valley, the zero-order mechanisms without a sub-valley map (Kaasbjerg set, supported part)
    "mos2_kaasbjerg_subset": dict(material="mos2k", cells=5, box=5e-8, field=2e6, fdir="1,0.5,0", dt=1e-15, steps=300, seed=29),
    # the whole Kaasbjerg set (setKaasbjergParameter): + Froehlich and piezoelectric single-layer mechanisms, unscreened as
    # the example calls them ...
    "mos2_kaasbjerg": dict(material="mos2kf", cells=5, box=5e-8, field=2e6, fdir="1,0.5,0", dt=1e-15, steps=300, seed=31),
    # ... and screened by a 2-D carrier gas of 5e16 1/m^2 (the optional argument of the example's helper functions)
    "mos2_kaasbjerg_screened": {"material": "mos2kf", "cells": 5, "box": 5e-8, "field": 2e6, "fdir": "-0.3,1,0", "dt": 1e-15,
                                "steps": 300, "seed": 37, "sheet-density": 5e16},
    # a supported, doped film: + charged impurities, interface roughness, remote surface-optical phonons of HfO2 (the example's
    # optional extrinsic helpers), everything screened by 5e16 1/m^2
    "mos2_kaasbjerg_supported": {"material": "mos2kx", "cells": 5, "box": 5e-8, "field": 2e6, "fdir": "1,-0.4,0", "dt": 1e-15,
                                 "steps": 300, "seed": 41, "sheet-density": 5e16},
    # Pilotto set with the K -> K Gamma-phonon pair screened (emcScreenedIntravalleyOpticalMechanism), 1e15 1/m^2
    "mos2_pilotto_screened": {"material": "mos2ps", "cells": 5, "box": 5e-8, "field": 4e6, "fdir": "1,0.2,0", "dt": 5e-16,
                              "steps": 400, "seed": 43, "sheet-density": 1e15},
}
MOS2_LZ = 0.65e-9


# golden cases: name -> (ref driver args, model builder kwargs)
GOLDEN_CASES = {
    "si_bulk": dict(
        args=dict(material="si", mechs="acoustic,zero,first", cells=2, box=1e-7, doping=1e23, field=1e6,
                  fdir="-1,0,0", dt=1e-16, steps=1200, seed=7, levels=1000, emax=1.0),
        builder="si", kwargs=dict(mechs=("acoustic", "zero", "first"), n_levels=1000, max_energy=1.0)),
    "si_coulomb_bigdt": dict(
        args=dict(material="si", mechs="acoustic,zero,first,coulomb", cells=3, box=1.2e-7, doping=1e23,
                  field=3e6, fdir="1,2,-0.5", dt=1e-15, steps=25, seed=12345, levels=500, emax=4.0),
        builder="si", kwargs=dict(mechs=("acoustic", "zero", "first", "coulomb"), n_levels=500, max_energy=4.0)),
    "mixed": dict(
        args=dict(material="mixed", mechs="acoustic,zero,first,coulomb", cells=2, box=1e-7, doping=1e23,
                  field=2e6, fdir="0.3,-1,0.2", dt=2e-15, steps=80, seed=11, levels=250, emax=2.0),
        builder="mixed", kwargs=dict(mechs=("acoustic", "zero", "first", "coulomb"), n_levels=250, max_energy=2.0)),
    # the shipped bulkSimulation example at its own size: 12 500 electrons (box 5e-7 m, 5 cells per edge), 1000 time steps.
    # The fixture does not store the ~0.8 million raw draws (incompressible): they ARE the mt19937_64 stream of the seed
    # (tests/test_oracle_golden.py checks that for every fixture); it stores their count and a digest of the recorded
    # stream, helpers.load_golden() regenerates and verifies them.
    "si_bulk_config1": dict(
        args=dict(material="si", mechs="acoustic,zero,first", cells=5, box=5e-7, doping=1e23, field=1e6,
                  fdir="-1,0,0", dt=1e-16, steps=1000, seed=2026, levels=1000, emax=1.0),
        builder="si", kwargs=dict(mechs=("acoustic", "zero", "first"), n_levels=1000, max_energy=1.0), strip_draws=True),
    # grain boundaries: a second free-flight clock with a reflect / transmit hemisphere sampler (emcGrainScatterMechanism)
    "si_grain": dict(
        args={"material": "si", "mechs": "acoustic,zero,first", "cells": 2, "box": 1e-7, "doping": 1e23, "field": 2e6,
              "fdir": "-1,0.5,0", "dt": 2e-16, "steps": 300, "seed": 31, "levels": 1000, "emax": 1.0, "grain-rate": 2e13,
              "grain-prob": 0.4},
        builder="si", kwargs=dict(mechs=("acoustic", "zero", "first"), n_levels=1000, max_energy=1.0), grain=(0.4, 2e13)),
}


def build_model(case: str):
    c = GOLDEN_CASES[case]
    m = (build_si if c["builder"] == "si" else build_mixed)(**c["kwargs"])
    if "grain" in c:
        m.set_grain(*c["grain"])
    return m


# device-run golden cases (oracle/_ref/ref_device_driver): name -> driver args.  2-D silicon bars with ohmic
# contacts on the XMIN / XMAX faces (resistor2D.cpp scaled down); "device_gate" adds a gate contact on the
# middle third of YMIN and a second doping region (Robin boundary term, region look-ups, per-region tables).
DEVICE_CASES = {
    "device_bar": dict(lx=2e-7, ly=1e-7, hx=1e-8, hy=2.5e-8, doping=1e22, doping2=0, voltage=0.05, dt=1e-15,
                       steps=10, levels=1000, emax=4.0, gate=0, seed=5),
    "device_gate": dict(lx=3e-7, ly=1e-7, hx=1e-8, hy=2e-8, doping=2e22, doping2=5e21, voltage=0.3, dt=5e-16,
                        steps=6, levels=500, emax=4.0, gate=1, seed=9),
    # plug-in variants: particle-mesh scheme (emcCICScheme / emcNECScheme / mosfet2D's NECSchemeVWD), the electron
    # flavour of mosfet2D (electronVWD) and rough walls (constant specularity on YMIN, momentum dependent on YMAX)
    "device_cic": {"lx": 2e-7, "ly": 1e-7, "hx": 1e-8, "hy": 2.5e-8, "doping": 1e22, "doping2": 0, "voltage": 0.05,
                   "dt": 1e-15, "steps": 6, "levels": 1000, "emax": 4.0, "gate": 0, "seed": 21, "scheme": "cic",
                   "surface-ymin-const": 0.3, "surface-ymax-mom": 2e-9},
    "device_nec": {"lx": 3e-7, "ly": 1e-7, "hx": 1e-8, "hy": 2e-8, "doping": 2e22, "doping2": 5e21, "voltage": 0.3,
                   "dt": 5e-16, "steps": 5, "levels": 500, "emax": 4.0, "gate": 1, "seed": 22, "scheme": "nec"},
    "device_vwd": {"lx": 3e-7, "ly": 1e-7, "hx": 1e-8, "hy": 2e-8, "doping": 2e22, "doping2": 5e21, "voltage": 0.3,
                   "dt": 2e-15, "steps": 4, "levels": 500, "emax": 4.0, "gate": 1, "seed": 23, "scheme": "vwd",
                   "electron": "vwd", "surface-ymin-const": 0.5},
}
DEVICE_CASES["device_grain"] = {"lx": 2e-7, "ly": 1e-7, "hx": 1e-8, "hy": 2.5e-8, "doping": 1e22, "doping2": 0, "voltage": 0.05,
                               "dt": 1e-15, "steps": 8, "levels": 1000, "emax": 4.0, "gate": 0, "seed": 41, "grain-rate": 3e13,
                               "grain-prob": 0.6}
# 3-D boxes (oracle/_ref/ref_device3d_driver): contacts over whole faces, gate strip over the full depth
DEVICE_CASES["device_3d_ngp"] = {"lx": 1.6e-7, "ly": 8e-8, "lz": 6e-8, "hx": 1e-8, "hy": 2e-8, "hz": 2e-8, "doping": 3e22, "doping2": 1e22,
                                 "voltage": 0.2, "dt": 1e-15, "steps": 5, "levels": 500, "emax": 4.0, "gate": 1, "seed": 51}
DEVICE_CASES["device_3d_cic"] = {"lx": 1.6e-7, "ly": 8e-8, "lz": 6e-8, "hx": 1e-8, "hy": 2e-8, "hz": 2e-8, "doping": 3e22, "doping2": 0,
                                 "voltage": 0.1, "dt": 1e-15, "steps": 4, "levels": 500, "emax": 4.0, "gate": 0, "seed": 52,
                                 "scheme": "cic", "surface-ymin-const": 0.4}
PM_SCHEMES = {"ngp": po.PM_NGP, "cic": po.PM_CIC, "nec": po.PM_NEC, "vwd": po.PM_NEC_VWD}


# long CHAINED device run ("a device replay >= 200 steps"): the bar of device_bar, twice as long, 250 time steps;
# the recorder keeps the grids and ensembles of every 25th step and of the last one, the per-contact counters of every step
DEVICE_LONG_CASES = {
    "device_bar_long": {"lx": 4e-7, "ly": 1e-7, "hx": 1e-8, "hy": 2.5e-8, "doping": 1e22, "doping2": 0, "voltage": 0.2,
                        "dt": 1e-15, "steps": 250, "levels": 1000, "emax": 4.0, "gate": 0, "seed": 77, "snap-every": 25},
}


def build_device(case: str):
    """oracle-side model + device of a DEVICE_CASES entry (mirrors oracle/ref_device_driver.cpp)"""
    a = DEVICE_CASES[case] if case in DEVICE_CASES else DEVICE_LONG_CASES[case]
    regions = (0, 1) if a["doping2"] else (0,)
    if "lz" in a:  # 3-D box
        size, h = [a["lx"], a["ly"], a["lz"]], [a["hx"], a["hy"], a["hz"]]
        dev = po.Device(size, h)
        dev.add_doping_region([0, 0, 0], size, a["doping"])
        if a["doping2"]:
            dev.add_doping_region([a["lx"] / 2, 0, 0], size, a["doping2"])
        dev.add_contact(1, po.CONTACT_OHMIC, 0.0, [0.0, 0.0], [a["ly"], a["lz"]])
        dev.add_contact(0, po.CONTACT_OHMIC, a["voltage"], [0.0, 0.0], [a["ly"], a["lz"]])
        if a["gate"]:
            dev.add_contact(2, po.CONTACT_GATE, 0.5, [a["lx"] / 3, 0.0], [2 * a["lx"] / 3, a["lz"]], 3.9, 1.2e-9, 1.15 / 2)
    else:
        dev = po.Device([a["lx"], a["ly"]], [a["hx"], a["hy"]], device_width=1e-6)
        dev.add_doping_region([0, 0], [a["lx"], a["ly"]], a["doping"])
        if a["doping2"]:
            dev.add_doping_region([a["lx"] / 2, 0], [a["lx"], a["ly"]], a["doping2"])
        dev.add_contact(1, po.CONTACT_OHMIC, 0.0, [0.0], [a["ly"]])
        dev.add_contact(0, po.CONTACT_OHMIC, a["voltage"], [0.0], [a["ly"]])
        if a["gate"]:
            dev.add_contact(2, po.CONTACT_GATE, 0.5, [a["lx"] / 3], [2 * a["lx"] / 3], 3.9, 1.2e-9, 1.15 / 2)
    dev.pm_scheme = PM_SCHEMES[a.get("scheme", "ngp")]
    dev.electron_kind = po.ELECTRON_VWD if a.get("electron") == "vwd" else po.ELECTRON_EMC
    if "surface-ymin-const" in a:
        dev.surface_kind[2], dev.surface_param[2] = po.SURFACE_CONSTANT, a["surface-ymin-const"]
    if "surface-ymax-mom" in a:
        dev.surface_kind[3], dev.surface_param[3] = po.SURFACE_MOMENTUM, a["surface-ymax-mom"]
    m = po.Model(a["levels"], a["emax"], 300.0, 2329.0, 9040.0)
    m.add_valley(po.VALLEY_NONPARABOLIC_ANISO, [0.916, 0.196, 0.196], 3, 0.5, 0.0, SI_DIRS)
    dop = [a["doping"], a["doping2"]]
    # mechanism order of the driver: Acoustic, Zero x4, First x4, Coulomb -- each added for all regions at once
    for reg in regions:
        m.add_acoustic(0, reg, 9.0)
    for emission, sub in ((False, SI_F), (True, SI_F), (False, SI_G), (True, SI_G)):
        for reg in regions:
            m.add_intervalley(0, emission, 0, 0, reg, 5.23e10, 0.06, sub)
    for (dp, hw), sub in (((2.5, 0.023), SI_F), ((4.0, 0.018), SI_G)):
        for emission in (False, True):
            for reg in regions:
                m.add_intervalley(1, emission, 0, 0, reg, dp, hw, sub)
    for reg in regions:
        m.add_coulomb(0, reg, 11.8, dop[reg])
    m.build_tables()
    if "grain-rate" in a:
        m.set_grain(a["grain-prob"], a["grain-rate"])
    return m, dev


# ---- config 5: beta-Ga2O3 bulk with polar-optical (Froehlich) scattering and hot phonons -------------------------
# (oracle/_ref/ref_ga2o3_driver around examples/hotPhononGa2O3/Ga2O3Functions.hpp).  Parameter values:
# Ga2O3Functions.hpp:45-88.
GA2O3 = dict(eps_lo=10.2, eps_hi=3.573, rho=5880.0, v_sound=6800.0, rel_mass=0.284, alpha=0.106, hw_pop=0.044,
             mode_energy=[0.0302, 0.0429, 0.0646, 0.0796, 0.0936], mode_weight=[0.1382, 0.0326, 0.2448, 0.1510, 0.4334],
             hw_npo=0.090, d_npo=8.05e10, sigma_ac=4.8, n_bins=300, dq=1e7)
GA2O3_CASES = {
    # the shipped default of hotPhononGa2O3.cpp: screened-hot classes with screening off, mean-field occupation
    "ga2o3_hot": dict(polar="screened_hot", steps=160, field=2e7, seed=1),
    # everything on: 5 polar modes, Debye screening updated from the carrier temperature, q-resolved rate and angle
    "ga2o3_qres": dict(polar="screened_hot", steps=60, field=3e7, seed=2, multimode=1, screening=1, qresolved=1, impurity=1,
                       box=2e-7),
    # the unscreened classes: emcHotPhononFroehlich*3D and emcFroehlich*3D
    "ga2o3_plain_hot": dict(polar="hot", steps=80, field=2e7, seed=3, box=2e-7),
    "ga2o3_eq": dict(polar="eq", steps=60, field=1e7, seed=4, box=2e-7),
    "ga2o3_screened_eq": dict(polar="screened_eq", steps=60, field=1e7, seed=5, box=2e-7, screening=1),
}
GA2O3_DEFAULTS = dict(box=3e-7, doping=1e23, dt=1e-16, temperature=300.0, tau_lo=5e-12, tau_ac=20e-12, emax=5.0, levels=2000,
                      multimode=0, screening=0, qresolved=0, qres_angle=1, acoustic_bath=1, impurity=0, reinit_every=1)


def ga2o3_args(case):
    a = dict(GA2O3_DEFAULTS)
    a.update(GA2O3_CASES[case])
    return a


def build_ga2o3(case):
    """oracle-side model + phonon baths of a GA2O3_CASES entry (mirrors oracle/ref_ga2o3_driver.cpp)"""
    a = ga2o3_args(case)
    g = GA2O3
    m = po.Model(a["levels"], a["emax"], a["temperature"], g["rho"], g["v_sound"])
    m.add_valley(po.VALLEY_NONPARABOLIC_ISO, g["rel_mass"], 1, g["alpha"], 0.0)
    m.add_acoustic(0, 0, g["sigma_ac"])
    m.add_intervalley(0, False, 0, 0, 0, g["d_npo"], g["hw_npo"], [[0]])
    m.add_intervalley(0, True, 0, 0, 0, g["d_npo"], g["hw_npo"], [[0]])
    if a["impurity"]:
        m.add_coulomb(0, 0, g["eps_lo"], a["doping"])
    if a["multimode"]:
        energies = g["mode_energy"]
        inv_hi = 1.0 / g["eps_hi"]
        total = inv_hi - 1.0 / g["eps_lo"]
        eps_lo = [1.0 / (inv_hi - w * total) for w in g["mode_weight"]]
    else:
        energies, eps_lo = [g["hw_pop"]], [g["eps_lo"]]
    qs2 = po.plasmon_qs2(a["doping"], a["temperature"], g["eps_lo"]) if a["screening"] else 0.0
    m.set_qs2(qs2)
    hot = a["polar"] in ("hot", "screened_hot")
    baths = []
    if hot:
        v_sim = a["box"] * a["box"] * a["box"]
        for hw in energies:
            b = po.PhononBath(g["n_bins"], g["dq"], a["tau_lo"], hw, a["temperature"], v_sim, bool(a["acoustic_bath"]), hw / 2.0,
                              a["tau_ac"])
            b.set_qs2(qs2)
            m.add_bath(b)
            baths.append(b)
    variant = {"eq": po.FROEHLICH_EQ, "hot": po.FROEHLICH_HOT, "screened_eq": po.FROEHLICH_SCREENED_EQ,
               "screened_hot": po.FROEHLICH_SCREENED_HOT}[a["polar"]]
    for i, hw in enumerate(energies):
        # the unscreened helpers of Ga2O3Functions.hpp always use the full static permittivity (:173-187, :206-222)
        el = eps_lo[i] if variant >= po.FROEHLICH_SCREENED_EQ else g["eps_lo"]
        for emission in (False, True):
            m.add_froehlich(variant, emission, 0, 0, hw, g["rel_mass"], g["eps_hi"], el, a["temperature"], i if hot else -1,
                            bool(a["qresolved"]), bool(a["qres_angle"]))
    m.build_tables()
    return m, baths, a


# ---- hot carriers in a metal-halide perovskite: electrons AND holes on shared phonon baths ---------------
# oracle/_ref/ref_mhp_driver: the data-parallel part of examples/hotCarrierMHP/hotCarrierMHP.cpp (MAPbI3 preset)
MHP = dict(eps_hi=5.0, eps_lo=33.5, mass_e=0.20, mass_h=0.25, hw_lo=0.0115, gap=1.60, photon=3.1, rho=4000.0, v_sound=2000.0)
MHP_CASES = {
    # the example's default polar model: screened hot-phonon classes, q-resolved rate and angle, Debye screening from both species
    "mhp_qres": dict(polar="screened_hot", steps=60, seed=1),
    # unscreened hot-phonon classes, mean-field occupation
    "mhp_hot": dict(polar="hot", steps=60, seed=2, screening=0, qresolved=0),
    # parabolic bands (--alpha 0), equilibrium phonons with screening
    "mhp_parabolic_eq": dict(polar="screened_eq", steps=40, seed=3, qresolved=0, alpha_e=0.0, alpha_h=0.0),
}
MHP_DEFAULTS = dict(box=8e-8, density=1e24, dt=5e-15, temperature=300.0, tau_lo=0.6e-12, tau_ac=30e-12, emax=4.0, levels=1000,
                    screening=1, qresolved=1, acoustic_bath=1, bins=40, dq=5e7, alpha_e=-1.0, alpha_h=-1.0)


def mhp_args(case):
    a = dict(MHP_DEFAULTS)
    a.update(MHP_CASES[case])
    g = MHP
    if a["alpha_e"] < 0:
        a["alpha_e"] = (1.0 / g["gap"]) * (1.0 - g["mass_e"]) * (1.0 - g["mass_e"])
    if a["alpha_h"] < 0:
        a["alpha_h"] = (1.0 / g["gap"]) * (1.0 - g["mass_h"]) * (1.0 - g["mass_h"])
    excess = g["photon"] - g["gap"]
    a["energy_e"] = g["mass_h"] / (g["mass_e"] + g["mass_h"]) * excess
    a["energy_h"] = g["mass_e"] / (g["mass_e"] + g["mass_h"]) * excess
    return a


def build_mhp(case):
    """oracle-side models of the two species (electrons, holes) + the shared phonon baths (mirrors oracle/ref_mhp_driver.cpp)"""
    a = mhp_args(case)
    g = MHP
    hot = a["polar"] in ("hot", "screened_hot")
    v_sim = a["box"] * a["box"] * a["box"]
    baths = []
    if hot:
        baths.append(po.PhononBath(a["bins"], a["dq"], a["tau_lo"], g["hw_lo"], a["temperature"], v_sim, bool(a["acoustic_bath"]),
                                   g["hw_lo"] / 2.0, a["tau_ac"], 0.0, 0.00781, a["tau_ac"]))
    variant = {"hot": po.FROEHLICH_HOT, "screened_eq": po.FROEHLICH_SCREENED_EQ, "screened_hot": po.FROEHLICH_SCREENED_HOT}[a["polar"]]
    models = []
    for mass, alpha, energy in ((g["mass_e"], a["alpha_e"], a["energy_e"]), (g["mass_h"], a["alpha_h"], a["energy_h"])):
        m = po.Model(a["levels"], a["emax"], a["temperature"], g["rho"], g["v_sound"])
        m.add_valley(po.VALLEY_NONPARABOLIC_ISO, mass, 1, alpha, 0.0)
        m.set_init_energy(energy)
        for b in baths:
            m.add_bath(b)
        for emission in (False, True):
            m.add_froehlich(variant, emission, 0, 0, g["hw_lo"], mass, g["eps_hi"], g["eps_lo"], a["temperature"], 0 if hot else -1,
                            bool(a["qresolved"]), True)
        m.build_tables()
        models.append(m)
    return models, baths, a
