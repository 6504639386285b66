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
}


def build_model(case: str):
    c = GOLDEN_CASES[case]
    return (build_si if c["builder"] == "si" else build_mixed)(**c["kwargs"])


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
PM_SCHEMES = {"ngp": po.PM_NGP, "cic": po.PM_CIC, "nec": po.PM_NEC, "vwd": po.PM_NEC_VWD}


def build_device(case: str):
    """oracle-side model + device of a DEVICE_CASES entry (mirrors oracle/ref_device_driver.cpp)"""
    a = DEVICE_CASES[case]
    regions = (0, 1) if a["doping2"] else (0,)
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
    return m, dev
