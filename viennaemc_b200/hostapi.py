"""ctypes binding of libemchost.so (include/emchost.h): the drop-in C++ host API instantiated
for the silicon model.  Builds rate tables on the host and uploads them through the C ABI."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import capi

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "lib", "libemchost.so")

ACOUSTIC, ZERO_ORDER, FIRST_ORDER, COULOMB = 1, 2, 4, 8
BULK_EXAMPLE = ACOUSTIC | ZERO_ORDER | FIRST_ORDER  # examples/bulkSimulation/bulkSimulation.cpp:100-103


class SiSpec(C.Structure):
    _fields_ = [("nLevels", C.c_int32), ("coulombSecond", C.c_int32), ("mechanisms", C.c_uint32),
                ("reserved", C.c_uint32), ("maxEnergy", C.c_double), ("temperature", C.c_double),
                ("doping", C.c_double), ("box", C.c_double * 3), ("spacing", C.c_double * 3)]


class Ga2O3Spec(C.Structure):
    _fields_ = [("polar", C.c_int32), ("multimode", C.c_int32), ("screening", C.c_int32), ("qResolved", C.c_int32),
                ("qResolvedAngle", C.c_int32), ("acousticBath", C.c_int32), ("impurity", C.c_int32), ("nLevels", C.c_int32),
                ("maxEnergy", C.c_double), ("temperature", C.c_double), ("doping", C.c_double), ("box", C.c_double),
                ("tauLO", C.c_double), ("tauAc", C.c_double)]


POLAR = {"eq": 0, "hot": 1, "screened_eq": 2, "screened_hot": 3}
EXPORTED_SYMBOLS = ["emchost_si_upload", "emchost_si_tables", "emchost_si_initial_ensemble", "emchost_ga2o3_host_loop",
                    "emchost_ga2o3_upload"]
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `python -m viennaemc_b200.build`")
        capi.load()  # libemchost links against libemcgpu
        L = C.CDLL(LIB_PATH)
        L.emchost_si_upload.argtypes = [C.c_void_p, C.POINTER(SiSpec)]
        L.emchost_si_tables.argtypes = [C.POINTER(SiSpec), C.POINTER(C.c_double), C.c_int64, C.POINTER(C.c_double),
                                        C.POINTER(C.c_int32)]
        L.emchost_si_initial_ensemble.argtypes = [C.POINTER(SiSpec), C.c_uint64, C.c_int64,
                                                  C.POINTER(C.POINTER(C.c_double)), C.POINTER(C.c_uint32),
                                                  C.POINTER(C.c_double)]
        L.emchost_si_initial_ensemble.restype = C.c_int64
        dp = C.POINTER(C.c_double)
        L.emchost_ga2o3_host_loop.argtypes = [C.POINTER(Ga2O3Spec), C.c_int, C.c_double, dp, dp, dp, dp, dp, dp, dp]
        L.emchost_ga2o3_upload.argtypes = [C.c_void_p, C.POINTER(Ga2O3Spec)]
        _lib = L
    return _lib


def si_spec(n_levels=1000, max_energy=1.0, temperature=300.0, doping=1e23, box=(5e-7,) * 3, spacing=(1e-7,) * 3,
            mechanisms=BULK_EXAMPLE, coulomb_second=False) -> SiSpec:
    s = SiSpec()
    s.nLevels, s.coulombSecond, s.mechanisms = n_levels, int(coulomb_second), mechanisms
    s.maxEnergy, s.temperature, s.doping = max_energy, temperature, doping
    for i in range(3):
        s.box[i], s.spacing[i] = box[i], spacing[i]
    return s


def si_upload(ctx: capi.Context, spec: SiSpec):
    """valleys + host-built cumulative rate tables of the silicon model -> GPU context"""
    rc = load().emchost_si_upload(ctx.h, C.byref(spec))
    if rc != capi.OK:
        raise capi.EmcGpuError(rc, ctx.L.emcgpu_last_error(ctx.h).decode())
    ctx.n_valleys = 1


def si_tables(spec: SiSpec):
    L = load()
    n_mech = C.c_int32()
    tau = C.c_double()
    L.emchost_si_tables(C.byref(spec), None, 0, C.byref(tau), C.byref(n_mech))
    cum = np.zeros((n_mech.value, spec.nLevels))
    rc = L.emchost_si_tables(C.byref(spec), cum.ctypes.data_as(C.POINTER(C.c_double)), cum.size, C.byref(tau),
                             C.byref(n_mech))
    assert rc == 0
    return cum, tau.value


def si_initial_ensemble(spec: SiSpec, seed: int):
    L = load()
    n = L.emchost_si_initial_ensemble(C.byref(spec), seed, 0, None, None, None)
    streams = [np.zeros(n) for _ in range(capi.N_STREAMS)]
    packed = np.zeros(n, dtype=np.uint32)
    grain = np.zeros(n)
    ptrs = (C.POINTER(C.c_double) * capi.N_STREAMS)(*[a.ctypes.data_as(C.POINTER(C.c_double)) for a in streams])
    got = L.emchost_si_initial_ensemble(C.byref(spec), seed, n, ptrs, packed.ctypes.data_as(C.POINTER(C.c_uint32)),
                                        grain.ctypes.data_as(C.POINTER(C.c_double)))
    assert got == n
    return streams, packed, grain


def ga2o3_spec(polar="screened_hot", multimode=0, screening=0, qresolved=0, qres_angle=1, acoustic_bath=1, impurity=0, levels=2000,
               emax=5.0, temperature=300.0, doping=1e23, box=3e-7, tau_lo=5e-12, tau_ac=20e-12, **_ignored) -> Ga2O3Spec:
    s = Ga2O3Spec()
    s.polar, s.multimode, s.screening, s.qResolved, s.qResolvedAngle = POLAR[polar], multimode, screening, qresolved, qres_angle
    s.acousticBath, s.impurity, s.nLevels = acoustic_bath, impurity, levels
    s.maxEnergy, s.temperature, s.doping, s.box, s.tauLO, s.tauAc = emax, temperature, doping, box, tau_lo, tau_ac
    return s


def ga2o3_host_loop(spec: Ga2O3Spec, dt, counts, mean_energy, n_baths, n_bins=300):
    """the host side of the hot-phonon loop driven with recorded inputs; see include/emchost.h"""
    L = load()
    n_steps = len(mean_energy)
    n_mech = L.emchost_ga2o3_host_loop(C.byref(spec), 0, dt, None, None, None, None, None, None, None)
    assert n_mech > 0
    dp = C.POINTER(C.c_double)
    cum0, cum1 = np.zeros((n_mech, spec.nLevels)), np.zeros((n_mech, spec.nLevels))
    tau, mean_nq = np.zeros(n_steps), np.zeros((n_steps, max(1, n_baths)))
    final_nq = np.zeros((max(1, n_baths), n_bins))
    c = np.ascontiguousarray(counts, dtype=np.float64) if n_baths else None
    e = np.ascontiguousarray(mean_energy, dtype=np.float64)
    rc = L.emchost_ga2o3_host_loop(C.byref(spec), n_steps, dt, c.ctypes.data_as(dp) if c is not None else None, e.ctypes.data_as(dp),
                                   cum0.ctypes.data_as(dp), cum1.ctypes.data_as(dp), tau.ctypes.data_as(dp),
                                   mean_nq.ctypes.data_as(dp), final_nq.ctypes.data_as(dp))
    assert rc == n_mech
    return dict(cum_initial=cum0, cum_final=cum1, tau=tau, mean_nq=mean_nq[:, :n_baths], final_nq=final_nq[:n_baths])


def ga2o3_upload(ctx: capi.Context, spec: Ga2O3Spec):
    rc = load().emchost_ga2o3_upload(ctx.h, C.byref(spec))
    if rc != capi.OK:
        raise capi.EmcGpuError(rc, ctx.L.emcgpu_last_error(ctx.h).decode())
    ctx.n_valleys = 1
