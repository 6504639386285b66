"""TEST INFRASTRUCTURE: ctypes binding of oracle/liboracle.so (the CPU restatement).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import
this module, and only as the checker.  The product (viennaemc_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

MAX_SUB = 8
MAX_FINAL = 8
VALLEY_PARABOLIC_ISO, VALLEY_NONPARABOLIC_ISO, VALLEY_PARABOLIC_ANISO, VALLEY_NONPARABOLIC_ANISO = range(4)
VALLEY_PARABOLIC_ISO_SL, VALLEY_NONPARABOLIC_ISO_SL, VALLEY_NONPARABOLIC_ANISO_SL = 4, 5, 7  # single-layer classes
SAMPLER_NONE, SAMPLER_ISOTROPIC_ELASTIC, SAMPLER_INTERVALLEY, SAMPLER_COULOMB = range(4)
RNG_MT_GLOBAL, RNG_STREAMS, RNG_PHILOX = range(3)

ME = 9.11e-31
Q = 1.60219e-19


class Valley(C.Structure):
    _fields_ = [("kind", C.c_int32), ("deg", C.c_int32), ("mCond", C.c_double), ("mDos", C.c_double),
                ("alpha", C.c_double), ("eBottom", C.c_double), ("vogt", C.c_double * 3),
                ("rot", (C.c_double * 9) * MAX_SUB)]


class Mech(C.Structure):
    _fields_ = [("sampler", C.c_int32), ("finalValley", C.c_int32), ("nFinal", C.c_int32),
                ("globalId", C.c_int32), ("p", C.c_double * 4),
                ("finalSub", (C.c_int32 * MAX_FINAL) * MAX_SUB)]


_DP = C.POINTER(C.c_double)
_IP = C.POINTER(C.c_int32)


class EnsembleC(C.Structure):
    _fields_ = [("n", C.c_int64)] + [(f, _DP) for f in
                                     ("kx", "ky", "kz", "energy", "tau", "grainTau", "x", "y", "z")] + \
               [(f, _IP) for f in ("valley", "sub", "region")]


class RngCfg(C.Structure):
    _fields_ = [("mode", C.c_int32), ("mtState", C.POINTER(C.c_uint64)),
                ("draws", C.POINTER(C.c_uint64)), ("offsets", C.POINTER(C.c_int64)),
                ("cursor", C.POINTER(C.c_int64)), ("philoxSeed", C.c_uint64),
                ("particleIdBase", C.c_int64)]


def build(force: bool = False) -> str:
    """Compile liboracle.so (gcc only).  Building the checker is not using it."""
    so = os.path.join(_HERE, "liboracle.so")
    src = [os.path.join(_HERE, f) for f in ("emc_oracle.c", "emc_oracle.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_model_create.restype = C.c_void_p
        L.orc_model_create.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double]
        L.orc_model_destroy.argtypes = [C.c_void_p]
        L.orc_model_set_init_energy.argtypes = [C.c_void_p, C.c_double]
        L.orc_model_set_electron2d.argtypes = [C.c_void_p, C.c_int]
        L.orc_add_valley.argtypes = [C.c_void_p, C.c_int, _DP, C.c_double, C.c_int, C.c_double, C.c_double, _DP]
        L.orc_add_acoustic.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double]
        L.orc_add_intervalley.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                          C.c_double, C.c_int, C.c_int, _IP]
        L.orc_add_coulomb.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double]
        L.orc_add_acoustic_sl.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]
        L.orc_add_intervalley_sl.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                             C.c_double, C.c_int, C.c_int, _IP]
        L.orc_add_froehlich_sl.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double]
        L.orc_add_piezo_sl.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double]
        L.orc_add_charged_impurity_sl.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_double] * 6
        L.orc_add_surface_roughness_sl.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_double] * 4
        L.orc_add_remote_so_sl.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_double] * 4
        L.orc_add_screened_optical_sl.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_double] * 4
        L.orc_build_tables.argtypes = [C.c_void_p]
        L.orc_model_set_grain.argtypes = [C.c_void_p, C.c_double, C.c_double]
        # Froehlich family + phonon bath
        L.orc_bath_create.restype = C.c_void_p
        L.orc_bath_create.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_double,
                                      C.c_double, C.c_double, C.c_double, C.c_double]
        L.orc_bath_destroy.argtypes = [C.c_void_p]
        L.orc_bath_set_qs2.argtypes = [C.c_void_p, C.c_double]
        L.orc_bath_update.argtypes = [C.c_void_p, C.c_double]
        for name in ("orc_bath_mean_nq", "orc_bath_acoustic_temp", "orc_bath_n0"):
            getattr(L, name).restype = C.c_double
            getattr(L, name).argtypes = [C.c_void_p]
        L.orc_bath_nq_window.restype = C.c_double
        L.orc_bath_nq_window.argtypes = [C.c_void_p, C.c_double, C.c_double]
        L.orc_set_limit_flags.argtypes = [C.c_void_p]
        L.orc_bath_sample_q.restype = C.c_double
        L.orc_bath_sample_q.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_double]
        L.orc_bath_copy.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
        L.orc_bath_add_counts.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.orc_model_add_bath.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_model_set_qs2.argtypes = [C.c_void_p, C.c_double]
        L.orc_plasmon_qs2.restype = C.c_double
        L.orc_plasmon_qs2.argtypes = [C.c_double, C.c_double, C.c_double]
        L.orc_add_froehlich.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                        C.c_double, C.c_double, C.c_int, C.c_int, C.c_int]
        L.orc_n_valleys.argtypes = [C.c_void_p]
        L.orc_get_valley.argtypes = [C.c_void_p, C.c_int, C.POINTER(Valley)]
        L.orc_n_mechanisms.argtypes = [C.c_void_p]
        L.orc_raw_rate.restype = C.c_double
        L.orc_raw_rate.argtypes = [C.c_void_p, C.c_int, C.c_double]
        L.orc_n_tablesets.argtypes = [C.c_void_p]
        L.orc_tableset_info.argtypes = [C.c_void_p, C.c_int, _IP, _IP, _IP, _DP]
        L.orc_tableset_copy.argtypes = [C.c_void_p, C.c_int, _DP, C.POINTER(Mech)]
        L.orc_tau.restype = C.c_double
        L.orc_tau.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_dE.restype = C.c_double
        L.orc_dE.argtypes = [C.c_void_p]
        L.orc_energy.restype = C.c_double
        L.orc_energy.argtypes = [C.POINTER(Valley), _DP]
        L.orc_norm_wave_vec.restype = C.c_double
        L.orc_norm_wave_vec.argtypes = [C.POINTER(Valley), C.c_double]
        L.orc_velocity.argtypes = [C.POINTER(Valley), _DP, C.c_double, C.c_int, _DP]
        L.orc_to_ellipse.argtypes = [C.POINTER(Valley), C.c_int, _DP, _DP]
        L.orc_to_device.argtypes = [C.POINTER(Valley), C.c_int, _DP, _DP]
        L.orc_drift.argtypes = [C.POINTER(Valley), C.c_double, _DP, _DP, C.c_int, _DP, C.c_int, _DP]
        L.orc_uniform.restype = C.c_double
        L.orc_uniform.argtypes = [C.c_uint64, C.c_double, C.c_double]
        L.orc_energy_level.argtypes = [C.c_void_p, C.c_double]
        L.orc_select.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double]
        L.orc_mt_seed.argtypes = [C.POINTER(C.c_uint64), C.c_uint64]
        L.orc_mt_next.restype = C.c_uint64
        L.orc_mt_next.argtypes = [C.POINTER(C.c_uint64)]
        L.orc_mt_fill.argtypes = [C.c_uint64, C.POINTER(C.c_uint64), C.c_int64]
        L.orc_philox_draw.restype = C.c_uint64
        L.orc_philox_draw.argtypes = [C.c_uint64] * 4
        L.orc_generate_initial.restype = C.c_int64
        L.orc_generate_initial.argtypes = [C.c_void_p, _DP, _IP, C.c_double, C.POINTER(C.c_uint64),
                                           C.POINTER(EnsembleC), C.c_int64, C.POINTER(C.c_int64)]
        L.orc_bulk_steps.argtypes = [C.c_void_p, C.POINTER(EnsembleC), _DP, _DP, C.c_double, C.c_double,
                                     C.c_double, C.c_int, C.c_int64, C.POINTER(RngCfg), _DP, _IP, C.c_int64,
                                     C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_int64,
                                     C.POINTER(C.c_int64)]
        L.orc_bulk_observables.argtypes = [C.c_void_p, C.POINTER(EnsembleC), _DP, _DP]
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(_DP)


def _ip(a):
    return a.ctypes.data_as(_IP)


class Ensemble:
    """SoA ensemble held in numpy arrays (the layout shared with the GPU path)."""
    F64 = ("kx", "ky", "kz", "energy", "tau", "grainTau", "x", "y", "z")
    I32 = ("valley", "sub", "region")

    def __init__(self, capacity: int):
        self.n = 0
        for f in self.F64:
            setattr(self, f, np.zeros(capacity, dtype=np.float64))
        for f in self.I32:
            setattr(self, f, np.zeros(capacity, dtype=np.int32))

    @classmethod
    def from_arrays(cls, k, pos, energy, tau, grain_tau, idx):
        n = len(energy)
        e = cls(n)
        e.n = n
        e.kx[:], e.ky[:], e.kz[:] = k[:, 0], k[:, 1], k[:, 2]
        e.x[:], e.y[:], e.z[:] = pos[:, 0], pos[:, 1], pos[:, 2]
        e.energy[:], e.tau[:], e.grainTau[:] = energy, tau, grain_tau
        e.valley[:], e.sub[:], e.region[:] = idx[:, 0], idx[:, 1], idx[:, 2]
        return e

    def copy(self):
        o = Ensemble(len(self.kx))
        o.n = self.n
        for f in self.F64 + self.I32:
            getattr(o, f)[:] = getattr(self, f)
        return o

    def subset(self, mask):
        """the particles selected by a boolean mask over [0, n), as an ensemble of their own"""
        mask = np.asarray(mask, dtype=bool)[: self.n]
        o = Ensemble(max(1, int(mask.sum())))
        o.n = int(mask.sum())
        for f in self.F64 + self.I32:
            getattr(o, f)[: o.n] = getattr(self, f)[: self.n][mask]
        return o

    def trim(self):
        for f in self.F64 + self.I32:
            setattr(self, f, np.ascontiguousarray(getattr(self, f)[: self.n]))
        return self

    def c(self) -> EnsembleC:
        s = EnsembleC()
        s.n = self.n
        for f in self.F64:
            setattr(s, f, _dp(getattr(self, f)))
        for f in self.I32:
            setattr(s, f, _ip(getattr(self, f)))
        return s

    def packed(self) -> np.ndarray:
        """u32 {valley:8, sub:8, region:16}, the device packing."""
        return (self.valley.astype(np.uint32) | (self.sub.astype(np.uint32) << 8)
                | (self.region.astype(np.uint32) << 16))


class Model:
    def __init__(self, n_levels=1000, max_energy=1.0, temperature=300.0, rho=2329.0, v_sound=9040.0):
        self.L = lib()
        self.n_levels = n_levels
        self.max_energy = max_energy
        self.temperature = temperature
        self.h = C.c_void_p(self.L.orc_model_create(n_levels, max_energy, temperature, rho, v_sound))

    def __del__(self):
        try:
            self.L.orc_model_destroy(self.h)
        except Exception:
            pass

    def add_valley(self, kind, rel_mass, deg, alpha=0.0, e_bottom=0.0, dirs=None, particle_mass=ME, angles=None):
        """angles: the in-plane rotation angle of every sub-valley of the anisotropic single-layer class"""
        rm = np.ascontiguousarray(np.broadcast_to(np.asarray(rel_mass, dtype=np.float64), (3,)))
        d = None
        if angles is not None:
            dirs = np.zeros((deg, 3, 3))
            dirs[:, 0, 0] = angles
        if dirs is not None:
            d = np.ascontiguousarray(np.asarray(dirs, dtype=np.float64).reshape(deg, 3, 3))
        r = self.L.orc_add_valley(self.h, kind, _dp(rm), particle_mass, deg, alpha, e_bottom,
                                  _dp(d) if d is not None else None)
        assert r >= 0
        return r

    def add_acoustic(self, valley, region, sigma):
        return self.L.orc_add_acoustic(self.h, valley, region, sigma)

    def add_intervalley(self, order, emission, valley, final_valley, region, def_pot, phonon_energy, final_sub):
        fs = np.ascontiguousarray(np.asarray(final_sub, dtype=np.int32))
        r = self.L.orc_add_intervalley(self.h, order, int(emission), valley, final_valley, region, def_pot,
                                       phonon_energy, fs.shape[0], fs.shape[1], _ip(fs))
        assert r >= 0
        return r

    def add_acoustic_sl(self, valley, region, sigma, density_2d, v_sound):
        """emcAcousticSingleLayerMechanism"""
        return self.L.orc_add_acoustic_sl(self.h, valley, region, sigma, density_2d, v_sound)

    def add_intervalley_sl(self, emission, valley, final_valley, region, sigma, density_2d, phonon_energy, final_sub=None, order=0):
        """emc{Zero,First}OrderSingleLayerInterValley{Absorption,Emission}ScatterMechanism; final_sub None: the one-valley form"""
        if final_sub is None:
            r = self.L.orc_add_intervalley_sl(self.h, order, int(emission), valley, final_valley, region, sigma, density_2d,
                                              phonon_energy, 0, 0, None)
        else:
            fs = np.ascontiguousarray(np.asarray(final_sub, dtype=np.int32))
            r = self.L.orc_add_intervalley_sl(self.h, order, int(emission), valley, final_valley, region, sigma, density_2d,
                                              phonon_energy, fs.shape[0], fs.shape[1], _ip(fs))
        assert r >= 0
        return r

    def add_froehlich_sl(self, emission, valley, region, phonon_energy, coupling_const, width, qs=0.0):
        """emcFroehlichInteraction{Absorption,Emission}SL"""
        return self.L.orc_add_froehlich_sl(self.h, int(emission), valley, region, phonon_energy, coupling_const, width, qs)

    def add_piezo_sl(self, valley, region, piezo_const, width, density_2d, v_sound, qs=0.0):
        """emcPiezoelectricSingleLayerMechanism"""
        return self.L.orc_add_piezo_sl(self.h, valley, region, piezo_const, width, density_2d, v_sound, qs)

    def add_charged_impurity_sl(self, valley, region, impurity_density, eps_avg, qs, rytova_keldysh_length=0.0, remote_distance=0.0,
                                charge_number=1.0):
        """emc2DChargedImpurityScatterMechanism.hpp"""
        return self.L.orc_add_charged_impurity_sl(self.h, valley, region, impurity_density, eps_avg, qs, rytova_keldysh_length,
                                                  remote_distance, charge_number)

    def add_surface_roughness_sl(self, valley, region, effective_field, roughness_amplitude, correlation_length, qs):
        """emcSurfaceRoughnessScatterMechanism.hpp"""
        return self.L.orc_add_surface_roughness_sl(self.h, valley, region, effective_field, roughness_amplitude, correlation_length, qs)

    def add_remote_so_sl(self, emission, valley, region, phonon_energy, coupling_d, remote_distance, qs=0.0):
        """emcRemoteSurfaceOpticalPhononMechanism.hpp"""
        return self.L.orc_add_remote_so_sl(self.h, int(emission), valley, region, phonon_energy, coupling_d, remote_distance, qs)

    def add_screened_optical_sl(self, emission, valley, region, sigma, density_2d, phonon_energy, qs=0.0):
        """emcScreenedIntravalleyOpticalMechanism.hpp"""
        return self.L.orc_add_screened_optical_sl(self.h, int(emission), valley, region, sigma, density_2d, phonon_energy, qs)

    def add_coulomb(self, valley, region, eps_r, region_doping):
        return self.L.orc_add_coulomb(self.h, valley, region, eps_r, region_doping)

    def add_froehlich(self, variant, emission, valley, region, phonon_energy, rel_eff_mass, eps_hi, eps_lo, temperature=300.0,
                      bath=-1, q_resolved=False, q_resolved_angle=True):
        """variant: FROEHLICH_EQ / _HOT / _SCREENED_EQ / _SCREENED_HOT"""
        r = self.L.orc_add_froehlich(self.h, variant, int(emission), valley, region, phonon_energy, rel_eff_mass, eps_hi,
                                     eps_lo, temperature, bath, int(q_resolved), int(q_resolved_angle))
        assert r >= 0
        return r

    def add_bath(self, bath: "PhononBath"):
        self._baths = getattr(self, "_baths", []) + [bath]  # keep alive
        r = self.L.orc_model_add_bath(self.h, bath.h)
        assert r >= 0
        return r

    def set_grain(self, transmission_prob, scatter_rate):
        """emcGrainScatterMechanism(transmissionProb, scatterRate); rate <= 0 removes it"""
        self.grain = (transmission_prob, scatter_rate) if scatter_rate > 0 else None
        self.L.orc_model_set_grain(self.h, transmission_prob, scatter_rate)

    def set_qs2(self, qs2):
        self.L.orc_model_set_qs2(self.h, qs2)

    def build_tables(self):
        assert self.L.orc_build_tables(self.h) == 0

    @property
    def n_valleys(self):
        return self.L.orc_n_valleys(self.h)

    def valley(self, v) -> Valley:
        out = Valley()
        assert self.L.orc_get_valley(self.h, v, C.byref(out)) == 0
        return out

    def valleys(self):
        return [self.valley(v) for v in range(self.n_valleys)]

    @property
    def n_mechanisms(self):
        return self.L.orc_n_mechanisms(self.h)

    def raw_rates(self):
        dE = self.dE
        out = np.zeros((self.n_mechanisms, self.n_levels))
        for g in range(self.n_mechanisms):
            for lvl in range(self.n_levels):
                out[g, lvl] = self.L.orc_raw_rate(self.h, g, (lvl + 1) * dE)
        return out

    @property
    def dE(self):
        return self.L.orc_dE(self.h)

    def tau(self, valley, region):
        return self.L.orc_tau(self.h, valley, region)

    def tablesets(self):
        """list of dicts: valley, region, tau, cum [nMech][nLevels], mech (ctypes array)"""
        out = []
        for i in range(self.L.orc_n_tablesets(self.h)):
            v, r, nm = C.c_int32(), C.c_int32(), C.c_int32()
            tau = C.c_double()
            self.L.orc_tableset_info(self.h, i, C.byref(v), C.byref(r), C.byref(nm), C.byref(tau))
            cum = np.zeros((nm.value, self.n_levels))
            mech = (Mech * nm.value)()
            self.L.orc_tableset_copy(self.h, i, _dp(cum), mech)
            out.append(dict(valley=v.value, region=r.value, tau=tau.value, cum=cum, mech=mech))
        return out

    # ---- particle loop
    def bulk_observables(self, ens: "Ensemble", field_dir):
        """[valley][sum E, sum v.dir, count] of a resting ensemble (basicBulkParticleHandler.hpp:289-347)"""
        obs = np.zeros((self.n_valleys, 3))
        s = ens.c()
        self.L.orc_bulk_observables(self.h, C.byref(s), _dp(np.asarray(field_dir, dtype=np.float64)), _dp(obs))
        return obs

    def set_electron2d(self, per_grid_point=4):
        """initial ensemble of examples/singleLayerMoS2/electron2D.hpp"""
        self.L.orc_model_set_electron2d(self.h, per_grid_point)

    def set_init_energy(self, energy_ev):
        """mono-energetic initial ensemble (emcElectron / emcHole initEnergyEV)"""
        self.L.orc_model_set_init_energy(self.h, energy_ev)

    def generate_initial(self, box, cells, doping, mt_state, capacity=None):
        box = np.asarray(box, dtype=np.float64)
        cells = np.asarray(cells, dtype=np.int32)
        if capacity is None:
            capacity = int(doping * np.prod(box) * 1.05) + 64
        ens = Ensemble(capacity)
        s = ens.c()
        used = C.c_int64()
        n = self.L.orc_generate_initial(self.h, _dp(box), _ip(cells), doping, mt_state, C.byref(s), capacity,
                                        C.byref(used))
        assert n >= 0, "capacity too small"
        ens.n = int(n)
        return ens.trim(), used.value

    def bulk_steps(self, ens: Ensemble, box, field_dir, field_strength, dt, n_steps, rng: RngCfg,
                   first_step=0, charge=-Q, record=False, log_events=False):
        box = np.asarray(box, dtype=np.float64)
        fd = np.asarray(field_dir, dtype=np.float64)
        obs = np.zeros((n_steps, self.n_valleys, 3))
        s = ens.c()
        rec_cap = 0
        rec = None
        if record:
            rec_cap = max(1024, int(ens.n) * n_steps * 64 + 1024)
            rec = np.zeros(rec_cap, dtype=np.int32)
        ev_cap = 0
        ev = None
        if log_events:
            ev_cap = max(1024, int(ens.n) * n_steps * 4 + 1024)
            ev = np.zeros((ev_cap, 4), dtype=np.int64)
        rc, ec = C.c_int64(), C.c_int64()
        r = self.L.orc_bulk_steps(self.h, C.byref(s), _dp(box), _dp(fd), field_strength, charge, dt, n_steps,
                                  first_step, C.byref(rng), _dp(obs), _ip(rec) if rec is not None else None,
                                  rec_cap, C.byref(rc), ev.ctypes.data_as(C.POINTER(C.c_int64)) if ev is not None
                                  else None, ev_cap, C.byref(ec))
        assert r == 0
        res = dict(obs=obs, n_draws=rc.value, n_events=ec.value)
        if record:
            assert rc.value <= rec_cap
            res["rec_pid"] = rec[: rc.value]
        if log_events:
            assert ec.value <= ev_cap
            res["events"] = ev[: ec.value]
        return res


FROEHLICH_EQ, FROEHLICH_HOT, FROEHLICH_SCREENED_EQ, FROEHLICH_SCREENED_HOT = 0, 1, 2, 3


class PhononBath:
    """emcPhononBath (include/emcPhononBath.hpp)"""

    def __init__(self, n_bins, dq, tau_lo, phonon_energy, lattice_temp, v_sim, acoustic=False, ac_energy=0.0, tau_ac=0.0,
                 w_ridley=0.0, to_energy=0.0, tau_to=0.0):
        self.L = lib()
        self.n_bins, self.dq = n_bins, dq
        self.h = self.L.orc_bath_create(n_bins, dq, tau_lo, phonon_energy, lattice_temp, v_sim, int(acoustic), ac_energy,
                                        tau_ac, w_ridley, to_energy, tau_to)

    def __del__(self):
        try:
            self.L.orc_bath_destroy(self.h)
        except Exception:
            pass

    def set_qs2(self, qs2):
        self.L.orc_bath_set_qs2(self.h, qs2)

    def update(self, dt):
        self.L.orc_bath_update(self.h, dt)

    def mean_nq(self):
        return self.L.orc_bath_mean_nq(self.h)

    def acoustic_temp(self):
        return self.L.orc_bath_acoustic_temp(self.h)

    def n0(self):
        return self.L.orc_bath_n0(self.h)

    def nq_window(self, q_min, q_max):
        return self.L.orc_bath_nq_window(self.h, q_min, q_max)

    def sample_q(self, q_min, q_max, emission, r):
        return self.L.orc_bath_sample_q(self.h, q_min, q_max, int(emission), r)

    def _copy(self, which, n):
        out = np.zeros(n)
        assert self.L.orc_bath_copy(self.h, which, _dp(out)) == 0
        return out

    nq = property(lambda self: self._copy(0, self.n_bins))
    n_em = property(lambda self: self._copy(1, self.n_bins))
    n_abs = property(lambda self: self._copy(2, self.n_bins))
    cum_w = property(lambda self: self._copy(3, self.n_bins + 1))
    cum_wn = property(lambda self: self._copy(4, self.n_bins + 1))

    def add_counts(self, emission, absorption):
        em = np.ascontiguousarray(emission, dtype=np.int64)
        ab = np.ascontiguousarray(absorption, dtype=np.int64)
        self.L.orc_bath_add_counts(self.h, em.ctypes.data_as(C.POINTER(C.c_int64)), ab.ctypes.data_as(C.POINTER(C.c_int64)))


def plasmon_qs2(density, carrier_temp, eps_static):
    return lib().orc_plasmon_qs2(density, carrier_temp, eps_static)


def mt_state(seed: int):
    st = (C.c_uint64 * 313)()
    lib().orc_mt_seed(st, seed)
    return st


def mt_fill(seed: int, n: int) -> np.ndarray:
    out = np.zeros(n, dtype=np.uint64)
    lib().orc_mt_fill(seed, out.ctypes.data_as(C.POINTER(C.c_uint64)), n)
    return out


_limit_flags = None


def set_limit_flags(n):
    """test aid: a fresh uint8[n] that the q-resolved sampler marks per particle index when its |q| sample lies ON a
    kinematic limit (emc_oracle.c:g_limitFlags); n = 0 switches the marking off"""
    global _limit_flags
    _limit_flags = np.zeros(int(n), dtype=np.uint8) if n else None
    lib().orc_set_limit_flags(_limit_flags.ctypes.data if n else None)
    return _limit_flags


def rng_mt(state) -> RngCfg:
    r = RngCfg()
    r.mode = RNG_MT_GLOBAL
    r.mtState = C.cast(state, C.POINTER(C.c_uint64))
    return r


def rng_streams(draws: np.ndarray, offsets: np.ndarray, cursor: np.ndarray) -> RngCfg:
    r = RngCfg()
    r.mode = RNG_STREAMS
    r.draws = draws.ctypes.data_as(C.POINTER(C.c_uint64))
    r.offsets = offsets.ctypes.data_as(C.POINTER(C.c_int64))
    r.cursor = cursor.ctypes.data_as(C.POINTER(C.c_int64))
    r._keep = (draws, offsets, cursor)
    return r


def rng_philox(seed: int, particle_id_base: int = 0) -> RngCfg:
    r = RngCfg()
    r.mode = RNG_PHILOX
    r.philoxSeed = seed
    r.particleIdBase = particle_id_base
    return r


def streams_from_record(draws: np.ndarray, rec_pid: np.ndarray, n: int):
    """Turn a global draw log + per-draw particle attribution into per-particle
    replay streams (CSR): returns (draws_sorted, offsets[n+1])."""
    order = np.argsort(rec_pid, kind="stable")
    counts = np.bincount(rec_pid, minlength=n).astype(np.int64)
    offsets = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(counts, out=offsets[1:])
    return np.ascontiguousarray(draws[order]), offsets


# ======================================================================================
# Device-run path (orc_device_t and friends)
CONTACT_OHMIC, CONTACT_SCHOTTKY, CONTACT_GATE = range(3)
_I8P = C.POINTER(C.c_int8)


class DeviceC(C.Structure):
    _fields_ = [("dim", C.c_int32), ("extent", C.c_int32 * 3), ("spacing", C.c_double * 3),
                ("maxPos", C.c_double * 3), ("thermalVoltage", C.c_double), ("debyeLength", C.c_double),
                ("ni", C.c_double), ("cellVolume", C.c_double), ("epsR", C.c_double), ("nContacts", C.c_int32),
                ("contactType", _IP), ("contactVoltage", _DP), ("gateEpsOx", _DP), ("gateThickness", _DP),
                ("gateBarrier", _DP), ("region", _IP), ("faceContact", _I8P), ("doping", _DP),
                ("pmScheme", C.c_int32), ("electronKind", C.c_int32), ("surfaceKind", C.c_int32 * 6),
                ("surfaceParam", C.c_double * 6)]


PM_NGP, PM_CIC, PM_NEC, PM_NEC_VWD = 0, 1, 2, 3
SURFACE_SPECULAR, SURFACE_CONSTANT, SURFACE_MOMENTUM = 0, 1, 2
ELECTRON_EMC, ELECTRON_VWD = 0, 1


_DEV_BOUND = False


def _bind_device(L):
    global _DEV_BOUND
    if _DEV_BOUND:
        return
    dp = C.POINTER(DeviceC)
    L.orc_dev_cells.restype = C.c_int64
    L.orc_dev_cells.argtypes = [dp]
    L.orc_initial_potential.argtypes = [dp, _DP]
    L.orc_sor.argtypes = [dp, _DP, _DP, C.c_double, C.c_double, C.c_int, C.c_int]
    L.orc_efield.argtypes = [dp, _DP, _DP]
    L.orc_ngp_assign.argtypes = [dp, C.c_int64, _DP, _DP, _DP, C.c_double, _DP]
    L.orc_assign.argtypes = [dp, C.c_int64, _DP, _DP, _DP, C.c_double, _DP]
    L.orc_initial_nr_particles.restype = C.c_double
    L.orc_initial_nr_particles.argtypes = [dp, C.c_int64, _DP]
    L.orc_device_generate_initial_pot.restype = C.c_int64
    L.orc_device_generate_initial_pot.argtypes = [C.c_void_p, dp, C.c_double, C.POINTER(C.c_uint64), _DP,
                                                  C.POINTER(EnsembleC), C.c_int64]
    L.orc_concentration.argtypes = [dp, _DP, _DP]
    L.orc_expected_at_contact.argtypes = [dp, _DP]
    L.orc_device_generate_initial.restype = C.c_int64
    L.orc_device_generate_initial.argtypes = [C.c_void_p, dp, C.c_double, C.POINTER(C.c_uint64),
                                              C.POINTER(EnsembleC), C.c_int64]
    L.orc_device_step.argtypes = [C.c_void_p, dp, C.POINTER(EnsembleC), _DP, C.c_double, C.c_double, C.c_int64,
                                  C.POINTER(RngCfg), _I8P, _IP, _IP, C.c_int64, C.POINTER(C.c_int64),
                                  C.POINTER(C.c_int64), C.c_int64, C.POINTER(C.c_int64)]
    L.orc_compact.restype = C.c_int64
    L.orc_compact.argtypes = [C.POINTER(EnsembleC), _I8P]
    L.orc_contacts.restype = C.c_int64
    L.orc_contacts.argtypes = [C.c_void_p, dp, C.POINTER(EnsembleC), C.c_int64, _DP, C.c_double,
                               C.POINTER(C.c_uint64), _IP]
    _DEV_BOUND = True


KB, EPS0 = 1.38066e-23, 8.85419e-12


class Device:
    """Box device with doping regions and contacts, flattened the way the C ABI takes it
    (reference: emcDevice.hpp, emcDopingProfile.hpp, emcSurface.hpp)."""

    def __init__(self, max_pos, spacing, temperature=300.0, eps_r=11.8, ni=1.45e16, device_width=1e-6):
        self.L = lib()
        _bind_device(self.L)
        self.dim = len(max_pos)
        self.max_pos = [float(x) for x in max_pos]
        self.spacing = [float(x) for x in spacing]
        self.extent = [int(round(m / h)) + 1 for m, h in zip(max_pos, spacing)]  # emcUtil.hpp:109-119
        self.cells = int(np.prod(self.extent))
        self.vt = KB / Q * temperature  # emcDevice.hpp:87
        self.ni, self.eps_r = ni, eps_r
        self.debye = float(np.sqrt(EPS0 * eps_r * self.vt / Q / ni))  # emcDevice.hpp:88-90
        vol = 1.0
        for h in self.spacing:
            vol = vol * h
        if self.dim == 2:
            vol *= device_width
        self.cell_volume = vol
        self.doping = np.full(self.cells, ni, dtype=np.float64)
        self.region = np.full(self.cells, -1, dtype=np.int32)
        self.n_regions = 0
        self.face_contact = np.full((self.cells, 2 * self.dim), -2, dtype=np.int8)
        coords = self.coords()
        for d in range(self.dim):
            self.face_contact[coords[:, d] == 0, 2 * d] = -1
            self.face_contact[coords[:, d] == self.extent[d] - 1, 2 * d + 1] = -1
        self.contact_type, self.contact_voltage = [], []
        self.gate_eps, self.gate_thick, self.gate_barrier = [], [], []
        # plug-in variants: particle-mesh scheme, electron flavour, surface scatter mechanism per face
        self.pm_scheme = PM_NGP
        self.electron_kind = ELECTRON_EMC
        self.surface_kind = [SURFACE_SPECULAR] * 6
        self.surface_param = [0.0] * 6

    def coords(self):
        idx = np.arange(self.cells)
        out = np.zeros((self.cells, self.dim), dtype=np.int64)
        for d in range(self.dim):
            out[:, d] = idx % self.extent[d]
            idx = idx // self.extent[d]
        return out

    def _to_coord(self, pos, spacing):
        return int(np.floor(abs(pos / spacing) + 0.5) * (1 if pos >= 0 else -1))  # std::round

    def add_doping_region(self, lo, hi, doping):
        c = self.coords()
        inside = np.ones(self.cells, dtype=bool)
        for d in range(self.dim):
            a, b = self._to_coord(lo[d], self.spacing[d]), self._to_coord(hi[d], self.spacing[d])
            inside &= (c[:, d] >= min(a, b)) & (c[:, d] <= max(a, b))
        self.doping[inside] = doping
        self.region[inside] = self.n_regions
        self.n_regions += 1

    def add_contact(self, face, kind, voltage, lo, hi, eps_ox=0.0, thickness=0.0, barrier=0.0):
        """face: 0 XMIN, 1 XMAX, 2 YMIN, ...; lo/hi: positions along the face's own axes (emcDevice.hpp:178-215)"""
        idx = len(self.contact_type)
        fixed = face // 2
        axes = [d for d in range(self.dim) if d != fixed]
        c = self.coords()
        on_face = c[:, fixed] == (0 if face % 2 == 0 else self.extent[fixed] - 1)
        sel = on_face.copy()
        for a, l, h in zip(axes, lo, hi):
            ca, cb = self._to_coord(l, self.spacing[a]), self._to_coord(h, self.spacing[a])
            sel &= (c[:, a] >= ca) & (c[:, a] <= cb)
        self.face_contact[sel, face] = idx
        if kind in (CONTACT_OHMIC, CONTACT_SCHOTTKY):
            # emcSurface.hpp:351-381 updateAllOccurences: the contact also owns those cells on the other faces
            for f in range(2 * self.dim):
                on_f = self.face_contact[:, f] != -2
                self.face_contact[sel & on_f, f] = idx
        self.contact_type.append(kind)
        self.contact_voltage.append(float(voltage))
        self.gate_eps.append(float(eps_ox))
        self.gate_thick.append(float(thickness))
        self.gate_barrier.append(float(barrier))
        return idx

    def c(self) -> DeviceC:
        d = DeviceC()
        d.dim = self.dim
        for i in range(3):
            d.extent[i] = self.extent[i] if i < self.dim else 1
            d.spacing[i] = self.spacing[i] if i < self.dim else 1.0
            d.maxPos[i] = self.max_pos[i] if i < self.dim else 0.0
        d.thermalVoltage, d.debyeLength, d.ni, d.cellVolume, d.epsR = self.vt, self.debye, self.ni, self.cell_volume, self.eps_r
        d.nContacts = len(self.contact_type)
        self._keep = [np.asarray(self.contact_type, dtype=np.int32), np.asarray(self.contact_voltage, dtype=np.float64),
                      np.asarray(self.gate_eps, dtype=np.float64), np.asarray(self.gate_thick, dtype=np.float64),
                      np.asarray(self.gate_barrier, dtype=np.float64), np.ascontiguousarray(self.region),
                      np.ascontiguousarray(self.face_contact), np.ascontiguousarray(self.doping)]
        k = self._keep
        d.contactType, d.contactVoltage, d.gateEpsOx, d.gateThickness, d.gateBarrier = _ip(k[0]), _dp(k[1]), _dp(k[2]), _dp(k[3]), _dp(k[4])
        d.region, d.faceContact, d.doping = _ip(k[5]), k[6].ctypes.data_as(_I8P), _dp(k[7])
        d.pmScheme, d.electronKind = self.pm_scheme, self.electron_kind
        for f in range(6):
            d.surfaceKind[f] = self.surface_kind[f]
            d.surfaceParam[f] = self.surface_param[f]
        return d

    # ---- grid operations
    def initial_potential(self):
        pot = np.zeros(self.cells)
        self.L.orc_initial_potential(C.byref(self.c()), _dp(pot))
        return pot

    def sor(self, pot, conc=None, accuracy=1e-4, omega=1.8, reset_bc=True, max_sweeps=0):
        sweeps = self.L.orc_sor(C.byref(self.c()), _dp(pot), _dp(conc) if conc is not None else None, accuracy, omega,
                                int(reset_bc), max_sweeps)
        return sweeps

    def efield(self, pot):
        e = np.zeros((self.dim, self.cells))
        self.L.orc_efield(C.byref(self.c()), _dp(pot), _dp(e))
        return e

    def ngp_assign(self, ens, nr_carriers=1.0):
        count = np.zeros(self.cells)
        self.L.orc_ngp_assign(C.byref(self.c()), ens.n, _dp(ens.x), _dp(ens.y), _dp(ens.z), nr_carriers, _dp(count))
        return count

    def assign(self, ens, nr_carriers=1.0):
        """assignToMesh of the device's PM scheme"""
        count = np.zeros(self.cells)
        self.L.orc_assign(C.byref(self.c()), ens.n, _dp(ens.x), _dp(ens.y), _dp(ens.z), nr_carriers, _dp(count))
        return count

    def concentration(self, count):
        conc = np.zeros(self.cells)
        self.L.orc_concentration(C.byref(self.c()), _dp(count), _dp(conc))
        return conc

    def expected_at_contact(self):
        e = np.zeros(self.cells)
        self.L.orc_expected_at_contact(C.byref(self.c()), _dp(e))
        return e

    def generate_initial(self, model, mt, nr_carriers=1.0, capacity=None, pot=None):
        """pot: normalised potential for the potential-based initial density (always used by electronVWD)"""
        if capacity is None:
            dev = self.c()
            total = sum(self.L.orc_initial_nr_particles(C.byref(dev), i, _dp(pot) if pot is not None else None)
                        for i in range(self.cells))
            capacity = int(abs(total) * 1.2) + 1000
        cap = capacity
        ens = Ensemble(cap)
        c = ens.c()
        n = self.L.orc_device_generate_initial_pot(model.h, C.byref(self.c()), nr_carriers,
                                                   C.cast(mt, C.POINTER(C.c_uint64)),
                                                   _dp(pot) if pot is not None else None, C.byref(c), cap)
        assert n >= 0, "capacity too small"
        ens.n = int(n)
        return ens

    def step(self, model, ens, efield, dt, rng, step_index=1, charge=-Q, record=False, log_events=False):
        n = ens.n
        removed = np.zeros(max(1, n), dtype=np.int8)
        per_contact = np.zeros(max(1, len(self.contact_type)), dtype=np.int32)
        rec_cap = 64 * n + 1024 if record else 0
        rec = np.zeros(max(1, rec_cap), dtype=np.int32)
        rec_count = C.c_int64(0)
        ev_cap = 16 * n + 1024 if log_events else 0
        ev = np.zeros((max(1, ev_cap), 4), dtype=np.int64)
        ev_count = C.c_int64(0)
        cfg = rng
        c = ens.c()
        ef = np.ascontiguousarray(efield, dtype=np.float64)
        self.L.orc_device_step(model.h, C.byref(self.c()), C.byref(c), _dp(ef), charge, dt, step_index, C.byref(cfg),
                               removed.ctypes.data_as(_I8P), _ip(per_contact), _ip(rec) if record else None, rec_cap,
                               C.byref(rec_count), ev.ctypes.data_as(C.POINTER(C.c_int64)) if log_events else None, ev_cap,
                               C.byref(ev_count))
        out = dict(removed=removed[:n], removed_per_contact=per_contact[: len(self.contact_type)])
        if record:
            assert rec_count.value <= rec_cap
            out["rec_pid"] = rec[: rec_count.value]
        if log_events:
            out["events"] = ev[: ev_count.value]
        return out

    def compact(self, ens, removed):
        c = ens.c()
        ens.n = int(self.L.orc_compact(C.byref(c), np.ascontiguousarray(removed, dtype=np.int8).ctypes.data_as(_I8P)))
        return ens

    def contacts(self, model, ens, expected, mt, nr_carriers=1.0):
        net = np.zeros(max(1, len(self.contact_type)), dtype=np.int32)
        cap = len(ens.kx)
        c = ens.c()
        n = self.L.orc_contacts(model.h, C.byref(self.c()), C.byref(c), cap, _dp(expected), nr_carriers,
                                C.cast(mt, C.POINTER(C.c_uint64)), _ip(net))
        assert n >= 0, "ensemble capacity too small for the injected particles"
        ens.n = int(n)
        return net[: len(self.contact_type)]
