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
        L.orc_add_valley.argtypes = [C.c_void_p, C.c_int, _DP, C.c_double, C.c_int, C.c_double, C.c_double, _DP]
        L.orc_add_acoustic.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double]
        L.orc_add_intervalley.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                          C.c_double, C.c_int, C.c_int, _IP]
        L.orc_add_coulomb.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double]
        L.orc_build_tables.argtypes = [C.c_void_p]
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

    def add_valley(self, kind, rel_mass, deg, alpha=0.0, e_bottom=0.0, dirs=None, particle_mass=ME):
        rm = np.ascontiguousarray(np.broadcast_to(np.asarray(rel_mass, dtype=np.float64), (3,)))
        d = None
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

    def add_coulomb(self, valley, region, eps_r, region_doping):
        return self.L.orc_add_coulomb(self.h, valley, region, eps_r, region_doping)

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
            rec_cap = max(1024, int(ens.n) * n_steps * 8 + 1024)
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


def mt_state(seed: int):
    st = (C.c_uint64 * 313)()
    lib().orc_mt_seed(st, seed)
    return st


def mt_fill(seed: int, n: int) -> np.ndarray:
    out = np.zeros(n, dtype=np.uint64)
    lib().orc_mt_fill(seed, out.ctypes.data_as(C.POINTER(C.c_uint64)), n)
    return out


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
