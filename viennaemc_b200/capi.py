"""ctypes binding of libemcgpu.so (the C ABI declared in include/emcgpu.h).

This is the thin Python view of the drop-in boundary used by tests/, bench.py and
__graft_entry__.py.  It loads the in-tree CUDA library and fails loudly if it is
missing: there is no CPU fallback anywhere in the product path.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
# EMCGPU_LIB: developer override to A/B-test another build of the same library (never a CPU path)
LIB_PATH = os.environ.get("EMCGPU_LIB") or os.path.join(_PKG, "lib", "libemcgpu.so")

MAX_VALLEYS, MAX_SUB, MAX_FINAL, MAX_MECH_PER_SET, MAX_TABLESETS, NAME_LEN = 8, 8, 8, 32, 32, 48
N_STREAMS = 8
KX, KY, KZ, ENERGY, TAU, X, Y, Z = range(8)
STREAM_NAMES = ("kx", "ky", "kz", "energy", "tau", "x", "y", "z")

OK, E_INVALID, E_CUDA, E_UNSUPPORTED_MECHANISM, E_UNSUPPORTED_VALLEY, E_CAPACITY, E_REPLAY_EXHAUSTED = range(7)
SAMPLER_NONE, SAMPLER_ISOTROPIC_ELASTIC, SAMPLER_INTERVALLEY, SAMPLER_COULOMB = range(4)
MATH_EXACT, MATH_FAST = 0, 1
PM_NGP, PM_CIC, PM_NEC, PM_NEC_VWD = 0, 1, 2, 3
SURFACE_SPECULAR, SURFACE_CONSTANT, SURFACE_MOMENTUM_DEPENDENT = 0, 1, 2
PARTICLE_ELECTRON, PARTICLE_ELECTRON_VWD = 0, 1

_DP = C.POINTER(C.c_double)


class ValleyC(C.Structure):
    _fields_ = [("kind", C.c_int32), ("degeneracy", C.c_int32), ("effMassCond", C.c_double),
                ("effMassDOS", C.c_double), ("alpha", C.c_double), ("bottomEnergy", C.c_double),
                ("vogt", C.c_double * 3), ("rot", (C.c_double * 9) * MAX_SUB)]


class MechC(C.Structure):
    _fields_ = [("sampler", C.c_int32), ("finalValley", C.c_int32), ("nFinal", C.c_int32), ("mechId", C.c_int32),
                ("param", C.c_double * 4), ("finalSub", (C.c_uint8 * MAX_FINAL) * MAX_SUB),
                ("name", C.c_char * NAME_LEN)]


class TableSetC(C.Structure):
    _fields_ = [("valley", C.c_int32), ("region", C.c_int32), ("nMech", C.c_int32), ("reserved", C.c_int32),
                ("tau", C.c_double), ("cum", _DP), ("mech", C.POINTER(MechC))]


class EmcGpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"emcgpu error {code}: {msg}")
        self.code = code


_lib = None


def load():
    """Load libemcgpu.so; raises if the CUDA extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -m viennaemc_b200.build` "
                          "(the product path has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.emcgpu_abi_version.restype = C.c_int
    L.emcgpu_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.emcgpu_destroy.argtypes = [vp]
    L.emcgpu_destroy.restype = None
    L.emcgpu_last_error.argtypes = [vp]
    L.emcgpu_last_error.restype = C.c_char_p
    L.emcgpu_launch_count.argtypes = [vp]
    L.emcgpu_launch_count.restype = C.c_int64
    L.emcgpu_set_stream.argtypes = [vp, vp]
    L.emcgpu_synchronize.argtypes = [vp]
    L.emcgpu_set_option.argtypes = [vp, C.c_char_p, C.c_int64]
    L.emcgpu_set_valleys.argtypes = [vp, C.POINTER(ValleyC), C.c_int]
    L.emcgpu_set_tables.argtypes = [vp, C.POINTER(TableSetC), C.c_int, C.c_int, C.c_double]
    L.emcgpu_set_grain.argtypes = [vp, C.c_double, C.c_double]
    L.emcgpu_set_grain_clock.argtypes = [vp, _DP]
    L.emcgpu_get_grain_clock.argtypes = [vp, _DP]
    L.emcgpu_set_phonon_baths.argtypes = [vp, C.c_int, C.c_int, C.c_double, _DP, _DP]
    L.emcgpu_get_phonon_counts.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_int]
    L.emcgpu_set_ensemble.argtypes = [vp, C.c_int64, C.POINTER(_DP), C.POINTER(C.c_uint32), C.c_int64]
    L.emcgpu_get_ensemble.argtypes = [vp, C.POINTER(_DP), C.POINTER(C.c_uint32)]
    L.emcgpu_ensemble_size.argtypes = [vp]
    L.emcgpu_ensemble_size.restype = C.c_int64
    L.emcgpu_generate_bulk_ensemble.argtypes = [vp, C.c_int64, _DP, C.c_double, C.c_int32, C.c_uint64, C.c_int64]
    L.emcgpu_ensemble_device_ptrs.argtypes = [vp, C.POINTER(_DP), C.POINTER(C.POINTER(C.c_uint32))]
    L.emcgpu_rng_philox.argtypes = [vp, C.c_uint64]
    L.emcgpu_rng_replay.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_int64), C.c_int64]
    L.emcgpu_bulk_configure.argtypes = [vp, _DP, _DP, C.c_double, C.c_double, C.c_int]
    L.emcgpu_bulk_step.argtypes = [vp, C.c_double, C.c_int, C.c_int, _DP]
    L.emcgpu_bulk_run_host.argtypes = [vp, C.c_int64, C.POINTER(_DP), C.POINTER(C.c_uint32), C.c_int64, C.c_double,
                                       C.c_int, C.c_int, C.c_int64, _DP]
    L.emcgpu_bulk_step_device.argtypes = [vp, C.c_double, C.c_int, C.c_int, vp]
    L.emcgpu_bulk_step_ahead.argtypes = [vp, C.c_double, C.c_int, C.c_int, _DP]
    L.emcgpu_bulk_rewind.argtypes = [vp]
    L.emcgpu_kernel_times.argtypes = [vp, _DP, C.POINTER(C.c_int64), C.c_int]
    L.emcgpu_bulk_record_velocities.argtypes = [vp, C.c_int, _DP, C.c_int64]
    L.emcgpu_bulk_observables.argtypes = [vp, _DP]
    L.emcgpu_set_step_index.argtypes = [vp, C.c_int64]
    L.emcgpu_get_step_index.argtypes = [vp]
    L.emcgpu_get_step_index.restype = C.c_int64
    L.emcgpu_event_log_enable.argtypes = [vp, C.c_int64]
    L.emcgpu_event_log_read.argtypes = [vp, C.POINTER(C.c_int64), C.c_int64]
    L.emcgpu_event_log_read.restype = C.c_int64
    IP32 = C.POINTER(C.c_int32)
    L.emcgpu_device_configure.argtypes = [vp, C.POINTER(DeviceC), C.c_double, C.c_double, _DP, C.c_int]
    L.emcgpu_device_set_surface.argtypes = [vp, C.c_int, C.c_int, C.c_double]
    L.emcgpu_device_set_particle_kind.argtypes = [vp, C.c_int]
    L.emcgpu_device_set_sharding.argtypes = [vp, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.emcgpu_device_set_grid.argtypes = [vp, C.c_int, _DP]
    L.emcgpu_device_get_grid.argtypes = [vp, C.c_int, _DP]
    L.emcgpu_device_reserve.argtypes = [vp, C.c_int64]
    L.emcgpu_device_poisson.argtypes = [vp, C.c_int, C.c_double, C.c_double, C.c_int, IP32]
    L.emcgpu_device_efield.argtypes = [vp]
    L.emcgpu_device_assign.argtypes = [vp]
    L.emcgpu_device_concentration.argtypes = [vp]
    L.emcgpu_device_step.argtypes = [vp, C.c_double, IP32]
    L.emcgpu_device_contacts.argtypes = [vp, IP32, C.POINTER(C.c_uint64), C.c_int64]
    L.emcgpu_device_run.argtypes = [vp, C.c_double, C.c_int, C.c_double, C.c_double, C.c_int, IP32, IP32]
    L.emcgpu_device_run_averaging.argtypes = [vp, C.c_double, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, IP32,
                                              IP32]
    _lib = L
    return L


class DeviceC(C.Structure):
    _fields_ = [("dim", C.c_int32), ("nContacts", C.c_int32), ("extent", C.c_int32 * 3), ("pmScheme", C.c_int32),
                ("spacing", C.c_double * 3), ("maxPos", C.c_double * 3), ("thermalVoltage", C.c_double),
                ("debyeLength", C.c_double), ("ni", C.c_double), ("cellVolume", C.c_double), ("epsR", C.c_double),
                ("contactType", C.POINTER(C.c_int32)), ("contactVoltage", _DP), ("gateEpsOx", _DP),
                ("gateThickness", _DP), ("gateBarrier", _DP), ("region", C.POINTER(C.c_int32)),
                ("faceContact", C.POINTER(C.c_int8)), ("doping", _DP)]


(GRID_POTENTIAL, GRID_CONCENTRATION, GRID_COUNT, GRID_EFIELD_X, GRID_EFIELD_Y, GRID_EFIELD_Z,
 GRID_EXPECTED, GRID_SUM_POTENTIAL, GRID_SUM_CONCENTRATION) = range(9)
CONTACT_OHMIC, CONTACT_SCHOTTKY, CONTACT_GATE = range(3)

EXPORTED_SYMBOLS = [
    "emcgpu_abi_version", "emcgpu_create", "emcgpu_destroy", "emcgpu_last_error", "emcgpu_launch_count",
    "emcgpu_set_stream", "emcgpu_synchronize", "emcgpu_set_option", "emcgpu_set_valleys", "emcgpu_set_tables", "emcgpu_set_grain", "emcgpu_set_grain_clock", "emcgpu_get_grain_clock", "emcgpu_set_phonon_baths", "emcgpu_get_phonon_counts", "emcgpu_set_ensemble",
    "emcgpu_get_ensemble", "emcgpu_ensemble_size", "emcgpu_generate_bulk_ensemble",
    "emcgpu_ensemble_device_ptrs", "emcgpu_rng_philox", "emcgpu_rng_replay", "emcgpu_bulk_configure",
    "emcgpu_bulk_step", "emcgpu_bulk_run_host", "emcgpu_bulk_step_device", "emcgpu_bulk_step_ahead", "emcgpu_bulk_rewind", "emcgpu_kernel_times", "emcgpu_bulk_record_velocities", "emcgpu_bulk_observables", "emcgpu_set_step_index",
    "emcgpu_get_step_index", "emcgpu_event_log_enable", "emcgpu_event_log_read",
    "emcgpu_device_configure", "emcgpu_device_set_surface", "emcgpu_device_set_particle_kind", "emcgpu_device_set_sharding", "emcgpu_device_set_grid", "emcgpu_device_get_grid", "emcgpu_device_reserve",
    "emcgpu_device_poisson", "emcgpu_device_efield", "emcgpu_device_assign", "emcgpu_device_concentration",
    "emcgpu_device_step", "emcgpu_device_contacts", "emcgpu_device_run", "emcgpu_device_run_averaging",
]


def make_valley(kind, degeneracy, eff_mass_cond, eff_mass_dos, alpha, bottom_energy, vogt, rot) -> ValleyC:
    v = ValleyC()
    v.kind, v.degeneracy = int(kind), int(degeneracy)
    v.effMassCond, v.effMassDOS, v.alpha, v.bottomEnergy = eff_mass_cond, eff_mass_dos, alpha, bottom_energy
    for i in range(3):
        v.vogt[i] = vogt[i]
    rot = np.asarray(rot, dtype=np.float64).reshape(-1, 9)
    for s in range(MAX_SUB):
        for j in range(9):
            v.rot[s][j] = rot[s][j] if s < len(rot) else (1.0 if j in (0, 4, 8) else 0.0)
    return v


def make_mech(sampler, name="", mech_id=0, final_valley=0, final_sub=None, params=()) -> MechC:
    m = MechC()
    m.sampler, m.finalValley, m.mechId = int(sampler), int(final_valley), int(mech_id)
    m.name = name.encode()[: NAME_LEN - 1]
    for i, p in enumerate(params):
        m.param[i] = p
    if final_sub is not None:
        fs = np.asarray(final_sub)
        m.nFinal = fs.shape[1]
        for s in range(fs.shape[0]):
            for f in range(fs.shape[1]):
                m.finalSub[s][f] = int(fs[s, f])
    return m


class Context:
    """One GPU context (one per process / per GPU)."""

    def __init__(self, device: int = 0):
        self.L = load()
        h = C.c_void_p()
        rc = self.L.emcgpu_create(device, C.byref(h))
        if rc != OK:
            raise EmcGpuError(rc, self.L.emcgpu_last_error(None).decode())
        self.h = h
        self.device = int(device)
        self.n_valleys = 0
        self._keep = []

    def close(self):
        if getattr(self, "h", None):
            self.L.emcgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc != OK:
            raise EmcGpuError(rc, self.L.emcgpu_last_error(self.h).decode())

    # -- model
    def set_valleys(self, valleys):
        arr = (ValleyC * len(valleys))(*valleys)
        self._chk(self.L.emcgpu_set_valleys(self.h, arr, len(valleys)))
        self.n_valleys = len(valleys)

    def set_tables(self, sets, n_levels, max_energy):
        """sets: list of dicts(valley, region, tau, cum ndarray [nMech][nLevels], mech list[MechC])"""
        arr = (TableSetC * max(1, len(sets)))()
        keep = []
        for i, s in enumerate(sets):
            cum = np.ascontiguousarray(s["cum"], dtype=np.float64)
            mech = (MechC * len(s["mech"]))(*s["mech"])
            keep += [cum, mech]
            arr[i].valley, arr[i].region, arr[i].nMech, arr[i].tau = s["valley"], s["region"], cum.shape[0], s["tau"]
            arr[i].cum = cum.ctypes.data_as(_DP)
            arr[i].mech = mech
        self._chk(self.L.emcgpu_set_tables(self.h, arr, len(sets), int(n_levels), float(max_energy)))

    # -- ensemble
    def set_grain(self, transmission_prob, scatter_rate):
        self._chk(self.L.emcgpu_set_grain(self.h, transmission_prob, scatter_rate))
        self.grain_on = scatter_rate > 0

    def set_grain_clock(self, grain_tau):
        g = np.ascontiguousarray(grain_tau, dtype=np.float64)
        assert g.size == self.size
        self._chk(self.L.emcgpu_set_grain_clock(self.h, g.ctypes.data_as(_DP)))

    def get_grain_clock(self):
        g = np.zeros(self.size)
        self._chk(self.L.emcgpu_get_grain_clock(self.h, g.ctypes.data_as(_DP)))
        return g

    def set_phonon_baths(self, n_baths, n_bins, dq, cum_w=None, cum_wn=None):
        """cum_w / cum_wn: [n_baths][n_bins + 1] prefix sums of emcPhononBath (only needed for q-resolved angles)"""
        cw = np.ascontiguousarray(cum_w, dtype=np.float64) if cum_w is not None else None
        cwn = np.ascontiguousarray(cum_wn, dtype=np.float64) if cum_wn is not None else None
        self._chk(self.L.emcgpu_set_phonon_baths(self.h, n_baths, n_bins, dq, cw.ctypes.data_as(_DP) if cw is not None else None,
                                                 cwn.ctypes.data_as(_DP) if cwn is not None else None))
        self._bath_shape = (n_baths, n_bins)

    def get_phonon_counts(self, reset=True):
        nb, bins = self._bath_shape
        em = np.zeros((nb, bins), dtype=np.int64)
        ab = np.zeros((nb, bins), dtype=np.int64)
        self._chk(self.L.emcgpu_get_phonon_counts(self.h, em.ctypes.data_as(C.POINTER(C.c_int64)),
                                                  ab.ctypes.data_as(C.POINTER(C.c_int64)), int(reset)))
        return em, ab

    def set_ensemble(self, streams, packed, particle_id_base=0):
        streams = [np.ascontiguousarray(a, dtype=np.float64) for a in streams]
        packed = np.ascontiguousarray(packed, dtype=np.uint32)
        n = len(packed)
        ptrs = (_DP * N_STREAMS)(*[a.ctypes.data_as(_DP) for a in streams])
        self._chk(self.L.emcgpu_set_ensemble(self.h, n, ptrs, packed.ctypes.data_as(C.POINTER(C.c_uint32)),
                                             particle_id_base))

    def get_ensemble(self):
        n = self.size
        streams = [np.empty(n, dtype=np.float64) for _ in range(N_STREAMS)]
        packed = np.empty(n, dtype=np.uint32)
        ptrs = (_DP * N_STREAMS)(*[a.ctypes.data_as(_DP) for a in streams])
        self._chk(self.L.emcgpu_get_ensemble(self.h, ptrs, packed.ctypes.data_as(C.POINTER(C.c_uint32))))
        return streams, packed

    def set_ensemble_from(self, streams, packed, particle_id_base=0):
        """upload from caller-owned host arrays WITHOUT copying them first (e.g. views of pinned memory)"""
        n = len(packed)
        assert all(a.dtype == np.float64 and a.flags.c_contiguous and len(a) == n for a in streams)
        assert packed.dtype == np.uint32 and packed.flags.c_contiguous
        ptrs = (_DP * N_STREAMS)(*[a.ctypes.data_as(_DP) for a in streams])
        self._chk(self.L.emcgpu_set_ensemble(self.h, n, ptrs, packed.ctypes.data_as(C.POINTER(C.c_uint32)),
                                             particle_id_base))

    def get_ensemble_into(self, streams, packed):
        """download into caller-owned host arrays (e.g. views of pinned memory)"""
        n = self.size
        assert all(a.dtype == np.float64 and a.flags.c_contiguous and len(a) == n for a in streams)
        assert packed.dtype == np.uint32 and packed.flags.c_contiguous and len(packed) == n
        ptrs = (_DP * N_STREAMS)(*[a.ctypes.data_as(_DP) for a in streams])
        self._chk(self.L.emcgpu_get_ensemble(self.h, ptrs, packed.ctypes.data_as(C.POINTER(C.c_uint32))))

    @property
    def size(self):
        return int(self.L.emcgpu_ensemble_size(self.h))

    def generate_bulk_ensemble(self, n, box, temperature=300.0, region=0, seed=12345, particle_id_base=0):
        b = (C.c_double * 3)(*box)
        self._chk(self.L.emcgpu_generate_bulk_ensemble(self.h, n, b, temperature, region, seed, particle_id_base))

    def device_ptrs(self):
        ptrs = (_DP * N_STREAMS)()
        packed = C.POINTER(C.c_uint32)()
        self._chk(self.L.emcgpu_ensemble_device_ptrs(self.h, ptrs, C.byref(packed)))
        return [C.cast(p, C.c_void_p).value for p in ptrs], C.cast(packed, C.c_void_p).value

    # -- rng
    def rng_philox(self, seed):
        self._chk(self.L.emcgpu_rng_philox(self.h, seed))

    def rng_replay(self, draws, offsets):
        draws = np.ascontiguousarray(draws, dtype=np.uint64)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        self._chk(self.L.emcgpu_rng_replay(self.h, draws.ctypes.data_as(C.POINTER(C.c_uint64)),
                                           offsets.ctypes.data_as(C.POINTER(C.c_int64)), len(offsets) - 1))

    # -- bulk
    def bulk_configure(self, box, field_dir, field_strength, charge=-1.60219e-19, math_mode=MATH_EXACT):
        b = (C.c_double * 3)(*box)
        d = (C.c_double * 3)(*field_dir)
        self._chk(self.L.emcgpu_bulk_configure(self.h, b, d, field_strength, charge, math_mode))

    def bulk_step(self, dt, n_steps=1, steps_per_launch=1, want_obs=True):
        obs = np.zeros((n_steps, self.n_valleys, 3)) if want_obs else None
        self._chk(self.L.emcgpu_bulk_step(self.h, dt, n_steps, steps_per_launch,
                                          obs.ctypes.data_as(_DP) if want_obs else None))
        return obs

    def bulk_step_ahead(self, dt, n_steps, steps_per_launch):
        """like bulk_step, and the ensemble as it was before the call survives (see bulk_rewind)"""
        obs = np.zeros((n_steps, self.n_valleys, 3))
        self._chk(self.L.emcgpu_bulk_step_ahead(self.h, dt, n_steps, steps_per_launch, obs.ctypes.data_as(_DP)))
        return obs

    def bulk_rewind(self):
        self._chk(self.L.emcgpu_bulk_rewind(self.h))

    def record_velocities(self, components, n_steps=0, n_particles=0):
        """per-particle velocities of the steps of the following step calls; returns the host array they land in"""
        if not components:
            self._chk(self.L.emcgpu_bulk_record_velocities(self.h, 0, None, 0))
            self._vel = None
            return None
        self._vel = np.zeros((n_steps, n_particles, components))
        self._chk(self.L.emcgpu_bulk_record_velocities(self.h, components, self._vel.ctypes.data_as(_DP), n_steps))
        return self._vel

    def kernel_times(self, reset=True):
        """(ms, launches) of the flight kernel, the event kernel and the other bulk kernels (option kernel_timing)"""
        ms = np.zeros(3)
        n = np.zeros(3, dtype=np.int64)
        self._chk(self.L.emcgpu_kernel_times(self.h, ms.ctypes.data_as(_DP), n.ctypes.data_as(C.POINTER(C.c_int64)),
                                             1 if reset else 0))
        return ms, n

    def bulk_run_host(self, streams, packed, dt, n_steps, steps_per_launch=8, slice_particles=0, particle_id_base=0,
                      want_obs=True):
        """advance a HOST-resident ensemble in place, slice by slice (copies overlap the step kernels when the
        arrays are views of pinned memory)"""
        n = len(packed)
        assert all(a.dtype == np.float64 and a.flags.c_contiguous and len(a) == n for a in streams)
        assert packed.dtype == np.uint32 and packed.flags.c_contiguous
        ptrs = (_DP * N_STREAMS)(*[a.ctypes.data_as(_DP) for a in streams])
        obs = np.zeros((n_steps, self.n_valleys, 3)) if want_obs else None
        self._chk(self.L.emcgpu_bulk_run_host(self.h, n, ptrs, packed.ctypes.data_as(C.POINTER(C.c_uint32)),
                                              particle_id_base, dt, n_steps, steps_per_launch, slice_particles,
                                              obs.ctypes.data_as(_DP) if want_obs else None))
        return obs

    def bulk_step_device(self, dt, n_steps, steps_per_launch, obs_device_ptr):
        self._chk(self.L.emcgpu_bulk_step_device(self.h, dt, n_steps, steps_per_launch, obs_device_ptr))

    def bulk_observables(self):
        obs = np.zeros((self.n_valleys, 3))
        self._chk(self.L.emcgpu_bulk_observables(self.h, obs.ctypes.data_as(_DP)))
        return obs

    # -- device run
    def device_configure(self, dim, extent, spacing, max_pos, thermal_voltage, debye_length, ni, cell_volume, eps_r,
                         contact_type, contact_voltage, gate_eps, gate_thickness, gate_barrier, region, face_contact,
                         doping, charge=-1.60219e-19, nr_carriers=1.0, expected=None, math_mode=MATH_EXACT, pm_scheme=PM_NGP):
        d = DeviceC()
        d.dim, d.nContacts, d.pmScheme = dim, len(contact_type), pm_scheme
        for i in range(3):
            d.extent[i] = extent[i] if i < dim else 1
            d.spacing[i] = spacing[i] if i < dim else 1.0
            d.maxPos[i] = max_pos[i] if i < dim else 0.0
        d.thermalVoltage, d.debyeLength, d.ni, d.cellVolume, d.epsR = thermal_voltage, debye_length, ni, cell_volume, eps_r
        keep = [np.ascontiguousarray(contact_type, dtype=np.int32), np.ascontiguousarray(contact_voltage, dtype=np.float64),
                np.ascontiguousarray(gate_eps, dtype=np.float64), np.ascontiguousarray(gate_thickness, dtype=np.float64),
                np.ascontiguousarray(gate_barrier, dtype=np.float64), np.ascontiguousarray(region, dtype=np.int32),
                np.ascontiguousarray(face_contact, dtype=np.int8), np.ascontiguousarray(doping, dtype=np.float64)]
        d.contactType = keep[0].ctypes.data_as(C.POINTER(C.c_int32))
        d.contactVoltage, d.gateEpsOx, d.gateThickness, d.gateBarrier = [k.ctypes.data_as(_DP) for k in keep[1:5]]
        d.region = keep[5].ctypes.data_as(C.POINTER(C.c_int32))
        d.faceContact = keep[6].ctypes.data_as(C.POINTER(C.c_int8))
        d.doping = keep[7].ctypes.data_as(_DP)
        exp = np.ascontiguousarray(expected, dtype=np.float64) if expected is not None else None
        self._chk(self.L.emcgpu_device_configure(self.h, C.byref(d), charge, nr_carriers,
                                                 exp.ctypes.data_as(_DP) if exp is not None else None, math_mode))
        self.n_cells = int(np.prod(extent[:dim]))
        self.n_contacts = len(contact_type)

    def device_set_surface(self, face, kind, parameter):
        self._chk(self.L.emcgpu_device_set_surface(self.h, face, kind, parameter))

    def device_set_sharding(self, rank, world, callback, user=None):
        """callback: ctypes CFUNCTYPE(None, c_void_p user, c_void_p deviceBuffer, c_int64 count, c_void_p stream), or the
        address of a C function with that signature (libemcnccl: emcnccl_allreduce_sum_f64, user = the communicator)"""
        self._sharding_cb = callback
        fn = callback if isinstance(callback, (int, C.c_void_p)) or callback is None else C.cast(callback, C.c_void_p)
        self._chk(self.L.emcgpu_device_set_sharding(self.h, rank, world, fn, user))

    def device_set_particle_kind(self, kind):
        self._chk(self.L.emcgpu_device_set_particle_kind(self.h, kind))

    def device_set_grid(self, grid, values):
        v = np.ascontiguousarray(values, dtype=np.float64).ravel()
        assert v.size == self.n_cells
        self._chk(self.L.emcgpu_device_set_grid(self.h, grid, v.ctypes.data_as(_DP)))

    def device_get_grid(self, grid):
        v = np.zeros(self.n_cells)
        self._chk(self.L.emcgpu_device_get_grid(self.h, grid, v.ctypes.data_as(_DP)))
        return v

    def device_reserve(self, capacity):
        self._chk(self.L.emcgpu_device_reserve(self.h, capacity))

    def device_poisson(self, equilibrium, accuracy=1e-4, omega=1.8, reset_bc=True):
        sweeps = C.c_int32(0)
        self._chk(self.L.emcgpu_device_poisson(self.h, int(equilibrium), accuracy, omega, int(reset_bc), C.byref(sweeps)))
        return sweeps.value

    def device_efield(self):
        self._chk(self.L.emcgpu_device_efield(self.h))

    def device_assign(self):
        self._chk(self.L.emcgpu_device_assign(self.h))

    def device_concentration(self):
        self._chk(self.L.emcgpu_device_concentration(self.h))

    def device_step(self, dt):
        rem = np.zeros(max(1, self.n_contacts), dtype=np.int32)
        self._chk(self.L.emcgpu_device_step(self.h, dt, rem.ctypes.data_as(C.POINTER(C.c_int32))))
        return rem[: self.n_contacts]

    def device_contacts(self, replay_draws=None):
        net = np.zeros(max(1, self.n_contacts), dtype=np.int32)
        if replay_draws is not None:
            rd = np.ascontiguousarray(replay_draws, dtype=np.uint64)
            self._chk(self.L.emcgpu_device_contacts(self.h, net.ctypes.data_as(C.POINTER(C.c_int32)),
                                                    rd.ctypes.data_as(C.POINTER(C.c_uint64)), len(rd)))
        else:
            self._chk(self.L.emcgpu_device_contacts(self.h, net.ctypes.data_as(C.POINTER(C.c_int32)), None, 0))
        return net[: self.n_contacts]

    def device_run(self, dt, n_steps, accuracy=1e-4, omega=1.8, reset_bc_first=True, n_average=0):
        counters = np.zeros((n_steps, 2, max(1, self.n_contacts)), dtype=np.int32)
        sweeps = np.zeros(n_steps, dtype=np.int32)
        self._chk(self.L.emcgpu_device_run_averaging(self.h, dt, n_steps, n_average, accuracy, omega, int(reset_bc_first),
                                           counters.ctypes.data_as(C.POINTER(C.c_int32)),
                                           sweeps.ctypes.data_as(C.POINTER(C.c_int32))))
        return counters[:, :, : self.n_contacts], sweeps

    def set_step_index(self, s):
        self._chk(self.L.emcgpu_set_step_index(self.h, s))

    @property
    def step_index(self):
        return int(self.L.emcgpu_get_step_index(self.h))

    def set_stream(self, cuda_stream_ptr):
        self._chk(self.L.emcgpu_set_stream(self.h, cuda_stream_ptr))

    def set_option(self, name, value):
        self._chk(self.L.emcgpu_set_option(self.h, name.encode(), int(value)))

    def synchronize(self):
        self._chk(self.L.emcgpu_synchronize(self.h))

    @property
    def launch_count(self):
        return int(self.L.emcgpu_launch_count(self.h))

    def event_log_enable(self, capacity):
        self._chk(self.L.emcgpu_event_log_enable(self.h, capacity))

    def event_log_read(self, capacity):
        out = np.zeros((capacity, 4), dtype=np.int64)
        n = self.L.emcgpu_event_log_read(self.h, out.ctypes.data_as(C.POINTER(C.c_int64)), capacity)
        if n < 0:
            raise EmcGpuError(E_CUDA, self.L.emcgpu_last_error(self.h).decode())
        return out[: min(n, capacity)], int(n)
