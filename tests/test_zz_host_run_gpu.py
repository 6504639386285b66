"""Host-resident bulk runs (emcgpu_bulk_run_host) against resident runs of the same ensemble."""
import numpy as np
import pytest

from helpers import (STATE_RTOL, assert_state_close, download_ensemble, field_dir_of, golden_ensemble, load_golden,
                     upload_ensemble, upload_model)
from oracle import pyoracle as po
from scenarios import GOLDEN_CASES, build_model, build_si
from viennaemc_b200 import capi

pytestmark = pytest.mark.gpu


@pytest.mark.gpu
@pytest.mark.parametrize("slice_particles", [0, 3000, 4096, 20000], ids=["auto", "ragged", "even", "one_slice"])
def test_host_resident_run_equals_resident_run(gpu_ctx_factory, slice_particles):
    """emcgpu_bulk_run_host (ensemble in host memory, advanced slice by slice with the copies overlapping the step
    kernels) = emcgpu_set_ensemble + emcgpu_bulk_step + emcgpu_get_ensemble: particle states bit for bit for every
    slice size (Philox key = global particle id), observables the same sums, step index advanced, the context's
    resident ensemble untouched."""
    m = build_si()
    box = [5e-7] * 3
    ens, _ = m.generate_initial(box, [5, 5, 5], 1e23, po.mt_state(5))

    def fresh():
        ctx = gpu_ctx_factory()
        ctx.set_option("multi_kernel", 2)
        upload_model(ctx, m)
        ctx.rng_philox(42)
        ctx.bulk_configure(box, [-1, 0, 0], 1e6, math_mode=capi.MATH_FAST)
        ctx.set_step_index(1)
        return ctx

    ctx = fresh()
    upload_ensemble(ctx, ens, particle_id_base=7)
    obs_ref = ctx.bulk_step(1e-15, 40, 8)
    ref = download_ensemble(ctx)

    ctx = fresh()
    keep = ens.copy()
    for f in po.Ensemble.F64 + po.Ensemble.I32:
        setattr(keep, f, getattr(ens, f)[:100].copy())
    keep.n = 100
    upload_ensemble(ctx, keep)  # a resident ensemble that the host run must leave alone
    streams = [np.ascontiguousarray(a[: ens.n], dtype=np.float64).copy()
               for a in (ens.kx, ens.ky, ens.kz, ens.energy, ens.tau, ens.x, ens.y, ens.z)]
    packed = ens.packed().copy()
    obs = ctx.bulk_run_host(streams, packed, 1e-15, 40, 8, slice_particles, particle_id_base=7)
    assert ctx.step_index == 41
    for got, f in zip(streams, ("kx", "ky", "kz", "energy", "tau", "x", "y", "z")):
        assert np.array_equal(got, getattr(ref, f)[: ens.n]), f
    assert np.array_equal(packed, ref.packed())
    assert np.array_equal(obs[:, :, 2], obs_ref[:, :, 2])
    assert np.allclose(obs, obs_ref, rtol=1e-12)
    assert ctx.size == 100
    still = download_ensemble(ctx)
    assert np.array_equal(still.kx, keep.kx[:100]) and np.array_equal(still.energy, keep.energy[:100])


@pytest.mark.gpu
def test_host_resident_run_rejects_replay_and_bad_arguments(gpu_ctx_factory):
    m = build_si()
    ctx = gpu_ctx_factory()
    upload_model(ctx, m)
    z = [np.zeros(4) for _ in range(8)]
    with pytest.raises(capi.EmcGpuError, match="emcgpu_bulk_configure"):
        ctx.bulk_run_host(z, np.zeros(4, dtype=np.uint32), 1e-15, 1)
    ctx.bulk_configure([1e-6] * 3, [-1, 0, 0], 1e6)
    with pytest.raises(capi.EmcGpuError, match="bad arguments"):
        ctx.bulk_run_host(z, np.zeros(4, dtype=np.uint32), -1.0, 1)
    obs = ctx.bulk_run_host([np.zeros(0) for _ in range(8)], np.zeros(0, dtype=np.uint32), 1e-15, 3)
    assert obs.shape[0] == 3 and not obs.any()
