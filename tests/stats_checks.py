"""Statistical comparison of the GPU path with the reference (BASELINE.json north_star: "with independent RNG,
steady-state observables agree with the reference within 3 sigma of the ensemble statistical error at matched particle
count").

The reference side is a sample of >= 30 runs of the UNMODIFIED example programs (tests/golden/ref_*_stats.json, scripts
oracle/make_ref_*_stats.py), so its mean and its run-to-run scatter are known to a few per cent and the bars below are
plain multiples of sigma -- no Student-t widening, no floors on sigma, no absolute floors.

  scalar, one GPU run x:        |x - mean_ref| <= 3 sigma_ref sqrt(1 + 1/n_ref)
  scalar, several GPU seeds:    |mean_gpu - mean_ref| <= 3 s_pooled sqrt(1/n_gpu + 1/n_ref)   (two-sample test: it sees a
                                bias of ~1 sigma of a single run that one run could never resolve)
  profile of m points:          the same per point with 4 sigma instead of 3.  A family of m = 100 independent points at
                                4 sigma raises a false alarm with probability 100 x 6.3e-5 = 0.6 %, i.e. the level of ONE
                                3-sigma test (0.27 %) within a factor of two; neighbouring points of an averaged profile
                                are positively correlated, which only lowers it (Bonferroni bound).
  points whose scatter is exactly zero in both samples (Dirichlet nodes, reservoir cells) must agree to rounding.

All GPU runs are seeded (own drivers: --seed; unmodified reference mains: EMCGPU_SEED), and a seeded GPU run is
reproducible bit for bit, so the outcome of these tests does not fluctuate from one run of the suite to the next."""
import numpy as np

N_SIGMA = 3.0
N_SIGMA_PROFILE = 4.0


def pooled_sigma(gpu, ref):
    gpu, ref = np.atleast_2d(np.asarray(gpu, float).T).T, np.atleast_2d(np.asarray(ref, float).T).T
    ng, nr = gpu.shape[0], ref.shape[0]
    if ng < 2:
        return ref.std(axis=0, ddof=1)
    ss = ref.var(axis=0, ddof=1) * (nr - 1) + gpu.var(axis=0, ddof=1) * (ng - 1)
    return np.sqrt(ss / (ng + nr - 2))


def z_scores(gpu, ref):
    """(mean_gpu - mean_ref) in units of its standard error; gpu: [n_gpu, ...] (or a single run [...]), ref: [n_ref, ...]"""
    ref = np.asarray(ref, float)
    gpu = np.asarray(gpu, float)
    if gpu.ndim == ref.ndim - 1:
        gpu = gpu[None]
    ng, nr = gpu.shape[0], ref.shape[0]
    s = pooled_sigma(gpu, ref)
    diff = gpu.mean(axis=0) - ref.mean(axis=0)
    se = s * np.sqrt(1.0 / ng + 1.0 / nr)
    # a "scatter" at rounding level (means of identical numbers) is no scatter: such points must simply agree
    scale = np.maximum(np.abs(ref.mean(axis=0)), 1e-300)
    pinned = se <= 1e-12 * scale
    with np.errstate(divide="ignore", invalid="ignore"):
        z = np.where(pinned, np.where(np.abs(diff) <= 1e-9 * scale, 0.0, np.inf), diff / np.where(pinned, 1.0, se))
    return z


def assert_scalar(gpu, ref, what, n_sigma=N_SIGMA):
    z = float(np.asarray(z_scores(gpu, ref)).reshape(-1)[0])
    assert abs(z) <= n_sigma, f"{what}: {z:+.2f} sigma (gpu mean {np.mean(gpu):.6g}, reference {np.mean(ref):.6g} +- {np.std(ref, ddof=1):.3g})"
    return z


def assert_profile(gpu, ref, what, n_sigma=N_SIGMA_PROFILE):
    z = np.asarray(z_scores(gpu, ref))
    worst = int(np.argmax(np.abs(z)))
    assert np.all(np.abs(z) <= n_sigma), f"{what}: point {worst} at {z[worst]:+.2f} sigma (max over {z.size} points, bar {n_sigma})"
    return z
