"""GPU, N > 1: the sharded device run (tests/mgpu_device_run.py) under torchrun on 2 GPUs of the box.  Skipped on a
single-GPU box; the host-side rules it relies on are covered on CPU by tests/test_sharding.py (gloo, world size 2)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_device_run_on_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "mgpu_device_run.py"), "--steps",
                        "300"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and line, r.stdout[-2000:] + r.stderr[-2000:]
    out = json.loads(line[-1])
    assert out["ok"] and all(out["checks"].values()), out


def _launch_sharded(exe, args, workdir, world, extra_env=None, timeout=900):
    """`world` processes of a drop-in C++ driver, one per GPU, the way a launcher would start them: EMCGPU_SHARD=1, RANK,
    WORLD_SIZE, LOCAL_RANK, EMCNCCL_ID_FILE (host/include/ParticleHandler/emcBasicParticleHandler.hpp)"""
    id_file = os.path.join(workdir, "nccl.id")
    procs = []
    for r in range(world):
        d = os.path.join(workdir, f"rank{r}")
        os.makedirs(d)
        env = dict(os.environ, EMCGPU_SHARD="1", RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), EMCNCCL_ID_FILE=id_file,
                   **(extra_env or {}))
        procs.append((d, subprocess.Popen([exe, *args], cwd=d, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    outs = []
    for d, p in procs:
        out, _ = p.communicate(timeout=timeout)
        assert p.returncode == 0, out[-3000:]
        outs.append((d, out))
    return outs


def test_sharded_resistor_driver_on_two_gpus_within_3_sigma_of_the_reference(tmp_path):
    """the drop-in C++ resistor driver (emcSimulation + emcBasicParticleHandler + emcSORSolver) started once per GPU:
    particles split over the ranks, grids replicated, ncclAllReduce through libemcnccl.  Grids bitwise identical on the
    ranks (asserted by emcSimulation itself), terminal currents and profiles against the reference's distribution."""
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from test_dropin_gpu import BIN, _check_resistor, _load, _resistor_summary
    outs = _launch_sharded(os.path.join(BIN, "resistor2D"), ["--seed", "7"], str(tmp_path), 2)
    for d, out in outs:
        assert "Sharded run: 2 ranks, potential and averaged concentration identical on all ranks" in out, out[-2000:]
    root = outs[0][0]
    # rank 0 wrote the (replicated) grids and the summed currents; the particle files are per rank
    n = 0
    for r, (d, _) in enumerate(outs):
        with open(os.path.join(d, f"resistorElectronsFinal.rank{r}.txt")) as f:
            n += sum(1 for _ in f) - 1
    assert not os.path.exists(os.path.join(outs[1][0], "resistorPotentialAvg.txt"))
    os.symlink(os.path.join(root, "resistorElectronsFinal.rank0.txt"), os.path.join(root, "resistorElectronsFinal.txt"))
    s = _resistor_summary(root, "resistor")
    s["n_final"] = n
    _check_resistor([s], _load("ref_resistor_stats.json"), "sharded resistor driver, 2 GPUs")


def _run_single(exe, args, workdir, timeout=900):
    os.makedirs(workdir)
    r = subprocess.run([exe, *args], cwd=workdir, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


def test_sharded_bulk_driver_on_two_gpus_moves_every_particle_as_the_one_gpu_run(tmp_path):
    """the drop-in bulk handler (basicBulkParticleHandler) started once per GPU: every rank creates the same ensemble and keeps
    a block of it; the Philox streams are keyed by the position in the WHOLE ensemble, so the particle files of the ranks,
    put together, are the particle file of the one-GPU run of the same seed, character by character; the per-step averages
    (summed over the ranks with ncclAllReduce once per look-ahead window) agree to the digits the driver prints."""
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from test_dropin_gpu import BIN
    args = ["--seed", "5", "--particles", "50001", "--steps", "200", "--dt", "1e-15", "--prefix", "b", "--print-at", "137"]
    single = os.path.join(str(tmp_path), "single")
    out1 = _run_single(os.path.join(BIN, "bulkSimulation"), args, single)
    outs = _launch_sharded(os.path.join(BIN, "bulkSimulation"), args, str(tmp_path), 2)
    n_total = int(out1.split(" Electrons")[0].split()[-1])
    for d, out in outs:
        assert f"{n_total} Electrons" in out  # getNrParticles() is the size of the whole ensemble on every rank
    ref_lines = open(os.path.join(single, "bElectrons137.txt")).read().splitlines()
    got = []
    for r, (d, _) in enumerate(outs):
        lines = open(os.path.join(d, f"bElectrons137.rank{r}.txt")).read().splitlines()
        assert lines[0] == ref_lines[0]
        got += lines[1:]
    assert len(got) == n_total and abs(len(got) // 2 - (len(open(os.path.join(outs[0][0], "bElectrons137.rank0.txt")).read().splitlines()) - 1)) <= 1
    assert got == ref_lines[1:]
    for k in ("AvgEnergy", "AvgDriftVelocity", "valleyOccupation"):
        ref = np.loadtxt(os.path.join(single, f"b{k}.txt"))
        for d, _ in outs:  # the sums are the same on every rank
            a = np.loadtxt(os.path.join(d, f"b{k}.txt"))
            assert a.shape == ref.shape == (201, 2)
            assert np.allclose(a, ref, rtol=2e-5, atol=0), k


def test_sharded_hot_phonon_driver_sums_the_bath_counters_over_the_ranks(tmp_path):
    """config 5 on two GPUs: the emission / absorption counters per |q| bin of both ranks are added (all-reduce) before the
    phonon baths are updated, so occupations, screening and the rebuilt rate tables are the same on every rank -- and the
    same as in the one-GPU run of the same seed: velocity, mean energy and the phonon occupation of the summary file."""
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from test_dropin_gpu import BIN
    args = ["--fields", "200", "--time", "1.5e-12", "--seed", "11", "--use-hpb", "1", "--outdir", ".", "--tag", "t"]
    single = os.path.join(str(tmp_path), "single")
    _run_single(os.path.join(BIN, "hotPhononGa2O3"), args, single)
    outs = _launch_sharded(os.path.join(BIN, "hotPhononGa2O3"), args, str(tmp_path), 2)
    ref = np.loadtxt(os.path.join(single, "ga2o3_vE_t.txt"))
    assert ref[4] > 1.05  # the LO phonons are driven out of equilibrium (N_LO / N_0): the counters matter
    for d, _ in outs:
        a = np.loadtxt(os.path.join(d, "ga2o3_vE_t.txt"))
        assert np.allclose(a, ref, rtol=1e-3, atol=0), (a, ref)
