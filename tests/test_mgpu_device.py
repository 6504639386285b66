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
