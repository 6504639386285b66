"""TEST INFRASTRUCTURE.  Statistical fixture of the UNMODIFIED reference bulk example
(examples/bulkSimulation/bulkSimulation.cpp as shipped: Si, 12 500 electrons, 10 kV/cm, 40 000
steps of 1e-16 s, clock-seeded): runs it several times here (only where /root/reference exists) and
stores, per run, the mean of the ensemble-average energy / drift velocity over the last 1 ps, plus
the transient at a few times.  tests/test_dropin_gpu.py compares the GPU-backed drop-in against it
within 3 sigma (BASELINE.json north_star).  Output: tests/golden/ref_bulk_stats.json"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUNS = int(sys.argv[1]) if len(sys.argv) > 1 else 6


def main():
    work = tempfile.mkdtemp(prefix="refbulk")
    exe = os.path.join(work, "ref_bulk")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fopenmp", "-I", REF + "/include",
                           REF + "/examples/bulkSimulation/bulkSimulation.cpp", "-o", exe])
    runs = []
    for r in range(RUNS):
        subprocess.check_call([exe], cwd=work, stdout=subprocess.DEVNULL)
        e = np.loadtxt(os.path.join(work, "bulkSimulationAvgEnergy.txt"))
        v = np.loadtxt(os.path.join(work, "bulkSimulationAvgDriftVelocity.txt"))
        runs.append(dict(energy_last_ps=float(e[-10000:, 1].mean()), drift_last_ps=float(v[-10000:, 1].mean()),
                         energy_at=[float(e[i, 1]) for i in (0, 2000, 5000, 10000)],
                         drift_at=[float(v[i, 1]) for i in (0, 2000, 5000, 10000)]))
        print(r, runs[-1], flush=True)
    out = dict(config="examples/bulkSimulation/bulkSimulation.cpp as shipped (12500 e-, 1e6 V/m along -x, dt 1e-16 s, "
                      "40000 steps, 4 OpenMP threads, clock seed)",
               n_runs=RUNS, runs=runs,
               energy_mean=float(np.mean([r["energy_last_ps"] for r in runs])),
               energy_std=float(np.std([r["energy_last_ps"] for r in runs], ddof=1)),
               drift_mean=float(np.mean([r["drift_last_ps"] for r in runs])),
               drift_std=float(np.std([r["drift_last_ps"] for r in runs], ddof=1)))
    with open(os.path.join(ROOT, "tests", "golden", "ref_bulk_stats.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps({k: v for k, v in out.items() if k != "runs"}, indent=1))


if __name__ == "__main__":
    main()
