"""TEST INFRASTRUCTURE.  Statistical fixture of the UNMODIFIED reference device example
(examples/resistor2D/resistor2D.cpp as shipped: Si bar 1 um x 1 um, 101 x 21 grid, 1e22 m^-3 donors,
two ohmic contacts at 0 / 50 mV, ~1e4 electrons, 50 000 steps of 1 fs, clock-seeded, NGP + SOR): runs
it several times here (only where /root/reference exists) and stores, per run, the terminal currents,
the y-averaged final potential / concentration profiles and the ensemble size.
tests/test_dropin_gpu.py compares the GPU-backed drop-in against it within 3 sigma of the reference's
own run-to-run scatter (BASELINE.json north_star).  Output: tests/golden/ref_resistor_stats.json"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUNS = int(sys.argv[1]) if len(sys.argv) > 1 else 4
PARALLEL = int(sys.argv[2]) if len(sys.argv) > 2 else 2
PREFIX = "resistorV50as1000"


def read_grid(path):
    with open(path) as f:
        f.readline()  # extent line
        return np.loadtxt(f)


def summarise(work):
    cur = np.loadtxt(os.path.join(work, PREFIX + "ElectronsCurrent.txt"))
    pot = read_grid(os.path.join(work, PREFIX + "PotentialAvg.txt"))
    conc = read_grid(os.path.join(work, PREFIX + "ElectronsConcAvg.txt"))
    with open(os.path.join(work, PREFIX + "ElectronsFinal.txt")) as f:
        n_final = sum(1 for _ in f) - 1
    return dict(current=[float(cur[-1, 3]), float(cur[-1, 4])], netto_sum=[int(cur[:, 1].sum()), int(cur[:, 2].sum())],
                pot_x=[float(v) for v in pot.mean(axis=0)], conc_x=[float(v) for v in conc.mean(axis=0)], n_final=n_final)


def main():
    base = tempfile.mkdtemp(prefix="refres")
    exe = os.path.join(base, "ref_resistor")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fopenmp", "-I", REF + "/include",
                           REF + "/examples/resistor2D/resistor2D.cpp", "-o", exe])
    runs = []
    for first in range(0, RUNS, PARALLEL):
        procs = []
        for r in range(first, min(RUNS, first + PARALLEL)):
            work = os.path.join(base, f"run{r}")
            os.makedirs(work)
            procs.append((work, subprocess.Popen([exe], cwd=work, stdout=subprocess.DEVNULL)))
        for work, p in procs:
            assert p.wait() == 0
            runs.append(summarise(work))
            print(len(runs), runs[-1]["current"], runs[-1]["n_final"], flush=True)
    cur = np.array([r["current"] for r in runs])
    out = dict(config="examples/resistor2D/resistor2D.cpp as shipped (101x21 grid, 1e22 m^-3, 50 mV, dt 1e-15 s, 50000 steps, "
                      "30000 non-transient, 4 OpenMP threads, clock seed)",
               n_runs=len(runs), runs=runs, current_mean=cur.mean(axis=0).tolist(), current_std=cur.std(axis=0, ddof=1).tolist(),
               pot_x_mean=np.mean([r["pot_x"] for r in runs], axis=0).tolist(),
               pot_x_std=np.std([r["pot_x"] for r in runs], axis=0, ddof=1).tolist(),
               conc_x_mean=np.mean([r["conc_x"] for r in runs], axis=0).tolist(),
               conc_x_std=np.std([r["conc_x"] for r in runs], axis=0, ddof=1).tolist())
    with open(os.path.join(ROOT, "tests", "golden", "ref_resistor_stats.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps({k: v for k, v in out.items() if k in ("current_mean", "current_std", "n_runs")}, indent=1))


if __name__ == "__main__":
    main()
