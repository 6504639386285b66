"""TEST INFRASTRUCTURE.  Statistical fixture of the reference MOSFET example (examples/mosfet2D/mosfet2D.cpp: 126 x 101
grid, 4 doping regions, gate contact, NEC-VWD scheme, electronVWD, ~1.5e5 electrons, Vd = Vg = 1 V, clock-seeded).
The example as shipped runs 66 667 steps (about an hour of CPU here); the fixture uses the same main() with ONLY its run
length changed (regex on a temporary copy, plus an absolute path for its "../SiliconFunctions.hpp" include: 2000 steps of 0.15 fs, 1000 of them transient, final average over 500) and
stores, per run, terminal currents, ensemble size and the averaged potential / concentration along the channel.
tests/test_dropin_gpu.py runs the GPU-backed drop-in with the same run length and compares within 3 sigma of the
reference's run-to-run scatter.  Output: tests/golden/ref_mosfet_stats.json"""
import json
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUNS = int(sys.argv[1]) if len(sys.argv) > 1 else 4
PARALLEL = int(sys.argv[2]) if len(sys.argv) > 2 else 2
PREFIX = "mosfetVd1000Vg1000"
STEPS, TRANSIENT, AVG, DT = 2000, 1000, 500, 1.5e-16


def read_grid(path):
    with open(path) as f:
        f.readline()
        return np.loadtxt(f)


def summarise(work):
    cur = np.loadtxt(os.path.join(work, PREFIX + "ElectronsCurrent.txt"))
    pot = read_grid(os.path.join(work, PREFIX + "PotentialAvg.txt"))
    conc = read_grid(os.path.join(work, PREFIX + "ElectronsConcAvg.txt"))
    with open(os.path.join(work, PREFIX + "ElectronsFinal.txt")) as f:
        n_final = sum(1 for _ in f) - 1
    # rows are y (depth from the gate side), columns x (source -> drain)
    return dict(current=[float(v) for v in cur[-1, 5:9]], netto_sum=[int(v) for v in cur[:, 1:5].sum(axis=0)],
                n_final=n_final, pot_surface=[float(v) for v in pot[1]], pot_depth=[float(v) for v in pot[:, 63]],
                conc_surface=[float(v) for v in conc[1:4].mean(axis=0)], conc_depth=[float(v) for v in conc[:, 63]],
                conc_total=float(conc.sum()))


def main():
    base = tempfile.mkdtemp(prefix="refmos")
    src = open(os.path.join(REF, "examples", "mosfet2D", "mosfet2D.cpp")).read()
    short, n1 = re.subn(r"param\.setTimes\(10e-12, 1\.5e-16, 5e-12\);",
                        f"param.setTimes({(STEPS - 0.5) * DT!r}, {DT!r}, {(TRANSIENT - 0.5) * DT!r});", src)
    short, n2 = re.subn(r"param\.setNrStepsForFinalAvg\(6667\);", f"param.setNrStepsForFinalAvg({AVG});", short)
    short, n3 = re.subn(r'#include "\.\./SiliconFunctions\.hpp"', f'#include "{REF}/examples/SiliconFunctions.hpp"', short)
    assert n1 == 1 and n2 == 1 and n3 == 1
    main_cpp = os.path.join(base, "mosfet2D_short.cpp")
    with open(main_cpp, "w") as f:
        f.write(short)
    exe = os.path.join(base, "ref_mosfet")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fopenmp", "-I", REF + "/include", "-I", REF + "/examples/mosfet2D",
                           "-o", exe, main_cpp], stderr=subprocess.DEVNULL)
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

    def stat(key):
        a = np.array([r[key] for r in runs], dtype=float)
        return a.mean(axis=0).tolist(), a.std(axis=0, ddof=1).tolist()

    out = dict(config=f"examples/mosfet2D/mosfet2D.cpp with its run length changed to {STEPS} steps of {DT} s "
                      f"({TRANSIENT} transient, final average over {AVG}); 4 OpenMP threads, clock seed",
               steps=STEPS, transient=TRANSIENT, avg=AVG, dt=DT, n_runs=len(runs), runs=runs)
    for key in ("current", "n_final", "pot_surface", "pot_depth", "conc_surface", "conc_depth", "conc_total"):
        out[key + "_mean"], out[key + "_std"] = stat(key)
    with open(os.path.join(ROOT, "tests", "golden", "ref_mosfet_stats.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps({k: out[k] for k in ("current_mean", "current_std", "n_final_mean", "n_final_std")}, indent=1))


if __name__ == "__main__":
    main()
