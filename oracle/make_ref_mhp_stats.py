"""TEST INFRASTRUCTURE.  Statistical fixture of the UNMODIFIED reference hot-carrier example
(examples/hotCarrierMHP/hotCarrierMHP.cpp, built as shipped with OpenMP) with its pairwise host steps switched off
(--use_cc 0 --use_recomb 0 --use_esc 0): electrons and holes on a shared hot-phonon bath, screened q-resolved Froehlich
coupling (the example's defaults), 2 ps.  Per seed: mean energy of both species, LO occupation, acoustic temperature and
screening wave vector at 0.5 / 1 / 2 ps.  tests/test_dropin_gpu.py compares the GPU-backed drop-in (the same unmodified
main) against it within 3 sigma of the seed-to-seed scatter.  Output: tests/golden/ref_mhp_stats.json"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SEEDS = list(range(1, (int(sys.argv[1]) if len(sys.argv) > 1 else 32) + 1))
ARGS = ["--use_cc", "0", "--use_recomb", "0", "--use_esc", "0", "--total_time", "2e-12"]
SUFFIX = "HPB_AC_SCR_QR"
ROWS = (9, 19, 39)  # output every 10 steps of 5 fs: 0.5 ps, 1 ps, 2 ps


def summarise(work):
    e = np.loadtxt(os.path.join(work, f"avgEnergyElectrons{SUFFIX}.txt"))
    h = np.loadtxt(os.path.join(work, f"avgEnergyHoles{SUFFIX}.txt"))
    ph = np.loadtxt(os.path.join(work, f"phononOccupation{SUFFIX}.txt"))
    nc = np.loadtxt(os.path.join(work, f"nrCarriers{SUFFIX}.txt"))
    rows = [r + 1 for r in ROWS]  # row 0 is the initial state
    return dict(energy_e=[float(e[r, 1]) for r in rows], energy_h=[float(h[r, 1]) for r in rows],
                n_lo=[float(ph[r, 1]) for r in rows], t_ac=[float(ph[r, 3]) for r in rows], q_s=[float(ph[r, 4]) for r in rows],
                n_e=int(nc[0, 1]), n_h=int(nc[0, 2]))


def main():
    base = tempfile.mkdtemp(prefix="refmhp")
    exe = os.path.join(base, "ref_mhp")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fopenmp", "-I", REF + "/include",
                           REF + "/examples/hotCarrierMHP/hotCarrierMHP.cpp", "-o", exe], stderr=subprocess.DEVNULL)
    runs = []
    for first in range(0, len(SEEDS), 4):
        procs = []
        for seed in SEEDS[first:first + 4]:
            work = os.path.join(base, f"seed{seed}")
            os.makedirs(work)
            procs.append((work, subprocess.Popen([exe, *ARGS, "--seed", str(seed)], cwd=work, stdout=subprocess.DEVNULL,
                                                 env=dict(os.environ, OMP_NUM_THREADS="2"))))
        for work, p in procs:
            assert p.wait() == 0
            runs.append(summarise(work))
            print(len(runs), runs[-1]["energy_e"], runs[-1]["n_lo"], flush=True)
    out = dict(config="examples/hotCarrierMHP/hotCarrierMHP.cpp as shipped, " + " ".join(ARGS) + " --seed s (2 OpenMP threads); "
                      "values at 0.5 / 1 / 2 ps", args=ARGS, suffix=SUFFIX, rows=list(ROWS), n_runs=len(runs), runs=runs)
    with open(os.path.join(ROOT, "tests", "golden", "ref_mhp_stats.json"), "w") as f:
        json.dump(out, f, indent=1)
    print({k: np.mean([r[k] for r in runs], axis=0).tolist() for k in ("energy_e", "energy_h", "n_lo")})


if __name__ == "__main__":
    main()
