"""TEST INFRASTRUCTURE.  Statistical fixture of the UNMODIFIED reference hot-phonon example
(examples/hotPhononGa2O3/hotPhononGa2O3.cpp, built as shipped with OpenMP): velocity, mean energy and LO occupation at two
fields for several seeds, with non-equilibrium phonons (--use_hpb 1) and with equilibrium phonons (--use_hpb 0).  Run
length 1 ps per field (the example's own --time option), everything else at the example's defaults.
tests/test_dropin_gpu.py compares the GPU-backed drop-in against it within 3 sigma of the seed-to-seed scatter.
Output: tests/golden/ref_ga2o3_stats.json"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SEEDS = list(range(1, (int(sys.argv[1]) if len(sys.argv) > 1 else 6) + 1))
FIELDS, TIME = "100,300", 1e-12


def main():
    base = tempfile.mkdtemp(prefix="refga")
    exe = os.path.join(base, "ref_ga2o3")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fopenmp", "-I", REF + "/include",
                           REF + "/examples/hotPhononGa2O3/hotPhononGa2O3.cpp", "-o", exe], stderr=subprocess.DEVNULL)
    out = dict(config=f"examples/hotPhononGa2O3/hotPhononGa2O3.cpp as shipped, --fields {FIELDS} --time {TIME} --seed s "
                      "--use_hpb {0,1}; columns: F[kV/cm] v[cm/s] <E>[eV] N_LO N_LO/N_0 T_LO[K] T_ac[K]",
               fields=[float(f) for f in FIELDS.split(",")], time=TIME, seeds=SEEDS)
    for hpb in (1, 0):
        rows = []
        procs = []
        for seed in SEEDS:
            work = os.path.join(base, f"hpb{hpb}_seed{seed}")
            os.makedirs(work)
            procs.append((work, subprocess.Popen([exe, "--fields", FIELDS, "--time", str(TIME), "--seed", str(seed), "--use_hpb",
                                                  str(hpb), "--threads", "2"], cwd=work, stdout=subprocess.DEVNULL)))
            if len(procs) == 4:
                for w, p in procs:
                    assert p.wait() == 0
                    rows.append(np.loadtxt(os.path.join(w, "ga2o3_vE_" + ("hpb" if hpb else "eq") + ".txt")).tolist())
                procs = []
        for w, p in procs:
            assert p.wait() == 0
            rows.append(np.loadtxt(os.path.join(w, "ga2o3_vE_" + ("hpb" if hpb else "eq") + ".txt")).tolist())
        a = np.array(rows)  # [seed][field][column]
        key = "hpb" if hpb else "eq"
        out[key] = dict(runs=rows, mean=a.mean(axis=0).tolist(), std=a.std(axis=0, ddof=1).tolist())
        print(key, "mean", a.mean(axis=0).tolist(), flush=True)
    with open(os.path.join(ROOT, "tests", "golden", "ref_ga2o3_stats.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
