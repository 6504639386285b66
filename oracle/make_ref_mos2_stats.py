"""TEST INFRASTRUCTURE.  Statistical fixture of the reference's single-layer MoS2 example
(examples/singleLayerMoS2/singleLayerMoS2.cpp: electron2D, 20 808 electrons, Pilotto parameter set, dt 1e-16 s, 20 000 steps
per field, clock-seeded, 4 OpenMP threads).  The example as shipped sweeps 15 fields (~5 min per run here); for the fixture
only its field LIST is shortened (sed on a generated copy under the git-ignored oracle/_ref/gen: CUSTOM fields 2e5 and 4e6
V/m, both members of the shipped HIGH list) -- every field of the example is an independent run from a fresh ensemble.
Runs it RUNS times and stores, per run and field, the means over the last 1 ps of the per-valley average energy, drift
velocity and occupation.  tests/test_dropin_gpu.py compares the GPU-backed UNMODIFIED example against it.
Output: tests/golden/ref_mos2_stats.json"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
RUNS = int(sys.argv[1]) if len(sys.argv) > 1 else 32
# second argument "kaasbjerg": the example with its Kaasbjerg parameter set and ONE field (its CUSTOM list), see oracle/Makefile
KAASBJERG = len(sys.argv) > 2 and sys.argv[2] == "kaasbjerg"
FIELDS = (4000000,) if KAASBJERG else (200000, 4000000)


def summary(work, field, n=20808):
    tag = f"E{field}T300N{n}.txt"
    e = np.loadtxt(os.path.join(work, "singleLayerMoS2AvgEnergy" + tag), ndmin=2)[-10000:, 1:]
    v = np.loadtxt(os.path.join(work, "singleLayerMoS2AvgDriftVelocity" + tag), ndmin=2)[-10000:, 1:]
    o = np.loadtxt(os.path.join(work, "singleLayerMoS2valleyOccupation" + tag), ndmin=2)[-10000:, 1:]
    return dict(energy=[float(x) for x in e.mean(0)], drift=[float(x) for x in v.mean(0)], occupation=[float(x) for x in o.mean(0)],
                energy_all=float((e * o).sum(1).mean()), drift_all=float((v * o).sum(1).mean()))


def rate_files(work):
    """the per-mechanism rate files the scatter handler writes ("<name><valley><region>ScatterMechanism.txt": energy, rate):
    every 50th of the 5000 levels, as printed (6 significant digits)"""
    out = {}
    for name in sorted(os.listdir(work)):
        if name.endswith("ScatterMechanism.txt"):
            out[name] = [float(x) for x in np.loadtxt(os.path.join(work, name))[::50, 1]]
    return out


def main():
    name = "ref_singleLayerMoS2_kaasbjerg" if KAASBJERG else "ref_singleLayerMoS2_two_fields"
    subprocess.check_call(["make", "-C", HERE, "_ref/" + name], stdout=subprocess.DEVNULL)
    exe = os.path.join(HERE, "_ref", name)
    runs, rates = [], None
    for r in range(RUNS):
        with tempfile.TemporaryDirectory() as work:
            subprocess.check_call([exe], cwd=work, stdout=subprocess.DEVNULL)
            runs.append({str(f): summary(work, f) for f in FIELDS})
            rates = rates or rate_files(work)
        print(r, runs[-1], flush=True)
    if KAASBJERG:
        config = ("examples/singleLayerMoS2/singleLayerMoS2.cpp with selectedPaperForParameter = KAASBJERG and appliedFields = CUSTOM "
                  "(one field, 4e6 V/m): 20808 e-, one parabolic single-layer valley, acoustic + zero- and first-order intervalley + "
                  "Froehlich + piezoelectric single-layer mechanisms, dt 1e-16 s, 20000 steps, 4 OpenMP threads, clock seed; means "
                  "over the last 1 ps")
    else:
        config = ("examples/singleLayerMoS2/singleLayerMoS2.cpp with the field list shortened to {2e5, 4e6} V/m (20808 e-, "
                  "Pilotto parameters, dt 1e-16 s, 20000 steps per field, 4 OpenMP threads, clock seed); means over the last 1 ps")
    out = dict(config=config, n_runs=RUNS, fields=list(FIELDS), runs=runs)
    if KAASBJERG:
        out["rate_files_every_50th_level"] = rates
    with open(os.path.join(ROOT, "tests", "golden", "ref_mos2k_stats.json" if KAASBJERG else "ref_mos2_stats.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
