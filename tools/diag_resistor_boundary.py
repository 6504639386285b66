import json, os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from stats_checks import z_scores
st = json.load(open(os.path.join(ROOT, "tests/golden/ref_resistor_stats.json")))
ref = np.array([r["pot_x"] for r in st["runs"]])
refc = np.array([r["conc_x"] for r in st["runs"]])
runs, concs = [], []
with tempfile.TemporaryDirectory() as tmp:
    for seed in range(1, 9):
        subprocess.check_call([os.path.join(ROOT, "viennaemc_b200/bin/resistor2D"), "--seed", str(seed), "--red-black", "1", "--prefix", f"r{seed}"], cwd=tmp, stdout=subprocess.DEVNULL)
        a = np.loadtxt(os.path.join(tmp, f"r{seed}PotentialAvg.txt"), skiprows=1)
        c = np.loadtxt(os.path.join(tmp, f"r{seed}ElectronsConcAvg.txt"), skiprows=1)
        runs.append(a.mean(axis=0)); concs.append(c.mean(axis=0))
        print(seed, "col100 unique", np.unique(a[:, 100]), "col0 unique", np.unique(a[:, 0]), "mean100 %.17g" % a[:, 100].mean())
z = z_scores(runs, ref)
print("z pot", np.round(z, 2).tolist())
print("z conc", np.round(z_scores(concs, refc), 2).tolist())
print("ref mean100 %.17g" % ref[0, 100])
