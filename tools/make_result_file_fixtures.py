"""Result files as OUR drivers write them (GPU box): short runs of viennaemc_b200/bin/resistor2D and bulkSimulation with the
reference examples' geometry; long files cut after a few hundred lines.  Output: tests/golden/result_files/ours/ (or
gpurun_out/result_files when run through gpurun; copy it over).  See oracle/make_ref_result_files.py."""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "result_files", "ours")
BIN = os.path.join(ROOT, "viennaemc_b200", "bin")
KEEP_LINES = 300


def cut(src, dst):
    with open(src) as f, open(dst, "w") as g:
        for i, line in enumerate(f):
            if i >= KEEP_LINES:
                break
            g.write(line)


def main():
    os.makedirs(OUT, exist_ok=True)
    # one working directory per driver: both write the per-mechanism rate files (tabulated up to the particle type's
    # maximal energy, 4 eV in the resistor, 1 eV in the bulk example); the resistor's are kept, like in the reference set
    with tempfile.TemporaryDirectory() as work:
        subprocess.check_call([os.path.join(BIN, "resistor2D"), "--steps", "3000", "--transient", "1000", "--avg", "1000", "--seed", "4",
                               "--prefix", "resistorV50as1000"], cwd=work, stdout=subprocess.DEVNULL)
        for name in sorted(os.listdir(work)):
            cut(os.path.join(work, name), os.path.join(OUT, name))
    with tempfile.TemporaryDirectory() as work:
        subprocess.check_call([os.path.join(BIN, "bulkSimulation"), "--steps", "300", "--seed", "4", "--print-at", "300"], cwd=work,
                              stdout=subprocess.DEVNULL)
        for name in sorted(os.listdir(work)):
            if name.startswith("bulkSimulation"):
                cut(os.path.join(work, name), os.path.join(OUT, name))
    print(sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
