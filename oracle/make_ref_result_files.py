"""TEST INFRASTRUCTURE.  Result files as the UNMODIFIED reference writes them (only the run lengths shortened on a
temporary copy): the device-run set of examples/resistor2D (particle, grid, current and scatter-rate files) and the bulk
set of examples/bulkSimulation (time series).  Long files are cut after a few hundred lines -- the format is per line.
tests/test_result_files.py holds our drivers' files (tests/golden/result_files/ours, written on the GPU box by
tools/make_result_file_fixtures.py) against these: same layout, and both go through the reference's own readers
(helper/emcPlottingFiles/emcPlottingFiles/readResultFile.py).  Output: tests/golden/result_files/reference/"""
import os
import re
import shutil
import subprocess
import tempfile

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "result_files", "reference")
KEEP_LINES = 300


def cut(src, dst):
    with open(src) as f, open(dst, "w") as g:
        for i, line in enumerate(f):
            if i >= KEEP_LINES:
                break
            g.write(line)


def main():
    os.makedirs(OUT, exist_ok=True)
    base = tempfile.mkdtemp(prefix="refresult")
    # device run: the short resistor of oracle/Makefile
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "_ref/ref_resistor2D_short"], stdout=subprocess.DEVNULL)
    work = os.path.join(base, "resistor")
    os.makedirs(work)
    subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", "ref_resistor2D_short")], cwd=work, stdout=subprocess.DEVNULL)
    for name in sorted(os.listdir(work)):
        cut(os.path.join(work, name), os.path.join(OUT, name))
    # bulk run: examples/bulkSimulation with totalTime 4e-12 -> 3e-14 (300 steps)
    src = open(os.path.join(REF, "examples", "bulkSimulation", "bulkSimulation.cpp")).read()
    short, n = re.subn(r"const NumType totalTime = 4e-12;", "const NumType totalTime = 3e-14;", src)
    assert n == 1
    work = os.path.join(base, "bulk")
    os.makedirs(work)
    main_cpp = os.path.join(work, "bulk_short.cpp")
    open(main_cpp, "w").write(short)
    exe = os.path.join(work, "bulk_short")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fopenmp", "-I", REF + "/include", "-I", REF + "/examples/bulkSimulation",
                           "-I", REF + "/examples", "-o", exe, main_cpp])
    subprocess.check_call([exe], cwd=work, stdout=subprocess.DEVNULL)
    for name in sorted(os.listdir(work)):
        if name.startswith("bulkSimulation") and name.endswith(".txt"):
            cut(os.path.join(work, name), os.path.join(OUT, name))
    shutil.rmtree(base)
    print(sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
