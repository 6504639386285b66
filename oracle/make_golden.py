"""TEST INFRASTRUCTURE: generate tests/golden/*.npz from the UNMODIFIED reference.

Runs oracle/_ref/ref_bulk_driver (built by oracle/Makefile from /root/reference,
development container only) for every case in tests/scenarios.py::GOLDEN_CASES and
stores its dumps as compressed numpy archives.  The fixtures travel to the GPU
box; the reference tree does not.

    python oracle/make_golden.py
"""
import os
import struct
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from scenarios import DEVICE_CASES, DEVICE_LONG_CASES, GA2O3_CASES, GOLDEN_CASES, MHP_CASES, MOS2_CASES, ga2o3_args, mhp_args  # noqa: E402

DT = {"d": np.float64, "q": np.int64, "Q": np.uint64}


def read_blob(path):
    out = {}
    with open(path, "rb") as f:
        data = f.read()
    o = 0
    while o < len(data):
        (nl,) = struct.unpack_from("<I", data, o); o += 4
        name = data[o:o + nl].decode(); o += nl
        dt = chr(data[o]); o += 1
        (nd,) = struct.unpack_from("<I", data, o); o += 4
        dims = struct.unpack_from("<%dQ" % nd, data, o); o += 8 * nd
        n = int(np.prod(dims)) if nd else 1
        arr = np.frombuffer(data, dtype=DT[dt], count=n, offset=o).reshape(dims).copy(); o += 8 * n
        out[name] = arr
    return out


def main():
    subprocess.check_call(["make", "-C", HERE, "_ref/ref_bulk_driver"])
    drv = os.path.join(HERE, "_ref", "ref_bulk_driver")
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    only = sys.argv[2:] if len(sys.argv) > 2 else None
    for name, case in GOLDEN_CASES.items():
        if only and name not in only:
            continue
        with tempfile.TemporaryDirectory() as tmp:  # the reference writes rate files into the CWD
            out = os.path.join(tmp, "ref.bin")
            cmd = [drv, "--out", out]
            for k, v in case["args"].items():
                cmd += ["--" + k, str(v)]
            subprocess.check_call(cmd, cwd=tmp, stdout=subprocess.DEVNULL)
            blob = read_blob(out)
        dst = os.path.join(ROOT, "tests", "golden", name + ".npz")
        n_draws, n_events = len(blob["draws"]), len(blob["events"])
        if case.get("strip_draws"):
            import hashlib
            blob["draws_count"] = np.array([n_draws], dtype=np.int64)
            blob["draws_sha256"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(blob["draws"]).tobytes()).digest(), dtype=np.uint8)
            del blob["draws"]
            blob["events"] = blob["events"].astype(np.int32)  # (step, particle, mechanism): small integers
        np.savez_compressed(dst, **blob)
        print(name, "->", dst, os.path.getsize(dst) // 1024, "KiB; particles", int(blob["params"][-1]),
              "draws", n_draws, "events", n_events)


def main_mos2(only=None):
    """single-layer MoS2 (examples/singleLayerMoS2: Pilotto and Kaasbjerg parameter sets, the optional extrinsic mechanisms)
    through the same recorder; only: comma-separated case names (third command-line argument)"""
    import hashlib
    subprocess.check_call(["make", "-C", HERE, "_ref/ref_bulk_driver"], stdout=subprocess.DEVNULL)
    drv = os.path.join(HERE, "_ref", "ref_bulk_driver")
    for name, args in MOS2_CASES.items():
        if only and name not in only.split(","):
            continue
        with tempfile.TemporaryDirectory() as tmp:
            out = os.path.join(tmp, "ref.bin")
            cmd = [drv, "--out", out]
            for k, v in args.items():
                cmd += ["--" + k, str(v)]
            subprocess.check_call(cmd, cwd=tmp, stdout=subprocess.DEVNULL)
            blob = read_blob(out)
        n_draws, n_events = len(blob["draws"]), len(blob["events"])
        blob["draws_count"] = np.array([n_draws], dtype=np.int64)
        blob["draws_sha256"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(blob["draws"]).tobytes()).digest(), dtype=np.uint8)
        del blob["draws"]
        blob["events"] = blob["events"].astype(np.int32)
        blob["raw_rates"] = blob["raw_rates"][:, ::25].copy()  # every 25th of the 5000 levels: the full tables are in cum_*
        if name in ("mos2_pilotto_bigdt", "mos2_kaasbjerg_subset"):  # the tables are pinned by another case of the model
            for k in [k for k in blob if k.startswith("cum_") or k == "raw_rates"]:
                del blob[k]
        dst = os.path.join(ROOT, "tests", "golden", name + ".npz")
        np.savez_compressed(dst, **blob)
        print(name, "->", dst, os.path.getsize(dst) // 1024, "KiB; particles", int(blob["params"][-1]), "draws", n_draws,
              "events", n_events)


def main_device():
    subprocess.check_call(["make", "-C", HERE, "_ref/ref_device_driver"], stdout=subprocess.DEVNULL,
                          stderr=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", HERE, "_ref/ref_device3d_driver"], stdout=subprocess.DEVNULL,
                          stderr=subprocess.DEVNULL)
    for name, args in DEVICE_CASES.items():
        drv = os.path.join(HERE, "_ref", "ref_device3d_driver" if "lz" in args else "ref_device_driver")
        with tempfile.TemporaryDirectory() as tmp:
            out = os.path.join(tmp, "ref.bin")
            cmd = [drv, "--out", out]
            for k, v in args.items():
                cmd += ["--" + k, str(v)]
            subprocess.check_call(cmd, cwd=tmp, stdout=subprocess.DEVNULL)
            blob = read_blob(out)
        dst = os.path.join(ROOT, "tests", "golden", name + ".npz")
        np.savez_compressed(dst, **blob)
        print(name, "->", dst, os.path.getsize(dst) // 1024, "KiB; draws", len(blob["draws"]))


def main_device_long():
    """long chained device runs: snapshots only, the raw draws as count + digest (they are the mt19937_64 stream of the seed)"""
    import hashlib
    subprocess.check_call(["make", "-C", HERE, "_ref/ref_device_driver"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    drv = os.path.join(HERE, "_ref", "ref_device_driver")
    for name, args in DEVICE_LONG_CASES.items():
        with tempfile.TemporaryDirectory() as tmp:
            out = os.path.join(tmp, "ref.bin")
            cmd = [drv, "--out", out]
            for k, v in args.items():
                cmd += ["--" + k, str(v)]
            subprocess.check_call(cmd, cwd=tmp, stdout=subprocess.DEVNULL)
            blob = read_blob(out)
        n_draws = len(blob["draws"])
        blob["draws_count"] = np.array([n_draws], dtype=np.int64)
        blob["draws_sha256"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(blob["draws"]).tobytes()).digest(), dtype=np.uint8)
        del blob["draws"]
        for k in [k for k in blob if k.endswith("_pre_k") or "_pre_" in k]:  # pre = post of the step before, relabelled
            del blob[k]
        dst = os.path.join(ROOT, "tests", "golden", name + ".npz")
        np.savez_compressed(dst, **blob)
        print(name, "->", dst, os.path.getsize(dst) // 1024, "KiB; draws", n_draws, "final size", int(blob["size_all"][-1]))


def main_ga2o3():
    subprocess.check_call(["make", "-C", HERE, "_ref/ref_ga2o3_driver"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    drv = os.path.join(HERE, "_ref", "ref_ga2o3_driver")
    for name in GA2O3_CASES:
        with tempfile.TemporaryDirectory() as tmp:
            out = os.path.join(tmp, "ref.bin")
            cmd = [drv, "--out", out]
            for k, v in ga2o3_args(name).items():
                cmd += ["--" + k.replace("_", "-"), str(v)]
            subprocess.check_call(cmd, cwd=tmp, stdout=subprocess.DEVNULL)
            blob = read_blob(out)
        dst = os.path.join(ROOT, "tests", "golden", name + ".npz")
        np.savez_compressed(dst, **blob)
        print(name, "->", dst, os.path.getsize(dst) // 1024, "KiB; particles", int(blob["params"][-1]), "draws", len(blob["draws"]))


def main_mhp():
    subprocess.check_call(["make", "-C", HERE, "_ref/ref_mhp_driver"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    drv = os.path.join(HERE, "_ref", "ref_mhp_driver")
    keys = ("polar", "box", "density", "dt", "steps", "levels", "emax", "screening", "qresolved", "acoustic_bath", "bins", "dq",
            "alpha_e", "alpha_h", "seed")
    for name in MHP_CASES:
        a = mhp_args(name)
        with tempfile.TemporaryDirectory() as tmp:
            out = os.path.join(tmp, "ref.bin")
            cmd = [drv, "--out", out]
            for k in keys:
                cmd += ["--" + k.replace("_", "-"), repr(a[k]) if isinstance(a[k], float) else str(a[k])]
            subprocess.check_call(cmd, cwd=tmp, stdout=subprocess.DEVNULL)
            blob = read_blob(out)
        # the raw draws are the mt19937_64 stream of the seed: keep count + digest (helpers.load_golden regenerates them)
        import hashlib
        blob["draws_count"] = np.array([len(blob["draws"])], dtype=np.int64)
        blob["draws_sha256"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(blob["draws"]).tobytes()).digest(), dtype=np.uint8)
        n_draws = len(blob.pop("draws"))
        dst = os.path.join(ROOT, "tests", "golden", name + ".npz")
        np.savez_compressed(dst, **blob)
        print(name, "->", dst, os.path.getsize(dst) // 1024, "KiB; electrons", int(blob["params"][-2]), "holes", int(blob["params"][-1]),
              "draws", n_draws)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "bulk"):
        main()
    if which in ("all", "device"):
        main_device()
    if which in ("all", "device_long"):
        main_device_long()
    if which in ("all", "ga2o3"):
        main_ga2o3()
    if which in ("all", "mhp"):
        main_mhp()
    if which in ("all", "mos2"):
        main_mos2(sys.argv[2] if len(sys.argv) > 2 else None)
