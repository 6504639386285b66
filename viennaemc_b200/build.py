"""Build the sm_100a shared libraries of the package in-tree with nvcc/g++.

    python -m viennaemc_b200.build

The built .so files are git-ignored but travel with the working tree to the GPU
box.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
LIBDIR = os.path.join(PKG, "lib")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-fmad=false",
              "-Xcompiler", "-fPIC", "-shared"]


def _newer(target: str, sources) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _sources(*dirs):
    out = []
    for d in dirs:
        for base, _, files in os.walk(d):
            out += [os.path.join(base, f) for f in files if f.endswith((".cu", ".cuh", ".h", ".hpp", ".cpp"))]
    return out


def build_emcgpu(force: bool = False, verbose: bool = False) -> str:
    """libemcgpu.so: the CUDA kernels + the C ABI of include/emcgpu.h."""
    os.makedirs(LIBDIR, exist_ok=True)
    target = os.path.join(LIBDIR, "libemcgpu.so")
    src = os.path.join(PKG, "csrc", "emcgpu.cu")
    deps = _sources(os.path.join(PKG, "csrc"), os.path.join(ROOT, "include"))
    if force or _newer(target, deps):
        cmd = ["nvcc", *NVCC_FLAGS, "-o", target, src]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        subprocess.check_call(cmd)
    return target


def build_all(force: bool = False) -> dict:
    return {"emcgpu": build_emcgpu(force)}


if __name__ == "__main__":
    print(build_all(force=True))
