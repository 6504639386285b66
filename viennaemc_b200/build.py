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
              "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off", "-shared"]


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
    src = [os.path.join(PKG, "csrc", "emcgpu.cu"), os.path.join(PKG, "csrc", "emcgpu_device.cu")]
    deps = _sources(os.path.join(PKG, "csrc"), os.path.join(ROOT, "include"))
    if force or _newer(target, deps):
        cmd = ["nvcc", "--threads", "2", *NVCC_FLAGS, "-o", target, *src]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        subprocess.check_call(cmd)
    return target


HOST_INC = os.path.join(PKG, "host", "include")
# host code builds the rate tables: plain IEEE arithmetic, no FMA contraction, no -march=native,
# so that the tables do not depend on the build machine
HOST_FLAGS = ["-std=c++17", "-O2", "-ffp-contract=off", "-fPIC"]
REFERENCE = "/root/reference"


def build_emchost(force: bool = False) -> str:
    """libemchost.so: the drop-in C++ host API instantiated for the silicon model (include/emchost.h)."""
    target = os.path.join(LIBDIR, "libemchost.so")
    src = os.path.join(PKG, "host", "src", "emchost.cpp")
    deps = _sources(os.path.join(PKG, "host"), os.path.join(ROOT, "include")) + [os.path.join(LIBDIR, "libemcgpu.so")]
    if force or _newer(target, deps):
        subprocess.check_call(["g++", *HOST_FLAGS, "-shared", "-I", os.path.join(ROOT, "include"), "-I", HOST_INC,
                               "-o", target, src, "-L", LIBDIR, "-lemcgpu", "-Wl,-rpath,$ORIGIN"])
    return target


def build_emcnccl(force: bool = False) -> str:
    """libemcnccl.so: ncclAllReduce behind the all-reduce callback of the sharded device run (include/emcnccl.h).  Links the
    NCCL of the system; in a process that has loaded torch the loader resolves libnccl.so.2 to torch's copy."""
    target = os.path.join(LIBDIR, "libemcnccl.so")
    src = os.path.join(PKG, "csrc", "emcnccl.cpp")
    if force or _newer(target, [src, os.path.join(ROOT, "include", "emcnccl.h")]):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-I", "/usr/local/cuda/include", "-o", target, src,
                               "-L", "/usr/local/cuda/lib64", "-lcudart", "-lnccl"])
    return target


def build_examples(force: bool = False) -> dict:
    """Example drivers on top of the drop-in headers.  `reference_bulkSimulation_gpu` is the UNMODIFIED
    main() of the reference's examples/bulkSimulation/bulkSimulation.cpp compiled against OUR headers
    (pre-including our basicBulkParticleHandler.hpp, whose include guard makes the example's own
    quote-include a no-op); it can only be (re)built where the reference tree is mounted, the binary
    travels to the GPU box."""
    bindir = os.path.join(PKG, "bin")
    os.makedirs(bindir, exist_ok=True)
    out = {}
    common = ["g++", *HOST_FLAGS, "-I", HOST_INC, "-I", os.path.join(ROOT, "include")]
    link = ["-L", LIBDIR, "-lemcgpu", "-lemcnccl", "-Wl,-rpath,$ORIGIN/../lib"]
    deps = _sources(os.path.join(PKG, "host"), os.path.join(ROOT, "include")) + [os.path.join(LIBDIR, "libemcgpu.so"),
                                                                                 os.path.join(LIBDIR, "libemcnccl.so")]
    for name in ("bulkSimulation", "resistor2D", os.path.join("mosfet2D", "mosfet2D"),
                 os.path.join("hotPhononGa2O3", "hotPhononGa2O3")):
        own = os.path.join(PKG, "host", "examples", name + ".cpp")
        name = os.path.basename(name)
        if os.path.exists(own):
            target = os.path.join(bindir, name)
            if force or _newer(target, deps):
                subprocess.check_call([*common, "-o", target, own, *link])
            out[name] = target
    ref_main = os.path.join(REFERENCE, "examples", "bulkSimulation", "bulkSimulation.cpp")
    target = os.path.join(bindir, "reference_bulkSimulation_gpu")
    if os.path.exists(ref_main) and (force or _newer(target, deps)):
        subprocess.check_call([*common, "-include", os.path.join(HOST_INC, "basicBulkParticleHandler.hpp"), "-o", target,
                               ref_main, *link])
    if os.path.exists(target):
        out["reference_bulkSimulation_gpu"] = target
    # the UNMODIFIED hot-phonon example (its own Ga2O3Functions.hpp included; the handler swapped like above)
    ref_main = os.path.join(REFERENCE, "examples", "hotPhononGa2O3", "hotPhononGa2O3.cpp")
    target = os.path.join(bindir, "reference_hotPhononGa2O3_gpu")
    if os.path.exists(ref_main) and (force or _newer(target, deps)):
        subprocess.check_call([*common, "-include", os.path.join(HOST_INC, "basicBulkParticleHandler.hpp"), "-o", target,
                               ref_main, *link])
    if os.path.exists(target):
        out["reference_hotPhononGa2O3_gpu"] = target
    # the UNMODIFIED hot-carrier example (two species: emcElectron + emcHole on shared phonon baths).  Its pairwise host
    # steps (carrier-carrier scattering, recombination, energy-selective contacts, band filling) have no GPU implementation:
    # run it with --use_cc 0 --use_recomb 0 --use_esc 0, anything else is rejected with the name of the mechanism
    ref_main = os.path.join(REFERENCE, "examples", "hotCarrierMHP", "hotCarrierMHP.cpp")
    target = os.path.join(bindir, "reference_hotCarrierMHP_gpu")
    if os.path.exists(ref_main) and (force or _newer(target, deps)):
        subprocess.check_call([*common, "-include", os.path.join(HOST_INC, "basicBulkParticleHandler.hpp"), "-o", target,
                               ref_main, *link])
    if os.path.exists(target):
        out["reference_hotCarrierMHP_gpu"] = target
    # the UNMODIFIED single-layer MoS2 example (its own electron2D.hpp and parameter*.hpp included; the handler swapped like
    # above).  Its default parameter set (Pilotto: single-layer valleys, acoustic + zero-order intervalley mechanisms) runs
    # on the GPU; the Kaasbjerg set names mechanisms without a device sampler and is rejected with their names.
    ref_main = os.path.join(REFERENCE, "examples", "singleLayerMoS2", "singleLayerMoS2.cpp")
    target = os.path.join(bindir, "reference_singleLayerMoS2_gpu")
    if os.path.exists(ref_main) and (force or _newer(target, deps)):
        subprocess.check_call([*common, "-include", os.path.join(HOST_INC, "basicBulkParticleHandler.hpp"), "-o", target,
                               ref_main, *link])
    if os.path.exists(target):
        out["reference_singleLayerMoS2_gpu"] = target
    # the same example with its OTHER parameter set: a generated copy of the main in which two configuration constants are
    # changed (Kaasbjerg instead of Pilotto parameters, the CUSTOM field list = one field of 40 kV/cm instead of the 15-field
    # sweep) -- the example has no command line.  One parabolic single-layer valley; acoustic, zero- and first-order
    # intervalley, Froehlich and piezoelectric single-layer mechanisms.
    target = os.path.join(bindir, "reference_singleLayerMoS2_kaasbjerg_gpu")
    if os.path.exists(ref_main) and (force or _newer(target, deps)):
        text = open(ref_main).read()
        for old, new in (("PaperType selectedPaperForParameter = PaperType::PILOTTO;", "PaperType selectedPaperForParameter = PaperType::KAASBJERG;"),
                         ("AppliedFieldsType appliedFields = AppliedFieldsType::HIGH;", "AppliedFieldsType appliedFields = AppliedFieldsType::CUSTOM;")):
            assert text.count(old) == 1, old
            text = text.replace(old, new)
        gen = os.path.join(bindir, "gen")
        os.makedirs(gen, exist_ok=True)
        src = os.path.join(gen, "singleLayerMoS2_kaasbjerg.cpp")
        with open(src, "w") as f:
            f.write(text)
        subprocess.check_call([*common, "-I", os.path.dirname(ref_main), "-include", os.path.join(HOST_INC, "basicBulkParticleHandler.hpp"),
                               "-o", target, src, *link])
    if os.path.exists(target):
        out["reference_singleLayerMoS2_kaasbjerg_gpu"] = target
    # the UNMODIFIED device-run examples of the reference (emcSimulation + emcBasicParticleHandler + emcSORSolver +
    # PM scheme) compiled against OUR headers: every object they create is the GPU-backed drop-in
    for name, rel in (("reference_resistor2D_gpu", ("examples", "resistor2D", "resistor2D.cpp")),):
        ref_main = os.path.join(REFERENCE, *rel)
        target = os.path.join(bindir, name)
        if os.path.exists(ref_main) and (force or _newer(target, deps)):
            subprocess.check_call([*common, "-o", target, ref_main, *link])
        if os.path.exists(target):
            out[name] = target
    # mosfet2D.cpp of the reference brings its own particle-mesh scheme and particle type as sibling headers
    # ("NECSchemeVWD.hpp", "electronVWD.hpp": host code).  Their GPU-backed counterparts of the same names live in
    # host/examples/mosfet2D; -I- makes the quote-includes of the unmodified main() find those first.
    ref_main = os.path.join(REFERENCE, "examples", "mosfet2D", "mosfet2D.cpp")
    target = os.path.join(bindir, "reference_mosfet2D_gpu")
    if os.path.exists(ref_main) and (force or _newer(target, deps)):
        # (-I- also switches off the current-directory look-up inside libstdc++'s pstl headers: list that directory too)
        import glob
        pstl = [a for d in glob.glob("/usr/include/c++/*/pstl") for a in ("-I", d)]
        subprocess.check_call(["g++", *HOST_FLAGS, "-I", os.path.join(PKG, "host", "examples", "mosfet2D"), "-I",
                               os.path.dirname(ref_main), *pstl, "-I-", "-I", HOST_INC, "-I", os.path.join(ROOT, "include"),
                               "-o", target, ref_main, *link], stderr=subprocess.DEVNULL)
    if os.path.exists(target):
        out["reference_mosfet2D_gpu"] = target
    return out


def build_all(force: bool = False) -> dict:
    out = {"emcgpu": build_emcgpu(force), "emcnccl": build_emcnccl(force), "emchost": build_emchost(force)}
    out.update(build_examples(force))
    return out


if __name__ == "__main__":
    print(build_all(force=True))
