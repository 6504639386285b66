"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol
include/emcgpu.h declares, and fails loudly (no CPU fallback) without a GPU."""
import ctypes
import os
import re

import pytest

from viennaemc_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "emcgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(emcgpu_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(capi.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/emcgpu.h but not exported"
    assert sorted(capi.EXPORTED_SYMBOLS) == syms
    assert lib.emcgpu_abi_version() == 1


def test_struct_layouts_match_header():
    # sizes implied by include/emcgpu.h (natural alignment)
    assert ctypes.sizeof(capi.ValleyC) == 8 + 4 * 8 + 3 * 8 + 8 * 9 * 8
    assert ctypes.sizeof(capi.MechC) == 16 + 32 + 64 + 48
    assert ctypes.sizeof(capi.TableSetC) == 16 + 8 + 8 + 8


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.EmcGpuError) as ei:
        capi.Context(0)
    assert ei.value.code == capi.E_CUDA
    assert "no CPU fallback" in str(ei.value)


def test_product_does_not_import_oracle():
    """The product path must never route through oracle/ (test infrastructure)."""
    bad = []
    for base in ("viennaemc_b200", "include"):
        for d, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                    txt = open(os.path.join(d, f), errors="ignore").read()
                    if re.search(r"(import|from)\s+oracle|#include\s*[<\"].*oracle|liboracle", txt):
                        bad.append(os.path.join(d, f))
    assert not bad, bad
