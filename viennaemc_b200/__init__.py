"""B200-native implementation of ViennaEMC's per-time-step particle loop.

The product is the CUDA library behind the C ABI of include/emcgpu.h plus the
reference-compatible C++17 host headers in viennaemc_b200/host/.  This Python
package only binds that ABI for tests, benchmarks and multi-GPU plumbing.
"""
from . import capi  # noqa: F401

__all__ = ["capi"]
