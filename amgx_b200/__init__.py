"""amgx_b200 -- B200-native AMG solve-phase engine behind the AMGX C API.

The product is ``libamgx_b200.so`` (hand-written sm_100a CUDA + C++ host code, built by
``amgx_b200.build``).  This package is the thin host-side mirror used by the tests and the bench:
``capi`` binds the C-ABI with ctypes, ``gallery`` builds the synthetic matrices BASELINE.json names.
"""
from . import gallery  # noqa: F401

__all__ = ["capi", "gallery", "build"]
