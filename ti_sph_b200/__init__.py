"""ti_sph_b200 -- B200-native (sm_100a) engine for Ti-SPH's per-step WCSPH loop.

The package holds only what the hot path needs: the CUDA kernels and the C ABI (csrc/,
include/tisph.h), their ctypes binding (_capi), the Engine wrapper the drop-in classes in
core/ and utils/ are built on, and scene helpers.  Importing it never touches oracle/.
"""
from ._capi import TisphError, load, library_path          # noqa: F401
from .engine import Engine                                  # noqa: F401
from . import scene                                         # noqa: F401

__all__ = ["Engine", "TisphError", "load", "library_path", "scene"]
