"""star_b200 — B200-native implementation of STAR's per-read alignment hot path.

The product is the C-ABI shared library star_b200/lib/libstar_b200.so (hand-written sm_100a CUDA kernels,
include/star_b200.h) and the drop-in command line star_b200/bin/STAR.  This Python package is a thin
ctypes binding used by tests/ and bench.py.
"""
from . import capi  # noqa: F401
from .capi import Engine, Index, StarError, default_params, load_library, pack_reads  # noqa: F401
