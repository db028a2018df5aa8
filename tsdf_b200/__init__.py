"""tsdf_b200 — B200-native TSDF integrate + raycast hot path (hand-written sm_100a CUDA).

The product is ``libtsdf_b200.so`` (C-ABI in ``include/tsdf_b200.h``) plus the drop-in C++
classes under ``tsdf_b200/include``.  This Python package is only the ctypes binding the
tests and ``bench.py`` drive it through; there is no CPU or PyTorch fallback: importing
:mod:`tsdf_b200.capi` raises if the CUDA library has not been built.
"""
from .capi import lib, LIB_PATH, check, Volume  # noqa: F401
