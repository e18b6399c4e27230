"""aqsis_b200 -- B200-native REYES hider + pixel filter (drop-in for the post-shading hot path of aqsis).

The product is the native library aqsis_b200/_lib/libaqsis_b200_hider.so (C ABI in
include/aqsis_b200_hider.h, sm_100a kernels in aqsis_b200/csrc).  This package holds the
ctypes face of that ABI, the synthetic scene generators of the benchmark and the build script.
"""
from . import _abi as abi
from .hider import Hider, HiderError, GridArrays, default_params, display_info, lib

__all__ = ["abi", "Hider", "HiderError", "GridArrays", "default_params", "display_info", "lib"]
