"""mpc_b200: B200 (sm_100a) implementation of MPC's per-block compression loop behind a C ABI.

The product is cal_22-mpc_b200/libmpc_b200.so (CUDA kernels + include/mpc_capi.h) and the C++ host
mirror of the reference interface in cal_22-mpc_b200/host/.  This package is the ctypes binding
used by the tests and bench.py.  Import with importlib (the directory name has a hyphen):

    mpcb = importlib.import_module("cal_22-mpc_b200")
"""
from .capi import (ConfigPod, ModulePod, Mpc, MpcError, Stats, StatsPod, SYN, LIB_PATH, SYMBOLS, lib, load_config,
                   unpack, variant_run, VariantStats, sc2_run, sc2_sampling_lines, cpack_run, jit_compile_check, pattern_run, PatternStats)

from .shard import allreduce_stats, shard_range

__all__ = ["allreduce_stats", "shard_range", "ConfigPod", "ModulePod", "Mpc", "MpcError", "Stats", "StatsPod", "SYN", "LIB_PATH", "SYMBOLS", "lib",
           "load_config", "unpack", "variant_run", "VariantStats", "sc2_run", "sc2_sampling_lines", "cpack_run", "jit_compile_check", "pattern_run", "PatternStats"]
