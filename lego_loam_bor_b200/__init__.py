"""B200-native implementation of LeGO-LOAM-BOR's per-scan hot path.

Layout (DESIGN.md has the full map):
  csrc/    hand-written sm_100a CUDA kernels + the C ABI of include/lego_loam_b200.h
  host/    C++ mirror of the reference's ImageProjection / FeatureAssociation /
           MapOptimization class interfaces on top of the C ABI
  capi.py  ctypes binding of the C ABI (what tests/ and bench.py call)
  synth.py deterministic synthetic lidar sequences (inputs only; not on the hot path)

The package never imports anything under oracle/ (the CPU oracle is test infrastructure).
"""
from .params import LegoLoamParams, default_params, config_params  # noqa: F401
