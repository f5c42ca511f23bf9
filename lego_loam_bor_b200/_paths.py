import os
PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
SYNTH_DIR = os.path.join(PKG, "synth")
INCLUDE = os.path.join(ROOT, "include")
LIB_CUDA = os.path.join(PKG, "liblego_loam_b200.so")
LIB_SYNTH = os.path.join(SYNTH_DIR, "libsynth_lidar.so")
