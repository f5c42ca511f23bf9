import os
PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
SYNTH_DIR = os.path.join(PKG, "synth")
INCLUDE = os.path.join(ROOT, "include")
# LEGO_LOAM_B200_LIB: alternative build of the library (A/B experiments)
LIB_CUDA = os.environ.get("LEGO_LOAM_B200_LIB") or os.path.join(PKG, "liblego_loam_b200.so")
LIB_SYNTH = os.path.join(SYNTH_DIR, "libsynth_lidar.so")
