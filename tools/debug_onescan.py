import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from lego_loam_bor_b200 import config_params, synth
from lego_loam_bor_b200.capi import LegoLoam
from oracle.oracle_py import Oracle
from parity_utils import EXACT_PROJECTION, EXACT_FEATURES, same_bits, describe_mismatch, curvature_ties
cfgname, K, kf = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
p = config_params(cfgname); cfg = synth.make_arena(p, n_keyframes=K)
sc = synth.arena_scan(cfg, 0, synth.KEYFRAME, kf)
gpu = LegoLoam(p, batch=1); o = Oracle(p)
gpu.set_scans_host([sc]); gpu.image_projection(); o.image_projection(sc)
for name in EXACT_PROJECTION + ["SEG_CLOUD"]:
    a, b = gpu.download(name), o.download(name)
    if not same_bits(a, b): print("PROJ", describe_mismatch(name, a, b))
gpu.feature_association(); o.feature_association()
S = len(o.download("SEG_CLOUD"))
print("S", S, "ties", curvature_ties(o, S))
ca, cb = gpu.download("CLOUD_CURVATURE"), o.download("CLOUD_CURVATURE")
print("curv maxrel", np.max(np.abs(ca[:S]-cb[:S])/np.maximum(np.abs(cb[:S]),1e-30)))
for name in EXACT_FEATURES + ["CORNER_LESS_SHARP", "SURF_LESS_FLAT", "CORNER_LAST", "SURF_LAST", "OUTLIER_LAST"]:
    a, b = gpu.download(name), o.download(name)
    if not same_bits(a, b): print("FEAT", describe_mismatch(name, a, b)[:600])
gpu.map_downsample_current_scan(); o.map_downsample_current_scan()
for name in ("SCAN_CORNER_DS", "SCAN_SURF_DS", "SCAN_OUTLIER_DS", "SCAN_SURF_TOTAL_DS"):
    a, b = gpu.download(name), o.download(name)
    if not same_bits(a, b): print("DS", describe_mismatch(name, a, b)[:600])
print("done")
