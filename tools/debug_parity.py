"""Developer tool: run GPU and oracle side by side and print where they first diverge."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from parity_utils import EXACT_FEATURES, EXACT_PROJECTION, describe_mismatch, make_scans, same_bits
from lego_loam_bor_b200.capi import LegoLoam
from oracle.oracle_py import Oracle

cfgname = sys.argv[1] if len(sys.argv) > 1 else "T"
nframes = int(sys.argv[2]) if len(sys.argv) > 2 else 4
seqs = [0, 1]
p, cfg, scans = make_scans(cfgname, seqs, range(nframes))
gpu = LegoLoam(p, batch=len(seqs))
oracles = [Oracle(p) for _ in seqs]
for f in range(nframes):
    gpu.set_scans_host([scans[(s, f)] for s in seqs])
    gpu.image_projection()
    for k, s in enumerate(seqs):
        oracles[k].image_projection(scans[(s, f)])
    for k in range(len(seqs)):
        for name in EXACT_PROJECTION + ["SEG_CLOUD"]:
            a, b = gpu.download(name, k), oracles[k].download(name)
            if not same_bits(a, b):
                print(f"[frame {f} seq {k}] MISMATCH", describe_mismatch(name, a, b))
    gpu.feature_association()
    for k in range(len(seqs)):
        oracles[k].feature_association()
    for k in range(len(seqs)):
        o = oracles[k]
        for name in EXACT_FEATURES + ["CLOUD_CURVATURE", "SURF_LESS_FLAT", "CORNER_LAST", "SURF_LAST", "OUTLIER_LAST", "ODOM_ITERS"]:
            a, b = gpu.download(name, k), o.download(name)
            if name == "SURF_LESS_FLAT" and f > 0:
                continue  # swapped into *_LAST on the oracle side
            if not same_bits(a, b):
                print(f"[frame {f} seq {k}] MISMATCH", describe_mismatch(name, a, b))
        for name in ("TRANSFORM_CUR", "TRANSFORM_SUM"):
            a, b = gpu.download(name, k), o.download(name)
            print(f"[frame {f} seq {k}] {name} gpu {a} oracle {b} maxdiff {np.abs(a-b).max():.3e}")
print("launches", gpu.kernel_launches())
