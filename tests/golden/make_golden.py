"""Generates tests/golden/golden_T.json: SHA-256 digests of the synthetic inputs and of every parity-relevant
oracle output for the tiny 16x450 sensor (sequence 0, frames 0..4), plus the poses in clear.
The reference ships no golden vectors, so these pin OUR oracle (portable math
backend) against regressions; tests/test_oracle_pins.py ties the oracle to the reference with the
hand-derivable known answers.  Run from the repo root:  python tests/golden/make_golden.py"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lego_loam_bor_b200 import config_params, synth  # noqa: E402
from oracle.oracle_py import Oracle  # noqa: E402

BUFS = ["RANGE_MAT", "GROUND_MAT", "LABEL_MAT", "SEG_COL_IND", "SEG_GROUND_FLAG", "SEG_RANGE", "START_RING_INDEX",
        "END_RING_INDEX", "ORIENTATION", "OUTLIER_CLOUD", "CORNER_SHARP_IND", "CORNER_LESS_SHARP_IND", "SURF_FLAT_IND",
        "NEIGHBOR_PICKED", "CLOUD_LABEL", "CORNER_LAST", "SURF_LAST"]


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def compute(n_frames=5):
    p = config_params("T")
    cfg = synth.make_config(p)
    o = Oracle(p, libm=False, nanoflann=False, prefer_ref=False)
    out = {"config": "T", "seq": 0, "frames": []}
    for f in range(n_frames):
        scan = synth.scan(cfg, 0, f)
        o.image_projection(scan)
        handed = o.feature_association()
        rec = {"input": digest(scan), "n_points": int(len(scan)), "handed_to_mapping": int(handed),
               "buffers": {b: digest(o.download(b)) for b in BUFS},
               "counts": {b: int(len(o.download(b))) for b in ("SEG_COL_IND", "CORNER_SHARP_IND", "SURF_FLAT_IND", "SURF_LAST")},
               "transform_cur": [float(x) for x in o.download("TRANSFORM_CUR")],
               "transform_sum": [float(x) for x in o.download("TRANSFORM_SUM")],
               "odom_iters": [int(x) for x in o.download("ODOM_ITERS")]}
        out["frames"].append(rec)
    return out


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_T.json")
    with open(path, "w") as f:
        json.dump(compute(), f, indent=1)
    print("wrote", path)
