"""Shared helpers of the parity tests: run the CUDA path (through the C ABI) and the CPU oracle on the
same seeded synthetic scans and compare buffer by buffer."""
import numpy as np

from lego_loam_bor_b200 import config_params, synth
from lego_loam_bor_b200.capi import LegoLoam
from oracle.oracle_py import Oracle

# buffers that must match bit for bit after ImageProjection / feature extraction
EXACT_PROJECTION = ["RANGE_MAT", "FULL_CLOUD", "GROUND_MAT", "LABEL_MAT", "SEG_GROUND_FLAG", "SEG_COL_IND",
                    "SEG_RANGE", "START_RING_INDEX", "END_RING_INDEX", "ORIENTATION", "OUTLIER_CLOUD"]
EXACT_FEATURES = ["CORNER_SHARP_IND", "CORNER_LESS_SHARP_IND", "SURF_FLAT_IND", "NEIGHBOR_PICKED", "CLOUD_LABEL",
                  "SURF_LESS_FLAT_RAW_COUNT", "SEG_CLOUD", "CORNER_SHARP", "SURF_FLAT"]


def same_bits(a, b):
    """Equality that treats NaN == NaN (the empty cells of _full_cloud are NaN)."""
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    if a.shape != b.shape or a.dtype != b.dtype:
        return False
    return a.tobytes() == b.tobytes() or bool(np.array_equal(a, b, equal_nan=True))


def describe_mismatch(name, a, b):
    if a.shape != b.shape:
        return f"{name}: shape {a.shape} vs {b.shape}"
    bad = np.flatnonzero(~((a == b) | ((a != a) & (b != b))).reshape(len(a), -1).all(axis=1))
    return f"{name}: {len(bad)} of {len(a)} rows differ, first at {bad[:5]}: gpu {a[bad[:3]]} oracle {b[bad[:3]]}"


def curvature_ties(oracle, S):
    """True if two curvature values inside one sorted sextant are equal (std::sort is unstable there)."""
    curv = oracle.download("CLOUD_CURVATURE")
    sr, er = oracle.download("START_RING_INDEX"), oracle.download("END_RING_INDEX")
    for i in range(len(sr)):
        for j in range(6):
            sp = (sr[i] * (6 - j) + er[i] * j) // 6
            ep = (sr[i] * (5 - j) + er[i] * (j + 1)) // 6 - 1
            if sp >= ep:
                continue
            v = np.sort(curv[sp:ep])
            if np.any(v[1:] == v[:-1]):
                return True
    return False


def make_scans(cfgname, seqs, frames, seed=synth.SEED):
    p = config_params(cfgname)
    cfg = synth.make_config(p, seed)
    return p, cfg, synth.scans(cfg, seqs, frames)
