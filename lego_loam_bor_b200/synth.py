"""ctypes wrapper of synth/synth_lidar.c: deterministic synthetic lidar sequences
(SURVEY.md section 8d).  Inputs only -- not on the hot path."""
import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from ._paths import LIB_SYNTH, SYNTH_DIR

SEED = 20181001


class SynthConfig(C.Structure):
    _fields_ = [("V", C.c_int32), ("H", C.c_int32), ("bottom_deg", C.c_float), ("top_deg", C.c_float),
                ("seed", C.c_uint64), ("range_sigma", C.c_float), ("jitter_cells", C.c_float),
                ("min_range", C.c_float), ("max_range", C.c_float), ("room_half_x", C.c_float),
                ("room_half_y", C.c_float), ("n_pillars", C.c_int32), ("speed", C.c_float),
                ("radius", C.c_float), ("dt", C.c_float)]


def build(force=False):
    src = os.path.join(SYNTH_DIR, "synth_lidar.c")
    if force or not os.path.exists(LIB_SYNTH) or os.path.getmtime(LIB_SYNTH) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", LIB_SYNTH, src, "-lm"])
    return LIB_SYNTH


_lib = None


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_SYNTH)
        _lib.synth_scan.argtypes = [C.POINTER(SynthConfig), C.c_int, C.c_int, C.c_void_p]
        _lib.synth_scan.restype = C.c_int
        _lib.synth_pose.argtypes = [C.POINTER(SynthConfig), C.c_int, C.c_int, C.c_void_p]
        _lib.synth_local_map.argtypes = [C.POINTER(SynthConfig), C.c_int, C.c_int, C.c_float, C.c_float,
                                         C.c_float, C.c_void_p, C.c_int]
        _lib.synth_local_map.restype = C.c_int
    return _lib


def make_config(params, seed=SEED):
    return SynthConfig(V=params.num_vertical_scans, H=params.num_horizontal_scans,
                       bottom_deg=params.vertical_angle_bottom, top_deg=params.vertical_angle_top,
                       seed=seed, range_sigma=0.01, jitter_cells=0.3, min_range=0.5, max_range=100.0,
                       room_half_x=30.0, room_half_y=20.0, n_pillars=30, speed=1.0, radius=10.0, dt=0.1)


def scan(cfg, seq, frame):
    """One scan as float32 [n, 4] (x, y, z, intensity) in firing order."""
    lib = _load()
    buf = np.empty((cfg.V * cfg.H, 4), np.float32)
    n = lib.synth_scan(C.byref(cfg), seq, frame, buf.ctypes.data)
    return buf[:n].copy()


def scans(cfg, seqs, frames, threads=8):
    """Dict (seq, frame) -> scan, generated on a thread pool (the C call releases the GIL)."""
    _load()
    jobs = [(s, f) for s in seqs for f in frames]
    with ThreadPoolExecutor(threads) as ex:
        out = list(ex.map(lambda sf: scan(cfg, sf[0], sf[1]), jobs))
    return dict(zip(jobs, out))


def pose(cfg, seq, frame):
    lib = _load()
    p = np.zeros(6, np.float64)
    lib.synth_pose(C.byref(cfg), seq, frame, p.ctypes.data)
    return p


def local_map(cfg, seq, kind, step, sigma=0.01, radius_limit=1e9, cap=4_000_000):
    """Synthetic down-sampled local map (camera axes): kind 0 = surfaces, 1 = vertical edges."""
    lib = _load()
    buf = np.empty((cap, 4), np.float32)
    n = lib.synth_local_map(C.byref(cfg), seq, kind, step, sigma, radius_limit, buf.ctypes.data, cap)
    return buf[:n].copy()
