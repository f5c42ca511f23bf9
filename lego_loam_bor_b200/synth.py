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


class ArenaConfig(C.Structure):
    """synth/synth_arena.h: the 120 m arena with 500 key-frame poses on a spiral (BASELINE.json configs[3]/[4])."""
    _fields_ = [("V", C.c_int32), ("H", C.c_int32), ("bottom_deg", C.c_float), ("top_deg", C.c_float),
                ("seed", C.c_uint64), ("range_sigma", C.c_float), ("jitter_cells", C.c_float),
                ("min_range", C.c_float), ("max_range", C.c_float), ("half", C.c_float),
                ("n_pillars", C.c_int32), ("n_walls", C.c_int32), ("n_keyframes", C.c_int32),
                ("spiral_pitch", C.c_float), ("spiral_r0", C.c_float),
                ("speed", C.c_float), ("radius", C.c_float), ("dt", C.c_float)]


ARENA_MAX_BOXES = 320
ARENA_CTX_BYTES = 9 * 8 + 3 * 8 + 3 * 8 + 2 * 4 + 6 * 8   # sizeof(ArenaScanCtx)
KEYFRAME, DRIVE = 0, 1


def build(force=False):
    srcs = [os.path.join(SYNTH_DIR, "synth_lidar.c"), os.path.join(SYNTH_DIR, "synth_arena.c")]
    deps = srcs + [os.path.join(SYNTH_DIR, "synth_arena.h")]
    if force or not os.path.exists(LIB_SYNTH) or os.path.getmtime(LIB_SYNTH) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", LIB_SYNTH] + srcs + ["-lm"])
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
        for name in ("arena_scan", "arena_scan_bruteforce"):
            getattr(_lib, name).argtypes = [C.POINTER(ArenaConfig), C.c_int, C.c_int, C.c_int, C.c_void_p]
            getattr(_lib, name).restype = C.c_int
        _lib.arena_pose.argtypes = [C.POINTER(ArenaConfig), C.c_int, C.c_int, C.c_int, C.c_void_p]
        _lib.arena_pose.restype = None
        _lib.arena_build_world.argtypes = [C.POINTER(ArenaConfig), C.c_int, C.c_void_p]
        _lib.arena_build_world.restype = C.c_int
        _lib.arena_scan_ctx.argtypes = [C.POINTER(ArenaConfig), C.c_int, C.c_uint64, C.c_void_p, C.c_void_p]
        _lib.arena_scan_ctx.restype = None
        _lib.arena_frame_key.argtypes = [C.c_int, C.c_int]
        _lib.arena_frame_key.restype = C.c_uint64
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


# ---- the arena world (synth/synth_arena.h): key frames on a spiral + a driving circle -------------------------


def make_arena(params, n_keyframes=500, seed=SEED):
    return ArenaConfig(V=params.num_vertical_scans, H=params.num_horizontal_scans,
                       bottom_deg=params.vertical_angle_bottom, top_deg=params.vertical_angle_top, seed=seed,
                       range_sigma=0.01, jitter_cells=0.3, min_range=0.5, max_range=100.0, half=60.0,
                       n_pillars=180, n_walls=16, n_keyframes=n_keyframes, spiral_pitch=2.4, spiral_r0=3.0,
                       speed=1.0, radius=10.0, dt=0.1)


def arena_scan(cfg, seq, kind, index, bruteforce=False):
    """One scan of the arena world as float32 [n, 4] in firing order; kind KEYFRAME (pose `index` of the spiral) or
    DRIVE (frame `index` of the driving circle)."""
    lib = _load()
    buf = np.empty((cfg.V * cfg.H, 4), np.float32)
    fn = lib.arena_scan_bruteforce if bruteforce else lib.arena_scan
    n = fn(C.byref(cfg), seq, kind, index, buf.ctypes.data)
    return buf[:n].copy()


def arena_scans(cfg, seqs, kind, indices, threads=8):
    """Dict (seq, index) -> scan.  Jobs of one sequence stay on one thread (the world is cached per thread)."""
    _load()
    seqs = list(seqs)
    indices = list(indices)
    with ThreadPoolExecutor(threads) as ex:
        rows = list(ex.map(lambda s: [arena_scan(cfg, s, kind, i) for i in indices], seqs))
    return {(s, i): a for s, row in zip(seqs, rows) for i, a in zip(indices, row)}


def arena_pose(cfg, seq, kind, index):
    """x, y, z, roll, pitch, yaw (world = Rz Ry Rx)."""
    lib = _load()
    p = np.zeros(6, np.float64)
    lib.arena_pose(C.byref(cfg), seq, kind, index, p.ctypes.data)
    return p


def arena_world(cfg, seq):
    """The boxes of sequence seq's world as float64 [n, 6] (lo xyz, hi xyz); the first four are the outer walls."""
    lib = _load()
    buf = np.zeros((ARENA_MAX_BOXES, 6), np.float64)
    n = lib.arena_build_world(C.byref(cfg), seq, buf.ctypes.data)
    return buf[:n].copy()


def pose_to_transform(pose6):
    """Sensor pose (x, y, z, roll, pitch, yaw; z up) -> the reference's camera-axes 6-vector
    (rx, ry, rz, tx, ty, tz) = (pitch, yaw, roll, y, z, x): the axes are permuted cyclically (x, y, z) <- (y, z, x)
    (featureAssociation.cpp:165-167) and pointAssociateToMap applies Ry(t[1]) Rx(t[0]) Rz(t[2]) (mapOptmization.cpp:412-426)."""
    x, y, z, roll, pitch, yaw = [float(v) for v in pose6]
    return np.array([pitch, yaw, roll, y, z, x], np.float32)


LIB_SYNTH_CUDA = os.path.join(SYNTH_DIR, "libsynth_cuda.so")


def build_cuda(force=False):
    """The device version of the arena generator (synth/synth_arena_cuda.cu), a library of its own: test / bench data
    only, never linked into the hot-path library."""
    src = os.path.join(SYNTH_DIR, "synth_arena_cuda.cu")
    deps = [src, os.path.join(SYNTH_DIR, "synth_arena.h")]
    if force or not os.path.exists(LIB_SYNTH_CUDA) or os.path.getmtime(LIB_SYNTH_CUDA) < max(os.path.getmtime(d) for d in deps):
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        subprocess.check_call([nvcc, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false",
                               "-Xcompiler", "-fPIC", "-shared", "-o", LIB_SYNTH_CUDA, src, "-lcudart"])
    return LIB_SYNTH_CUDA


class ArenaDeviceGenerator:
    """Scans of the arena world for a batch of sequences, ray-cast on the GPU (bit-identical to arena_scan).
    scans(kind, index) returns (float32 device tensor [B, V*H, 4] with the valid points of every sequence packed at
    the front in firing order, int32 numpy counts [B])."""

    def __init__(self, cfg, seq_ids, device):
        import torch
        self.torch = torch
        self.cfg, self.seq_ids, self.device = cfg, list(seq_ids), device
        self.lib = C.CDLL(build_cuda())
        self.lib.arena_scans_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        self.lib.arena_scans_device.restype = C.c_int
        B = len(self.seq_ids)
        self.N = cfg.V * cfg.H
        boxes = np.zeros((B, ARENA_MAX_BOXES, 6), np.float64)
        nbs = np.zeros(B, np.int32)
        host = _load()
        for k, s in enumerate(self.seq_ids):
            nbs[k] = host.arena_build_world(C.byref(cfg), s, boxes[k].ctypes.data)
        self.boxes = torch.from_numpy(boxes).to(device)
        self.nbs = torch.from_numpy(nbs).to(device)
        self.ctx_host = np.zeros((B, ARENA_CTX_BYTES), np.uint8)
        self.dense = torch.empty((B, self.N, 4), dtype=torch.float32, device=device)
        self.rows = torch.arange(B, device=device).unsqueeze(1).expand(B, self.N)

    def scans(self, kind, index, out=None):
        torch = self.torch
        host = _load()
        B = len(self.seq_ids)
        key = host.arena_frame_key(kind, index)
        pose = np.zeros(6, np.float64)
        for k, s in enumerate(self.seq_ids):
            host.arena_pose(C.byref(self.cfg), s, kind, index, pose.ctypes.data)
            host.arena_scan_ctx(C.byref(self.cfg), s, key, pose.ctypes.data, self.ctx_host[k].ctypes.data)
        ctx = torch.from_numpy(self.ctx_host).to(self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        rc = self.lib.arena_scans_device(ctx.data_ptr(), self.boxes.data_ptr(), self.nbs.data_ptr(), B, ARENA_MAX_BOXES, self.N,
                                         self.dense.data_ptr(), stream)
        if rc != 0:
            raise RuntimeError(f"arena_scans_device: cudaError {rc}")
        mask = ~torch.isnan(self.dense[..., 0])
        counts = mask.sum(1)
        pos = torch.cumsum(mask, 1) - 1
        if out is None:
            out = torch.zeros((B, self.N, 4), dtype=torch.float32, device=self.device)
        out[self.rows[mask], pos[mask]] = self.dense[mask]
        return out, counts.to(torch.int32).cpu().numpy()
