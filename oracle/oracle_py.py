"""TEST INFRASTRUCTURE: ctypes wrapper of the CPU oracle (oracle/lego_oracle.h).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_REF = os.path.join(HERE, "_ref", "liblego_oracle_ref.so")
LIB_PORT = os.path.join(HERE, "build", "liblego_oracle_port.so")
REFERENCE = "/root/reference"

# (ctype, numpy dtype, floats-per-element) per ll_buffer id; mirrors include/lego_loam_b200.h
PT = ("pt", np.float32, 4)
BUFFERS = {
    "RANGE_MAT": (0, np.float32, 1), "FULL_CLOUD": (1,) + PT[1:], "GROUND_MAT": (2, np.int8, 1),
    "LABEL_MAT": (3, np.int32, 1), "SEG_CLOUD": (4,) + PT[1:], "SEG_GROUND_FLAG": (5, np.uint8, 1),
    "SEG_COL_IND": (6, np.uint32, 1), "SEG_RANGE": (7, np.float32, 1), "START_RING_INDEX": (8, np.int32, 1),
    "END_RING_INDEX": (9, np.int32, 1), "ORIENTATION": (10, np.float32, 1), "OUTLIER_CLOUD": (11,) + PT[1:],
    "CLOUD_CURVATURE": (12, np.float32, 1), "NEIGHBOR_PICKED": (13, np.int32, 1), "CLOUD_LABEL": (14, np.int32, 1),
    "CORNER_SHARP": (15,) + PT[1:], "CORNER_LESS_SHARP": (16,) + PT[1:], "SURF_FLAT": (17,) + PT[1:],
    "SURF_LESS_FLAT": (18,) + PT[1:], "CORNER_SHARP_IND": (19, np.int32, 1),
    "CORNER_LESS_SHARP_IND": (20, np.int32, 1), "SURF_FLAT_IND": (21, np.int32, 1),
    "CORNER_LAST": (22,) + PT[1:], "SURF_LAST": (23,) + PT[1:], "TRANSFORM_CUR": (24, np.float32, 1),
    "TRANSFORM_SUM": (25, np.float32, 1), "ODOM_ITERS": (26, np.int32, 1), "MAP_CORNER": (27,) + PT[1:],
    "MAP_SURF": (28,) + PT[1:], "SCAN_CORNER_DS": (29,) + PT[1:], "SCAN_SURF_TOTAL_DS": (30,) + PT[1:],
    "TRANSFORM_TOBE_MAPPED": (31, np.float32, 1), "MAP_ITERS": (32, np.int32, 1),
    "OUTLIER_LAST": (33,) + PT[1:], "SURF_LESS_FLAT_RAW_COUNT": (34, np.int32, 1),
    "MAP_TRACE": (35, np.float64, 1), "TRANSFORM_BEF_MAPPED": (36, np.float32, 1),
    "TRANSFORM_AFT_MAPPED": (37, np.float32, 1), "SCAN_SURF_DS": (38, np.float32, 4),
    "SCAN_OUTLIER_DS": (39, np.float32, 4), "STAGE_CLOCKS": (40, np.int64, 1),
    "KEYFRAME_STATE": (41, np.int32, 1), "KEY_POSES_6D": (42, np.float32, 6), "SURROUNDING_KEY_IDS": (43, np.int32, 1),
    "MAP_KNN_IDX": (45, np.int32, 5), "ODOM_SEARCH_IDX": (46, np.int32, 3),
}


def build(force=False):
    """make port (+ ref when /root/reference is present).  Prebuilt files are used as they are
    when the sources are not newer (the GPU box has no /root/reference)."""
    have_ref = os.path.exists(os.path.join(REFERENCE, "LeGO-LOAM/include/lego_loam/nanoflann.hpp"))
    targets = ["all"] if have_ref else ["port", "smallmat"]
    args = ["make", "-C", HERE] + targets + (["-B"] if force else [])
    subprocess.check_call(args, stdout=subprocess.DEVNULL)


_libs = {}


def load(prefer_ref=True):
    key = "ref" if (prefer_ref and os.path.exists(LIB_REF)) else "port"
    if key not in _libs:
        path = LIB_REF if key == "ref" else LIB_PORT
        if not os.path.exists(path):
            build()
        lib = C.CDLL(path)
        lib.lo_create.restype = C.c_void_p
        lib.lo_create.argtypes = [C.c_void_p]
        for name in ("lo_destroy", "lo_reset", "lo_reset_timers", "lo_reset_feature_association"):
            getattr(lib, name).argtypes = [C.c_void_p]
            getattr(lib, name).restype = None
        lib.lo_image_projection.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        lib.lo_feature_association.argtypes = [C.c_void_p]
        lib.lo_map_set_local.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        lib.lo_map_set_scan.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        lib.lo_map_downsample_current_scan.argtypes = [C.c_void_p]
        lib.lo_map_set_initial_guess.argtypes = [C.c_void_p, C.c_void_p]
        lib.lo_scan_to_map.argtypes = [C.c_void_p]
        lib.lo_map_set_poses.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.lo_map_predict_pose.argtypes = [C.c_void_p]
        lib.lo_download.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        lib.lo_upload.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
        lib.lo_voxel_grid.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_void_p]
        lib.lo_knn.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        lib.lo_get_timers.argtypes = [C.c_void_p, C.c_void_p]
        lib.lo_get_timers.restype = None
        for name in ("lo_map_extract_surrounding_keyframes", "lo_map_save_keyframe", "lo_mapping_cycle"):
            getattr(lib, name).argtypes = [C.c_void_p]
        lib.lo_map_download_keyframe.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        lib.lo_get_timer_map_assembly.argtypes = [C.c_void_p]
        lib.lo_get_timer_map_assembly.restype = C.c_double
        lib.lo_run_pipeline.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        lib.lo_decode_pointcloud2.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        _libs[key] = lib
    return _libs[key]


def kind(prefer_ref=True):
    """'reference-knn' when the reference's nanoflann is compiled in, else 'port'."""
    return "reference-knn" if load(prefer_ref).lo_has_nanoflann() else "port"


class Oracle:
    """One sequence through the CPU restatement."""

    def __init__(self, params, libm=False, nanoflann=True, prefer_ref=True, float_accum=0, stable_sort=False):
        self.lib = load(prefer_ref)
        self.lib.lo_set_accum_backend(int(float_accum))
        self.lib.lo_set_sort_backend(1 if stable_sort else 0)
        self.lib.lo_set_math_backend(1 if libm else 0)
        self.lib.lo_set_knn_backend(1 if nanoflann else 0)
        self._libm, self._nf = libm, nanoflann
        self.params = params
        self.h = self.lib.lo_create(C.addressof(params))
        self.lib.lo_set_accum_backend(0)
        self.lib.lo_set_sort_backend(0)
        if not self.h:
            raise RuntimeError("lo_create failed")

    def _select(self):
        self.lib.lo_set_math_backend(1 if self._libm else 0)
        self.lib.lo_set_knn_backend(1 if self._nf else 0)

    def close(self):
        if self.h:
            self.lib.lo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        self.lib.lo_reset(self.h)

    def reset_feature_association(self):
        self.lib.lo_reset_feature_association(self.h)

    def image_projection(self, xyzi):
        self._select()
        xyzi = np.ascontiguousarray(xyzi, np.float32)
        return self.lib.lo_image_projection(self.h, xyzi.ctypes.data, xyzi.shape[0])

    def feature_association(self):
        self._select()
        return self.lib.lo_feature_association(self.h)

    def map_set_local(self, corner, surf):
        corner = np.ascontiguousarray(corner, np.float32)
        surf = np.ascontiguousarray(surf, np.float32)
        return self.lib.lo_map_set_local(self.h, corner.ctypes.data, len(corner), surf.ctypes.data, len(surf))

    def map_set_scan(self, corner, surf_total):
        corner = np.ascontiguousarray(corner, np.float32)
        surf_total = np.ascontiguousarray(surf_total, np.float32)
        return self.lib.lo_map_set_scan(self.h, corner.ctypes.data, len(corner), surf_total.ctypes.data, len(surf_total))

    def map_downsample_current_scan(self):
        return self.lib.lo_map_downsample_current_scan(self.h)

    def map_set_initial_guess(self, t6):
        t6 = np.ascontiguousarray(t6, np.float32)
        return self.lib.lo_map_set_initial_guess(self.h, t6.ctypes.data)

    def map_set_poses(self, aft6, bef6):
        aft6 = np.ascontiguousarray(aft6, np.float32)
        bef6 = np.ascontiguousarray(bef6, np.float32)
        return self.lib.lo_map_set_poses(self.h, aft6.ctypes.data, bef6.ctypes.data)

    def map_predict_pose(self):
        self._select()
        return self.lib.lo_map_predict_pose(self.h)

    def scan_to_map(self):
        self._select()
        return self.lib.lo_scan_to_map(self.h)

    def map_extract_surrounding_keyframes(self):
        self._select()
        return self.lib.lo_map_extract_surrounding_keyframes(self.h)

    def map_save_keyframe(self):
        self._select()
        return self.lib.lo_map_save_keyframe(self.h)

    def mapping_cycle(self):
        self._select()
        return self.lib.lo_mapping_cycle(self.h)

    def download_keyframe(self, kf, which):
        n = C.c_size_t(0)
        rc = self.lib.lo_map_download_keyframe(self.h, kf, which, None, 0, C.byref(n))
        if rc != 0:
            raise RuntimeError(f"lo_map_download_keyframe({kf}, {which}) -> {rc}")
        out = np.empty((n.value, 4), np.float32)
        self.lib.lo_map_download_keyframe(self.h, kf, which, out.ctypes.data, out.nbytes, C.byref(n))
        return out

    def timer_map_assembly(self):
        return self.lib.lo_get_timer_map_assembly(self.h)

    def download(self, name):
        bid, dt, w = BUFFERS[name]
        n = C.c_size_t(0)
        self.lib.lo_download(self.h, bid, None, 0, C.byref(n))
        out = np.empty((n.value, w) if w > 1 else (n.value,), dt)
        rc = self.lib.lo_download(self.h, bid, out.ctypes.data, out.nbytes, C.byref(n))
        if rc != 0:
            raise RuntimeError(f"lo_download({name}) -> {rc}")
        return out

    def upload(self, name, arr):
        bid, dt, w = BUFFERS[name]
        arr = np.ascontiguousarray(arr, dt)
        rc = self.lib.lo_upload(self.h, bid, arr.ctypes.data, arr.size // w)
        if rc != 0:
            raise RuntimeError(f"lo_upload({name}) -> {rc}")

    def timers(self):
        t = np.zeros(5, np.float64)
        self.lib.lo_get_timers(self.h, t.ctypes.data)
        return t

    def reset_timers(self):
        self.lib.lo_reset_timers(self.h)


def run_pipeline(o_ip, o_fa, o_mo, scans, first_timed_frame=0):
    """The reference's three stage threads + blocking one-slot channels (main.cpp:37-47) over one sequence.
    scans: list of float32 [n, 4]; returns (stage_ms float64 [n_frames, 3], wall seconds from first_timed_frame on)."""
    o_mo._select()
    scans = [np.ascontiguousarray(a, np.float32) for a in scans]
    n = len(scans)
    ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in scans])
    counts = np.array([len(a) for a in scans], np.int32)
    stage_ms = np.zeros((n, 3), np.float64)
    wall = C.c_double(0)
    o_ip.lib.lo_run_pipeline(o_ip.h, o_fa.h, o_mo.h, ptrs, counts.ctypes.data, n, first_timed_frame, stage_ms.ctypes.data, C.byref(wall))
    return stage_ms, wall.value


def decode_pointcloud2(data, n_points, point_step, off_x, off_y, off_z, off_intensity, is_dense, prefer_ref=True):
    """fromROSMsg + removeNaNFromPointCloud on raw message bytes (uint8 array) -> float32 [m, 4]."""
    lib = load(prefer_ref)
    data = np.ascontiguousarray(data, np.uint8)
    out = np.empty((max(1, n_points), 4), np.float32)
    m = lib.lo_decode_pointcloud2(data.ctypes.data, n_points, point_step, off_x, off_y, off_z, off_intensity, 1 if is_dense else 0,
                                  out.ctypes.data)
    return out[:m].copy()


def voxel_grid(xyzi, leaf, prefer_ref=True):
    lib = load(prefer_ref)
    xyzi = np.ascontiguousarray(xyzi, np.float32)
    out = np.empty_like(xyzi)
    n = lib.lo_voxel_grid(xyzi.ctypes.data, len(xyzi), leaf, out.ctypes.data)
    return out[:n].copy()


def knn(cloud, query, k, nanoflann=True, prefer_ref=True):
    lib = load(prefer_ref)
    lib.lo_set_knn_backend(1 if nanoflann else 0)
    cloud = np.ascontiguousarray(cloud, np.float32)
    query = np.ascontiguousarray(query, np.float32)
    idx = np.empty((len(query), k), np.int32)
    d2 = np.empty((len(query), k), np.float32)
    lib.lo_knn(cloud.ctypes.data, len(cloud), query.ctypes.data, len(query), k, idx.ctypes.data, d2.ctypes.data)
    return idx, d2


# ---- nav_msgs/Odometry form of a pose; tf is not vendored: parity unpinned ----
def _quat_mul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz], np.float64)


def transform_to_odometry(t6):
    """featureAssociation.cpp:1286-1298 / mapOptmization.cpp:510-522: position x, y, z + orientation x, y, z, w.
    tf::createQuaternionMsgFromRollPitchYaw(t[2], -t[0], -t[1]) written as the product Rz(yaw) Ry(pitch) Rx(roll)."""
    t = np.asarray(t6, np.float32)
    roll, pitch, yaw = np.float64(t[2]), np.float64(-t[0]), np.float64(-t[1])
    qx = np.array([np.sin(roll / 2), 0, 0, np.cos(roll / 2)])
    qy = np.array([0, np.sin(pitch / 2), 0, np.cos(pitch / 2)])
    qz = np.array([0, 0, np.sin(yaw / 2), np.cos(yaw / 2)])
    q = _quat_mul(_quat_mul(qz, qy), qx)
    return np.array([t[3], t[4], t[5], -q[1], -q[2], q[0], q[3]], np.float64)


def odometry_to_transform(o7):
    """utility.h:96-110: tf::Matrix3x3(tf::Quaternion(o.z, -o.x, -o.y, o.w)).getRPY, transform = (-pitch, -yaw, roll, pos)."""
    o = np.asarray(o7, np.float64)
    x, y, z, w = o[5], -o[3], -o[4], o[6]
    n = x * x + y * y + z * z + w * w
    x, y, z, w = np.array([x, y, z, w]) / np.sqrt(n)
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    pitch = -np.arcsin(R[2, 0])
    roll = np.arctan2(R[2, 1], R[2, 2])
    yaw = np.arctan2(R[1, 0], R[0, 0])
    return np.array([-pitch, -yaw, roll, o[0], o[1], o[2]]).astype(np.float32)
