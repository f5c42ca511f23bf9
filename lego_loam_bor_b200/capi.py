"""ctypes binding of the C ABI in include/lego_loam_b200.h (what tests/ and bench.py call).

There is no CPU fallback: importing works anywhere, but creating a `LegoLoam` needs the CUDA
shared library built by `lego_loam_bor_b200.build` and a CUDA device."""
import ctypes as C
import os

import numpy as np

from ._paths import LIB_CUDA
from .params import LegoLoamParams

# name -> (ll_buffer id, numpy dtype, columns); mirrors `enum ll_buffer`
BUFFERS = {
    "RANGE_MAT": (0, np.float32, 1), "FULL_CLOUD": (1, np.float32, 4), "GROUND_MAT": (2, np.int8, 1),
    "LABEL_MAT": (3, np.int32, 1), "SEG_CLOUD": (4, np.float32, 4), "SEG_GROUND_FLAG": (5, np.uint8, 1),
    "SEG_COL_IND": (6, np.uint32, 1), "SEG_RANGE": (7, np.float32, 1), "START_RING_INDEX": (8, np.int32, 1),
    "END_RING_INDEX": (9, np.int32, 1), "ORIENTATION": (10, np.float32, 1), "OUTLIER_CLOUD": (11, np.float32, 4),
    "CLOUD_CURVATURE": (12, np.float32, 1), "NEIGHBOR_PICKED": (13, np.int32, 1), "CLOUD_LABEL": (14, np.int32, 1),
    "CORNER_SHARP": (15, np.float32, 4), "CORNER_LESS_SHARP": (16, np.float32, 4), "SURF_FLAT": (17, np.float32, 4),
    "SURF_LESS_FLAT": (18, np.float32, 4), "CORNER_SHARP_IND": (19, np.int32, 1),
    "CORNER_LESS_SHARP_IND": (20, np.int32, 1), "SURF_FLAT_IND": (21, np.int32, 1),
    "CORNER_LAST": (22, np.float32, 4), "SURF_LAST": (23, np.float32, 4), "TRANSFORM_CUR": (24, np.float32, 1),
    "TRANSFORM_SUM": (25, np.float32, 1), "ODOM_ITERS": (26, np.int32, 1), "MAP_CORNER": (27, np.float32, 4),
    "MAP_SURF": (28, np.float32, 4), "SCAN_CORNER_DS": (29, np.float32, 4), "SCAN_SURF_TOTAL_DS": (30, np.float32, 4),
    "TRANSFORM_TOBE_MAPPED": (31, np.float32, 1), "MAP_ITERS": (32, np.int32, 1),
    "OUTLIER_LAST": (33, np.float32, 4), "SURF_LESS_FLAT_RAW_COUNT": (34, np.int32, 1),
    "MAP_TRACE": (35, np.float64, 1), "TRANSFORM_BEF_MAPPED": (36, np.float32, 1),
    "TRANSFORM_AFT_MAPPED": (37, np.float32, 1), "SCAN_SURF_DS": (38, np.float32, 4),
    "SCAN_OUTLIER_DS": (39, np.float32, 4), "STAGE_CLOCKS": (40, np.int64, 1),
    "KEYFRAME_STATE": (41, np.int32, 1), "KEY_POSES_6D": (42, np.float32, 6), "SURROUNDING_KEY_IDS": (43, np.int32, 1),
    "INPUT_CLOUD": (44, np.float32, 4), "MAP_KNN_IDX": (45, np.int32, 5), "ODOM_SEARCH_IDX": (46, np.int32, 3),
    "RING_CLOCKS": (47, np.int64, 10),
}

EXPORTS = [
    "ll_default_params", "ll_create", "ll_destroy", "ll_reset", "ll_reset_feature_association", "ll_last_error", "ll_kernel_launches",
    "ll_set_scans_host", "ll_set_scans_device", "ll_image_projection", "ll_feature_association",
    "ll_map_set_local", "ll_map_set_scan", "ll_map_downsample_current_scan", "ll_map_set_initial_guess",
    "ll_map_set_initial_guess_async", "ll_map_set_poses", "ll_map_set_odometry", "ll_map_predict_pose",
    "ll_scan_to_map", "ll_process_scans", "ll_get_poses", "ll_get_poses_async", "ll_wait_poses", "ll_download", "ll_upload", "ll_synchronize", "ll_join_mapping",
    "ll_enable_stage_timing", "ll_enable_index_trace", "ll_get_stage_times_ms", "ll_time_kernel", "ll_get_kernel_time",
    "ll_get_kernel_time_table",
    "ll_map_enable_keyframes", "ll_map_extract_surrounding_keyframes", "ll_map_save_keyframe", "ll_mapping_cycle",
    "ll_map_download_keyframe", "ll_set_scans_pointcloud2_host", "ll_set_scans_xyz_host",
    "ll_transform_to_odometry", "ll_odometry_to_transform", "ll_get_odometry",
    "ll_bag_open", "ll_bag_num_messages", "ll_bag_topic", "ll_bag_get_pointcloud2", "ll_bag_close", "ll_bag_last_error",
]

_lib = None


class LegoLoamError(RuntimeError):
    pass


class PointCloud2View(C.Structure):
    """ll_pointcloud2_view of include/lego_loam_b200.h"""
    _fields_ = [("bag_time_ns", C.c_uint64), ("stamp_sec", C.c_uint32), ("stamp_nsec", C.c_uint32),
                ("height", C.c_uint32), ("width", C.c_uint32), ("point_step", C.c_uint32), ("row_step", C.c_uint32),
                ("is_bigendian", C.c_int32), ("is_dense", C.c_int32),
                ("off_x", C.c_int32), ("off_y", C.c_int32), ("off_z", C.c_int32), ("off_intensity", C.c_int32),
                ("data", C.c_void_p), ("data_len", C.c_uint64)]


class RosBag:
    """ll_bag_*: the sensor_msgs/PointCloud2 messages of one topic of a rosbag v2.0 file, in time order (host code only)."""

    def __init__(self, path, topic=None):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.ll_bag_open(os.fsencode(path), topic.encode() if topic else None, C.byref(h))
        if rc != 0:
            raise LegoLoamError(f"ll_bag_open({path}): {self.lib.ll_bag_last_error().decode()}")
        self.h = h

    def __len__(self):
        return int(self.lib.ll_bag_num_messages(self.h))

    @property
    def topic(self):
        return self.lib.ll_bag_topic(self.h).decode()

    def message(self, i):
        """(view, data): the PointCloud2 header fields and a uint8 copy of its `data` array"""
        v = PointCloud2View()
        if self.lib.ll_bag_get_pointcloud2(self.h, i, C.byref(v)) != 0:
            raise LegoLoamError(f"ll_bag_get_pointcloud2({i}): {self.lib.ll_bag_last_error().decode()}")
        data = np.ctypeslib.as_array(C.cast(v.data, C.POINTER(C.c_uint8)), shape=(int(v.data_len),)).copy() if v.data_len else np.zeros(0, np.uint8)
        return v, data

    def close(self):
        if self.h:
            self.lib.ll_bag_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def transform_to_odometry(t6):
    """ll_transform_to_odometry: pose 6-vector -> position + orientation quaternion (host function, no GPU needed)."""
    t = np.ascontiguousarray(t6, np.float32)
    o = np.zeros(7, np.float64)
    load_library().ll_transform_to_odometry(t.ctypes.data, o.ctypes.data)
    return o


def odometry_to_transform(o7):
    """ll_odometry_to_transform: OdometryToTransform of utility.h:96-110."""
    o = np.ascontiguousarray(o7, np.float64)
    t = np.zeros(6, np.float32)
    load_library().ll_odometry_to_transform(o.ctypes.data, t.ctypes.data)
    return t


def load_library(path=None):
    """dlopen the CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_CUDA
    if not os.path.exists(path):
        raise LegoLoamError(f"{path} is missing: run `python -m lego_loam_bor_b200.build` (there is no CPU fallback)")
    lib = C.CDLL(path)
    vp, ip, sz = C.c_void_p, C.c_int, C.c_size_t
    lib.ll_default_params.argtypes = [vp]
    lib.ll_default_params.restype = None
    lib.ll_create.argtypes = [vp, ip, ip, ip, vp, C.POINTER(vp)]
    lib.ll_destroy.argtypes = [vp]
    lib.ll_reset.argtypes = [vp]
    lib.ll_reset_feature_association.argtypes = [vp]
    lib.ll_last_error.argtypes = [vp]
    lib.ll_last_error.restype = C.c_char_p
    lib.ll_kernel_launches.argtypes = [vp]
    lib.ll_kernel_launches.restype = C.c_int64
    lib.ll_set_scans_host.argtypes = [vp, vp, vp, ip]
    lib.ll_set_scans_xyz_host.argtypes = [vp, vp, vp, ip]
    lib.ll_transform_to_odometry.argtypes = [vp, vp]
    lib.ll_transform_to_odometry.restype = None
    lib.ll_odometry_to_transform.argtypes = [vp, vp]
    lib.ll_odometry_to_transform.restype = None
    lib.ll_get_odometry.argtypes = [vp, vp, vp]
    lib.ll_bag_open.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(vp)]
    lib.ll_bag_num_messages.argtypes = [vp]
    lib.ll_bag_topic.argtypes = [vp]
    lib.ll_bag_topic.restype = C.c_char_p
    lib.ll_bag_get_pointcloud2.argtypes = [vp, ip, vp]
    lib.ll_bag_close.argtypes = [vp]
    lib.ll_bag_close.restype = None
    lib.ll_bag_last_error.argtypes = []
    lib.ll_bag_last_error.restype = C.c_char_p
    lib.ll_set_scans_device.argtypes = [vp, vp, vp, ip]
    lib.ll_map_set_initial_guess_async.argtypes = [vp, vp]
    lib.ll_map_set_poses.argtypes = [vp, vp, vp]
    lib.ll_map_set_odometry.argtypes = [vp, vp]
    for name in ("ll_image_projection", "ll_feature_association", "ll_map_downsample_current_scan",
                 "ll_scan_to_map", "ll_process_scans", "ll_synchronize", "ll_map_predict_pose", "ll_join_mapping"):
        getattr(lib, name).argtypes = [vp]
    lib.ll_map_set_local.argtypes = [vp, ip, vp, ip, vp, ip]
    lib.ll_map_set_scan.argtypes = [vp, ip, vp, ip, vp, ip]
    lib.ll_map_set_initial_guess.argtypes = [vp, vp]
    lib.ll_get_poses.argtypes = [vp, vp, vp, vp]
    lib.ll_get_poses_async.argtypes = [vp, vp, vp, vp]
    lib.ll_wait_poses.argtypes = [vp]
    lib.ll_download.argtypes = [vp, ip, ip, vp, sz, C.POINTER(sz)]
    lib.ll_upload.argtypes = [vp, ip, ip, vp, sz]
    lib.ll_enable_stage_timing.argtypes = [vp, ip]
    lib.ll_enable_index_trace.argtypes = [vp, ip]
    lib.ll_get_stage_times_ms.argtypes = [vp, vp]
    lib.ll_time_kernel.argtypes = [vp, C.c_char_p]
    lib.ll_get_kernel_time.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_int)]
    lib.ll_get_kernel_time_table.argtypes = [vp, C.c_char_p, sz]
    lib.ll_map_enable_keyframes.argtypes = [vp, ip, ip, ip, ip]
    for name in ("ll_map_extract_surrounding_keyframes", "ll_map_save_keyframe", "ll_mapping_cycle"):
        getattr(lib, name).argtypes = [vp]
    lib.ll_map_download_keyframe.argtypes = [vp, ip, ip, ip, vp, sz, C.POINTER(sz)]
    lib.ll_set_scans_pointcloud2_host.argtypes = [vp, vp, vp, sz, ip, ip, ip, ip, ip, ip]
    if path == LIB_CUDA:
        _lib = lib
    return lib


class LegoLoam:
    """`batch` independent sequences advancing in lock step on one GPU."""

    def __init__(self, params: LegoLoamParams, batch=1, max_points=None, device=0, stream=None):
        self.lib = load_library()
        self.params = params
        self.batch = batch
        self.V, self.H = params.num_vertical_scans, params.num_horizontal_scans
        self.N = self.V * self.H
        self.max_points = max_points or self.N
        h = C.c_void_p()
        rc = self.lib.ll_create(C.addressof(params), batch, self.max_points, device, stream, C.byref(h))
        if rc != 0:
            raise LegoLoamError(f"ll_create failed with status {rc} (no CUDA device, or bad parameters)")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.ll_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc < 0:
            raise LegoLoamError(f"{what} -> {rc}: {self.lib.ll_last_error(self.h).decode()}")
        return rc

    def reset(self):
        self._ck(self.lib.ll_reset(self.h), "ll_reset")

    def reset_feature_association(self):
        self._ck(self.lib.ll_reset_feature_association(self.h), "ll_reset_feature_association")

    def set_scans_host(self, scans):
        """scans: list (len batch) of float32 [n_i, 4] arrays, or (packed [batch, stride, 4], counts)."""
        if isinstance(scans, tuple):
            packed, counts = scans
        else:
            stride = max(1, max(len(s) for s in scans))
            packed = np.zeros((self.batch, stride, 4), np.float32)
            counts = np.zeros(self.batch, np.int32)
            for i, s in enumerate(scans):
                packed[i, :len(s)] = s
                counts[i] = len(s)
        counts = np.ascontiguousarray(counts, np.int32)
        self._keep = (packed, counts)
        self._in_width = 4
        self._ck(self.lib.ll_set_scans_host(self.h, packed.ctypes.data, counts.ctypes.data, packed.shape[1]),
                 "ll_set_scans_host")

    def set_scans_host_ptr(self, ptr, counts, stride):
        counts = np.ascontiguousarray(counts, np.int32)
        self._in_width = 4
        self._ck(self.lib.ll_set_scans_host(self.h, ptr, counts.ctypes.data, stride), "ll_set_scans_host")

    def set_scans_xyz_host(self, scans):
        """Packed 12-byte points: list (len batch) of float32 [n_i, 3] arrays, or (packed [batch, stride, 3], counts)."""
        if isinstance(scans, tuple):
            packed, counts = scans
        else:
            stride = max(1, max(len(s) for s in scans))
            packed = np.zeros((self.batch, stride, 3), np.float32)
            counts = np.zeros(self.batch, np.int32)
            for i, s in enumerate(scans):
                packed[i, :len(s)] = np.asarray(s, np.float32)[:, :3]
                counts[i] = len(s)
        packed = np.ascontiguousarray(packed, np.float32)
        counts = np.ascontiguousarray(counts, np.int32)
        self._keep = (packed, counts)
        self._in_width = 3
        self._ck(self.lib.ll_set_scans_xyz_host(self.h, packed.ctypes.data, counts.ctypes.data, packed.shape[1]),
                 "ll_set_scans_xyz_host")

    def set_scans_xyz_host_ptr(self, ptr, counts, stride):
        counts = np.ascontiguousarray(counts, np.int32)
        self._in_width = 3
        self._ck(self.lib.ll_set_scans_xyz_host(self.h, ptr, counts.ctypes.data, stride), "ll_set_scans_xyz_host")

    def set_scans_pointcloud2(self, messages, point_step, off_x, off_y, off_z, off_intensity, is_dense=False):
        """messages: list (len batch) of uint8 arrays, the `data` of one sensor_msgs/PointCloud2 each."""
        stride = max(1, max(len(m) for m in messages))
        packed = np.zeros((self.batch, stride), np.uint8)
        counts = np.zeros(self.batch, np.int32)
        for i, m in enumerate(messages):
            packed[i, :len(m)] = m
            counts[i] = len(m) // point_step
        self._ck(self.lib.ll_set_scans_pointcloud2_host(self.h, packed.ctypes.data, counts.ctypes.data, stride, point_step,
                                                        off_x, off_y, off_z, off_intensity, 1 if is_dense else 0),
                 "ll_set_scans_pointcloud2_host")

    def set_scans_pointcloud2_ptr(self, ptr, counts, stride_bytes, point_step, off_x, off_y, off_z, off_intensity, is_dense=False):
        """Raw address of host bytes [batch][stride_bytes]: the `data` arrays of one PointCloud2 message per sequence."""
        counts = np.ascontiguousarray(counts, np.int32)
        self._in_width = 4
        self._ck(self.lib.ll_set_scans_pointcloud2_host(self.h, ptr, counts.ctypes.data, stride_bytes, point_step,
                                                        off_x, off_y, off_z, off_intensity, 1 if is_dense else 0),
                 "ll_set_scans_pointcloud2_host")

    def set_scans_device(self, dev_ptr, counts, stride):
        counts = np.ascontiguousarray(counts, np.int32)
        self._ck(self.lib.ll_set_scans_device(self.h, dev_ptr, counts.ctypes.data, stride), "ll_set_scans_device")

    def image_projection(self):
        return self._ck(self.lib.ll_image_projection(self.h), "ll_image_projection")

    def feature_association(self):
        return self._ck(self.lib.ll_feature_association(self.h), "ll_feature_association")

    def process_scans(self):
        return self._ck(self.lib.ll_process_scans(self.h), "ll_process_scans")

    def map_set_local(self, seq, corner, surf):
        corner = np.ascontiguousarray(corner, np.float32)
        surf = np.ascontiguousarray(surf, np.float32)
        self._ck(self.lib.ll_map_set_local(self.h, seq, corner.ctypes.data, len(corner), surf.ctypes.data, len(surf)),
                 "ll_map_set_local")

    def map_set_scan(self, seq, corner, surf_total):
        corner = np.ascontiguousarray(corner, np.float32)
        surf_total = np.ascontiguousarray(surf_total, np.float32)
        self._ck(self.lib.ll_map_set_scan(self.h, seq, corner.ctypes.data, len(corner), surf_total.ctypes.data,
                                          len(surf_total)), "ll_map_set_scan")

    def map_downsample_current_scan(self):
        self._ck(self.lib.ll_map_downsample_current_scan(self.h), "ll_map_downsample_current_scan")

    def map_set_initial_guess(self, t):
        t = np.ascontiguousarray(t, np.float32).reshape(self.batch, 6)
        self._ck(self.lib.ll_map_set_initial_guess(self.h, t.ctypes.data), "ll_map_set_initial_guess")

    def map_set_poses(self, aft, bef):
        aft = np.ascontiguousarray(aft, np.float32).reshape(self.batch, 6)
        bef = np.ascontiguousarray(bef, np.float32).reshape(self.batch, 6)
        self._ck(self.lib.ll_map_set_poses(self.h, aft.ctypes.data, bef.ctypes.data), "ll_map_set_poses")

    def map_predict_pose(self):
        self._ck(self.lib.ll_map_predict_pose(self.h), "ll_map_predict_pose")

    def map_set_odometry(self, transform_sum):
        t = np.ascontiguousarray(transform_sum, np.float32).reshape(self.batch, 6)
        self._ck(self.lib.ll_map_set_odometry(self.h, t.ctypes.data), "ll_map_set_odometry")

    def scan_to_map(self):
        self._ck(self.lib.ll_scan_to_map(self.h), "ll_scan_to_map")

    def map_enable_keyframes(self, max_keyframes=256, pool_points=None, max_map_corner=None, max_map_surf=None):
        """Key frames and the local map on the device (MapOptimization::saveKeyFramesAndFactor /
        extractSurroundingKeyFrames); after this, process_scans runs whole mapping cycles."""
        pool_points = pool_points or max_keyframes * (self.N // 4)
        max_map_corner = max_map_corner or self.N
        max_map_surf = max_map_surf or 2 * self.N
        self._ck(self.lib.ll_map_enable_keyframes(self.h, max_keyframes, pool_points, max_map_corner, max_map_surf),
                 "ll_map_enable_keyframes")

    def map_extract_surrounding_keyframes(self):
        return self._ck(self.lib.ll_map_extract_surrounding_keyframes(self.h), "ll_map_extract_surrounding_keyframes")

    def map_save_keyframe(self):
        return self._ck(self.lib.ll_map_save_keyframe(self.h), "ll_map_save_keyframe")

    def mapping_cycle(self):
        return self._ck(self.lib.ll_mapping_cycle(self.h), "ll_mapping_cycle")

    def download_keyframe(self, seq, keyframe, which):
        n = C.c_size_t(0)
        self._ck(self.lib.ll_map_download_keyframe(self.h, seq, keyframe, which, None, 0, C.byref(n)), "ll_map_download_keyframe")
        out = np.empty((n.value, 4), np.float32)
        self._ck(self.lib.ll_map_download_keyframe(self.h, seq, keyframe, which, out.ctypes.data, out.nbytes, C.byref(n)),
                 "ll_map_download_keyframe")
        return out

    def synchronize(self):
        self._ck(self.lib.ll_synchronize(self.h), "ll_synchronize")

    def join_mapping(self):
        self._ck(self.lib.ll_join_mapping(self.h), "ll_join_mapping")

    def poses(self):
        ts = np.zeros((self.batch, 6), np.float32)
        tc = np.zeros((self.batch, 6), np.float32)
        tm = np.zeros((self.batch, 6), np.float32)
        self._ck(self.lib.ll_get_poses(self.h, ts.ctypes.data, tc.ctypes.data, tm.ctypes.data), "ll_get_poses")
        return ts, tc, tm

    def odometry(self):
        """(laser_odometry f64[batch, 7], odom_aft_mapped f64[batch, 13]): the nav_msgs/Odometry fields of the path."""
        lo = np.zeros((self.batch, 7), np.float64)
        am = np.zeros((self.batch, 13), np.float64)
        self._ck(self.lib.ll_get_odometry(self.h, lo.ctypes.data, am.ctypes.data), "ll_get_odometry")
        return lo, am

    def poses_async(self, ts_ptr, tc_ptr, tm_ptr):
        """Enqueue the pose copies into caller-owned (pinned) float32 [batch, 6] buffers given as raw addresses."""
        self._ck(self.lib.ll_get_poses_async(self.h, ts_ptr, tc_ptr, tm_ptr), "ll_get_poses_async")

    def wait_poses(self):
        self._ck(self.lib.ll_wait_poses(self.h), "ll_wait_poses")

    def download(self, name, seq=0):
        bid, dt, w = BUFFERS[name]
        if name == "INPUT_CLOUD":
            w = getattr(self, "_in_width", 4)   # packed xyz scans come back as they went in
        n = C.c_size_t(0)
        self._ck(self.lib.ll_download(self.h, seq, bid, None, 0, C.byref(n)), f"ll_download({name})")
        out = np.empty((n.value, w) if w > 1 else (n.value,), dt)
        self._ck(self.lib.ll_download(self.h, seq, bid, out.ctypes.data, max(out.nbytes, 1), C.byref(n)),
                 f"ll_download({name})")
        return out

    def upload(self, name, arr, seq=0):
        bid, dt, w = BUFFERS[name]
        arr = np.ascontiguousarray(arr, dt)
        self._ck(self.lib.ll_upload(self.h, seq, bid, arr.ctypes.data, arr.size // w), f"ll_upload({name})")

    def kernel_launches(self):
        return int(self.lib.ll_kernel_launches(self.h))

    def enable_index_trace(self, on=True):
        self._ck(self.lib.ll_enable_index_trace(self.h, 1 if on else 0), "ll_enable_index_trace")

    def enable_stage_timing(self, on=True):
        self._ck(self.lib.ll_enable_stage_timing(self.h, 1 if on else 0), "ll_enable_stage_timing")

    def stage_times_ms(self):
        ms = np.zeros(5, np.float32)
        self._ck(self.lib.ll_get_stage_times_ms(self.h, ms.ctypes.data), "ll_get_stage_times_ms")
        return ms

    def time_kernel(self, name):
        self._ck(self.lib.ll_time_kernel(self.h, name.encode() if name else None), "ll_time_kernel")

    def kernel_time(self):
        """(total milliseconds, launches) of the kernel selected with time_kernel()."""
        ms, n = C.c_double(0), C.c_int(0)
        self._ck(self.lib.ll_get_kernel_time(self.h, C.byref(ms), C.byref(n)), "ll_get_kernel_time")
        return ms.value, n.value

    def kernel_time_table(self):
        """After time_kernel("*"): {kernel name: (total ms, launches)}."""
        buf = C.create_string_buffer(1 << 16)
        self._ck(self.lib.ll_get_kernel_time_table(self.h, buf, len(buf)), "ll_get_kernel_time_table")
        out = {}
        for line in buf.value.decode().splitlines():
            name, ms, n = line.split()
            out[name] = (float(ms), int(n))
        return out


class LegoLoamStreams:
    """`batch` independent sequences split over `n_streams` handles, each with its own CUDA stream.

    Sequences never interact, so the sub-batches may run concurrently: while the few slow sequences of
    one sub-batch finish their LM iterations (kernels that occupy a handful of SMs), the wide kernels of the
    other sub-batches fill the rest of the GPU.  Same call surface as `LegoLoam`; `streams` are the raw
    cudaStream_t handles (e.g. from torch.cuda.Stream().cuda_stream) or None to let each handle create one."""

    def __init__(self, params, batch, n_streams, max_points=None, device=0, streams=None):
        if batch % n_streams:
            raise ValueError("batch must be a multiple of n_streams")
        self.batch, self.n, self.sub = batch, n_streams, batch // n_streams
        self.parts = [LegoLoam(params, self.sub, max_points, device, streams[i] if streams else None)
                      for i in range(n_streams)]
        self.max_points = self.parts[0].max_points

    def _loc(self, seq):
        return self.parts[seq // self.sub], seq % self.sub

    def reset(self):
        for p in self.parts:
            p.reset()

    def reset_feature_association(self):
        for p in self.parts:
            p.reset_feature_association()

    def map_set_initial_guess(self, t):
        t = np.ascontiguousarray(t, np.float32).reshape(self.batch, 6)
        for i, p in enumerate(self.parts):
            p.map_set_initial_guess(t[i * self.sub:(i + 1) * self.sub])

    def map_save_keyframe(self):
        for p in self.parts:
            p.map_save_keyframe()

    def map_downsample_current_scan(self):
        for p in self.parts:
            p.map_downsample_current_scan()

    def image_projection(self):
        for p in self.parts:
            p.image_projection()

    def feature_association(self):
        return [p.feature_association() for p in self.parts][0]

    def set_scans_device(self, dev_ptr, counts, stride):
        for i, p in enumerate(self.parts):
            p.set_scans_device(dev_ptr + i * self.sub * stride * 16, counts[i * self.sub:(i + 1) * self.sub], stride)

    def set_scans_host_ptr(self, ptr, counts, stride):
        for i, p in enumerate(self.parts):
            p.set_scans_host_ptr(ptr + i * self.sub * stride * 16, counts[i * self.sub:(i + 1) * self.sub], stride)

    def set_scans_pointcloud2_ptr(self, ptr, counts, stride_bytes, *fmt):
        for i, p in enumerate(self.parts):
            p.set_scans_pointcloud2_ptr(ptr + i * self.sub * stride_bytes, counts[i * self.sub:(i + 1) * self.sub], stride_bytes, *fmt)

    def set_scans_xyz_host_ptr(self, ptr, counts, stride):
        for i, p in enumerate(self.parts):
            p.set_scans_xyz_host_ptr(ptr + i * self.sub * stride * 12, counts[i * self.sub:(i + 1) * self.sub], stride)

    def process_scans(self):
        return [p.process_scans() for p in self.parts][0]

    def map_set_local(self, seq, corner, surf):
        p, k = self._loc(seq)
        p.map_set_local(k, corner, surf)

    def map_enable_keyframes(self, *a, **kw):
        for p in self.parts:
            p.map_enable_keyframes(*a, **kw)

    def map_set_poses(self, aft, bef):
        aft = np.ascontiguousarray(aft, np.float32).reshape(self.batch, 6)
        bef = np.ascontiguousarray(bef, np.float32).reshape(self.batch, 6)
        for i, p in enumerate(self.parts):
            p.map_set_poses(aft[i * self.sub:(i + 1) * self.sub], bef[i * self.sub:(i + 1) * self.sub])

    def poses(self):
        out = [p.poses() for p in self.parts]
        return tuple(np.concatenate([o[j] for o in out]) for j in range(3))

    def synchronize(self):
        for p in self.parts:
            p.synchronize()

    def join_mapping(self):
        for p in self.parts:
            p.join_mapping()

    def odometry(self):
        """(laser_odometry f64[batch, 7], odom_aft_mapped f64[batch, 13]): the nav_msgs/Odometry fields of the path."""
        lo = np.zeros((self.batch, 7), np.float64)
        am = np.zeros((self.batch, 13), np.float64)
        self._ck(self.lib.ll_get_odometry(self.h, lo.ctypes.data, am.ctypes.data), "ll_get_odometry")
        return lo, am

    def poses_async(self, ts_ptr, tc_ptr, tm_ptr):
        for i, p in enumerate(self.parts):
            o = i * self.sub * 24
            p.poses_async(ts_ptr + o, tc_ptr + o, tm_ptr + o)

    def wait_poses(self):
        for p in self.parts:
            p.wait_poses()

    def download(self, name, seq=0):
        p, k = self._loc(seq)
        return p.download(name, k)

    def kernel_launches(self):
        return sum(p.kernel_launches() for p in self.parts)

    def time_kernel(self, name):
        for p in self.parts:
            p.time_kernel(name)

    def kernel_time(self):
        ms, n = 0.0, 0
        for p in self.parts:
            a, b = p.kernel_time()
            ms, n = ms + a, n + b
        return ms, n

    def kernel_time_table(self):
        out = {}
        for p in self.parts:
            for k, (ms, n) in p.kernel_time_table().items():
                a, b = out.get(k, (0.0, 0))
                out[k] = (a + ms, b + n)
        return out
