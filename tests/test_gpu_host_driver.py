"""The C++ host mirror of the reference's stage classes (ImageProjection / FeatureAssociation /
MapOptimization + Channels, lego_loam_bor_b200/host) driven by the bag-less sequence driver, against the
same sequence pushed through the C ABI from Python."""
import os
import subprocess

import numpy as np
import pytest

from parity_utils import make_scans

pytestmark = pytest.mark.gpu


def test_sequence_driver_matches_c_abi(built, tmp_path):
    from lego_loam_bor_b200._paths import PKG
    from lego_loam_bor_b200.capi import LegoLoam
    n_frames = 17
    p, cfg, scans = make_scans("T", [0], range(n_frames))
    path = tmp_path / "scans.bin"
    with open(path, "wb") as f:
        f.write(np.int32(n_frames).tobytes())
        for i in range(n_frames):
            a = scans[(0, i)]
            f.write(np.int32(len(a)).tobytes())
            f.write(np.ascontiguousarray(a, np.float32).tobytes())
    out = tmp_path / "poses.txt"
    exe = os.path.join(PKG, "host", "sequence_driver")
    r = subprocess.run([exe, "T", str(path), str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    rows = np.loadtxt(out)
    assert rows.shape == (n_frames, 15)
    gpu = LegoLoam(p, batch=1)
    gpu.map_enable_keyframes(max_keyframes=32)
    for i in range(n_frames):
        gpu.set_scans_host([scans[(0, i)]])
        gpu.process_scans()
        ts = gpu.download("TRANSFORM_SUM")
        assert np.array_equal(rows[i, 1:7].astype(np.float32), ts), f"frame {i}: odometry differs"
        # the stage classes run the same mapping cycle (key frames + local map on the device) as ll_process_scans
        am = gpu.download("TRANSFORM_AFT_MAPPED")
        assert np.array_equal(rows[i, 7:13].astype(np.float32), am), f"frame {i}: transformAftMapped differs"
        assert rows[i, 13] == gpu.download("KEYFRAME_STATE")[0], f"frame {i}: key-frame count differs"
    # mapping: frames 5, 10, 15 are handed over; the first cycle only stores a key frame (empty map)
    assert rows[-1, 14] == 3 and rows[-1, 13] >= 2
    aft = rows[-1, 7:13]
    assert np.all(np.isfinite(aft)) and np.abs(aft).max() > 0
    # scan-to-map against a map built from the sequence's own key frames stays close to the odometry pose
    assert np.abs(aft[3:] - rows[15, 4:7]).max() < 0.3 and np.abs(aft[:3] - rows[15, 1:4]).max() < 0.1


def test_stage_threads_overlap_without_changing_the_result(built, tmp_path):
    """--stream: scans are pushed without waiting for the stages, so FeatureAssociation integrates further frames while
    MapOptimization still works on an earlier hand-over (the reference's live mode).  The mapping cycle must use the
    odometry pose that was handed over WITH its scan (AssociationOut::laser_odometry, mapOptmization.cpp:1539), so the
    final poses and key frames equal those of the run that waits after every frame."""
    from lego_loam_bor_b200._paths import PKG
    n_frames = 27
    p, cfg, scans = make_scans("T", [1], range(n_frames))
    path = tmp_path / "scans.bin"
    with open(path, "wb") as f:
        f.write(np.int32(n_frames).tobytes())
        for i in range(n_frames):
            a = scans[(1, i)]
            f.write(np.int32(len(a)).tobytes())
            f.write(np.ascontiguousarray(a, np.float32).tobytes())
    exe = os.path.join(PKG, "host", "sequence_driver")
    rows = []
    for extra in ([], ["--stream"]):
        out = tmp_path / ("poses%d.txt" % len(rows))
        r = subprocess.run([exe, "T", str(path), str(out)] + extra, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        rows.append(np.atleast_2d(np.loadtxt(out)))
    assert rows[0].shape == (n_frames, 15) and rows[1].shape == (1, 15)
    assert np.array_equal(rows[0][-1], rows[1][0]), (rows[0][-1], rows[1][0])
    assert rows[1][0, 14] == 5 and rows[1][0, 13] >= 4


def test_sequence_driver_bag_mode(built, tmp_path):
    """The same sequence recorded as a rosbag (the reference's bag mode, main.cpp:60-76): the driver reads the
    PointCloud2 messages of the lidar topic and hands the raw message bytes to ImageProjection::cloudHandler; every pose
    line must equal the one of the scans.bin run."""
    from lego_loam_bor_b200._paths import PKG
    from test_rosbag import make_bag
    n_frames = 12
    p, cfg, scans = make_scans("T", [0], range(n_frames))
    clouds = [scans[(0, i)] for i in range(n_frames)]
    bin_path, bag_path = tmp_path / "scans.bin", tmp_path / "scans.bag"
    with open(bin_path, "wb") as f:
        f.write(np.int32(n_frames).tobytes())
        for a in clouds:
            f.write(np.int32(len(a)).tobytes())
            f.write(np.ascontiguousarray(a, np.float32).tobytes())
    make_bag(str(bag_path), clouds, np.random.default_rng(9))
    exe = os.path.join(PKG, "host", "sequence_driver")
    outs = []
    for src in (str(bin_path), str(bag_path) + ":/velodyne_points"):
        out = tmp_path / ("poses_%d.txt" % len(outs))
        r = subprocess.run([exe, "T", src, str(out)], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        outs.append(np.loadtxt(out))
    assert outs[0].shape == outs[1].shape == (n_frames, 15)
    assert np.array_equal(outs[0], outs[1])
    r = subprocess.run([exe, "T", str(tmp_path / "missing.bag"), str(tmp_path / "x.txt")], capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "Unable to open rosbag" in r.stderr
