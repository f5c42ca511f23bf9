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
