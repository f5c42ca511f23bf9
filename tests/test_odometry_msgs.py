"""nav_msgs/Odometry outputs of the path: FeatureAssociation::publishOdometry
(featureAssociation.cpp:1286-1298), MapOptimization::publishTF (mapOptmization.cpp:510-530) and the consumer's
OdometryToTransform (utility.h:96-110).  CPU: the oracle's restatement (rotation product, numpy) against
hand-derivable quaternions, the library's host functions against the oracle, and the round trip the reference sends every
odometry pose through.  GPU: ll_get_odometry against the oracle on the poses of a short run."""
import numpy as np
import pytest

from oracle import oracle_py


def test_oracle_known_quaternions():
    th = float(np.float32(0.3))
    s, c = np.sin(th / 2), np.cos(th / 2)
    # LOAM's camera-style axes: transform[0..2] are rotations about x, y, z of the odometry frame
    cases = [([0, 0, 0, 1, 2, 3], [1, 2, 3, 0, 0, 0, 1]),
             ([th, 0, 0, 0, 0, 0], [0, 0, 0, s, 0, 0, c]),
             ([0, th, 0, 0, 0, 0], [0, 0, 0, 0, s, 0, c]),
             ([0, 0, th, 0, 0, 0], [0, 0, 0, 0, 0, s, c])]
    for t, want in cases:
        got = oracle_py.transform_to_odometry(np.array(t, np.float32))
        assert np.allclose(got, np.array(want, np.float64), rtol=0, atol=1e-15), (t, got, want)
    # unit norm
    rng = np.random.default_rng(3)
    for _ in range(100):
        q = oracle_py.transform_to_odometry(rng.uniform(-1.4, 1.4, 6).astype(np.float32))[3:]
        assert abs(np.dot(q, q) - 1.0) < 1e-14


def test_library_host_functions_match_oracle(built):
    from lego_loam_bor_b200 import capi
    rng = np.random.default_rng(5)
    for _ in range(2000):
        t = np.concatenate([rng.uniform(-1.5, 1.5, 3), rng.uniform(-100, 100, 3)]).astype(np.float32)
        o_lib, o_ref = capi.transform_to_odometry(t), oracle_py.transform_to_odometry(t)
        assert np.allclose(o_lib, o_ref, rtol=0, atol=4e-16), (t, o_lib, o_ref)
        assert np.array_equal(o_lib[:3], t[3:].astype(np.float64))
        t_lib, t_ref = capi.odometry_to_transform(o_lib), oracle_py.odometry_to_transform(o_ref)
        # the round trip of featureAssociation.cpp:1287-1297 -> utility.h:96-110 is the identity to one float ulp
        # away from gimbal lock (|rx| < 1.5 here), which is what lets the oracle restate it as the identity
        assert np.all(np.abs(t_lib - t_ref) <= 2.4e-7 * np.maximum(1.0, np.abs(t_ref))), (t, t_lib, t_ref)
        assert np.all(np.abs(t_lib[:3] - t[:3]) <= 2.4e-7 * np.maximum(1.0, np.abs(t[:3])) * 4), (t, t_lib)
        assert np.array_equal(t_lib[3:], t[3:])


def test_small_angle_round_trip_is_exact_mostly(built):
    """On poses like the path's (rotations of a few degrees) the round trip returns the same floats almost always."""
    from lego_loam_bor_b200 import capi
    rng = np.random.default_rng(7)
    same = 0
    n = 2000
    for _ in range(n):
        t = np.concatenate([rng.uniform(-0.1, 0.1, 3), rng.uniform(-50, 50, 3)]).astype(np.float32)
        back = capi.odometry_to_transform(capi.transform_to_odometry(t))
        assert np.all(np.abs(back[:3] - t[:3]) <= 1.5e-8)  # two float ulps at 0.1 rad
        same += int(np.array_equal(back, t))
    assert same > 0.9 * n


@pytest.mark.gpu
def test_get_odometry_matches_oracle(built):
    from lego_loam_bor_b200 import synth
    from lego_loam_bor_b200.capi import LegoLoam
    from parity_utils import make_scans
    seqs = [0, 4]
    p, cfg, scans = make_scans("A", seqs, range(7))
    gpu = LegoLoam(p, batch=len(seqs))
    aft = np.zeros((len(seqs), 6), np.float32)
    for k, s in enumerate(seqs):
        x, y, z, roll, pitch, yaw = synth.pose(cfg, s, 0)
        aft[k] = [0, yaw, 0, y, z, x]
        gpu.map_set_local(k, synth.local_map(cfg, s, 1, 0.2), synth.local_map(cfg, s, 0, 0.4))
    gpu.map_set_poses(aft, np.zeros_like(aft))
    for f in range(7):
        gpu.set_scans_host([scans[(s, f)] for s in seqs])
        gpu.process_scans()
    lo, am = gpu.odometry()
    for k in range(len(seqs)):
        ts, ta, tb = gpu.download("TRANSFORM_SUM", k), gpu.download("TRANSFORM_AFT_MAPPED", k), gpu.download("TRANSFORM_BEF_MAPPED", k)
        assert np.any(ts != 0) and np.any(ta != 0)
        assert np.allclose(lo[k], oracle_py.transform_to_odometry(ts), rtol=0, atol=4e-16)
        assert np.allclose(am[k, :7], oracle_py.transform_to_odometry(ta), rtol=0, atol=4e-16)
        assert np.array_equal(am[k, 7:], tb.astype(np.float64))
