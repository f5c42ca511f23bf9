"""GPU parity of the scan-to-map slice (MapOptimization::downsampleCurrentScan and
scan2MapOptimization, mapOptmization.cpp:999-1026,1028-1332) against the CPU oracle."""
import numpy as np
import pytest

from parity_utils import make_scans

pytestmark = pytest.mark.gpu

POSE_TOL_M = 1e-4
POSE_TOL_RAD = 1e-5


def _true_guess(cfg, seq, frame, perturb):
    """transformTobeMapped (rx, ry, rz, tx, ty, tz) of the synthetic sensor pose in the reference's
    camera axes (x<-y, y<-z, z<-x), plus a perturbation so that the LM has work to do."""
    from lego_loam_bor_b200 import synth
    x, y, z, roll, pitch, yaw = synth.pose(cfg, seq, frame)
    return np.array([0.0, yaw, 0.0, y, z, x], np.float64) + perturb


def _run(cfgname, seqs, n_frames=6):
    from lego_loam_bor_b200 import synth
    from lego_loam_bor_b200.capi import LegoLoam
    from oracle.oracle_py import Oracle
    p, cfg, scans = make_scans(cfgname, seqs, range(n_frames))
    gpu = LegoLoam(p, batch=len(seqs))
    oracles = [Oracle(p) for _ in seqs]
    handed = 0
    for f in range(n_frames):
        gpu.set_scans_host([scans[(s, f)] for s in seqs])
        gpu.image_projection()
        rc = gpu.feature_association()
        ro = []
        for k, s in enumerate(seqs):
            oracles[k].image_projection(scans[(s, f)])
            ro.append(oracles[k].feature_association())
        assert all(r == rc for r in ro), "mapping hand-over cadence differs (featureAssociation.cpp:1432)"
        if rc != 1:
            continue
        handed += 1
        gpu.map_downsample_current_scan()
        guess = np.zeros((len(seqs), 6), np.float32)
        for k, s in enumerate(seqs):
            o = oracles[k]
            o.map_downsample_current_scan()
            for name in ("SCAN_CORNER_DS", "SCAN_SURF_TOTAL_DS"):
                a, b = gpu.download(name, k), o.download(name)
                assert a.shape == b.shape, f"{name}: {a.shape} vs {b.shape}"
                assert np.array_equal(a, b), f"{name} differs: max {np.abs(a - b).max()}"
            corner_map = synth.local_map(cfg, s, 1, 0.2)
            surf_map = synth.local_map(cfg, s, 0, 0.4)
            gpu.map_set_local(k, corner_map, surf_map)
            o.map_set_local(corner_map, surf_map)
            guess[k] = _true_guess(cfg, s, f, np.array([0.002, 0.01, -0.002, 0.05, 0.02, -0.05]))
            o.map_set_initial_guess(guess[k])
        gpu.map_set_initial_guess(guess)
        gpu.scan_to_map()
        for k, s in enumerate(seqs):
            o = oracles[k]
            o.scan_to_map()
            ia, ib = gpu.download("MAP_ITERS", k), o.download("MAP_ITERS")
            assert np.array_equal(ia, ib), f"scan-to-map iterations/rows {ia} vs {ib}"
            assert ib[0] >= 2 and ib[1] >= 50, f"test scene too weak: {ib}"
            a, b = gpu.download("TRANSFORM_TOBE_MAPPED", k), o.download("TRANSFORM_TOBE_MAPPED")
            assert np.all(np.abs(a[:3] - b[:3]) <= POSE_TOL_RAD), f"rot {a} vs {b}"
            assert np.all(np.abs(a[3:] - b[3:]) <= POSE_TOL_M), f"trans {a} vs {b}"
            # the LM must actually have pulled the perturbed guess back towards the truth
            truth = _true_guess(cfg, s, f, 0)
            assert np.linalg.norm(a[3:] - truth[3:]) < np.linalg.norm(guess[k][3:] - truth[3:])
    assert handed >= 1


def test_scan_to_map_vlp16(built):
    _run("A", [0, 2])


def test_scan_to_map_64_beam(built):
    _run("C", [1])


def test_full_pipeline_with_pose_chain(built):
    """ll_process_scans end to end (every stage, mapping every 5th odometry frame, odometry -> map pose
    chain on the device) against the oracle driven the same way, free-running over 12 frames."""
    from lego_loam_bor_b200 import synth
    from lego_loam_bor_b200.capi import LegoLoam
    from oracle.oracle_py import Oracle
    seqs = [0, 4]
    n_frames = 12
    p, cfg, scans = make_scans("A", seqs, range(n_frames))
    gpu = LegoLoam(p, batch=len(seqs))
    oracles = [Oracle(p) for _ in seqs]
    aft = np.zeros((len(seqs), 6), np.float32)
    for k, s in enumerate(seqs):
        x, y, z, roll, pitch, yaw = synth.pose(cfg, s, 0)
        aft[k] = [0, yaw, 0, y, z, x]
        cm, sm = synth.local_map(cfg, s, 1, 0.2), synth.local_map(cfg, s, 0, 0.4)
        gpu.map_set_local(k, cm, sm)
        oracles[k].map_set_local(cm, sm)
        oracles[k].map_set_poses(aft[k], np.zeros(6, np.float32))
    gpu.map_set_poses(aft, np.zeros_like(aft))
    cycles = 0
    for f in range(n_frames):
        gpu.set_scans_host([scans[(s, f)] for s in seqs])
        rc = gpu.process_scans()
        for k, s in enumerate(seqs):
            o = oracles[k]
            o.image_projection(scans[(s, f)])
            if o.feature_association() == 1:
                assert rc == 1
                o.map_downsample_current_scan()
                o.map_predict_pose()
                o.scan_to_map()
            else:
                assert rc == 0
        cycles += rc
        for k in range(len(seqs)):
            for name in ("TRANSFORM_SUM", "TRANSFORM_TOBE_MAPPED", "TRANSFORM_AFT_MAPPED", "TRANSFORM_BEF_MAPPED"):
                a, b = gpu.download(name, k), oracles[k].download(name)
                assert np.all(np.abs(a[:3] - b[:3]) <= POSE_TOL_RAD), f"frame {f} {name} rot {a} vs {b}"
                assert np.all(np.abs(a[3:] - b[3:]) <= POSE_TOL_M), f"frame {f} {name} trans {a} vs {b}"
    assert cycles == 2
    # the map-frame pose must track the synthetic ground truth to a few centimetres
    for k, s in enumerate(seqs):
        x, y, z, roll, pitch, yaw = synth.pose(cfg, s, 10)  # last mapped frame
        a = gpu.download("TRANSFORM_AFT_MAPPED", k)
        assert abs(a[3] - y) < 0.1 and abs(a[5] - x) < 0.1 and abs(a[4] - z) < 0.1, (a, (x, y, z))
