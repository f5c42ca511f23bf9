"""CPU tests of the arena world (lego_loam_bor_b200/synth/synth_arena.*): the 120 m arena with 500 key-frame poses on a
spiral that BASELINE.json configs[3]/[4] and SURVEY.md section 8d (way 1) ask for, and the key-frame map the oracle
builds from it."""
import numpy as np
import pytest

from lego_loam_bor_b200 import config_params, synth


def test_sector_culling_equals_every_box_for_every_ray(built):
    """The host generator culls boxes by azimuth sector; the device generator tests every box.  Same scans, bit for bit."""
    p = config_params("T")
    cfg = synth.make_arena(p, n_keyframes=60)
    for seq, kind, idx in [(0, synth.KEYFRAME, 0), (0, synth.KEYFRAME, 59), (3, synth.DRIVE, 0), (3, synth.DRIVE, 41), (7, synth.KEYFRAME, 17)]:
        a = synth.arena_scan(cfg, seq, kind, idx)
        b = synth.arena_scan(cfg, seq, kind, idx, bruteforce=True)
        assert len(a) > 1000 and np.array_equal(a, b), (seq, kind, idx)
    pc = config_params("C")
    cfgc = synth.make_arena(pc)
    a = synth.arena_scan(cfgc, 1, synth.KEYFRAME, 250)
    assert np.array_equal(a, synth.arena_scan(cfgc, 1, synth.KEYFRAME, 250, bruteforce=True))
    assert 60000 < len(a) <= 64 * 2048


def test_keyframe_poses_follow_survey_8d(built):
    """500 poses >= 1 m apart inside a 50 m radius (SURVEY 8d way 1), all within the 50 m key-frame search radius of
    every point of the driving circle, none inside an obstacle."""
    p = config_params("C")
    cfg = synth.make_arena(p)
    for seq in (0, 5):
        P = np.array([synth.arena_pose(cfg, seq, synth.KEYFRAME, i) for i in range(500)])
        d = np.linalg.norm(P[:, None, :2] - P[None, :, :2], axis=2) + np.eye(500) * 1e9
        assert d.min() >= 1.0 and d.min() > 2.0
        r = np.hypot(P[:, 0], P[:, 1])
        assert r.max() < 50.0 - cfg.radius  # inside the search radius from anywhere on the circle
        W = synth.arena_world(cfg, seq)
        assert len(W) > 150
        D = np.array([synth.arena_pose(cfg, seq, synth.DRIVE, f) for f in range(0, 700, 7)])
        for pts in (P, D):
            for b in W[4:]:
                dx = np.maximum(np.maximum(b[0] - pts[:, 0], pts[:, 0] - b[3]), 0)
                dy = np.maximum(np.maximum(b[1] - pts[:, 1], pts[:, 1] - b[4]), 0)
                assert np.hypot(dx, dy).min() > 0.8
        step = np.linalg.norm(np.diff(D[:, :2], axis=0), axis=1)
        assert np.allclose(step, 0.7, atol=0.01)  # 1 m/s at 10 Hz, every 7th frame


def test_scans_are_deterministic_and_sequence_specific(built):
    p = config_params("T")
    cfg = synth.make_arena(p, n_keyframes=20)
    a = synth.arena_scan(cfg, 2, synth.DRIVE, 5)
    assert np.array_equal(a, synth.arena_scan(cfg, 2, synth.DRIVE, 5))
    assert not np.array_equal(a[:100], synth.arena_scan(cfg, 3, synth.DRIVE, 5)[:100])
    assert not np.array_equal(a[:100], synth.arena_scan(cfg, 2, synth.KEYFRAME, 5)[:100])
    r = np.linalg.norm(a[:, :3], axis=1)
    assert r.min() >= 0.5 - 0.05 and r.max() <= 100.05 and np.all(a[:, 3] == 0)


def test_oracle_localises_against_the_keyframe_map(built):
    """The whole kf workload on the CPU oracle at the tiny sensor: key frames stored at the true spiral poses, then a drive
    whose scan-to-map result must stay near the true pose (the map frame is the world frame) -- this pins the pose
    convention of synth.pose_to_transform against mapOptmization.cpp:412-426 / featureAssociation.cpp:165-167."""
    from oracle.oracle_py import Oracle
    p = config_params("T")
    K = 40
    cfg = synth.make_arena(p, n_keyframes=K)
    o = Oracle(p, libm=True)
    zero = np.zeros(6, np.float32)
    for i in range(K):
        o.reset_feature_association()
        o.image_projection(synth.arena_scan(cfg, 0, synth.KEYFRAME, i))
        assert o.feature_association() == 0
        o.map_downsample_current_scan()
        T = synth.pose_to_transform(synth.arena_pose(cfg, 0, synth.KEYFRAME, i))
        o.map_set_poses(T, zero)
        o.map_set_initial_guess(T)
        o.map_save_keyframe()
    assert o.download("KEYFRAME_STATE")[0] == K
    assert np.allclose(o.download("KEY_POSES_6D")[7], synth.pose_to_transform(synth.arena_pose(cfg, 0, synth.KEYFRAME, 7)))
    o.reset_feature_association()
    o.map_set_poses(synth.pose_to_transform(synth.arena_pose(cfg, 0, synth.DRIVE, 0)), zero)
    cycles = 0
    for f in range(16):
        o.image_projection(synth.arena_scan(cfg, 0, synth.DRIVE, f))
        if o.feature_association() == 1:
            o.mapping_cycle()
            cycles += 1
            truth = synth.pose_to_transform(synth.arena_pose(cfg, 0, synth.DRIVE, f))
            aft = o.download("TRANSFORM_AFT_MAPPED")
            rot = np.abs((aft[:3] - truth[:3] + np.pi) % (2 * np.pi) - np.pi)
            assert rot.max() < 0.02 and np.abs(aft[3:] - truth[3:]).max() < 0.25, (f, aft, truth)
            assert o.download("MAP_ITERS")[0] >= 1 and len(o.download("MAP_SURF")) > 2000
    assert cycles == 3 and o.download("KEYFRAME_STATE")[0] == K + 3
