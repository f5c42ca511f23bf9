"""GPU parity of MapOptimization's key frames and local map: saveKeyFramesAndFactor,
extractSurroundingKeyFrames (loop closure off) and transformPointCloud, mapOptmization.cpp:428-473, 856-996,
1335-1474, against the CPU oracle -- through the C ABI, all sequences of a batch at once."""
import numpy as np
import pytest

from parity_utils import make_scans, same_bits

pytestmark = pytest.mark.gpu

POSE_TOL_M = 1e-4
POSE_TOL_RAD = 1e-5


def _compare_maps(gpu, o, k, where, exact):
    st_g, st_o = gpu.download("KEYFRAME_STATE", k), o.download("KEYFRAME_STATE")
    assert st_g[3] == 0, f"{where}: capacity error bits {st_g[3]}"
    assert np.array_equal(st_g[:3], st_o[:3]), f"{where}: key-frame state {st_g} vs {st_o}"
    ids_g, ids_o = gpu.download("SURROUNDING_KEY_IDS", k), o.download("SURROUNDING_KEY_IDS")
    assert np.array_equal(ids_g, ids_o), f"{where}: surroundingExistingKeyPosesID {ids_g} vs {ids_o}"
    for name in ("MAP_CORNER", "MAP_SURF"):
        a, b = gpu.download(name, k), o.download(name)
        assert a.shape == b.shape, f"{where}: {name} {a.shape} vs {b.shape}"
        if exact:
            assert same_bits(a, b), f"{where}: {name} differs, max {np.abs(a - b).max()}"
        else:
            assert np.abs(a - b).max() <= 2e-4, f"{where}: {name} max diff {np.abs(a - b).max()}"


def _compare_keyframes(gpu, o, k, where, first, exact):
    pg, po = gpu.download("KEY_POSES_6D", k), o.download("KEY_POSES_6D")
    assert pg.shape == po.shape, f"{where}: key poses {pg.shape} vs {po.shape}"
    if exact:
        assert same_bits(pg, po), f"{where}: cloudKeyPoses6D differ"
    for kf in range(first, len(po)):
        for which in range(3):
            a, b = gpu.download_keyframe(k, kf, which), o.download_keyframe(kf, which)
            assert a.shape == b.shape, f"{where}: key frame {kf} cloud {which}: {a.shape} vs {b.shape}"
            if exact:
                assert same_bits(a, b), f"{where}: key frame {kf} cloud {which} differs, max {np.abs(a - b).max()}"
            else:
                assert np.abs(a - b).max() <= 2e-4


def _live(cfgname, seqs, n_frames, min_cycles):
    from lego_loam_bor_b200.capi import LegoLoam
    from oracle.oracle_py import Oracle
    p, cfg, scans = make_scans(cfgname, seqs, range(n_frames))
    gpu = LegoLoam(p, batch=len(seqs))
    gpu.map_enable_keyframes(max_keyframes=64)
    oracles = [Oracle(p) for _ in seqs]
    cycles = 0
    saved = [0] * len(seqs)
    for f in range(n_frames):
        gpu.set_scans_host([scans[(s, f)] for s in seqs])
        rc = gpu.process_scans()
        for k, s in enumerate(seqs):
            o = oracles[k]
            o.image_projection(scans[(s, f)])
            r = o.feature_association()
            assert r == rc
            if r == 1:
                o.mapping_cycle()
        if rc != 1:
            continue
        cycles += 1
        for k in range(len(seqs)):
            o = oracles[k]
            where = f"frame {f} seq {seqs[k]}"
            poses_equal = True
            for name in ("TRANSFORM_SUM", "TRANSFORM_TOBE_MAPPED", "TRANSFORM_AFT_MAPPED", "TRANSFORM_BEF_MAPPED"):
                a, b = gpu.download(name, k), o.download(name)
                assert np.all(np.abs(a[:3] - b[:3]) <= POSE_TOL_RAD), f"{where} {name} rot {a} vs {b}"
                assert np.all(np.abs(a[3:] - b[3:]) <= POSE_TOL_M), f"{where} {name} trans {a} vs {b}"
                poses_equal = poses_equal and same_bits(a, b)
            ia, ib = gpu.download("MAP_ITERS", k), o.download("MAP_ITERS")
            assert np.array_equal(ia, ib), f"{where}: scan-to-map iterations/rows {ia} vs {ib}"
            _compare_maps(gpu, o, k, where, exact=poses_equal)
            _compare_keyframes(gpu, o, k, where, saved[k], exact=poses_equal)
            saved[k] = int(o.download("KEYFRAME_STATE")[0])
    assert cycles >= min_cycles
    for k in range(len(seqs)):
        st = oracles[k].download("KEYFRAME_STATE")
        assert st[0] >= min_cycles - 1 and st[1] >= 2, f"scene too weak: {st}"
        assert oracles[k].download("MAP_ITERS")[1] >= 50


def test_live_map_vlp16(built):
    """ll_process_scans with the local map built from the sequence's own key frames, 8 mapping cycles."""
    _live("A", [0, 3], 42, 8)


def test_live_map_64_beam(built):
    _live("C", [1], 22, 4)


def test_keyframe_erase_and_radius(built):
    """saveKeyFramesAndFactor + extractSurroundingKeyFrames driven with forced poses: three key poses per 1 m
    voxel make the averaged id move (erase + rebuild of the voxel sums), a 70 m jump leaves the 50 m radius."""
    from lego_loam_bor_b200.capi import LegoLoam
    from oracle.oracle_py import Oracle
    seqs = [0, 1]
    p, cfg, scans = make_scans("A", seqs, range(7))
    gpu = LegoLoam(p, batch=len(seqs))
    gpu.map_enable_keyframes(max_keyframes=64)
    oracles = [Oracle(p) for _ in seqs]
    for f in range(7):  # through the first mapping hand-over: the down-sampled scan clouds exist afterwards
        gpu.set_scans_host([scans[(s, f)] for s in seqs])
        gpu.process_scans()
        for k, s in enumerate(seqs):
            oracles[k].image_projection(scans[(s, f)])
            if oracles[k].feature_association() == 1:
                oracles[k].mapping_cycle()
    rebuilds = 0
    emptied = 0
    for step in range(1, 40):
        aft = np.zeros((len(seqs), 6), np.float32)
        for k in range(len(seqs)):
            along = 0.35 * step + (70.0 if 20 <= step < 26 else 0.0)
            aft[k] = [0.01 * k, 0.02 * step, -0.01, 0.3 * k + 0.05 * step, 0.1 * k, along]
        gpu.map_set_poses(aft, np.zeros_like(aft))
        gpu.map_save_keyframe()
        gpu.map_extract_surrounding_keyframes()
        for k in range(len(seqs)):
            o = oracles[k]
            o.map_set_poses(aft[k], np.zeros(6, np.float32))
            o.map_save_keyframe()
            o.map_extract_surrounding_keyframes()
            _compare_maps(gpu, o, k, f"step {step} seq {k}", exact=True)
            st = o.download("KEYFRAME_STATE")
            rebuilds += int(st[2])
            emptied += int(step == 20 and st[1] <= 2)
    assert rebuilds >= 4 and emptied == len(seqs), (rebuilds, emptied)
    for k in range(len(seqs)):
        _compare_keyframes(gpu, oracles[k], k, f"end seq {k}", 0, exact=True)


def test_full_size_batch_slot_invariance(built):
    """BASELINE configs[4] shape: 64 sequences of 64x2048 in one batch (4 sub-batches of 16 on 4 streams, like bench.py),
    whole pipeline with the live local map.  Size-independent properties: a sequence gives the same bits whatever
    slot / sub-batch it runs in (4 distinct sequences replicated over the 64 slots), and the four distinct ones agree
    with the oracle."""
    import torch
    from lego_loam_bor_b200.capi import LegoLoamStreams
    from oracle.oracle_py import Oracle
    B, U, n_frames = 64, 4, 11
    p, cfg, scans = make_scans("C", list(range(U)), range(n_frames))
    N = p.num_vertical_scans * p.num_horizontal_scans
    streams = [torch.cuda.Stream() for _ in range(4)]
    gpu = LegoLoamStreams(p, B, 4, max_points=N, device=0, streams=[s.cuda_stream for s in streams])
    gpu.map_enable_keyframes(max_keyframes=16)
    oracles = [Oracle(p) for _ in range(U)]
    stride = max(len(scans[(u, f)]) for u in range(U) for f in range(n_frames))
    host = torch.zeros((B, stride, 4), dtype=torch.float32).pin_memory()
    counts = np.zeros(B, np.int32)
    for f in range(n_frames):
        gpu.synchronize()
        for k in range(B):
            a = scans[(k % U, f)]
            host.numpy()[k, :len(a)] = a
            counts[k] = len(a)
        gpu.set_scans_host_ptr(host.data_ptr(), counts, stride)
        gpu.process_scans()
        for u in range(U):
            oracles[u].image_projection(scans[(u, f)])
            if oracles[u].feature_association() == 1:
                oracles[u].mapping_cycle()
    gpu.synchronize()
    names = ["TRANSFORM_SUM", "TRANSFORM_AFT_MAPPED", "CORNER_LAST", "SURF_LAST", "LABEL_MAT", "MAP_CORNER", "MAP_SURF",
             "SURROUNDING_KEY_IDS", "KEY_POSES_6D", "MAP_ITERS"]
    ref = {(u, n): gpu.download(n, u) for u in range(U) for n in names}
    for k in range(U, B):
        for n in names:
            assert same_bits(gpu.download(n, k), ref[(k % U, n)]), f"slot {k} differs from slot {k % U} in {n}"
    for u in range(U):
        o = oracles[u]
        assert int(gpu.download("KEYFRAME_STATE", u)[3]) == 0
        assert np.array_equal(gpu.download("LABEL_MAT", u), o.download("LABEL_MAT"))
        assert np.array_equal(gpu.download("SURROUNDING_KEY_IDS", u), o.download("SURROUNDING_KEY_IDS"))
        assert np.array_equal(gpu.download("MAP_ITERS", u), o.download("MAP_ITERS"))
        for name in ("TRANSFORM_SUM", "TRANSFORM_AFT_MAPPED"):
            a, b = gpu.download(name, u), o.download(name)
            assert np.all(np.abs(a[:3] - b[:3]) <= POSE_TOL_RAD) and np.all(np.abs(a[3:] - b[3:]) <= POSE_TOL_M), f"{name} {a} vs {b}"
        for name in ("MAP_CORNER", "MAP_SURF"):
            a, b = gpu.download(name, u), o.download(name)
            assert a.shape == b.shape and np.abs(a - b).max() <= 2e-4, f"{name}: {a.shape} vs {b.shape}"
