"""GPU parity at BASELINE.json configs[3]: the local map assembled from pre-stored key frames (SURVEY.md section 8d way 1,
mapOptmization.cpp:915-995), the mapping cycle against it (:1315-1332, :1335-1474), and index-level parity of both
search structures (nanoflann_pcl.h:131-152 for the map 5-NN; featureAssociation.h:87-93 for the scan-to-scan
correspondences) -- all through the C ABI, against the CPU oracle on identical scans.

Mapping poses are teacher-forced after every cycle (the oracle's transformAftMapped / BefMapped / TobeMapped are uploaded
before the key frame is saved): a pose that differs in the last bit, which the 1e-4 m / 1e-5 rad tolerance allows,
would otherwise move a stored point across a voxel boundary and end the bit-exact comparison of the maps.

The oracle runs with stable_sort=True: about one 64x2048 scan in a hundred has two EQUAL curvature values inside one
sextant at the edge of the 20-point selection (featureAssociation.cpp:285-304); which one std::sort ranks first is
libstdc++-internal, the device sorts by (value, position).  The other parity tests detect such scans and skip them; here
every key frame feeds the map, so the tie rule is fixed on the oracle's side instead."""
import numpy as np
import pytest

from lego_loam_bor_b200 import config_params, synth, workloads

pytestmark = pytest.mark.gpu

POSE_TOL_M = 1e-4
POSE_TOL_RAD = 1e-5


def _oracle_prebuild(o, cfg, seq, K):
    zero = np.zeros(6, np.float32)
    for i in range(K):
        o.reset_feature_association()
        o.image_projection(synth.arena_scan(cfg, seq, synth.KEYFRAME, i))
        o.feature_association()
        o.map_downsample_current_scan()
        T = synth.pose_to_transform(synth.arena_pose(cfg, seq, synth.KEYFRAME, i))
        o.map_set_poses(T, zero)
        o.map_set_initial_guess(T)
        o.map_save_keyframe()
    o.reset_feature_association()
    o.map_set_poses(synth.pose_to_transform(synth.arena_pose(cfg, seq, synth.DRIVE, 0)), zero)


def _gpu_prebuild(gpu, cfg, seqs, K, device_generator):
    import torch
    p_N = cfg.V * cfg.H
    if device_generator:
        dev = torch.device("cuda", 0)
        gen = synth.ArenaDeviceGenerator(cfg, seqs, dev)
        buf = torch.zeros((len(seqs), p_N, 4), dtype=torch.float32, device=dev)

        def scans_of(i):
            _, counts = gen.scans(synth.KEYFRAME, i, out=buf)
            return buf.data_ptr(), counts, p_N
    else:
        buf = torch.zeros((len(seqs), p_N, 4), dtype=torch.float32, device="cuda:0")

        def scans_of(i):
            host = np.zeros((len(seqs), p_N, 4), np.float32)
            counts = np.zeros(len(seqs), np.int32)
            for k, s in enumerate(seqs):
                a = synth.arena_scan(cfg, s, synth.KEYFRAME, i)
                host[k, :len(a)] = a
                counts[k] = len(a)
            buf.copy_(torch.from_numpy(host))
            return buf.data_ptr(), counts, p_N
    workloads.prebuild_keyframes(gpu, cfg, seqs, K, scans_of, sync=lambda: torch.cuda.synchronize())
    workloads.start_drive(gpu, cfg, seqs)
    return buf  # keep alive


def _pose_close(a, b, what):
    rot = np.abs((a[:3] - b[:3] + np.pi) % (2 * np.pi) - np.pi)
    assert np.all(rot <= POSE_TOL_RAD), f"{what}: rot {a} vs {b}"
    assert np.all(np.abs(a[3:] - b[3:]) <= POSE_TOL_M), f"{what}: trans {a} vs {b}"


def _run_kf(cfgname, seqs, K, frames, device_generator=False, oracle_seqs=None, libm=False, need_erase=False, min_cycles=1):
    from lego_loam_bor_b200.capi import LegoLoam
    from oracle.oracle_py import Oracle
    p = config_params(cfgname)
    cfg = synth.make_arena(p, n_keyframes=K)
    oracle_seqs = list(range(len(seqs))) if oracle_seqs is None else oracle_seqs
    gpu = LegoLoam(p, batch=len(seqs))
    gpu.map_enable_keyframes(*workloads.keyframe_capacities(p, K, extra_keyframes=len(frames) // 5 + 8))
    keep = _gpu_prebuild(gpu, cfg, seqs, K, device_generator)
    oracles = {}
    for k in oracle_seqs:
        oracles[k] = Oracle(p, libm=libm, stable_sort=True)
        _oracle_prebuild(oracles[k], cfg, seqs[k], K)
    for k, o in oracles.items():
        assert np.array_equal(gpu.download("KEY_POSES_6D", k), o.download("KEY_POSES_6D"))
        for which, kf in ((0, 0), (1, K // 2), (2, K - 1), (1, K - 1)):
            a, b = gpu.download_keyframe(k, kf, which), o.download_keyframe(kf, which)
            # the store keeps a key-frame cloud transformed by its key pose; the oracle transforms on demand
            assert a.shape == b.shape, f"key frame {kf} cloud {which}: {a.shape} vs {b.shape}"
    cycles, erases = 0, 0
    for f in frames:
        scans = [synth.arena_scan(cfg, s, synth.DRIVE, f) for s in seqs]
        gpu.set_scans_host(scans)
        gpu.image_projection()
        rc = gpu.feature_association()
        ro = {k: o.image_projection(scans[k]) or o.feature_association() for k, o in oracles.items()}
        assert all(r == rc for r in ro.values())
        for k, o in oracles.items():
            assert np.array_equal(gpu.download("ODOM_ITERS", k), o.download("ODOM_ITERS")), f"frame {f} seq {k}: LM iterations"
            _pose_close(gpu.download("TRANSFORM_SUM", k), o.download("TRANSFORM_SUM"), f"frame {f} seq {k} transformSum")
        if rc != 1:
            continue
        cycles += 1
        gpu.map_downsample_current_scan()
        gpu.map_predict_pose()
        gpu.map_extract_surrounding_keyframes()
        gpu.scan_to_map()
        for k, o in oracles.items():
            o.mapping_cycle()
            st_o = o.download("KEYFRAME_STATE")
            st_g = gpu.download("KEYFRAME_STATE", k)
            assert st_g[3] == 0, f"capacity error bits {st_g[3]}"
            assert np.array_equal(gpu.download("SURROUNDING_KEY_IDS", k), o.download("SURROUNDING_KEY_IDS")), f"frame {f} seq {k}"
            assert st_g[2] == st_o[2], f"frame {f} seq {k}: erase flag {st_g[2]} vs {st_o[2]}"
            erases += int(st_o[2])
            for name in ("MAP_CORNER", "MAP_SURF"):
                a, b = gpu.download(name, k), o.download(name)
                assert a.shape == b.shape, f"frame {f} seq {k}: {name} {a.shape} vs {b.shape}"
                if libm:
                    assert np.abs(a - b).max() < 1e-3
                else:
                    assert np.array_equal(a, b), f"frame {f} seq {k}: {name} differs, max {np.abs(a - b).max()}"
            assert np.array_equal(gpu.download("MAP_ITERS", k), o.download("MAP_ITERS")), f"frame {f} seq {k}: scan-to-map iterations / rows"
            assert o.download("MAP_ITERS")[0] >= 1
            for name in ("TRANSFORM_AFT_MAPPED", "TRANSFORM_BEF_MAPPED", "TRANSFORM_TOBE_MAPPED"):
                _pose_close(gpu.download(name, k), o.download(name), f"frame {f} seq {k} {name}")
                gpu.upload(name, o.download(name), k)   # teacher forcing (see the module docstring)
        gpu.map_save_keyframe()
        for k, o in oracles.items():
            assert gpu.download("KEYFRAME_STATE", k)[0] == o.download("KEYFRAME_STATE")[0]
            assert np.array_equal(gpu.download("KEY_POSES_6D", k), o.download("KEY_POSES_6D"))
    assert cycles >= min_cycles
    if need_erase:
        assert erases >= 1, "the run never erased a key frame: the erase path was not exercised"
    del keep
    return cycles, erases


def test_device_generator_is_bit_identical_to_host(built):
    import torch
    for cfgname, seqs in (("T", [0, 3]), ("C", [1])):
        p = config_params(cfgname)
        cfg = synth.make_arena(p, n_keyframes=50)
        gen = synth.ArenaDeviceGenerator(cfg, seqs, torch.device("cuda", 0))
        for kind, idx in ((synth.KEYFRAME, 0), (synth.KEYFRAME, 49), (synth.DRIVE, 13)):
            pts, counts = gen.scans(kind, idx)
            for k, s in enumerate(seqs):
                ref = synth.arena_scan(cfg, s, kind, idx)
                assert counts[k] == len(ref)
                assert np.array_equal(pts[k, :counts[k]].cpu().numpy(), ref), (cfgname, s, kind, idx)


def test_keyframe_map_tiny_sensor_long_drive_with_erases(built):
    """24 mapping cycles against a 40-key-frame map at the 16x450 sensor: the averaged-id rule of
    extractSurroundingKeyFrames (mapOptmization.cpp:940,962) erases key frames on the way, which exercises the
    unlink-and-re-sum path; maps stay bit-identical to the oracle's concatenate-and-VoxelGrid."""
    cycles, erases = _run_kf("T", [0, 1], 40, range(121), need_erase=True, min_cycles=24)
    assert cycles == 24


def test_keyframe_map_tiny_sensor_libm(built):
    """The same flow against the oracle on glibc's libm (what the reference calls): discrete outputs equal, maps and poses
    within tolerance."""
    _run_kf("T", [2], 30, range(41), libm=True, min_cycles=8)


def test_keyframe_map_64_beam_120(built):
    """64x2048, 120 key frames, two sequences in one batch, five mapping cycles."""
    _run_kf("C", [0, 6], 120, range(26), min_cycles=5)


def test_keyframe_map_64_beam_500(built):
    """BASELINE.json configs[3]: 64x2048 with the 500-key-frame local map (~124 k corner + ~146 k surf points from
    10.7 M stored points); key-frame scans from the device generator (bit-identical to the host's, tested above)."""
    _run_kf("C", [0, 9], 500, range(11), device_generator=True, oracle_seqs=[0], min_cycles=2)


def _jitter_frames(n, rng):
    """Drive frames with uneven gaps (1..4 frames = 0.1..0.4 m per step): the constant-velocity prior of transformCur
    (featureAssociation.cpp:907-909 keeps it across frames) is wrong at every step, so both LM stages run many iterations
    and the correspondences are searched again at iterations 5, 10, ..."""
    out, f = [0], 0
    while len(out) < n:
        f += int(rng.integers(1, 5))
        out.append(f)
    return out


def test_index_level_parity_of_both_search_structures(built):
    """LL_BUF_ODOM_SEARCH_IDX against the oracle's pointSearch*Ind arrays in every search round, LL_BUF_MAP_KNN_IDX against
    nanoflann's nearestKSearch in every scan-to-map iteration, on a jittered drive so that re-searches and the provable
    reuse of correspondences both happen."""
    from lego_loam_bor_b200.capi import LegoLoam
    from oracle.oracle_py import Oracle
    rng = np.random.default_rng(5)
    for cfgname, K, n in (("T", 30, 26), ("A", 30, 16)):
        p = config_params(cfgname)
        cfg = synth.make_arena(p, n_keyframes=K)
        frames = _jitter_frames(n, rng)
        gpu = LegoLoam(p, batch=1)
        gpu.enable_index_trace(True)
        gpu.map_enable_keyframes(*workloads.keyframe_capacities(p, K))
        keep = _gpu_prebuild(gpu, cfg, [0], K, False)
        o = Oracle(p, nanoflann=True, stable_sort=True)
        _oracle_prebuild(o, cfg, 0, K)
        rounds_seen, reuse_rounds, knn_iters = 0, 0, 0
        for f in frames:
            sc = synth.arena_scan(cfg, 0, synth.DRIVE, f)
            gpu.set_scans_host([sc])
            gpu.image_projection()
            rc = gpu.feature_association()
            o.image_projection(sc)
            assert o.feature_association() == rc
            it_g, it_o = gpu.download("ODOM_ITERS"), o.download("ODOM_ITERS")
            assert np.array_equal(it_g, it_o), f"frame {f}: LM iterations {it_g} vs {it_o}"
            if f != frames[0]:
                tg = gpu.download("ODOM_SEARCH_IDX").reshape(2, 5, 24 * p.num_vertical_scans, 3)
                to = o.download("ODOM_SEARCH_IDX").reshape(2, 5, 24 * p.num_vertical_scans, 3)
                nfeat = (len(o.download("SURF_FLAT")), len(o.download("CORNER_SHARP")))
                for stage in (0, 1):
                    n_rounds = (int(it_o[stage]) + 4) // 5
                    for r in range(n_rounds):
                        a, b = tg[stage, r, :nfeat[stage]], to[stage, r, :nfeat[stage]]
                        assert np.array_equal(a, b), f"frame {f} stage {stage} round {r}: {np.flatnonzero((a != b).any(1))[:5]}"
                        rounds_seen += 1
                        reuse_rounds += 1 if r > 0 else 0
            _pose_close(gpu.download("TRANSFORM_SUM"), o.download("TRANSFORM_SUM"), f"frame {f}")
            if rc == 1:
                gpu.map_downsample_current_scan(); gpu.map_predict_pose(); gpu.map_extract_surrounding_keyframes(); gpu.scan_to_map()
                o.mapping_cycle()
                mi = o.download("MAP_ITERS")
                assert np.array_equal(gpu.download("MAP_ITERS"), mi)
                kg, ko = gpu.download("MAP_KNN_IDX"), o.download("MAP_KNN_IDX")
                assert kg.shape == ko.shape and kg.shape[0] % 10 == 0
                Q = kg.shape[0] // 10
                for it in range(int(mi[0])):
                    a, b = kg[it * Q:(it + 1) * Q], ko[it * Q:(it + 1) * Q]
                    assert np.array_equal(a, b), f"frame {f} scan-to-map iteration {it}: queries {np.flatnonzero((a != b).any(1))[:5]} differ"
                    assert (b[:, 0] >= 0).sum() > 50
                    knn_iters += 1
                for name in ("TRANSFORM_AFT_MAPPED", "TRANSFORM_BEF_MAPPED", "TRANSFORM_TOBE_MAPPED"):
                    _pose_close(gpu.download(name), o.download(name), f"frame {f} {name}")
                    gpu.upload(name, o.download(name))
                gpu.map_save_keyframe()
        assert rounds_seen >= 2 * (len(frames) - 1) and reuse_rounds >= 3 and knn_iters >= 6, (rounds_seen, reuse_rounds, knn_iters)
        del keep


def test_knn_ties_on_a_lattice(built):
    """nanoflann keeps the first VISITED point among equal distances (nanoflann.hpp:175-202), the device the lowest
    index: on a map whose points sit on an exact lattice, with queries on lattice symmetry planes, the k-th neighbour is
    not unique.  What must hold: identical neighbour DISTANCES for every query, and identical index sets wherever the
    fifth and sixth distances differ."""
    from lego_loam_bor_b200.capi import LegoLoam
    from oracle import oracle_py
    p = config_params("T")
    g = np.arange(-4, 5, dtype=np.float32) * 0.5
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    lattice = np.stack([X.ravel(), Y.ravel(), Z.ravel(), np.zeros(X.size, np.float32)], 1).astype(np.float32)
    rng = np.random.default_rng(11)
    lattice = lattice[rng.permutation(len(lattice))]
    queries = []
    for _ in range(300):
        c = rng.integers(-2, 3, 3).astype(np.float32) * 0.5
        kind = rng.integers(0, 3)
        off = np.zeros(3, np.float32) if kind == 0 else (np.array([0.25, 0, 0], np.float32) if kind == 1 else np.array([0.25, 0.25, 0.25], np.float32))
        queries.append(np.concatenate([c + off, [0]]))
    queries = np.array(queries, np.float32)
    gpu = LegoLoam(p, batch=1)
    gpu.enable_index_trace(True)
    gpu.map_set_local(0, lattice, lattice)
    gpu.map_set_scan(0, queries[:150], queries[150:])
    gpu.map_set_initial_guess(np.zeros((1, 6), np.float32))
    gpu.scan_to_map()
    idx = gpu.download("MAP_KNN_IDX")[:300]
    ref_idx, ref_d2 = oracle_py.knn(lattice, queries, 6, nanoflann=True)
    pts = lattice[:, :3]
    n_tied = 0
    for q in range(300):
        assert idx[q, 0] >= 0, "every query has five lattice points within 1 m"
        d_gpu = np.sort(((pts[idx[q]] - queries[q, :3]) ** 2).sum(1).astype(np.float32))
        assert np.array_equal(d_gpu, np.sort(ref_d2[q, :5])), f"query {q}: neighbour distances differ"
        if ref_d2[q, 4] != ref_d2[q, 5]:
            assert set(idx[q]) == set(ref_idx[q, :5]), f"query {q}: unique 5-NN set differs"
        else:
            n_tied += 1
    assert n_tied > 100
