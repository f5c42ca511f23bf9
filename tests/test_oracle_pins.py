"""Pins of the CPU oracle (test infrastructure) -- all CPU, no GPU.

The reference has no tests, golden files or fixtures, so the oracle is tied to
the reference by the hand-derivable known answers listed there (each derived from the cited reference
lines), by the reference's own vendored nanoflann for k-NN, and by agreement between its two math
backends (glibc libm as the reference uses / the portable one the CUDA kernels share)."""
import json
import os

import numpy as np
import pytest

from lego_loam_bor_b200 import config_params, synth
from oracle import oracle_py
from oracle.oracle_py import Oracle

HERE = os.path.dirname(os.path.abspath(__file__))
DEG = np.pi / 180.0


def beam(p, row, col, rng, dcol=0.0):
    """A point that an ideal sensor returns on ring `row`, column `col` at range `rng`
    (inverse of imageProjection.cpp:190-200)."""
    V, H = p.num_vertical_scans, p.num_horizontal_scans
    elev = (p.vertical_angle_bottom + row * (p.vertical_angle_top - p.vertical_angle_bottom) / (V - 1)) * DEG
    ha = np.pi / 2 - (col + dcol - H / 2) * (2 * np.pi / H)
    return [rng * np.cos(elev) * np.sin(ha), rng * np.cos(elev) * np.cos(ha), rng * np.sin(elev), 0.0]


@pytest.fixture(scope="module")
def ora(built):
    return lambda p, **kw: Oracle(p, **kw)


def test_kat1_projection_indices_and_intensity(ora):
    """imageProjection.cpp:190-216: ring r, column c -> cell (r, c), intensity r + c/10000; +x axis -> column H/2."""
    p = config_params("A")
    o = ora(p)
    V, H = 16, 1800
    cells = [(0, 0), (3, 17), (15, 1799), (7, 900), (8, 1), (12, 1234)]
    pts = np.array([beam(p, r, c, 10.0 + r) for r, c in cells] + [[25.0, 0.0, 0.5, 0.0]], np.float32)
    o.image_projection(pts)
    rng = o.download("RANGE_MAT").reshape(V, H)
    full = o.download("FULL_CLOUD").reshape(V, H, 4)
    for r, c in cells:
        assert abs(rng[r, c] - (10.0 + r)) < 1e-4
        assert full[r, c, 3] == np.float32(np.float64(np.float32(r)) + np.float64(np.float32(c)) / 10000.0)
    filled = np.argwhere(rng != np.finfo(np.float32).max)
    assert len(filled) == len(cells) + 1
    # the point on the +x axis lands in column H/2 (imageProjection.cpp:198-200)
    assert rng[8, 900] != np.finfo(np.float32).max and abs(rng[8, 900] - np.hypot(25.0, 0.5)) < 1e-3
    assert np.isnan(full[0, 1, 0]) and full[0, 1, 3] == 0  # untouched cells: NaN xyz (imageProjection.cpp:109-132)


def test_kat2_flat_ground(ora):
    """imageProjection.cpp:259-300: a flat ground plane seen by rings 0..gsi marks exactly those rings."""
    p = config_params("A")
    o = ora(p)
    V, H, gsi = 16, 1800, p.ground_scan_index
    pts = []
    for c in range(H - 1, -1, -1):
        for r in range(V):
            elev = (-15 + 2 * r) * DEG
            if elev < 0:
                pts.append(beam(p, r, c, 1.5 / np.sin(-elev)))
            else:
                pts.append(beam(p, r, c, 30.0))  # a far cylinder wall above the horizon
    o.image_projection(np.array(pts, np.float32))
    g = o.download("GROUND_MAT").reshape(V, H)
    lab = o.download("LABEL_MAT").reshape(V, H)
    assert np.all(g[: gsi + 1] == 1) and np.all(g[gsi + 1:] == 0)
    assert np.all(lab[: gsi + 1] == -1)
    assert np.all(lab[gsi + 1:] == 1)  # the wall is one valid segment -> label 1 (imageProjection.cpp:489-490)
    # ground decimation of cloudSegmentation (imageProjection.cpp:376-378): kept iff j%5==0 or j<=5 or j>=H-5
    col = o.download("SEG_COL_IND")
    flag = o.download("SEG_GROUND_FLAG")
    gcols = col[flag == 1]
    assert np.all((gcols % 5 == 0) | (gcols <= 5) | (gcols >= H - 5))
    start, end = o.download("START_RING_INDEX"), o.download("END_RING_INDEX")
    per_ring_ground = len(set(list(range(0, H, 5)) + list(range(0, 6)) + list(range(H - 5, H))))
    assert start[0] == 4 and end[0] == per_ring_ground - 6 and start[1] == per_ring_ground + 4


def _blob_scan(p, blobs, rng=10.0):
    pts = []
    for cells in blobs:
        for r, c in cells:
            pts.append(beam(p, r, c, rng))
    return np.array(pts, np.float32)


def test_kat3_segment_validity_rules(ora):
    """imageProjection.cpp:469,476-495: >= 30 cells valid; >= 5 cells valid only if the NON-SEED cells touch
    >= 3 rows; invalid segments get 999999; valid ones are numbered in row-major seed order."""
    p = config_params("A")
    o = ora(p)
    V, H = 16, 1800
    big = [(9 + dr, 100 + dc) for dr in range(3) for dc in range(10)]            # 30 cells
    four = [(9, 300 + dc) for dc in range(4)]                                      # 4 cells
    five_row = [(9, 500 + dc) for dc in range(5)]                                  # 5 cells, 1 row
    five_col = [(9 + dr, 700) for dr in range(5)]                                  # 5 cells, 4 non-seed rows
    five_two = [(9, 900), (9, 901), (9, 902), (10, 900), (10, 901)]                # 5 cells, 2 rows
    wrap = [(12, H - 2), (12, H - 1), (12, 0), (12, 1), (13, 0), (14, 0)]          # columns wrap (imageProjection.cpp:446-451)
    o.image_projection(_blob_scan(p, [big, four, five_row, five_col, five_two, wrap]))
    lab = o.download("LABEL_MAT").reshape(V, H)
    assert {lab[r, c] for r, c in big} == {1}
    assert {lab[r, c] for r, c in four} == {999999}
    assert {lab[r, c] for r, c in five_row} == {999999}
    assert {lab[r, c] for r, c in five_col} == {2}          # second valid seed in row-major order
    assert {lab[r, c] for r, c in five_two} == {999999}
    assert {lab[r, c] for r, c in wrap} == {3}              # 6 cells over rows 12,13,14 joined across the column wrap
    # outliers: invalid cells above the ground rows at columns divisible by 5 (imageProjection.cpp:366-370)
    out = o.download("OUTLIER_CLOUD")
    assert len(out) == sum(1 for r, c in four + five_row + five_two if c % 5 == 0)


def test_kat4_constant_range_gives_zero_curvature(ora):
    """featureAssociation.cpp:203-215."""
    p = config_params("T")
    o = ora(p)
    V, H = p.num_vertical_scans, p.num_horizontal_scans
    pts = np.array([beam(p, r, c, 12.0) for c in range(H - 1, -1, -1) for r in range(10, 14)], np.float32)
    o.image_projection(pts)
    o.feature_association()
    S = len(o.download("SEG_CLOUD"))
    curv = o.download("CLOUD_CURVATURE")[:S]
    rng = o.download("SEG_RANGE")
    d = np.abs(rng - 12.0).max()
    assert S == 4 * H and d < 1e-5
    assert np.all(curv[5:S - 5] <= (11 * 2 * d) ** 2 + 1e-12)


def test_kat5_identity_motion(ora):
    """featureAssociation.cpp:917,1028: the same scan twice -> both LM stages stop after their first iteration
    with a (near) zero step."""
    p = config_params("T")
    o = ora(p)
    cfg = synth.make_config(p)
    scan = synth.scan(cfg, 0, 0)
    o.image_projection(scan); o.feature_association()
    o.image_projection(scan); o.feature_association()
    assert list(o.download("ODOM_ITERS")) == [1, 1]
    cur = o.download("TRANSFORM_CUR")
    assert np.all(np.abs(cur[:3]) < 2e-4) and np.all(np.abs(cur[3:]) < 2e-3)


def test_kat6_knn_backends_agree_with_brute_force(built):
    """nanoflann.hpp:1221-1241 (eps = 0): exact search.  Reference nanoflann == port kd-tree == brute force."""
    rng = np.random.default_rng(1)
    cloud = np.zeros((5000, 4), np.float32); cloud[:, :3] = rng.uniform(-20, 20, (5000, 3))
    query = np.zeros((300, 4), np.float32); query[:, :3] = rng.uniform(-22, 22, (300, 3))
    d = ((query[:, None, :3].astype(np.float32) - cloud[None, :, :3]) ** 2)
    d2 = (d[..., 0] + d[..., 1]) + d[..., 2]
    brute = np.argsort(d2, axis=1, kind="stable")[:, :5]
    idx_port, dist_port = oracle_py.knn(cloud, query, 5, nanoflann=False)
    assert np.array_equal(idx_port, brute)
    assert np.array_equal(dist_port, np.take_along_axis(d2, brute, 1))
    if oracle_py.load().lo_has_nanoflann():
        idx_nf, dist_nf = oracle_py.knn(cloud, query, 5, nanoflann=True)
        assert np.array_equal(idx_nf, brute) and np.array_equal(dist_nf, dist_port)


def test_voxel_grid_known_answer(built):
    """pcl::VoxelGrid restatement: ascending voxel index (x fastest), centroid of all fields."""
    pts = np.array([[0.05, 0.05, 0.05, 1.0], [0.15, 0.10, 0.02, 3.0],     # voxel (0,0,0)
                    [0.25, 0.05, 0.05, 5.0],                                 # voxel (1,0,0)
                    [0.05, 0.25, 0.05, 7.0],                                 # voxel (0,1,0)
                    [0.05, 0.05, 0.45, 9.0], [0.06, 0.07, 0.41, 11.0]], np.float32)  # voxel (0,0,2)
    out = oracle_py.voxel_grid(pts[::-1].copy(), 0.2)
    exp = np.array([[0.10, 0.075, 0.035, 2.0], [0.25, 0.05, 0.05, 5.0], [0.05, 0.25, 0.05, 7.0], [0.055, 0.06, 0.43, 10.0]])
    assert out.shape == (4, 4) and np.allclose(out, exp, atol=1e-6)
    assert len(oracle_py.voxel_grid(np.zeros((0, 4), np.float32), 0.2)) == 0


def test_math_backends_agree_on_discrete_outputs(built):
    """The portable math the kernels share must not change any discrete decision relative to glibc's libm
    (what the reference calls) on the fixture sequences; continuous outputs agree to float rounding."""
    for cfgname, frames in (("T", 6), ("A", 3)):
        p = config_params(cfgname)
        cfg = synth.make_config(p)
        a, b = Oracle(p, libm=True), Oracle(p, libm=False)
        for f in range(frames):
            scan = synth.scan(cfg, 1, f)
            for o in (a, b):
                o.image_projection(scan)
                o.feature_association()
            for name in ("GROUND_MAT", "LABEL_MAT", "SEG_COL_IND", "START_RING_INDEX", "END_RING_INDEX",
                         "CORNER_SHARP_IND", "CORNER_LESS_SHARP_IND", "SURF_FLAT_IND", "NEIGHBOR_PICKED", "ODOM_ITERS"):
                assert np.array_equal(a.download(name), b.download(name)), f"{cfgname} frame {f}: {name}"
            assert np.allclose(a.download("RANGE_MAT"), b.download("RANGE_MAT"), rtol=0, atol=0)
            assert np.allclose(a.download("TRANSFORM_SUM"), b.download("TRANSFORM_SUM"), rtol=0, atol=2e-6)


def test_golden_digests(built):
    """Regression pin: the committed digests of tests/golden/golden_T.json (made by make_golden.py)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    with open(os.path.join(HERE, "golden", "golden_T.json")) as f:
        gold = json.load(f)
    now = mg.compute(len(gold["frames"]))
    for f, (g, n) in enumerate(zip(gold["frames"], now["frames"])):
        assert g["input"] == n["input"], f"frame {f}: synthetic input changed"
        assert g["counts"] == n["counts"] and g["odom_iters"] == n["odom_iters"]
        for name, dig in g["buffers"].items():
            assert n["buffers"][name] == dig, f"frame {f}: {name} changed"
        assert np.allclose(g["transform_sum"], n["transform_sum"], rtol=0, atol=1e-7)


def test_first_frame_protocol_and_mapping_cadence(built):
    """featureAssociation.cpp:1414-1417,1429-1433: the first frame only initialises; every
    mapping_frequency_divider-th later frame is handed to MapOptimization."""
    p = config_params("T")
    cfg = synth.make_config(p)
    o = Oracle(p)
    handed = []
    for f in range(12):
        o.image_projection(synth.scan(cfg, 0, f))
        handed.append(o.feature_association())
        if f == 0:
            assert np.all(o.download("TRANSFORM_SUM") == 0) and len(o.download("CORNER_LAST")) > 0
    assert handed == [0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0]


def test_float_accumulation_of_the_normal_equations_stays_inside_the_tolerance(built):
    """The reference forms J^T J / J^T r with Eigen's float GEMM (featureAssociation.cpp:860-866, mapOptmization.cpp:1258),
    whose summation order is not reproducible; the oracle (and the kernels) accumulate exact double products.  With float
    accumulation instead -- row by row, or as eight interleaved partial sums like a SIMD kernel -- no LM iteration count
    changes and every pose stays inside the parity tolerance (1e-4 m / 1e-5 rad), over 40 frames with the live map."""
    from lego_loam_bor_b200 import config_params, synth
    from oracle.oracle_py import Oracle
    for cfgname, n_frames in (("T", 41), ("A", 16)):
        p = config_params(cfgname)
        cfg = synth.make_config(p)
        scans = synth.scans(cfg, [0], range(n_frames))
        runs = [Oracle(p, libm=True, float_accum=m) for m in (0, 1, 2)]
        worst_rot, worst_t, cycles = 0.0, 0.0, 0
        for f in range(n_frames):
            outs = []
            for o in runs:
                o.image_projection(scans[(0, f)])
                if o.feature_association() == 1:
                    o.mapping_cycle()
                outs.append((o.download("ODOM_ITERS"), o.download("MAP_ITERS"), o.download("TRANSFORM_SUM"), o.download("TRANSFORM_AFT_MAPPED")))
            cycles += 1 if f and f % 5 == 0 else 0
            for it, mi, ts, ta in outs[1:]:
                assert np.array_equal(it, outs[0][0]) and np.array_equal(mi, outs[0][1]), f"{cfgname} frame {f}: an LM exit flipped"
                for a, b in ((ts, outs[0][2]), (ta, outs[0][3])):
                    worst_rot = max(worst_rot, float(np.abs(a[:3] - b[:3]).max()))
                    worst_t = max(worst_t, float(np.abs(a[3:] - b[3:]).max()))
        assert worst_rot <= 1e-5 and worst_t <= 1e-4, (cfgname, worst_rot, worst_t)
        assert cycles >= 3


def test_three_stage_thread_pipeline_equals_inline_run(built):
    """lo_run_pipeline (the reference's threading: main.cpp:37-47, three stage threads + blocking one-slot channels, the
    payloads of utility.h:64-80 copied between three stage objects) gives the poses and key frames of the inline run."""
    from lego_loam_bor_b200 import config_params, synth
    from oracle.oracle_py import Oracle, run_pipeline
    p = config_params("T")
    cfg = synth.make_config(p)
    scans = [synth.scan(cfg, 3, f) for f in range(27)]
    ref = Oracle(p, libm=True)
    for a in scans:
        ref.image_projection(a)
        if ref.feature_association() == 1:
            ref.mapping_cycle()
    o_ip, o_fa, o_mo = Oracle(p, libm=True), Oracle(p, libm=True), Oracle(p, libm=True)
    ms, wall = run_pipeline(o_ip, o_fa, o_mo, scans, 1)
    assert ms.shape == (27, 3) and wall > 0 and (ms[:, 2] > 0).sum() == 5 and np.all(ms[:, :2] > 0)
    assert np.array_equal(ref.download("TRANSFORM_SUM"), o_fa.download("TRANSFORM_SUM"))
    for name in ("TRANSFORM_AFT_MAPPED", "KEYFRAME_STATE", "KEY_POSES_6D", "MAP_SURF"):
        assert np.array_equal(ref.download(name), o_mo.download(name)), name
