"""sensor_msgs/PointCloud2 ingest: pcl::fromROSMsg + pcl::removeNaNFromPointCloud of
imageProjection.cpp:159-161.  CPU: the oracle's restatement against numpy on a Velodyne-style record layout
(x, y, z, intensity float32, ring uint16, time float32: point_step 22, unaligned fields).  GPU: the device decode
through the C ABI against the oracle, then the projection of the decoded scan."""
import numpy as np
import pytest

from oracle import oracle_py

VELODYNE = np.dtype({"names": ["x", "y", "z", "intensity", "ring", "time"],
                     "formats": ["<f4", "<f4", "<f4", "<f4", "<u2", "<f4"], "offsets": [0, 4, 8, 12, 16, 18], "itemsize": 22})


def make_message(xyzi, rng, bad_fraction=0.05):
    """xyzi [n, 4] -> (raw uint8 bytes, expected xyzi after removeNaNFromPointCloud)"""
    n = len(xyzi)
    rec = np.zeros(n, VELODYNE)
    rec["x"], rec["y"], rec["z"], rec["intensity"] = xyzi[:, 0], xyzi[:, 1], xyzi[:, 2], xyzi[:, 3]
    rec["ring"] = rng.integers(0, 16, n)
    rec["time"] = rng.uniform(0, 0.1, n)
    bad = rng.random(n) < bad_fraction
    which = rng.integers(0, 4, n)
    vals = np.array([np.nan, np.inf, -np.inf], np.float32)[rng.integers(0, 3, n)]
    for k, name in enumerate(["x", "y", "z", "intensity"]):
        m = bad & (which == k)
        rec[name][m] = vals[m]
    keep = np.isfinite(rec["x"]) & np.isfinite(rec["y"]) & np.isfinite(rec["z"])  # a NaN intensity does not drop a point
    exp = np.stack([rec["x"], rec["y"], rec["z"], rec["intensity"]], 1)[keep].astype(np.float32)
    return np.frombuffer(rec.tobytes(), np.uint8).copy(), exp


def test_oracle_decode_matches_numpy():
    rng = np.random.default_rng(3)
    xyzi = rng.normal(0, 10, (5000, 4)).astype(np.float32)
    raw, exp = make_message(xyzi, rng)
    got = oracle_py.decode_pointcloud2(raw, len(xyzi), 22, 0, 4, 8, 12, is_dense=False)
    assert got.shape == exp.shape and np.array_equal(got, exp, equal_nan=True)
    # dense clouds are copied unchecked (removeNaNFromPointCloud's is_dense branch); a missing intensity field gives 0
    dense = oracle_py.decode_pointcloud2(raw, len(xyzi), 22, 0, 4, 8, -1, is_dense=True)
    assert len(dense) == len(xyzi) and np.all(dense[:, 3] == 0)
    assert len(oracle_py.decode_pointcloud2(raw[:0], 0, 22, 0, 4, 8, 12, is_dense=False)) == 0


@pytest.mark.gpu
def test_device_decode_and_projection(built):
    from lego_loam_bor_b200.capi import LegoLoam
    from oracle.oracle_py import Oracle
    from parity_utils import make_scans, same_bits
    rng = np.random.default_rng(11)
    p, cfg, scans = make_scans("A", [0, 1], [0])
    full = scans[(0, 0)]
    cases = [full, scans[(1, 0)][:9001], full[:0], full[:1]]  # ragged batch incl. an empty and a one-point message
    msgs, exps = zip(*[make_message(c, rng) for c in cases])
    nan_only = np.frombuffer(np.full(3, np.nan, np.float32).tobytes() + b"\0" * 10, np.uint8)  # one record, x = y = z = NaN
    msgs, exps = list(msgs) + [np.tile(nan_only, 7)], list(exps) + [np.zeros((0, 4), np.float32)]
    gpu = LegoLoam(p, batch=len(msgs))
    for is_dense in (False, True):
        gpu.set_scans_pointcloud2(msgs, 22, 0, 4, 8, 12, is_dense=is_dense)
        for k, m in enumerate(msgs):
            ref = oracle_py.decode_pointcloud2(m, len(m) // 22, 22, 0, 4, 8, 12, is_dense=is_dense)
            got = gpu.download("INPUT_CLOUD", k)
            assert got.shape == ref.shape and same_bits(got, ref), f"case {k} dense={is_dense}: {got.shape} vs {ref.shape}"
            if not is_dense:
                assert same_bits(ref, exps[k])
        if is_dense:
            continue  # NaN points in a "dense" cloud are undefined behaviour downstream in the reference too
        gpu.image_projection()
        for k in range(len(msgs)):
            o = Oracle(p)
            o.image_projection(exps[k])
            for name in ("RANGE_MAT", "LABEL_MAT", "SEG_COL_IND"):
                assert same_bits(gpu.download(name, k), o.download(name)), f"case {k}: {name}"
    # without the intensity field
    gpu.set_scans_pointcloud2(msgs, 22, 0, 4, 8, -1)
    assert np.all(gpu.download("INPUT_CLOUD", 0)[:, 3] == 0)


@pytest.mark.gpu
def test_packed_xyz_scans_match_xyzi(built):
    """ll_set_scans_xyz_host (12-byte points) against the oracle fed with the same scans as x, y, z, intensity: the
    reported intensity is overwritten by projectPointCloud (imageProjection.cpp:216) before anything reads it, so every
    output of the path is identical.  Ragged batch incl. an empty and a one-point scan, two frames (the input is double
    buffered), then the whole path through ll_process_scans."""
    from lego_loam_bor_b200.capi import LegoLoam
    from oracle.oracle_py import Oracle
    from parity_utils import make_scans, same_bits
    p, cfg, scans = make_scans("A", [0, 1], [0, 1, 2])
    seqs = [lambda f: scans[(0, f)], lambda f: scans[(1, f)][:9001], lambda f: scans[(0, f)][:0], lambda f: scans[(1, f)][:1]]
    gpu = LegoLoam(p, batch=len(seqs))
    oracles = [Oracle(p) for _ in seqs]
    for f in range(3):
        batch = [np.ascontiguousarray(s(f)) for s in seqs]
        gpu.set_scans_xyz_host([b[:, :3] for b in batch])
        for k, b in enumerate(batch):
            got = gpu.download("INPUT_CLOUD", k)
            assert got.shape == (len(b), 3) and same_bits(got, np.ascontiguousarray(b[:, :3]))
        gpu.process_scans()
        for k, b in enumerate(batch):
            o = oracles[k]
            noisy = b.copy()
            noisy[:, 3] = 255.0 * ((np.arange(len(b)) * 37) % 101) / 100.0   # whatever the sensor reported
            o.image_projection(noisy)
            o.feature_association()
            for name in ("RANGE_MAT", "FULL_CLOUD", "GROUND_MAT", "LABEL_MAT", "SEG_CLOUD", "SEG_COL_IND", "SURF_LAST", "CORNER_LAST"):
                assert same_bits(gpu.download(name, k), o.download(name)), f"frame {f} seq {k}: {name}"
            assert same_bits(gpu.download("TRANSFORM_SUM", k), o.download("TRANSFORM_SUM")), f"frame {f} seq {k}: pose"
    # switching back to 16-byte points on the same handle
    gpu.set_scans_host([np.ascontiguousarray(s(2)) for s in seqs])
    assert gpu.download("INPUT_CLOUD", 0).shape == (len(seqs[0](2)), 4)


def test_oracle_never_reads_the_input_intensity(built):
    """What ll_set_scans_xyz_host relies on, at the level of the reference's algorithm: projectPointCloud overwrites the
    intensity of every point it keeps (imageProjection.cpp:216) and nothing reads it before, so two scans that differ only
    in the sensor's intensity give identical outputs through projection, segmentation, features and odometry."""
    from oracle.oracle_py import Oracle
    from parity_utils import make_scans, same_bits
    p, cfg, scans = make_scans("T", [0], range(4))
    a, b = Oracle(p), Oracle(p)
    rng = np.random.default_rng(1)
    for f in range(4):
        s0 = np.ascontiguousarray(scans[(0, f)])
        s1 = s0.copy()
        s1[:, 3] = rng.uniform(0, 255, len(s1)).astype(np.float32)
        a.image_projection(s0)
        b.image_projection(s1)
        assert a.feature_association() == b.feature_association()
        for name in ("RANGE_MAT", "FULL_CLOUD", "LABEL_MAT", "SEG_CLOUD", "CORNER_SHARP", "SURF_FLAT", "CORNER_LAST", "SURF_LAST", "OUTLIER_CLOUD", "TRANSFORM_SUM"):
            assert same_bits(a.download(name), b.download(name)), f"frame {f}: {name}"
