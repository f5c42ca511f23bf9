"""CPU pins of the oracle's key-frame / local-map restatement (mapOptmization.cpp:856-996, 1335-1474; SURVEY.md
section 8 f2).  The reference has no fixtures for it, so the oracle is checked against (a) hand-derivable cases and
(b) an independent, loop-for-loop Python restatement of the same reference lines, and its two k-NN backends (the
reference's own nanoflann radiusSearch and the brute-force port) against each other."""
import numpy as np
import pytest

from lego_loam_bor_b200 import config_params
from oracle import oracle_py
from oracle.oracle_py import Oracle

F = np.float32


def _cloud(rng, n, spread):
    c = np.zeros((n, 4), F)
    c[:, :3] = rng.uniform(-spread, spread, (n, 3)).astype(F)
    c[:, 3] = rng.uniform(0, 16, n).astype(F)
    return c


class PyMapOptimization:
    """Python restatement, translation-only poses (cos = 1, sin = 0, so transformPointCloud is p + t in float32)."""

    def __init__(self, radius=50.0):
        self.r2 = F(radius * radius)
        self.poses = []  # (x, y, z) float32
        self.clouds = []  # (corner, surf, outlier)
        self.ids = []
        self.cur = np.zeros(3, F)
        self.prev = np.zeros(3, F)

    def save(self, aft, tobe, corner, surf, outlier):
        self.cur = np.array(aft[3:6], F)
        d = self.prev - self.cur
        dist = np.sqrt(F(F(d[0] * d[0]) + F(d[1] * d[1])) + F(d[2] * d[2]))
        if float(dist) < 0.3 and self.poses:
            return False
        self.prev = self.cur.copy()
        est = tobe if not self.poses else aft
        self.poses.append(np.array(est[3:6], F))
        self.clouds.append((corner, surf, outlier))
        return True

    def extract(self):
        if not self.poses:
            return np.zeros((0, 4), F), np.zeros((0, 4), F)
        P = np.array(self.poses, F)
        sel = []
        for i, q in enumerate(P):
            d2 = F(0)
            for a in range(3):
                df = F(self.cur[a] - q[a])
                d2 = F(d2 + F(df * df))
            if d2 < self.r2:
                sel.append((d2, i))
        sel.sort(key=lambda t: t[0])  # stable
        groups = {}
        for d2, i in sel:
            v = tuple(int(np.floor(F(P[i][a] * F(1.0)))) for a in range(3))
            groups.setdefault((v[2], v[1], v[0]), []).append(i)
        ds_ids = []
        for key in sorted(groups):
            acc = F(0)
            for i in groups[key]:
                acc = F(acc + F(i))
            ds_ids.append(int(F(acc / F(len(groups[key])))))
        erased = [i for i in self.ids if i not in ds_ids]
        self.ids = [i for i in self.ids if i in ds_ids]
        for i in ds_ids:
            if i not in self.ids:
                self.ids.append(i)
        corner, surf = [np.zeros((0, 4), F)], [np.zeros((0, 4), F)]
        for i in self.ids:
            t = np.array([*self.poses[i], 0], F)
            c, s, o = self.clouds[i]
            corner.append((c + t).astype(F))
            surf.append((s + t).astype(F))
            surf.append((o + t).astype(F))
        self.erased = bool(erased)
        return (oracle_py.voxel_grid(np.concatenate(corner), 0.2), oracle_py.voxel_grid(np.concatenate(surf), 0.4))


def _drive(o, py, rng, steps, nanoflann_note=""):
    rebuilds = 0
    for step in range(steps):
        # a path with ~0.35 m spacing (three key poses per 1 m voxel), one 70 m excursion, small y/z wobble
        along = 0.35 * step + (70.0 if 20 <= step < 26 else 0.0)
        aft = np.array([0, 0, 0, 0.05 * step, 0.02 * (step % 5), along], F)
        tobe = np.array([0, 0, 0, 0.01, 0.0, 0.02], F)
        corner, surf_total = _cloud(rng, 300, 8.0), _cloud(rng, 900, 8.0)
        o.map_set_scan(corner, surf_total)  # laserCloudCornerLastDS (+ an unused total cloud)
        o.map_set_poses(aft, np.zeros(6, F))
        o.upload("TRANSFORM_TOBE_MAPPED", tobe)
        o.map_save_keyframe()
        o.map_extract_surrounding_keyframes()
        py.save(aft, tobe, corner, np.zeros((0, 4), F), np.zeros((0, 4), F))
        pc, ps = py.extract()
        st = o.download("KEYFRAME_STATE")
        ids = o.download("SURROUNDING_KEY_IDS")
        assert st[0] == len(py.poses), f"step {step}{nanoflann_note}: {st} vs {len(py.poses)} key frames"
        assert list(ids) == py.ids, f"step {step}{nanoflann_note}: ids {list(ids)} vs {py.ids}"
        assert bool(st[2]) == py.erased
        rebuilds += int(st[2])
        mc = o.download("MAP_CORNER")
        assert mc.shape == pc.shape and np.array_equal(mc, pc), f"step {step}: corner map differs"
        assert len(o.download("MAP_SURF")) == len(ps) == 0
        if step == 20:
            assert len(ids) == 1, "the 70 m jump must leave every earlier key pose outside the 50 m radius"
    return rebuilds


@pytest.mark.parametrize("nanoflann", [True, False])
def test_keyframe_logic_against_python_restatement(nanoflann):
    if nanoflann and not oracle_py.load().lo_has_nanoflann():
        pytest.skip("reference nanoflann not compiled in (no /root/reference at build time)")
    p = config_params("A")
    o = Oracle(p, nanoflann=nanoflann)
    rebuilds = _drive(o, PyMapOptimization(), np.random.default_rng(7), 40, f" nanoflann={nanoflann}")
    assert rebuilds >= 4, "the averaged-id quirk (mapOptmization.cpp:940) must have erased key frames"


def test_known_answers():
    """Hand-derivable: key poses at z = 0.1, 0.5, 0.9 share one 1 m voxel, so the surrounding id is
    int((0+1+2)/3) = 1 -- not 0 and not 2 (mapOptmization.cpp:940,962) -- and the 0.3 m rule skips a 0.2 m move."""
    p = config_params("A")
    o = Oracle(p)
    corner = np.array([[1, 2, 3, 5]], F)
    o.map_set_scan(corner, np.zeros((0, 4), F))
    expect_ids = [[0], [0], [1]]
    for k, z in enumerate((0.1, 0.5, 0.9)):
        aft = np.array([0, 0, 0, 0, 0, z], F)
        o.map_set_poses(aft, np.zeros(6, F))
        o.upload("TRANSFORM_TOBE_MAPPED", aft)
        o.map_save_keyframe()
        o.map_extract_surrounding_keyframes()
        assert list(o.download("SURROUNDING_KEY_IDS")) == expect_ids[k]
    st = o.download("KEYFRAME_STATE")
    assert list(st[:3]) == [3, 1, 1]
    # local corner map = the single point of key frame 1 shifted by its pose
    assert np.array_equal(o.download("MAP_CORNER"), np.array([[1, 2, F(3) + F(0.5), 5]], F))
    # a 0.2 m move is not a new key frame
    o.map_set_poses(np.array([0, 0, 0, 0, 0, 1.1], F), np.zeros(6, F))
    o.map_save_keyframe()
    assert o.download("KEYFRAME_STATE")[0] == 3
    # key poses are (roll, pitch, yaw, x, y, z) of transformAftMapped; the first one comes from transformTobeMapped
    assert np.array_equal(o.download("KEY_POSES_6D")[:, 5], np.array([0.1, 0.5, 0.9], F))
