import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from lego_loam_bor_b200 import config_params, synth, workloads
from lego_loam_bor_b200.capi import LegoLoam
from oracle.oracle_py import Oracle
import test_gpu_kf500 as T
cfgname, K = sys.argv[1], int(sys.argv[2])
p = config_params(cfgname); cfg = synth.make_arena(p, n_keyframes=K)
gpu = LegoLoam(p, batch=1); gpu.map_enable_keyframes(*workloads.keyframe_capacities(p, K))
keep = T._gpu_prebuild(gpu, cfg, [0], K, False)
o = Oracle(p); T._oracle_prebuild(o, cfg, 0, K)
bad = 0
for kf in range(K):
    for which in range(3):
        a, b = gpu.download_keyframe(0, kf, which), o.download_keyframe(kf, which)
        if a.shape != b.shape or not np.array_equal(a, b):
            bad += 1
            if bad < 6: print("kf", kf, which, a.shape, b.shape, (np.abs(a-b).max() if a.shape == b.shape else None))
print("bad clouds", bad)
for f in range(6):
    sc = synth.arena_scan(cfg, 0, synth.DRIVE, f)
    gpu.set_scans_host([sc]); gpu.image_projection(); rc = gpu.feature_association()
    o.image_projection(sc); o.feature_association()
gpu.map_downsample_current_scan(); gpu.map_predict_pose(); gpu.map_extract_surrounding_keyframes()
o.mapping_cycle()
print("ids equal", np.array_equal(gpu.download("SURROUNDING_KEY_IDS"), o.download("SURROUNDING_KEY_IDS")), gpu.download("KEYFRAME_STATE"), o.download("KEYFRAME_STATE"))
for name, leaf in (("MAP_CORNER", 0.2), ("MAP_SURF", 0.4)):
    a, b = gpu.download(name), o.download(name)
    print(name, a.shape, b.shape)
    inv = np.float32(1.0) / np.float32(leaf)
    va = np.floor(a[:, :3] * inv).astype(np.int64); vb = np.floor(b[:, :3] * inv).astype(np.int64)
    sa = set(map(tuple, va)); sb = set(map(tuple, vb))
    print(" only gpu", list(sa - sb)[:5], " only oracle", list(sb - sa)[:5], "dups gpu", len(va) - len(sa), "dups oracle", len(vb) - len(sb))
    for v in list(sa - sb)[:3]:
        i = np.flatnonzero((va == v).all(1)); print("  gpu point", a[i])
