"""LM steps per scan-to-map iteration (MAP_TRACE) for a few sequences of the bench workload, to size the neighbour reuse.
  python tools/map_steps.py [B] [F] [synthetic|live]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from lego_loam_bor_b200 import config_params, synth
from lego_loam_bor_b200.capi import LegoLoam
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
F = int(sys.argv[2]) if len(sys.argv) > 2 else 11
live = len(sys.argv) > 3 and sys.argv[3] == "live"
params = config_params("C")
seq_ids = list(range(B))
cfg, scans, counts, N, _ = bench.gen_dataset(params, seq_ids, F)
gpu = LegoLoam(params, batch=B, max_points=N, device=0)
if live:
    gpu.map_enable_keyframes(max_keyframes=F // 5 + 8)
else:
    for k in range(B):
        gpu.map_set_local(k, *bench.local_maps(cfg, seq_ids[k]))
    aft = np.zeros((B, 6), np.float32)
    for k, s in enumerate(seq_ids):
        x, y, z, r, p, yaw = synth.pose(cfg, s, 0); aft[k] = [0, yaw, 0, y, z, x]
    gpu.map_set_poses(aft, np.zeros((B, 6), np.float32))
np.set_printoptions(linewidth=200, precision=5, suppress=True)
for f in range(F):
    gpu.set_scans_host([scans[(s, f)] for s in seq_ids])
    if gpu.process_scans() == 1:
        for k in range(min(B, 3)):
            it = gpu.download("MAP_ITERS", k)
            tr = gpu.download("MAP_TRACE", k).reshape(10, 34)
            print(f"frame {f} seq {k} iters/rows {it}")
            for i in range(int(it[0])):
                X = tr[i, 28:]
                print(f"   iter {i}: rows {int(tr[i, 27])}  |rot| {np.linalg.norm(X[:3]):.5f} rad  |trans| {np.linalg.norm(X[3:]):.4f} m")
