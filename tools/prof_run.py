"""Short run for ncu captures: B sequences of config C, one stream, F frames (default 16 x 8, map on).
  python tools/prof_run.py [B] [F] [synthetic|live]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from lego_loam_bor_b200 import config_params, synth
from lego_loam_bor_b200.capi import LegoLoam

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
F = int(sys.argv[2]) if len(sys.argv) > 2 else 8
params = config_params("C")
dev = torch.device("cuda", 0)
seq_ids = list(range(B))
cfg, scans, counts, N, _ = bench.gen_dataset(params, seq_ids, F)
host = torch.zeros((F, B, N, 4), dtype=torch.float32)
hv = host.numpy()
for f in range(F):
    for k, s in enumerate(seq_ids):
        a = scans[(s, f)]; hv[f, k, :len(a)] = a; counts[f, k] = len(a)
devdata = host.to(dev)
gpu = LegoLoam(params, batch=B, max_points=N, device=0)
if len(sys.argv) > 3 and sys.argv[3] == "live":
    gpu.map_enable_keyframes(max_keyframes=F // 5 + 8, pool_points=(F // 5 + 8) * (N // 8), max_map_corner=N // 2, max_map_surf=N)
else:
    for k in range(B):
        gpu.map_set_local(k, *bench.local_maps(cfg, seq_ids[k]))
    aft = np.zeros((B, 6), np.float32)
    for k, s in enumerate(seq_ids):
        x, y, z, r, p, yaw = synth.pose(cfg, s, 0); aft[k] = [0, yaw, 0, y, z, x]
    gpu.map_set_poses(aft, np.zeros((B, 6), np.float32))
for f in range(F):
    gpu.set_scans_device(devdata.data_ptr() + f * B * N * 16, counts[f], N)
    gpu.process_scans()
gpu.synchronize()
print("ok", gpu.kernel_launches(), gpu.download("ODOM_ITERS", 0), gpu.download("TRANSFORM_SUM", 0))
