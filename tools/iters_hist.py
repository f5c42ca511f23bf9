"""Histogram of LM iteration counts over B sequences and F frames + per-kernel time table (one stream)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from lego_loam_bor_b200 import config_params, synth
from lego_loam_bor_b200.capi import LegoLoam

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
F = int(sys.argv[2]) if len(sys.argv) > 2 else 12
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
for k in range(B):
    gpu.map_set_local(k, *bench.local_maps(cfg, seq_ids[k]))
aft = np.zeros((B, 6), np.float32)
for k, s in enumerate(seq_ids):
    x, y, z, r, p, yaw = synth.pose(cfg, s, 0); aft[k] = [0, yaw, 0, y, z, x]
gpu.map_set_poses(aft, np.zeros((B, 6), np.float32))
hist_s, hist_c, hist_m = np.zeros(26, int), np.zeros(26, int), np.zeros(11, int)
for f in range(F):
    if f == 2:
        gpu.time_kernel("*")
    gpu.set_scans_device(devdata.data_ptr() + f * B * N * 16, counts[f], N)
    rc = gpu.process_scans()
    if f >= 1:
        for k in range(B):
            it = gpu.download("ODOM_ITERS", k)
            hist_s[it[0]] += 1; hist_c[it[1]] += 1
            if f == F - 1 and (it[1] >= 20 or k < 3):
                print("seq", k, "iters", it, "stage clocks surf", gpu.download("STAGE_CLOCKS", k)[:6].tolist(), "corner", gpu.download("STAGE_CLOCKS", k)[8:16].tolist(), gpu.download("TRANSFORM_CUR", k))
            if rc == 1:
                hist_m[gpu.download("MAP_ITERS", k)[0]] += 1
print("surf iters hist", hist_s.tolist())
print("corner iters hist", hist_c.tolist())
print("map iters hist", hist_m.tolist())
tab = gpu.kernel_time_table()
steps = F - 2
for k, (ms, n) in sorted(tab.items(), key=lambda kv: -kv[1][0]):
    print(f"  {k:26s} {1e3*ms/steps:8.1f} us/step  {n/steps:5.1f} launches/step  avg {1e3*ms/max(n,1):8.1f} us")
print("total us/step", 1e3 * sum(v[0] for v in tab.values()) / steps)
