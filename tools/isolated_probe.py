"""How many scan-to-scan feature points have NO last-frame point inside the acceptance radius (their search walks the whole
11^3-cell block, and the LM stage searches them again at iterations 5, 10, ...), and how the stage time relates:
python tools/isolated_probe.py [B=16] [frames=14]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from lego_loam_bor_b200 import config_params, synth
from lego_loam_bor_b200.capi import LegoLoam
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
F = int(sys.argv[2]) if len(sys.argv) > 2 else 14
p = config_params("C"); cfg = synth.make_arena(p, n_keyframes=10)
dev = torch.device("cuda", 0)
gen = synth.ArenaDeviceGenerator(cfg, list(range(B)), dev)
gpu = LegoLoam(p, batch=B)
gpu.enable_stage_timing(True)
gpu.enable_index_trace(True)
cap = 24 * p.num_vertical_scans
rows = []
for f in range(F):
    pts, counts = gen.scans(synth.DRIVE, f)
    torch.cuda.synchronize()
    gpu.set_scans_device(pts.data_ptr(), counts, p.num_vertical_scans * p.num_horizontal_scans)
    gpu.process_scans(); gpu.synchronize()
    if f < 3:
        continue
    for s in range(B):
        tr = gpu.download("ODOM_SEARCH_IDX", s).reshape(2, 5, cap, 3)
        it = gpu.download("ODOM_ITERS", s)
        clk = gpu.download("STAGE_CLOCKS", s)
        n_sharp, n_flat = len(gpu.download("CORNER_SHARP", s)), len(gpu.download("SURF_FLAT", s))
        c0 = tr[1, 0, :n_sharp, 0]; s0 = tr[0, 0, :n_flat, 0]
        changed = [(int(np.sum(np.any(tr[1, r, :n_sharp] != tr[1, r - 1, :n_sharp], axis=1)))) for r in range(1, 5)]
        rows.append((int(it[1]), n_sharp, int(np.sum(c0 < 0)), int(np.sum(tr[1, 0, :n_sharp, 1] < 0)), n_flat, int(np.sum(s0 < 0)), clk[8] / 1e3, clk[10] / 1e3, int(clk[13]), changed))
r = rows
print("corner: iters, sharp, no closest, no second, | flat, no closest | stage us, re-search us, re-searched points, correspondences changed per round")
for x in sorted(r, key=lambda x: -x[6])[:12]: print(x)
a = np.array([x[:9] for x in r], dtype=np.float64)
print("means", a.mean(axis=0).round(1))
