"""Per-phase nanoseconds of the two scan-to-scan LM stages (LL_BUF_STAGE_CLOCKS) on the arena drive: python tools/stage_clocks.py [B=16] [frames=20]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from lego_loam_bor_b200 import config_params, synth
from lego_loam_bor_b200.capi import LegoLoam
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
F = int(sys.argv[2]) if len(sys.argv) > 2 else 20
p = config_params("C"); cfg = synth.make_arena(p, n_keyframes=10)
dev = torch.device("cuda", 0)
gen = synth.ArenaDeviceGenerator(cfg, list(range(B)), dev)
gpu = LegoLoam(p, batch=B)
gpu.enable_stage_timing(True)
rows = []
for f in range(F):
    pts, counts = gen.scans(synth.DRIVE, f)
    torch.cuda.synchronize()
    gpu.set_scans_device(pts.data_ptr(), counts, p.num_vertical_scans * p.num_horizontal_scans)
    gpu.process_scans(); gpu.synchronize()
    if f >= 3:
        for s in range(B):
            c = gpu.download("STAGE_CLOCKS", s); it = gpu.download("ODOM_ITERS", s)
            rows.append(np.concatenate([c, it, [len(gpu.download("SURF_FLAT", s)), len(gpu.download("CORNER_SHARP", s))]]))
r = np.array(rows, dtype=np.float64)
names = ["total", "prologue", "re-search", "reduce", "solve", "researched_pts", "rows", "-"]
for st, nm in ((0, "SURF"), (1, "CORNER")):
    c = r[:, st * 8:(st + 1) * 8]
    it = r[:, 16 + st]
    print(nm, "iters mean %.2f max %d hist %s" % (it.mean(), it.max(), np.bincount(it.astype(int))[:26].tolist()), "features", r[:, 18 + st].mean())
    for k, n in enumerate(names[:7]):
        print("   %-14s mean %9.1f us   p95 %9.1f us  max %9.1f us" % (n, c[:, k].mean() / 1e3, np.percentile(c[:, k], 95) / 1e3, c[:, k].max() / 1e3) if k != 5 else "   %-14s mean %9.1f" % (n, c[:, k].mean()))
    tot = c[:, 0]
    per_frame_max = tot.reshape(-1, B).max(axis=1)
    print("   launch time ~ max over sequences: mean %.1f us" % (per_frame_max.mean() / 1e3))
