"""Per-phase SM cycles of k_feature_ring (LL_BUF_RING_CLOCKS) on the arena drive: python tools/ring_clocks.py [B=1] [frames=12]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from lego_loam_bor_b200 import config_params, synth
from lego_loam_bor_b200.capi import LegoLoam
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
F = int(sys.argv[2]) if len(sys.argv) > 2 else 12
p = config_params("C"); cfg = synth.make_arena(p, n_keyframes=10)
dev = torch.device("cuda", 0)
gen = synth.ArenaDeviceGenerator(cfg, list(range(B)), dev)
gpu = LegoLoam(p, batch=B)
rows = []
for f in range(F):
    pts, counts = gen.scans(synth.DRIVE, f)
    torch.cuda.synchronize()
    gpu.set_scans_device(pts.data_ptr(), counts, p.num_vertical_scans * p.num_horizontal_scans)
    gpu.process_scans(); gpu.synchronize()
    if f >= 3:
        for s in range(B):
            rows.append(gpu.download("RING_CLOCKS", s))
r = np.concatenate(rows).astype(np.float64)          # [frames * B * V][10]
names = ["total", "load+keys", "spec picks", "boundary/reruns", "persist+collect", "bbox+voxkeys", "run heads", "run sort", "centroids"]
ghz = 1.965
info = r[:, 9].astype(np.int64)
span, nraw, nruns, rounds = info & 0xffff, (info >> 16) & 0xffff, (info >> 32) & 0xffff, (info >> 48) & 0xff
order = np.argsort(-r[:, 0])
print("rings %d   span mean %.0f max %d   n_raw mean %.0f max %d   runs mean %.0f max %d" % (len(r), span.mean(), span.max(), nraw.mean(), nraw.max(), nruns.mean(), nruns.max()))
for k, n in enumerate(names):
    print("  %-16s mean %8.1f us   p95 %8.1f us   max %8.1f us   in the 16 slowest rings %8.1f us" % (n, r[:, k].mean() / ghz / 1e3, np.percentile(r[:, k], 95) / ghz / 1e3, r[:, k].max() / ghz / 1e3, r[order[:16], k].mean() / ghz / 1e3))
print("re-run rounds histogram", np.bincount(rounds).tolist())
print("slowest rings: (span, n_raw, runs, total us)", [(int(span[i]), int(nraw[i]), int(nruns[i]), round(r[i, 0] / ghz / 1e3, 1)) for i in order[:8]])
