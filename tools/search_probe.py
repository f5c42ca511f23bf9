"""How the scan-to-scan correspondence searches depend on the acceptance radius (nearest_feature_search_distance): per-kernel
mean microseconds at B sequences.  python tools/search_probe.py [B=16] [frames=12]"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from lego_loam_bor_b200 import config_params, synth
from lego_loam_bor_b200.capi import LegoLoam
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
F = int(sys.argv[2]) if len(sys.argv) > 2 else 12
dev = torch.device("cuda", 0)
for radius in (5.0, 3.0, 2.0, 1.0):
    p = config_params("C"); p.nearest_feature_search_distance = radius
    cfg = synth.make_arena(p, n_keyframes=10)
    gen = synth.ArenaDeviceGenerator(cfg, list(range(B)), dev)
    gpu = LegoLoam(p, batch=B)
    for f in range(F):
        pts, counts = gen.scans(synth.DRIVE, f)
        torch.cuda.synchronize()
        if f == 3:
            gpu.time_kernel("*")
        gpu.set_scans_device(pts.data_ptr(), counts, p.num_vertical_scans * p.num_horizontal_scans)
        gpu.process_scans(); gpu.synchronize()
    tab = gpu.kernel_time_table()
    print(radius, {k: round(1e3 * v[0] / max(1, v[1]), 1) for k, v in tab.items() if "odom" in k}, flush=True)
    gpu.close()
