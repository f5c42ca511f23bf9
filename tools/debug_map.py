import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lego_loam_bor_b200 import config_params
from lego_loam_bor_b200.capi import LegoLoam
from oracle.oracle_py import Oracle
p = config_params("T")
rng = np.random.default_rng(0)
# plane z(cam y) = -1.5 sampled at 0.3 m, plus vertical lines
xs, zs = np.meshgrid(np.arange(-10, 10, 0.3), np.arange(-10, 10, 0.3))
surf = np.stack([xs.ravel(), np.full(xs.size, -1.5), zs.ravel(), np.zeros(xs.size)], 1).astype(np.float32)
surf[:, :3] += rng.normal(0, 0.01, (len(surf), 3)).astype(np.float32)
lines = []
for (cx, cz) in [(3, 4), (-5, 2), (6, -3), (-2, -6), (1, 8)]:
    ys = np.arange(-1.5, 3.0, 0.1)
    lines.append(np.stack([np.full(len(ys), cx), ys, np.full(len(ys), cz), np.zeros(len(ys))], 1))
corner = np.concatenate(lines).astype(np.float32)
corner[:, :3] += rng.normal(0, 0.005, (len(corner), 3)).astype(np.float32)
gpu = LegoLoam(p, batch=1)
o = Oracle(p)
gpu.map_set_local(0, corner, surf); o.map_set_local(corner, surf)
print("map roundtrip", np.array_equal(gpu.download("MAP_CORNER"), corner), np.array_equal(gpu.download("MAP_SURF"), surf))
qs, qc = surf[::3].copy(), corner[::2].copy()
gpu.map_set_scan(0, qc, qs); o.map_set_scan(qc, qs)
print("scan roundtrip", np.array_equal(gpu.download("SCAN_CORNER_DS"), qc), np.array_equal(gpu.download("SCAN_SURF_TOTAL_DS"), qs))
g = np.array([[0.002, 0.01, -0.003, 0.05, 0.02, -0.04]], np.float32)
gpu.map_set_initial_guess(g); o.map_set_initial_guess(g[0])
gpu.scan_to_map(); o.scan_to_map()
print("gpu iters", gpu.download("MAP_ITERS"), gpu.download("TRANSFORM_TOBE_MAPPED"))
print("ora iters", o.download("MAP_ITERS"), o.download("TRANSFORM_TOBE_MAPPED"))

tg, to = gpu.download("MAP_TRACE").reshape(10, 34), o.download("MAP_TRACE").reshape(10, 34)
np.set_printoptions(linewidth=200, precision=5, suppress=False)
for it in range(3):
    print("iter", it, "rows", tg[it, 27], to[it, 27])
    print(" gpu AtA", tg[it, :21]); print(" ora AtA", to[it, :21])
    print(" gpu AtB", tg[it, 21:27], "X", tg[it, 28:]); print(" ora AtB", to[it, 21:27], "X", to[it, 28:])
