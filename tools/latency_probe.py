"""Where the single-sequence scan latency goes: per-kernel mean microseconds at batch 1 (config C, arena drive, live key-frame
map), plain frames and mapping frames apart.  python tools/latency_probe.py [K=60] [frames=41] [config=C]"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from lego_loam_bor_b200 import config_params, synth, workloads
from lego_loam_bor_b200.capi import LegoLoam
K = int(sys.argv[1]) if len(sys.argv) > 1 else 60
F = int(sys.argv[2]) if len(sys.argv) > 2 else 41
cfgname = sys.argv[3] if len(sys.argv) > 3 else "C"
p = config_params(cfgname)
N = p.num_vertical_scans * p.num_horizontal_scans
cfg = synth.make_arena(p, n_keyframes=K)
dev = torch.device("cuda", 0)
gen = synth.ArenaDeviceGenerator(cfg, [0], dev)
stream = torch.cuda.Stream(device=dev)
gpu = LegoLoam(p, batch=1, max_points=N, device=0, stream=stream.cuda_stream)
gpu.map_enable_keyframes(*workloads.keyframe_capacities(p, K))
buf = torch.zeros((1, N, 4), dtype=torch.float32, device=dev)
def scans_of(i):
    _, counts = gen.scans(synth.KEYFRAME, i, out=buf)
    return buf.data_ptr(), counts, N
workloads.prebuild_keyframes(gpu, cfg, [0], K, scans_of, sync=lambda: torch.cuda.synchronize(dev))
workloads.start_drive(gpu, cfg, [0])
frames = [gen.scans(synth.DRIVE, f) for f in range(F)]
torch.cuda.synchronize(dev)
# pass 1: whole-frame device time, frames enqueued back to back (what bench.py reports)
evs, kinds = [], []
with torch.cuda.stream(stream):
    for f in range(F):
        pts, counts = frames[f]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gpu.set_scans_device(pts.data_ptr(), counts, N)
        e0.record(stream); rc = gpu.process_scans(); e1.record(stream)
        evs.append((e0, e1)); kinds.append(rc)
torch.cuda.synchronize(dev)
ms = np.array([a.elapsed_time(b) for a, b in evs])[7:]; kinds = np.array(kinds)[7:]
out = {"frame_ms": {"plain_p50": float(np.percentile(ms[kinds == 0], 50)), "plain_p95": float(np.percentile(ms[kinds == 0], 95)),
                    "mapping_p50": float(np.percentile(ms[kinds == 1], 50)) if np.any(kinds == 1) else None}}
# pass 2: the same frames again with every kernel event-timed
gpu.reset_feature_association(); workloads.start_drive(gpu, cfg, [0])
gpu.time_kernel("*")
for f in range(F):
    pts, counts = frames[f]
    gpu.set_scans_device(pts.data_ptr(), counts, N); gpu.process_scans()
gpu.synchronize()
tab = gpu.kernel_time_table()
out["kernel_us_mean"] = {k: [round(1e3 * v[0] / max(1, v[1]), 1), v[1]] for k, v in sorted(tab.items(), key=lambda kv: -kv[1][0])}
out["sum_per_frame_us"] = round(1e3 * sum(v[0] for v in tab.values()) / F, 1)
out["odom_iters_last"] = [int(x) for x in gpu.download("ODOM_ITERS", 0)]
print(json.dumps(out))
