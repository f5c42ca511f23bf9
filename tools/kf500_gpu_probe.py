"""GPU probe of the kf500 workload: python tools/kf500_gpu_probe.py [B=16] [K=500] [frames=21] [--check]
Prebuilds K key frames per sequence from device-generated arena scans, drives `frames` frames, prints per-cycle device time
and the per-kernel time table; --check compares sequence 0 with the oracle (same flow on the CPU)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from lego_loam_bor_b200 import config_params, synth, workloads
from lego_loam_bor_b200.capi import LegoLoam

args = [a for a in sys.argv[1:] if not a.startswith("--")]
B = int(args[0]) if len(args) > 0 else 16
K = int(args[1]) if len(args) > 1 else 500
F = int(args[2]) if len(args) > 2 else 21
cfgname = args[3] if len(args) > 3 else "C"
check = "--check" in sys.argv
p = config_params(cfgname)
N = p.num_vertical_scans * p.num_horizontal_scans
cfg = synth.make_arena(p, n_keyframes=K)
dev = torch.device("cuda", 0)
seqs = list(range(B))
gen = synth.ArenaDeviceGenerator(cfg, seqs, dev)
stream = torch.cuda.Stream(device=dev)
gpu = LegoLoam(p, batch=B, max_points=N, device=0, stream=stream.cuda_stream)
gpu.map_enable_keyframes(*workloads.keyframe_capacities(p, K))
buf = torch.zeros((B, N, 4), dtype=torch.float32, device=dev)

def scans_of(i):
    _, counts = gen.scans(synth.KEYFRAME, i, out=buf)
    return buf.data_ptr(), counts, N

t0 = time.time()
workloads.prebuild_keyframes(gpu, cfg, seqs, K, scans_of, sync=lambda: torch.cuda.synchronize(dev))
t_pre = time.time() - t0
workloads.start_drive(gpu, cfg, seqs)
frames = [gen.scans(synth.DRIVE, f) for f in range(F)]
torch.cuda.synchronize(dev)
res = {"B": B, "K": K, "prebuild_s": t_pre, "state_after_prebuild": [int(x) for x in gpu.download("KEYFRAME_STATE")], "cycles": []}
gpu.time_kernel("*")
for f in range(F):
    if f == 12 and "--steady" in sys.argv:
        gpu.time_kernel("*")   # restart the table: only steady-state cycles (one appended key frame each) are counted
    if f == 12 and "--ncu-range" in sys.argv:
        torch.cuda.synchronize(); torch.cuda.profiler.start()   # ncu --profile-from-start off: only the steady-state frames
    pts, counts = frames[f]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gpu.set_scans_device(pts.data_ptr(), counts, N)
    with torch.cuda.stream(stream):
        e0.record(stream)
        rc = gpu.process_scans()
        e1.record(stream)
    torch.cuda.synchronize(dev)
    if rc == 1:
        res["cycles"].append({"frame": f, "ms": e0.elapsed_time(e1), "state": [int(x) for x in gpu.download("KEYFRAME_STATE")],
                              "map": [len(gpu.download("MAP_CORNER")), len(gpu.download("MAP_SURF"))],
                              "iters_rows": [int(x) for x in gpu.download("MAP_ITERS")]})
    else:
        res.setdefault("plain_ms", []).append(e0.elapsed_time(e1))
if "--ncu-range" in sys.argv:
    torch.cuda.synchronize(); torch.cuda.profiler.stop()
tab = gpu.kernel_time_table()
res["kernels_ms_total"] = {k: [round(v[0], 3), v[1]] for k, v in sorted(tab.items(), key=lambda kv: -kv[1][0])}
if check:
    from oracle.oracle_py import Oracle
    o = Oracle(p, libm=False, nanoflann=True)
    for i in range(K):
        o.reset_feature_association()
        o.image_projection(synth.arena_scan(cfg, 0, synth.KEYFRAME, i))
        o.feature_association()
        o.map_downsample_current_scan()
        T = synth.pose_to_transform(synth.arena_pose(cfg, 0, synth.KEYFRAME, i))
        o.map_set_poses(T, np.zeros(6, np.float32)); o.map_set_initial_guess(T); o.map_save_keyframe()
    o.reset_feature_association()
    o.map_set_poses(synth.pose_to_transform(synth.arena_pose(cfg, 0, synth.DRIVE, 0)), np.zeros(6, np.float32))
    same_input = True
    for f in range(F):
        sc = synth.arena_scan(cfg, 0, synth.DRIVE, f)
        pts, counts = frames[f]
        same_input &= bool(np.array_equal(sc, pts[0, :counts[0]].cpu().numpy()))
        o.image_projection(sc)
        if o.feature_association() == 1:
            o.mapping_cycle()
    par = {"device_scans_equal_host_scans": same_input}
    for name in ("MAP_CORNER", "MAP_SURF", "SURROUNDING_KEY_IDS", "MAP_ITERS", "KEYFRAME_STATE", "TRANSFORM_SUM", "TRANSFORM_AFT_MAPPED", "KEY_POSES_6D"):
        a, b = gpu.download(name, 0), o.download(name)
        par[name] = bool(a.shape == b.shape and np.array_equal(a, b))
        if not par[name] and a.shape == b.shape and a.dtype.kind == "f":
            par[name + "_maxdiff"] = float(np.abs(a - b).max())
    res["parity_seq0"] = par
print(json.dumps(res))
