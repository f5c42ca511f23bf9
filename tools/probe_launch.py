"""Probe: host enqueue time vs device time per step at several stream counts (is the path launch-bound?)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from lego_loam_bor_b200 import config_params, synth
from lego_loam_bor_b200.capi import LegoLoam, LegoLoamStreams

def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    params = config_params("C")
    dev = torch.device("cuda", 0)
    n_frames = 26
    seq_ids = list(range(B))
    cfg, scans, counts, N, _ = bench.gen_dataset(params, seq_ids, n_frames)
    host = torch.empty((n_frames, B, N, 4), dtype=torch.float32).pin_memory()
    hv = host.numpy()
    for f in range(n_frames):
        for k, s in enumerate(seq_ids):
            a = scans[(s, f)]; hv[f, k, :len(a)] = a; counts[f, k] = len(a)
    devdata = host.to(dev)
    maps = [bench.local_maps(cfg, s) for s in seq_ids]
    for ns in (1, 2, 4, 8):
        streams = [torch.cuda.Stream(device=dev) for _ in range(ns)]
        gpu = LegoLoamStreams(params, B, ns, max_points=N, device=0, streams=[s.cuda_stream for s in streams])
        for k in range(B):
            gpu.map_set_local(k, *maps[k])
        aft = np.zeros((B, 6), np.float32)
        for k, s in enumerate(seq_ids):
            x, y, z, r, p, yaw = synth.pose(cfg, s, 0); aft[k] = [0, yaw, 0, y, z, x]
        gpu.map_set_poses(aft, np.zeros((B, 6), np.float32))
        fb = B * N * 16
        def step(f):
            gpu.set_scans_device(devdata.data_ptr() + f * fb, counts[f], N); gpu.process_scans()
        for f in range(6):
            step(f)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for f in range(6, 26):
            step(f)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"streams={ns}: host enqueue {1e3*(t1-t0)/20:.3f} ms/step, total wall {1e3*(t2-t0)/20:.3f} ms/step", flush=True)
        del gpu

main()
