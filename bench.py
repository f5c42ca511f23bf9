#!/usr/bin/env python
"""bench.py -- scans/sec of the LeGO-LOAM-BOR per-scan hot path at 64x2048 on B200.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU algorithm (oracle)

Workload (BASELINE.json configs[3]/[4]): 64-beam 64x2048 synthetic lidar, `--batch` (default 64)
independent sequences per GPU advancing in lock step, scan-to-map against a synthetic local map every
`mapping_frequency_divider`-th odometry frame.  A "step" is one scan of every sequence through the
whole hot path: projection, ground removal, segmentation, feature extraction, scan-to-scan LM,
and (every 5th step) downsampleCurrentScan + scan-to-map.  Sequences are independent, so N GPUs run N
independent batches (weak scaling, no data-path collective).

One JSON line is printed by rank 0 (keys: see the task contract).  `value` is timed with inputs already
resident in HBM; `e2e` goes through the same C ABI with pinned HOST scans (H2D inside the timed
region, pose D2H every step).  `--e2e-input xyz` (default) hands the scans over as packed 12-byte points
(ll_set_scans_xyz_host: the path never reads the intensity a sensor reports, imageProjection.cpp:216), `xyzi`
as 16-byte points (ll_set_scans_host); the other form is measured too and reported as `e2e_alt`.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "scans/sec"
UNIT = "scans/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="sequences per GPU")
    ap.add_argument("--config", default="C", help="sensor config: A 16x1800, B 32x1800, C 64x2048")
    ap.add_argument("--unique-seqs", type=int, default=0, help="distinct synthetic sequences per GPU (0 = batch)")
    ap.add_argument("--no-map", action="store_true", help="skip scan-to-map")
    ap.add_argument("--map", default="synthetic", choices=["synthetic", "live"],
                    help="synthetic: scan-to-map against a fixed pre-built local map (headline workload); live: the "
                         "whole MapOptimization::run body -- key frames saved and the local map re-assembled from them "
                         "on the device every mapping cycle (SURVEY 8 f2)")
    ap.add_argument("--cpu-frames", type=int, default=11, help="frames per sequence of the cpu_baseline sample")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--time-kernel", default="", help="kernel to report in `roofline` (default: the slowest)")
    ap.add_argument("--e2e-input", default="xyz", choices=["xyz", "xyzi"],
                    help="host point format of the `e2e` leg: packed 12-byte xyz (ll_set_scans_xyz_host; the path never reads "
                         "the sensor's intensity) or 16-byte xyzi (ll_set_scans_host); the other one is reported as e2e_alt")
    ap.add_argument("--streams", type=int, default=4, help="split the batch over this many handles / CUDA streams")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------
# data


def gen_dataset(params, seqs, n_frames, seed_base=0):
    """scans[f][k] for sequence seqs[k]; returns packed float32 [F, len(seqs), stride, 4] + counts [F, len(seqs)]."""
    from lego_loam_bor_b200 import synth
    cfg = synth.make_config(params)
    N = params.num_vertical_scans * params.num_horizontal_scans
    uniq = sorted(set(seqs))
    t0 = time.time()
    scans = synth.scans(cfg, uniq, range(n_frames), threads=max(1, os.cpu_count() or 1))
    gen_s = time.time() - t0
    counts = np.zeros((n_frames, len(seqs)), np.int32)
    return cfg, scans, counts, N, gen_s


def local_maps(cfg, seq):
    from lego_loam_bor_b200 import synth
    return synth.local_map(cfg, seq, 1, 0.2), synth.local_map(cfg, seq, 0, 0.4)


# --------------------------------------------------------------------------------------------
# multi-GPU: independent sequences are sharded over ranks, no data-path collective (replicas only)


def shard_sequences(rank, batch, unique=0):
    """Sequence ids of the `batch` slots of `rank` (weak scaling: every rank owns its own sequences).
    With unique < batch the rank's distinct sequences are replicated over its slots."""
    u = unique or batch
    return [rank * batch + (k % u) for k in range(batch)]


def aggregate_throughput(world, batch, steps, local_ms, reduce_max=None):
    """Whole-job scans/s: all ranks' scans divided by the slowest rank's device time.
    reduce_max: callable mapping a local float to the max over ranks (identity for one rank)."""
    worst_ms = reduce_max(local_ms) if reduce_max else local_ms
    return world * batch * steps / (worst_ms * 1e-3), worst_ms


# --------------------------------------------------------------------------------------------
# clocks


class ClockSampler:
    """nvidia-smi sampled every 100 ms while the benchmark runs (B200_PROFILING.md clocks line)."""

    def __init__(self, uuid):
        self.rows = []
        self.proc = None
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap,utilization.gpu")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", uuid, f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t_begin=None, t_end=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, reasons, util = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            if t_begin is not None and not (t_begin - 0.15 <= ts <= t_end + 0.15):
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])); smax.append(float(f[1])); util.append(float(f[7]))
            except Exception:
                continue
            for i, n in enumerate(names):
                if f[3 + i].lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(smax)) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------
# algorithmic bytes per launch of every kernel (DESIGN.md "Kernels and rooflines"); all per sequence


def algorithmic_bytes(kernel, st):
    N, S, n_in, g = st["N"], st["S"], st["n_in"], st["g"]
    lf, ls, ff, fs = st["less_flat"], st["less_sharp"], st["flat"], st["sharp"]
    qs, qc, ms, mc = st["q_surf"], st["q_corner"], st["map_surf"], st["map_corner"]
    out = st["outlier"]
    table = {
        "k_project_scatter": 16 * n_in + 8 * n_in,
        "k_gather_ground": 29 * N + 16 * g * N,
        "k_ccl_rows": 12 * N, "k_ccl_merge": 8 * N, "k_ccl_flatten": 8 * N,
        "k_seg_count": 5 * N, "k_seg_emit": 5 * N + 45 * S + 4 * S, "k_label_final": 8 * N,
        "k_feature_prep": 61 * S + 4 * S,
        "k_feature_sort": 21 * S, "k_feature_pick": 8 * S, "k_feature_lessflat": 24 * S + 16 * lf,
        "k_feature_compact": 32 * lf + 20 * (ls + ff + fs),
        # correspondence search: feature point in, geometry out, the last-frame cloud read once
        "k_odom_search_surf": ff * 96 + lf * 16, "k_odom_search_corner": fs * 80 + ls * 16,
        # LM stage: features + their correspondence geometry once per launch (they then live in shared memory)
        "k_odom_stage_surf": ff * 69, "k_odom_stage_corner": fs * 85,
        "k_publish_clouds_last": 32 * (lf + ls + out),
        "k_grid_count": 16 * (lf + ls), "k_grid_tile_sums": 0, "k_grid_scan": 0, "k_grid_fill": 32 * (lf + ls),
        "k_voxel_grid": 16 * (lf + ls + out) * 2, "k_voxel_grid_total": 16 * qs * 2,
        # 5-NN: query + 10-candidate record written (full search) or read (reuse) + the map read once
        "k_map_knn": (qs + qc) * (16 + 176) + (ms + mc) * 16,
        "k_map_iter": (qs + qc) * (16 + 4 + 5 * 16), "k_map_solve": 0,
        # key frames / local map (--map live): one appended key frame per cycle (point + voxel key in, voxel sums
        # read-modify-written), every occupied voxel read and one centroid written, the key frame's clouds stored
        "k_kf_select": 0, "k_kf_decide": 0, "k_kf_accumulate": (qs + qc) * (16 + 8 + 2 * 28),
        "k_kf_extract": (ms + mc) * (8 + 20 + 16), "k_kf_store": (qs + qc) * (16 + 28),
    }
    return float(table.get(kernel, 0))


# --------------------------------------------------------------------------------------------
# CPU arm (oracle): the reference's algorithm on the host cores


def cpu_pipeline(params, cfg_name, n_threads, n_frames, use_map, frames_data=None, live=False):
    """Runs n_threads independent sequences of n_frames frames through the CPU oracle, one thread per
    sequence (the ctypes calls release the GIL).  Returns (scans/s, wall seconds, scans, per-stage seconds)."""
    from concurrent.futures import ThreadPoolExecutor
    from lego_loam_bor_b200 import synth
    from oracle import oracle_py
    cfg = synth.make_config(params)
    scans = frames_data or synth.scans(cfg, range(n_threads), range(n_frames), threads=n_threads)
    maps = {s: local_maps(cfg, s) for s in range(n_threads)} if use_map and not live else {}
    oracles = [oracle_py.Oracle(params, libm=True, nanoflann=True) for _ in range(n_threads)]
    for s, o in enumerate(oracles):  # frame 0 only initialises (featureAssociation.cpp:1414-1417)
        o.image_projection(scans[(s, 0)])
        o.feature_association()
        if use_map and not live:
            o.map_set_local(*maps[s])
            x, y, z, roll, pitch, yaw = synth.pose(cfg, s, 0)
            o.map_set_poses(np.array([0, yaw, 0, y, z, x], np.float32), np.zeros(6, np.float32))
        o.reset_timers()

    def run(s):
        o = oracles[s]
        for f in range(1, n_frames):
            o.image_projection(scans[(s, f)])
            if o.feature_association() == 1 and use_map:
                if live:
                    o.mapping_cycle()
                    continue
                o.map_downsample_current_scan()
                o.map_predict_pose()
                o.scan_to_map()
        return np.concatenate([o.timers(), [o.timer_map_assembly()]])

    t0 = time.time()
    with ThreadPoolExecutor(n_threads) as ex:
        timers = list(ex.map(run, range(n_threads)))
    wall = time.time() - t0
    n_scans = n_threads * (n_frames - 1)
    return n_scans / wall, wall, n_scans, np.sum(timers, axis=0), oracle_py.kind()


def run_reference_arm(args, params):
    """--impl reference: the reference's own CPU algorithm (oracle; k-NN by the reference's vendored
    nanoflann when it was compiled in) on all host cores; step = one scan on every core."""
    from lego_loam_bor_b200 import synth
    from oracle import oracle_py
    cores = max(1, os.cpu_count() or 1)
    cfg = synth.make_config(params)
    n_frames = 1 + args.warmup + args.steps
    scans = synth.scans(cfg, range(cores), range(n_frames), threads=cores)
    use_map = not args.no_map
    live = args.map == "live"
    maps = {s: local_maps(cfg, s) for s in range(cores)} if use_map and not live else {}
    oracles = [oracle_py.Oracle(params, libm=True, nanoflann=True) for _ in range(cores)]
    from concurrent.futures import ThreadPoolExecutor

    def step(s, f):
        o = oracles[s]
        o.image_projection(scans[(s, f)])
        if o.feature_association() == 1 and use_map:
            if live:
                o.mapping_cycle()
                return
            o.map_downsample_current_scan()
            o.map_predict_pose()
            o.scan_to_map()

    with ThreadPoolExecutor(cores) as ex:
        for s in range(cores):
            if use_map and not live:
                oracles[s].map_set_local(*maps[s])
                x, y, z, roll, pitch, yaw = synth.pose(cfg, s, 0)
                oracles[s].map_set_poses(np.array([0, yaw, 0, y, z, x], np.float32), np.zeros(6, np.float32))
        for f in range(0, 1 + args.warmup):
            list(ex.map(lambda s: step(s, f), range(cores)))
        t0 = time.time()
        for f in range(1 + args.warmup, n_frames):
            list(ex.map(lambda s: step(s, f), range(cores)))
        wall = time.time() - t0
    value = cores * args.steps / wall
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, params, cores, "cpu"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "knn": oracle_py.kind(),
                         "sample": f"{cores} sequences x {args.steps} scans of the same workload, one thread per sequence"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(args, params, batch, where):
    return {"workload": f"{params.num_vertical_scans}x{params.num_horizontal_scans} synthetic lidar, {batch} independent "
                        f"sequences per {'GPU' if where == 'gpu' else 'host'}, full hot path per scan (projection, ground, "
                        "segmentation, features, scan-to-scan LM" + ("" if args.no_map else (", scan-to-map every 5th scan vs synthetic local map" if args.map == "synthetic" else ", whole mapping cycle every 5th scan: key frames + local map assembled from them + scan-to-map")) + ")",
            "map": "none" if args.no_map else args.map,
            "sensor": args.config, "batch_per_gpu": batch, "streams_per_gpu": args.streams, "parallelism": f"replicas x{args.gpus} (independent sequences, no collective)",
            "l2": "every step reads a distinct set of scans (inputs per step ~ L2 size, dataset >> L2); no reuse between steps"}


# --------------------------------------------------------------------------------------------


def kernel_rooflines(alone, st, sub, peak):
    """Every kernel timed ALONE on the GPU (one sub-batch of `sub` sequences per launch): average launch duration, achieved
    algorithmic GB/s and fraction of the measured HBM peak; `psf_mean_frac` is the time-weighted mean over the
    projection / segmentation / feature kernels (the group north_star's 60 % target is about)."""
    out, t_psf, b_psf = {}, 0.0, 0.0
    for k, (ms, n) in sorted(alone.items(), key=lambda kv: -kv[1][0]):
        if n == 0:
            continue
        us = 1e3 * ms / n
        alg = algorithmic_bytes(k, st) * sub
        gbs = alg / (us * 1e-6) / 1e9 if us > 0 else 0.0
        out[k] = {"avg_us": round(us, 1), "launches": n, "alg_MB": round(alg / 1e6, 2), "GBps": round(gbs, 1),
                  "frac": round(gbs / peak, 4) if peak else None}
        if k in PSF_KERNELS:
            t_psf += us
            b_psf += alg
    out["psf_mean_frac"] = round(b_psf / (t_psf * 1e-6) / 1e9 / peak, 4) if t_psf > 0 and peak else None
    out["sequences_per_launch"] = sub
    return out


PSF_KERNELS = ["k_project_scatter", "k_gather_ground", "k_ccl_rows", "k_ccl_merge", "k_ccl_flatten", "k_seg_count",
               "k_seg_emit", "k_label_final", "k_feature_prep", "k_feature_sort", "k_feature_pick", "k_feature_lessflat",
               "k_feature_compact"]


def single_sequence_latency(params, devdata, counts, stride, n_frames, B, dev, torch, local_map=None, aft0=None, live=False):
    """p50 / p95 device latency of one scan of ONE sequence (batch 1, one stream; BASELINE.json configs[3]):
    the projection..odometry chain of every scan, scan-to-map on every 5th."""
    from lego_loam_bor_b200.capi import LegoLoam
    stream = torch.cuda.Stream(device=dev)
    one = LegoLoam(params, batch=1, max_points=stride, device=dev.index, stream=stream.cuda_stream)
    if live:
        N1 = params.num_vertical_scans * params.num_horizontal_scans
        one.map_enable_keyframes(max_keyframes=min(1024, n_frames // 5 + 8), pool_points=(n_frames // 5 + 8) * (N1 // 8),
                                 max_map_corner=N1 // 2, max_map_surf=N1)
    elif local_map is not None:
        one.map_set_local(0, *local_map)
        one.map_set_poses(aft0, np.zeros((1, 6), np.float32))
    frame_bytes = B * stride * 16
    evs = []
    with torch.cuda.stream(stream):
        for f in range(n_frames):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            one.set_scans_device(devdata.data_ptr() + f * frame_bytes, counts[f][:1], stride)
            e0.record(stream)
            one.process_scans()
            e1.record(stream)
            evs.append((e0, e1))
    torch.cuda.synchronize(dev)
    ms = np.array([a.elapsed_time(b) for a, b in evs[3:]])  # first frame only initialises; two more to warm up
    one.close()
    return {"batch": 1, "scans": int(len(ms)), "p50_ms": float(np.percentile(ms, 50)), "p95_ms": float(np.percentile(ms, 95)),
            "mean_ms": float(ms.mean()),
            "note": "device time per scan of one sequence, one stream; every 5th scan also runs scan-to-map"
                    + ("" if local_map is not None else " (disabled: --no-map)")}


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import __graft_entry__ as ge
    from lego_loam_bor_b200 import config_params
    params = config_params(args.config)
    if args.impl == "reference":
        if rank != 0:
            return 0
        ge.build()
        run_reference_arm(args, params)
        return 0

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    if rank == 0:
        ge.build()
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    from lego_loam_bor_b200 import synth
    from lego_loam_bor_b200.capi import LegoLoam

    B = args.batch
    U = args.unique_seqs or B
    use_map = not args.no_map
    prof_steps = 5                               # one mapping cycle with every kernel timed (picks the roofline kernel)
    n_frames = 1 + args.warmup + prof_steps + args.steps
    seq_ids = shard_sequences(rank, B, U)   # weak scaling: every rank has its own sequences
    cfg, scans, counts, N, gen_s = gen_dataset(params, seq_ids, n_frames)
    stride = N
    # pinned host dataset [F][B][stride][4] and a device-resident copy
    host = torch.empty((n_frames, B, stride, 4), dtype=torch.float32).pin_memory()
    hv = host.numpy()
    for f in range(n_frames):
        for k, s in enumerate(seq_ids):
            a = scans[(s, f)]
            hv[f, k, :len(a)] = a
            counts[f, k] = len(a)
    del scans
    devdata = host.to(dev, non_blocking=False)
    stream = torch.cuda.Stream(device=dev)
    sub_streams = [torch.cuda.Stream(device=dev) for _ in range(args.streams)] if args.streams > 1 else [stream]
    if args.streams > 1:
        from lego_loam_bor_b200.capi import LegoLoamStreams
        gpu = LegoLoamStreams(params, B, args.streams, max_points=stride, device=local_rank,
                              streams=[s.cuda_stream for s in sub_streams])
    else:
        gpu = LegoLoam(params, batch=B, max_points=stride, device=local_rank, stream=stream.cuda_stream)

    def fork(ev):
        """timed regions are bracketed on `stream`; the per-handle streams start after ev and are joined before the end event"""
        if args.streams > 1:
            for ss in sub_streams:
                ss.wait_event(ev)

    def join():
        if args.streams > 1:
            for ss in sub_streams:
                stream.wait_stream(ss)

    live = use_map and args.map == "live"
    if live:
        kf_cap = min(1024, max(16, n_frames // max(1, params.mapping_frequency_divider) + 8))
        gpu.map_enable_keyframes(max_keyframes=kf_cap, pool_points=kf_cap * (N // 8), max_map_corner=N // 2, max_map_surf=N)
    elif use_map:
        for k, s in enumerate(seq_ids):
            cm, sm = local_maps(cfg, s)
            gpu.map_set_local(k, cm, sm)
    frame_bytes = B * stride * 16

    def seed_map_poses():
        if live:
            return  # the map frame is the odometry frame of the first key frame, like the reference
        _seed_map_poses()

    def _seed_map_poses():
        """transformAftMapped = pose of frame 0 in the map frame, transformBefMapped = odometry origin; from
        then on the odometry -> map chain (transformAssociateToMap / transformUpdate) stays on the device."""
        aft = np.zeros((B, 6), np.float32)
        for k, s in enumerate(seq_ids):
            x, y, z, roll, pitch, yaw = synth.pose(cfg, s, 0)
            aft[k] = [0, yaw, 0, y, z, x]
        gpu.map_set_poses(aft, np.zeros((B, 6), np.float32))

    if use_map:
        seed_map_poses()

    def step_device(f):
        gpu.set_scans_device(devdata.data_ptr() + f * frame_bytes, counts[f], stride)
        gpu.process_scans()

    host_xyz = host[..., :3].contiguous().pin_memory()   # the same scans as packed 12-byte points
    e2e_kind = [args.e2e_input]

    def upload(f):
        if e2e_kind[0] == "xyz":
            gpu.set_scans_xyz_host_ptr(host_xyz.data_ptr() + f * (frame_bytes // 4 * 3), counts[f], stride)
        else:
            gpu.set_scans_host_ptr(host.data_ptr() + f * frame_bytes, counts[f], stride)

    pose_host = torch.empty((2, 3, B, 6), dtype=torch.float32).pin_memory()  # two steps in flight x (sum, cur, map)

    def step_host(f):
        """Scan f was staged by upload(f).  Enqueue its processing and the read-back of its poses, stage scan f+1 (its
        H2D copy overlaps the kernels of scan f: the library double-buffers the input), then collect the poses of scan
        f-1: a consumer one scan behind, like the reference's stage threads behind their Channels."""
        gpu.process_scans()
        slot = pose_host[f & 1]
        gpu.poses_async(slot[0].data_ptr(), slot[1].data_ptr(), slot[2].data_ptr())
        step_host.in_flight += 1
        if f + 1 < n_frames:
            upload(f + 1)
        if step_host.in_flight == 2:
            gpu.wait_poses()          # the poses of scan f-1 are on the host now
            step_host.in_flight -= 1
        return pose_host[(f - 1) & 1]

    step_host.in_flight = 0

    def drain_poses():
        while step_host.in_flight:
            gpu.wait_poses()
            step_host.in_flight -= 1

    uuid = str(torch.cuda.get_device_properties(dev).uuid)
    sampler = ClockSampler(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)

    # ---- warm-up: frame 0 initialises, then W untimed steps ----
    f = 0
    step_device(f); f += 1
    for _ in range(args.warmup):
        step_device(f); f += 1
    # ---- profiling pass (untimed): every kernel bracketed by events, to pick the dominant kernel ----
    gpu.time_kernel("*")
    launches0 = gpu.kernel_launches()
    for _ in range(prof_steps):
        step_device(f); f += 1
    table = gpu.kernel_time_table()
    launches_per_step = (gpu.kernel_launches() - launches0) / prof_steps
    dominant = args.time_kernel or max(table, key=lambda k: table[k][0])
    gpu.time_kernel(dominant)
    torch.cuda.synchronize(dev)
    # ---- timed region: K steps, inputs resident in HBM ----
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches1 = gpu.kernel_launches()
    t_begin = time.time()
    with torch.cuda.stream(stream):
        ev0.record(stream)
        fork(ev0)
        for _ in range(args.steps):
            step_device(f); f += 1
        join()
        ev1.record(stream)
    torch.cuda.synchronize(dev)
    t_end = time.time()
    if world > 1:
        dist.barrier()
    dev_ms = ev0.elapsed_time(ev1)
    gpu_launches = gpu.kernel_launches() - launches1
    k_ms, k_n = gpu.kernel_time()
    gpu.time_kernel("")
    # workload statistics for the algorithmic byte count (last processed frame, averaged over sequences)
    stat_keys = {"S": "SEG_CLOUD", "less_flat": "SURF_LAST", "less_sharp": "CORNER_LAST", "flat": "SURF_FLAT",
                 "sharp": "CORNER_SHARP", "outlier": "OUTLIER_LAST", "q_surf": "SCAN_SURF_TOTAL_DS",
                 "q_corner": "SCAN_CORNER_DS", "map_surf": "MAP_SURF", "map_corner": "MAP_CORNER"}
    st = {"N": N, "g": (params.ground_scan_index + 1) / params.num_vertical_scans,
          "n_in": float(np.mean(counts[f - 1]))}
    sample_seqs = range(0, B, max(1, B // 4))
    for key, buf in stat_keys.items():
        st[key] = float(np.mean([len(gpu.download(buf, k)) for k in sample_seqs]))
    odom_iters = np.mean([gpu.download("ODOM_ITERS", k) for k in sample_seqs], axis=0)
    map_iters = np.mean([gpu.download("MAP_ITERS", k) for k in sample_seqs], axis=0) if use_map else [0, 0]

    # ---- per-kernel pass: the sub-batches one after the other, so that every launch runs alone on the GPU ----
    gpu.time_kernel("*")
    psf_steps = 3
    f_psf = f - psf_steps  # re-run the last frames (results are not used)
    gpu.reset()
    if use_map:
        seed_map_poses()
    parts = getattr(gpu, "parts", [gpu])
    sub = B // len(parts)
    for ff in range(max(0, f_psf - 2), f_psf + psf_steps):
        if ff == f_psf:
            gpu.time_kernel("*")
        for i, part in enumerate(parts):
            part.set_scans_device(devdata.data_ptr() + ff * frame_bytes + i * sub * stride * 16, counts[ff][i * sub:(i + 1) * sub], stride)
            part.process_scans()
            part.synchronize()
    alone = gpu.kernel_time_table()
    gpu.time_kernel("")
    latency = None
    if rank == 0:
        aft0 = None
        if use_map:
            x, y, z, roll, pitch, yaw = synth.pose(cfg, seq_ids[0], 0)
            aft0 = np.array([[0, yaw, 0, y, z, x]], np.float32)
        latency = single_sequence_latency(params, devdata, counts, stride, n_frames, B, dev, torch,
                                          local_maps(cfg, seq_ids[0]) if use_map else None, aft0, live=live)

    # ---- end-to-end: same C ABI, pinned host scans, H2D + pose D2H inside the timed region ----
    def run_e2e(kind):
        e2e_kind[0] = kind
        gpu.reset()
        if use_map:
            seed_map_poses()
        f = 0
        upload(f)
        step_host(f); f += 1
        for _ in range(min(args.warmup, 3)):
            step_host(f); f += 1
        drain_poses()
        steps = min(args.steps, n_frames - f - 1)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h2d = 0
        with torch.cuda.stream(stream):
            e0.record(stream)
            fork(e0)
            for _ in range(steps):
                h2d += int(counts[f].sum()) * (12 if kind == "xyz" else 16) + B * 4
                step_host(f); f += 1
            drain_poses()
            join()
            e1.record(stream)
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1), steps, h2d

    alt_kind = "xyzi" if args.e2e_input == "xyz" else "xyz"
    alt_ms, alt_steps, alt_h2d = run_e2e(alt_kind)
    e2e_ms, e2e_steps, h2d = run_e2e(args.e2e_input)
    clocks = sampler.stop(t_begin, t_end)
    if clocks.get("samples", 0) == 0:
        clocks = sampler.stop()  # timed region shorter than one sample: report the whole run

    # ---- max over ranks ----
    times = torch.tensor([dev_ms, e2e_ms, alt_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    value, dev_ms = aggregate_throughput(world, B, args.steps, float(times[0]))
    e2e_value, e2e_ms = aggregate_throughput(world, B, e2e_steps, float(times[1]))
    alt_value, alt_ms = aggregate_throughput(world, B, alt_steps, float(times[2]))

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        per_launch_s = (k_ms * 1e-3 / k_n) if k_n else float("nan")
        alg = algorithmic_bytes(dominant, st) * sub   # one launch covers one sub-batch
        traffic = None
        try:  # dram bytes per launch from the committed ncu --set full capture (profiles/r1_traffic.json)
            tr = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
            if dominant in tr:
                traffic = tr[dominant]["dram_bytes_per_sequence"] * sub
        except Exception:
            pass
        achieved = alg / per_launch_s / 1e9 if k_n else 0.0
        shares = {k: round(v[0] / sum(x[0] for x in table.values()), 4) for k, v in sorted(table.items(), key=lambda kv: -kv[1][0])}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args, params, B, "gpu"),
            "p50_scan_latency_ms": latency["p50_ms"] if latency else None, "latency_single_sequence": latency,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d // max(1, e2e_steps),
                    "d2h_bytes_per_step": B * 6 * 4 * 3, "steps": e2e_steps, "ms_per_step": e2e_ms / max(1, e2e_steps),
                    "host_points": args.e2e_input},
            "e2e_alt": {"value": alt_value, "unit": UNIT, "h2d_bytes_per_step": alt_h2d // max(1, alt_steps),
                        "d2h_bytes_per_step": B * 6 * 4 * 3, "steps": alt_steps, "ms_per_step": alt_ms / max(1, alt_steps),
                        "host_points": alt_kind},
            "gpu_launches": int(gpu_launches),
            "roofline": {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak if peak else None, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg, "avg_launch_us": per_launch_s * 1e6, "launches_timed": k_n,
                         "kernel_time_share_profiling_pass": shares},
            "kernel_rooflines": kernel_rooflines(alone, st, sub, peak),
            "clocks": clocks,
            "stats": {**{k: round(v, 1) for k, v in st.items()}, "odom_iters": [float(x) for x in odom_iters],
                      "map_iters_rows": [float(x) for x in map_iters], "launches_per_step": launches_per_step,
                      "dataset_gen_s": round(gen_s, 1)},
        }
        if not args.skip_cpu_baseline:
            cores = max(1, min(os.cpu_count() or 1, 64))
            v, wall, n_scans, stage_s, kind = cpu_pipeline(params, args.config, cores, args.cpu_frames, use_map, live=live)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "knn": kind,
                                    "sample": f"{cores} sequences x {args.cpu_frames - 1} scans of the same workload on "
                                              f"{cores} host threads ({wall:.1f} s wall)",
                                    "stage_seconds_sum": {"image_projection": stage_s[0], "feature_extraction": stage_s[1],
                                                          "scan_to_scan": stage_s[2], "scan_to_map": stage_s[3],
                                                          "downsample": stage_s[4], "map_assembly_keyframes": stage_s[5]}}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
