#!/usr/bin/env python
"""bench.py -- scans/sec of the LeGO-LOAM-BOR per-scan hot path at 64x2048 on B200.

  python bench.py --gpus N --steps K --warmup W                      # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K --warmup W     # the reference's CPU algorithm (oracle)

Workload (BASELINE.json configs[3]/[4], SURVEY.md section 8d way 1), `--map kf500` (default): 64-beam 64x2048
synthetic lidar in the 120 m arena world; every sequence starts with a 500-key-frame local map (key frames made from
scans taken at 500 poses >= 2.4 m apart on a spiral, stored through ll_map_save_keyframe / the oracle's
saveKeyFramesAndFactor), then drives a circle through it.  `--batch` (default 64) independent sequences per GPU advance
in lock step.  A "step" is one scan of every sequence through the whole hot path: projection, ground removal,
segmentation, feature extraction, scan-to-scan LM, and on every `mapping_frequency_divider`-th scan one body of
MapOptimization::run (extractSurroundingKeyFrames over the 500+ key frames, downsampleCurrentScan,
scan2MapOptimization against the ~270 k-point local map, saveKeyFramesAndFactor).
Other workloads: `--map live` (standard 60x40 m room, the map grows from nothing), `--map synthetic` (round-1 headline:
fixed synthetic lattice map), `--map none`.

Sequences are independent: N GPUs run N independent shards, no data-path collective ("replicas only").
`--total-seqs T` splits T sequences over the ranks (strong scaling, SURVEY 8e: 64/32/16/8 per GPU), otherwise every
rank owns `--batch` sequences (weak scaling).

One JSON line is printed by rank 0 (keys: see the task contract).  `value`: inputs already resident in HBM.
`e2e`: the same C ABI fed from pinned HOST memory, H2D of every scan and the pose read-back (one scan behind) inside
the timed region; `--e2e-input pc2` (default) hands every scan over as the sensor_msgs/PointCloud2 `data` bytes that
ImageProjection::cloudHandler receives (point_step 16, NaN check on the device), `xyzi` as 16-byte points, `xyz` as
packed 12-byte points; the other two are measured too (`e2e_alt`).
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


def parse_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="sequences per GPU (weak scaling)")
    ap.add_argument("--total-seqs", type=int, default=0, help="split this many sequences over the GPUs (strong scaling)")
    ap.add_argument("--config", default="C", help="sensor config: A 16x1800, B 32x1800, C 64x2048")
    ap.add_argument("--map", default="kf500", choices=["kf500", "live", "synthetic", "none"])
    ap.add_argument("--keyframes", type=int, default=500, help="key frames of the pre-built local map (--map kf500)")
    ap.add_argument("--no-map", action="store_true", help="same as --map none")
    ap.add_argument("--cpu-steps", type=int, default=10, help="timed scans per sequence of the cpu_baseline sample")
    ap.add_argument("--cpu-threads", type=int, default=0, help="host threads of the CPU arms (0 = all usable cores)")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-latency", action="store_true")
    ap.add_argument("--time-kernel", default="", help="kernel to report in `roofline` (default: the largest time share)")
    ap.add_argument("--e2e-input", default="pc2", choices=["pc2", "xyz", "xyzi"])
    ap.add_argument("--e2e-alt-steps", type=int, default=10)
    ap.add_argument("--streams", type=int, default=4, help="split the batch over this many handles / CUDA streams")
    ap.add_argument("--ncu-range", action="store_true",
                    help="cudaProfilerStart/Stop around the timed `value` region, for `ncu --profile-from-start off` launch lists "
                         "(a number printed by a run under ncu is never a bench value)")
    a = ap.parse_args(argv)
    if a.no_map:
        a.map = "none"
    return a


# --------------------------------------------------------------------------------------------
# multi-GPU: independent sequences are sharded over ranks, no data-path collective (replicas only)


def shard_sequences(rank, batch, unique=0):
    """Sequence ids of the `batch` slots of `rank` (weak scaling: every rank owns its own sequences)."""
    u = unique or batch
    return [rank * batch + (k % u) for k in range(batch)]


def shard_total(rank, world, total):
    """Strong scaling: sequences 0..total-1 dealt out in contiguous shares, the first ranks one more
    when it does not divide."""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return list(range(lo, lo + base + (1 if rank < extra else 0)))


def aggregate_throughput(world, batch, steps, local_ms, reduce_max=None):
    """Whole-job scans/s: all ranks' scans divided by the slowest rank's device time.
    reduce_max: callable mapping a local float to the max over ranks (identity for one rank).
    `batch` is the per-rank count (weak) -- for a strong split pass total / world as a float."""
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
        if self.proc.poll() is None:
            self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            if t_begin is not None and not (t_begin - 0.15 <= ts <= t_end + 0.15):
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])); smax.append(float(f[1]))
            except Exception:
                continue
            for i, n in enumerate(names):
                if f[3 + i].lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(smax)) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------
# algorithmic bytes


def survey_bytes(st):
    """SURVEY 8(d) ALGORITHMIC bytes per scan of the projection / ground / segmentation / feature stages, from the
    measured n_in and S (the figure the north_star's 60 % target is judged against)."""
    N, S, n_in, g = st["N"], st["S"], st["n_in"], st["g"]
    P = 16 * n_in + 4 * N + 16 * N
    G = 16 * g * N + 4 * N + 1 * N + 4 * N
    Sg = 4 * N + 8 * N + 5 * N + S * 20 + S * 25
    F = S * 32 + S * 9 + S * 12 + S * 16 + S * 16
    return {"P": P, "G": G, "S": Sg, "F": F, "total": P + G + Sg + F}


KNN_REC_BYTES = (8 + 1) * 16   # LL_KNN_K + 1 float4 per query (ll_device.cuh)


def algorithmic_bytes(kernel, st):
    """Per-kernel, per-sequence bytes (each kernel's own compulsory reads + writes)."""
    N, S, n_in, g = st["N"], st["S"], st["n_in"], st["g"]
    lf, ls, ff, fs = st["less_flat"], st["less_sharp"], st["flat"], st["sharp"]
    qs, qc, ms, mc = st["q_surf"], st["q_corner"], st["map_surf"], st["map_corner"]
    out = st["outlier"]
    kfp = st.get("kf_points", qs + qc)
    table = {
        "k_project_scatter": 16 * n_in + 8 * n_in,
        "k_gather_ground": 29 * N + 16 * g * N,
        "k_ccl_rows": 12 * N, "k_ccl_merge": 8 * N, "k_ccl_flatten": 8 * N,
        "k_ccl_tile": 4 * N + 4 * N, "k_ccl_border": 8 * N / 8,
        "k_seg_count": 5 * N + N, "k_seg_emit": N + 45 * S + 4 * S, "k_label_final": 8 * N,
        "k_feature_prep": 61 * S + 4 * S,
        "k_feature_sort": 21 * S, "k_feature_pick": 8 * S, "k_feature_lessflat": 24 * S + 16 * lf,
        "k_feature_ring": 21 * S + 8 * S + 24 * S + 16 * lf,
        "k_feature_compact": 32 * lf + 20 * (ls + ff + fs),
        "k_odom_search_surf": ff * 96 + lf * 16, "k_odom_search_corner": fs * 80 + ls * 16,
        "k_odom_stage_surf": ff * 69, "k_odom_stage_corner": fs * 85,
        "k_publish_clouds_last": 32 * (lf + ls + out),
        "k_grid_count": 16 * (lf + ls), "k_grid_scan": 0, "k_grid_fill": 32 * (lf + ls),
        "k_voxel_grid": 16 * (lf + ls + out) * 2, "k_voxel_grid_total": 16 * qs * 2,
        "k_map_knn": (qs + qc) * (16 + KNN_REC_BYTES) + (ms + mc) * 16,
        "k_map_iter": (qs + qc) * (16 + 4 + 5 * 16), "k_map_solve": 0,
        "k_map_knn_reuse": (qs + qc) * (16 + KNN_REC_BYTES + 4), "k_map_iter_cached": (qs + qc) * (16 + 4 + 32),
        "k_kf_select": 0, "k_kf_decide": 0, "k_kf_accumulate": kfp * (16 + 8 + 2 * 28),
        "k_kf_extract": (ms + mc) * (8 + 20 + 16), "k_kf_store": kfp * (16 + 28),
        "k_kfx_sort_new": 0, "k_kfx_merge": (ms + mc) * 24, "k_kfx_alive": (ms + mc) * 8, "k_kfx_output": (ms + mc) * (4 + 20 + 16),
    }
    return float(table.get(kernel, 0))


PSF_KERNELS = ["k_project_scatter", "k_gather_ground", "k_ccl_rows", "k_ccl_merge", "k_ccl_flatten", "k_ccl_tile", "k_ccl_border",
               "k_seg_count", "k_seg_emit", "k_label_final", "k_feature_prep", "k_feature_sort", "k_feature_pick",
               "k_feature_lessflat", "k_feature_ring", "k_feature_compact"]
PSF_STAGE = {"k_project_scatter": "P", "k_gather_ground": "G", "k_ccl_rows": "S", "k_ccl_merge": "S", "k_ccl_flatten": "S",
             "k_ccl_tile": "S", "k_ccl_border": "S", "k_seg_count": "S", "k_seg_emit": "S", "k_label_final": "S",
             "k_feature_prep": "F", "k_feature_sort": "F", "k_feature_pick": "F", "k_feature_lessflat": "F",
             "k_feature_ring": "F", "k_feature_compact": "F"}


def kernel_rooflines(alone, st, sub, peak, frames, parts):
    """Every kernel timed ALONE on the GPU (one sub-batch of `sub` sequences per launch): average launch duration,
    achieved algorithmic GB/s and fraction of the measured HBM peak.  `psf_*`: the projection / segmentation /
    feature group: time-weighted mean with each kernel's own bytes (`psf_mean_frac`) and with SURVEY 8(d)'s bytes per
    scan over the group's whole time per scan (`psf_survey_frac`, per stage in `psf_survey_stage_frac`)."""
    out, t_psf, b_psf = {}, 0.0, 0.0
    stage_us = {"P": 0.0, "G": 0.0, "S": 0.0, "F": 0.0}
    for k, (ms, n) in sorted(alone.items(), key=lambda kv: -kv[1][0]):
        if n == 0:
            continue
        us = 1e3 * ms / n
        alg = algorithmic_bytes(k, st) * sub
        gbs = alg / (us * 1e-6) / 1e9 if us > 0 else 0.0
        out[k] = {"avg_us": round(us, 1), "launches": n, "alg_MB": round(alg / 1e6, 2), "GBps": round(gbs, 1),
                  "frac": round(gbs / peak, 4) if peak else None}
        if k in PSF_KERNELS:
            per_frame = n / (frames * parts)    # launches of this kernel per frame of one sub-batch
            t_psf += us * per_frame             # microseconds per frame of one sub-batch (`sub` scans)
            b_psf += alg * per_frame
            stage_us[PSF_STAGE[k]] += us * per_frame
    sb = survey_bytes(st)
    out["psf_mean_frac"] = round(b_psf / (t_psf * 1e-6) / 1e9 / peak, 4) if t_psf > 0 and peak else None
    out["psf_survey_frac"] = round(sb["total"] * sub / (t_psf * 1e-6) / 1e9 / peak, 4) if t_psf > 0 and peak else None
    out["psf_survey_stage_frac"] = {s: (round(sb[s] * sub / (stage_us[s] * 1e-6) / 1e9 / peak, 4) if stage_us[s] > 0 and peak else None)
                                    for s in ("P", "G", "S", "F")}
    out["psf_us_per_scan"] = round(t_psf / sub, 2) if sub else None
    out["survey_bytes_per_scan"] = {k: int(v) for k, v in sb.items()}
    out["sequences_per_launch"] = sub
    return out


# --------------------------------------------------------------------------------------------
# workloads


class Workload:
    """What the GPU arm and the CPU arm both run: which world, which frames, how the map side is set up."""

    def __init__(self, args, params):
        from lego_loam_bor_b200 import synth
        self.args, self.params, self.kind = args, params, args.map
        self.N = params.num_vertical_scans * params.num_horizontal_scans
        self.K = args.keyframes
        if self.kind == "kf500":
            self.cfg = synth.make_arena(params, n_keyframes=self.K)
        else:
            self.cfg = synth.make_config(params)

    # ---- scans ----
    def cpu_scan(self, seq, f):
        from lego_loam_bor_b200 import synth
        return synth.arena_scan(self.cfg, seq, synth.DRIVE, f) if self.kind == "kf500" else synth.scan(self.cfg, seq, f)

    def device_frames(self, seq_ids, n_frames, dev, torch):
        """float32 device tensor [F][B][N][4] (valid points packed at the front of every row) + counts [F][B]."""
        from lego_loam_bor_b200 import synth
        B = len(seq_ids)
        data = torch.zeros((n_frames, B, self.N, 4), dtype=torch.float32, device=dev)
        counts = np.zeros((n_frames, B), np.int32)
        t0 = time.time()
        if self.kind == "kf500":
            gen = self.generator(seq_ids, dev)
            for f in range(n_frames):
                _, counts[f] = gen.scans(synth.DRIVE, f, out=data[f])
            torch.cuda.synchronize(dev)
        else:
            uniq = sorted(set(seq_ids))
            scans = synth.scans(self.cfg, uniq, range(n_frames), threads=usable_cores())
            host = torch.zeros((B, self.N, 4), dtype=torch.float32).pin_memory()
            hv = host.numpy()
            for f in range(n_frames):
                for k, s in enumerate(seq_ids):
                    a = scans[(s, f)]
                    hv[k, :len(a)] = a
                    counts[f, k] = len(a)
                data[f].copy_(host)
        return data, counts, time.time() - t0

    def generator(self, seq_ids, dev):
        from lego_loam_bor_b200 import synth
        key = (tuple(seq_ids), str(dev))
        if getattr(self, "_gen_key", None) != key:
            self._gen = synth.ArenaDeviceGenerator(self.cfg, seq_ids, dev)
            self._gen_key = key
        return self._gen

    # ---- map side, GPU ----
    def setup_gpu(self, gpu, seq_ids, dev, torch):
        from lego_loam_bor_b200 import synth, workloads
        B = len(seq_ids)
        info = {}
        if self.kind == "kf500":
            gpu.map_enable_keyframes(*workloads.keyframe_capacities(self.params, self.K, extra_keyframes=96))
            gen = self.generator(seq_ids, dev)
            buf = torch.zeros((B, self.N, 4), dtype=torch.float32, device=dev)

            def scans_of(i):
                _, counts = gen.scans(synth.KEYFRAME, i, out=buf)
                return buf.data_ptr(), counts, self.N

            t0 = time.time()
            workloads.prebuild_keyframes(gpu, self.cfg, seq_ids, self.K, scans_of, sync=lambda: torch.cuda.synchronize(dev))
            info["prebuild_s"] = round(time.time() - t0, 2)
            workloads.start_drive(gpu, self.cfg, seq_ids)
        elif self.kind == "live":
            n_kf = 64
            gpu.map_enable_keyframes(max_keyframes=n_kf, pool_points=n_kf * (self.N // 8), max_map_corner=self.N // 2, max_map_surf=self.N)
        elif self.kind == "synthetic":
            for k, s in enumerate(seq_ids):
                gpu.map_set_local(k, synth.local_map(self.cfg, s, 1, 0.2), synth.local_map(self.cfg, s, 0, 0.4))
            aft = np.zeros((B, 6), np.float32)
            for k, s in enumerate(seq_ids):
                x, y, z, roll, pitch, yaw = synth.pose(self.cfg, s, 0)
                aft[k] = [0, yaw, 0, y, z, x]
            gpu.map_set_poses(aft, np.zeros((B, 6), np.float32))
        return info

    # ---- map side, CPU oracle (one sequence) ----
    def setup_oracle(self, o, seq):
        from lego_loam_bor_b200 import synth
        if self.kind == "kf500":
            zero = np.zeros(6, np.float32)
            for i in range(self.K):
                o.reset_feature_association()
                o.image_projection(synth.arena_scan(self.cfg, seq, synth.KEYFRAME, i))
                o.feature_association()
                o.map_downsample_current_scan()
                T = synth.pose_to_transform(synth.arena_pose(self.cfg, seq, synth.KEYFRAME, i))
                o.map_set_poses(T, zero)
                o.map_set_initial_guess(T)
                o.map_save_keyframe()
            o.reset_feature_association()
            o.map_set_poses(synth.pose_to_transform(synth.arena_pose(self.cfg, seq, synth.DRIVE, 0)), zero)
        elif self.kind == "synthetic":
            o.map_set_local(synth.local_map(self.cfg, seq, 1, 0.2), synth.local_map(self.cfg, seq, 0, 0.4))
            x, y, z, roll, pitch, yaw = synth.pose(self.cfg, seq, 0)
            o.map_set_poses(np.array([0, yaw, 0, y, z, x], np.float32), np.zeros(6, np.float32))

    def oracle_step(self, o, scan):
        """One scan through the oracle; returns per-stage seconds (ip, fa, mapping) of this scan."""
        t0 = time.perf_counter()
        o.image_projection(scan)
        t1 = time.perf_counter()
        handed = o.feature_association()
        t2 = time.perf_counter()
        if handed == 1 and self.kind != "none":
            if self.kind in ("kf500", "live"):
                o.mapping_cycle()
            else:
                o.map_downsample_current_scan()
                o.map_predict_pose()
                o.scan_to_map()
        t3 = time.perf_counter()
        return t1 - t0, t2 - t1, t3 - t2, handed == 1

    def describe(self, batch, where, args):
        p = self.params
        maps = {"kf500": f", whole mapping cycle every {p.mapping_frequency_divider}th scan against a local map assembled from {self.K}+ key frames "
                         "(extractSurroundingKeyFrames + downsampleCurrentScan + scan2MapOptimization + saveKeyFramesAndFactor)",
                "live": f", whole mapping cycle every {p.mapping_frequency_divider}th scan, the map grows from the first key frame",
                "synthetic": f", scan-to-map every {p.mapping_frequency_divider}th scan vs a fixed synthetic lattice map",
                "none": ""}[self.kind]
        world = "120 m arena, 500-key-frame local map (SURVEY 8d way 1)" if self.kind == "kf500" else "60x40 m room"
        return {"workload": f"{p.num_vertical_scans}x{p.num_horizontal_scans} synthetic lidar, {world}, {batch} independent sequences per "
                            f"{'GPU' if where == 'gpu' else 'host'}, full hot path per scan (projection, ground, segmentation, features, "
                            f"scan-to-scan LM{maps})",
                "map": self.kind, "keyframes": self.K if self.kind == "kf500" else None, "sensor": args.config,
                "batch_per_gpu": batch, "streams_per_gpu": args.streams,
                "parallelism": f"replicas x{args.gpus} (independent sequences, no collective)",
                "l2": "every step reads a distinct set of scans (inputs per step ~ L2 size, dataset >> L2); no reuse between steps"}


def usable_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


# --------------------------------------------------------------------------------------------
# CPU arm (oracle): the reference's algorithm on the host cores


def cpu_run(wl, n_threads, warmup, steps, pipeline_frames=12):
    """n_threads independent sequences, one per host thread (the ctypes calls release the GIL), no barrier between
    steps: every thread sets its sequence up (key frames, frame 0, `warmup` scans) and then times `steps` scans.
    Then the reference's OWN threading (main.cpp:37-47: ImageProjection / FeatureAssociation / MapOptimization on a thread
    each, blocking one-slot channels) on n_threads // 3 of those sequences at once, continuing where they are.
    Returns a dict with the aggregate scans/s and per-stage p50 / p95 milliseconds of both."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle_py

    def run(s):
        o = oracle_py.Oracle(wl.params, libm=True, nanoflann=True)
        wl.setup_oracle(o, s)
        scans = [wl.cpu_scan(s, f) for f in range(1 + warmup + steps)]
        for f in range(1 + warmup):
            wl.oracle_step(o, scans[f])
        rows = []
        t0 = time.perf_counter()
        for f in range(1 + warmup, 1 + warmup + steps):
            rows.append(wl.oracle_step(o, scans[f]))
        t1 = time.perf_counter()
        return t0, t1, rows, o

    t_all = time.time()
    with ThreadPoolExecutor(n_threads) as ex:
        res = list(ex.map(run, range(n_threads)))
    wall_all = time.time() - t_all
    # all threads run their timed part concurrently after a set-up of (nearly) equal length: the job's timed wall clock
    # is the span from the first start to the last end
    span = max(r[1] for r in res) - min(r[0] for r in res)
    per_thread = [steps / (r[1] - r[0]) for r in res]
    ip = np.array([x[0] for r in res for x in r[2]]) * 1e3
    fa = np.array([x[1] for r in res for x in r[2]]) * 1e3
    mo = np.array([x[2] for r in res for x in r[2] if x[3]]) * 1e3
    tot = np.array([x[0] + x[1] + x[2] for r in res for x in r[2]]) * 1e3

    def pc(a):
        return {"p50_ms": round(float(np.percentile(a, 50)), 2), "p95_ms": round(float(np.percentile(a, 95)), 2)} if len(a) else None

    out = {"value": float(sum(per_thread)), "value_span": n_threads * steps / span, "threads": n_threads, "steps": steps,
           "timed_wall_s": round(span, 2), "total_wall_s": round(wall_all, 1), "knn": oracle_py.kind(),
           "stage_ms": {"image_projection": pc(ip), "feature_association": pc(fa), "mapping_cycle": pc(mo), "scan_total": pc(tot)}}
    if pipeline_frames and wl.kind != "none" and wl.kind != "synthetic":
        n_pipe = max(1, n_threads // 3)
        f0 = 1 + warmup + steps
        skip = 2   # a fresh FeatureAssociation object: its first frame only initialises (featureAssociation.cpp:1414-1417)

        def pipe(s):
            o_mo = res[s][3]
            o_ip = oracle_py.Oracle(wl.params, libm=True, nanoflann=True)
            o_fa = oracle_py.Oracle(wl.params, libm=True, nanoflann=True)
            scans = [wl.cpu_scan(s, f0 + f) for f in range(pipeline_frames)]
            ms, wall = oracle_py.run_pipeline(o_ip, o_fa, o_mo, scans, skip)
            o_ip.close(); o_fa.close()
            return ms, wall

        with ThreadPoolExecutor(n_pipe) as ex:
            pres = list(ex.map(pipe, range(n_pipe)))
        timed = pipeline_frames - skip
        pms = np.concatenate([r[0][skip:] for r in pres])
        out["pipeline_3_threads"] = {
            "pipelines": n_pipe, "threads": 3 * n_pipe, "frames_timed": timed,
            "value": float(sum(timed / r[1] for r in pres)), "scans_per_s_one_pipeline": float(np.mean([timed / r[1] for r in pres])),
            "stage_ms": {"image_projection": pc(pms[:, 0]), "feature_association": pc(pms[:, 1]), "mapping_cycle": pc(pms[:, 2][pms[:, 2] > 0])},
            "note": "the reference's own threading (main.cpp:37-47): three stage threads per sequence, blocking one-slot channels"}
    for r in res:
        r[3].close()
    return out


def run_reference_arm(args, params):
    """--impl reference: the reference's own CPU algorithm (oracle restatement; k-NN by the reference's vendored
    nanoflann when it was compiled in) on all usable host cores, one sequence per thread, same workload."""
    wl = Workload(args, params)
    cores = args.cpu_threads or usable_cores()
    r = cpu_run(wl, cores, args.warmup, args.steps)
    value = r["value"]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * cores / value, "higher_is_better": True,
        "scaling": "strong" if args.total_seqs else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": wl.describe(cores, "cpu", args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "knn": r["knn"],
                         "sample": f"{cores} sequences x {args.steps} scans of the same workload, one thread per sequence, "
                                   f"no barrier between steps ({r['timed_wall_s']} s timed, {r['total_wall_s']} s with set-up)",
                         "stage_ms": r["stage_ms"], "value_first_start_to_last_end": r["value_span"],
                         "pipeline_3_threads": r.get("pipeline_3_threads")},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------


def single_sequence_latency(wl, params, seq, dev, torch, n_frames):
    """p50 / p95 device latency of one scan of ONE sequence (batch 1, one stream): projection..odometry of every
    scan, the mapping cycle on every 5th."""
    from lego_loam_bor_b200.capi import LegoLoam
    stream = torch.cuda.Stream(device=dev)
    one = LegoLoam(params, batch=1, max_points=wl.N, device=dev.index, stream=stream.cuda_stream)
    with torch.cuda.stream(stream):
        wl.setup_gpu(one, [seq], dev, torch)
        data, counts, _ = wl.device_frames([seq], n_frames, dev, torch)
    torch.cuda.synchronize(dev)
    evs, kinds = [], []
    with torch.cuda.stream(stream):
        for f in range(n_frames):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            one.set_scans_device(data[f].data_ptr(), counts[f], wl.N)
            e0.record(stream)
            rc = one.process_scans()
            e1.record(stream)
            evs.append((e0, e1)); kinds.append(rc)
    torch.cuda.synchronize(dev)
    ms = np.array([a.elapsed_time(b) for a, b in evs])
    skip = 1 + params.mapping_frequency_divider   # first frame only initialises; the first mapping cycle assembles the whole map
    ms, kinds = ms[skip + 1:], np.array(kinds[skip + 1:])
    one.close()
    out = {"batch": 1, "scans": int(len(ms)), "p50_ms": float(np.percentile(ms, 50)), "p95_ms": float(np.percentile(ms, 95)),
           "mean_ms": float(ms.mean()),
           "plain_scan_p50_ms": float(np.percentile(ms[kinds == 0], 50)) if np.any(kinds == 0) else None,
           "mapping_scan_p50_ms": float(np.percentile(ms[kinds == 1], 50)) if np.any(kinds == 1) else None,
           "note": "device time per scan of one sequence, one stream; every 5th scan also runs the mapping cycle"}
    return out


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from lego_loam_bor_b200 import config_params
    params = config_params(args.config)
    if args.impl == "reference":
        if rank != 0:
            return 0
        # only the oracle and the scan generator are built / loaded here: the CUDA library stays out of this process
        from lego_loam_bor_b200 import synth
        from oracle import oracle_py
        synth.build()
        oracle_py.build()
        run_reference_arm(args, params)
        return 0

    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    if rank == 0:
        ge.build()
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    from lego_loam_bor_b200.capi import LegoLoam, LegoLoamStreams

    wl = Workload(args, params)
    if args.total_seqs:
        seq_ids = shard_total(rank, world, args.total_seqs)
        total_seqs = args.total_seqs
    else:
        seq_ids = shard_sequences(rank, args.batch)
        total_seqs = args.batch * world
    B = len(seq_ids)
    n_streams = args.streams
    while n_streams > 1 and B % n_streams:
        n_streams -= 1
    N = wl.N
    stride = N
    use_map = wl.kind != "none"
    prof_steps, alone_steps, e2e_warm = 5, 5, 2
    # kf500: the first two mapping cycles sum the 500 pre-stored key frames into the voxel tables (the second one picks up
    # the ~80 the averaged-id rule of extractSurroundingKeyFrames leaves out of the first): extra untimed steps so that
    # they lie before the profiling pass whatever --warmup is
    settle = max(0, 2 * params.mapping_frequency_divider + 1 - args.warmup) if wl.kind == "kf500" else 0
    kinds = [k for k in ("pc2", "xyzi", "xyz") if k != args.e2e_input] + [args.e2e_input]   # the headline form runs last
    e2e_steps = {k: (args.steps if k == args.e2e_input else min(args.steps, args.e2e_alt_steps)) for k in kinds}
    f_e2e0 = 1 + args.warmup + settle + prof_steps + args.steps + alone_steps
    n_frames = f_e2e0 + sum(e2e_warm + e2e_steps[k] + 1 for k in kinds)
    devdata, counts, gen_s = wl.device_frames(seq_ids, n_frames, dev, torch)

    stream = torch.cuda.Stream(device=dev)
    sub_streams = [torch.cuda.Stream(device=dev) for _ in range(n_streams)] if n_streams > 1 else [stream]
    if n_streams > 1:
        gpu = LegoLoamStreams(params, B, n_streams, max_points=stride, device=local_rank, streams=[s.cuda_stream for s in sub_streams])
    else:
        gpu = LegoLoam(params, batch=B, max_points=stride, device=local_rank, stream=stream.cuda_stream)
    setup_info = wl.setup_gpu(gpu, seq_ids, dev, torch) if use_map else {}
    gpu.synchronize()
    frame_bytes = B * stride * 16

    def fork(ev):
        """timed regions are bracketed on `stream`; the per-handle streams start after ev and are joined before the end event"""
        if n_streams > 1:
            for ss in sub_streams:
                ss.wait_event(ev)

    def join():
        if n_streams > 1:
            for ss in sub_streams:
                stream.wait_stream(ss)

    def step_device(f):
        gpu.set_scans_device(devdata[f].data_ptr(), counts[f], stride)
        return gpu.process_scans()

    uuid = str(torch.cuda.get_device_properties(dev).uuid)
    sampler = ClockSampler(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)

    # ---- warm-up: frame 0 initialises, then W untimed steps (W >= 5 includes the first mapping cycle, which in the
    # kf500 workload sums all 500 key frames into the voxel tables) ----
    f = 0
    step_device(f); f += 1
    for _ in range(args.warmup + settle):
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
    step_ev = [[torch.cuda.Event(enable_timing=True) for _ in range(args.steps)] for _ in sub_streams]
    launches1 = gpu.kernel_launches()
    if args.ncu_range:
        torch.cuda.profiler.start()
    t_begin = time.time()
    mapping_steps = 0
    with torch.cuda.stream(stream):
        ev0.record(stream)
        fork(ev0)
        for i in range(args.steps):
            mapping_steps += 1 if step_device(f) == 1 else 0
            f += 1
            for si, ss in enumerate(sub_streams):
                step_ev[si][i].record(ss)
        gpu.join_mapping()   # the mapping cycles run on streams of their own: the region ends when they have finished too
        join()
        ev1.record(stream)
    torch.cuda.synchronize(dev)
    t_end = time.time()
    if args.ncu_range:
        torch.cuda.profiler.stop()
    if world > 1:
        dist.barrier()
    dev_ms = ev0.elapsed_time(ev1)
    gpu_launches = gpu.kernel_launches() - launches1
    k_ms, k_n = gpu.kernel_time()
    gpu.time_kernel("")
    # per-step completion times (max over the streams), their differences = the spread of the step time
    done = np.array([[ev0.elapsed_time(step_ev[si][i]) for i in range(args.steps)] for si in range(len(sub_streams))]).max(axis=0)
    step_ms = np.diff(np.concatenate([[0.0], done]))
    # workload statistics for the algorithmic byte count (last processed frame, averaged over sequences)
    stat_keys = {"S": "SEG_CLOUD", "less_flat": "SURF_LAST", "less_sharp": "CORNER_LAST", "flat": "SURF_FLAT",
                 "sharp": "CORNER_SHARP", "outlier": "OUTLIER_LAST", "q_surf": "SCAN_SURF_TOTAL_DS",
                 "q_corner": "SCAN_CORNER_DS", "map_surf": "MAP_SURF", "map_corner": "MAP_CORNER"}
    st = {"N": N, "g": (params.ground_scan_index + 1) / params.num_vertical_scans, "n_in": float(np.mean(counts[f - 1]))}
    sample_seqs = range(0, B, max(1, B // 4))
    for key, buf in stat_keys.items():
        st[key] = float(np.mean([len(gpu.download(buf, k)) for k in sample_seqs]))
    if wl.kind in ("kf500", "live"):
        st["kf_points"] = float(np.mean([len(gpu.download("SCAN_CORNER_DS", k)) + len(gpu.download("SCAN_SURF_DS", k)) +
                                         len(gpu.download("SCAN_OUTLIER_DS", k)) for k in sample_seqs]))
        kstate = np.array([gpu.download("KEYFRAME_STATE", k) for k in sample_seqs])
    odom_iters = np.mean([gpu.download("ODOM_ITERS", k) for k in sample_seqs], axis=0)
    map_iters = np.mean([gpu.download("MAP_ITERS", k) for k in sample_seqs], axis=0) if use_map else [0, 0]

    # ---- per-kernel pass: the sub-batches one after the other, so that every launch runs alone on the GPU ----
    gpu.time_kernel("*")
    parts = getattr(gpu, "parts", [gpu])
    sub = B // len(parts)
    for _ in range(alone_steps):
        for i, part in enumerate(parts):
            part.set_scans_device(devdata[f].data_ptr() + i * sub * stride * 16, counts[f][i * sub:(i + 1) * sub], stride)
            part.process_scans()
            part.synchronize()
        f += 1
    alone = gpu.kernel_time_table()
    gpu.time_kernel("")
    assert f == f_e2e0

    # ---- end-to-end: same C ABI, pinned host scans, H2D + pose D2H inside the timed region ----
    pose_host = torch.empty((2, 3, B, 6), dtype=torch.float32).pin_memory()  # two steps in flight x (sum, cur, map)

    def run_e2e(kind, f0):
        steps = e2e_steps[kind]
        nf = e2e_warm + steps + 1
        src = devdata[f0:f0 + nf]
        if kind == "xyz":
            host = src[..., :3].contiguous().cpu().pin_memory()
            fbytes = frame_bytes // 4 * 3
        else:
            host = src.cpu().pin_memory()
            fbytes = frame_bytes

        def upload(j):
            ptr = host.data_ptr() + j * fbytes
            if kind == "xyz":
                gpu.set_scans_xyz_host_ptr(ptr, counts[f0 + j], stride)
            elif kind == "xyzi":
                gpu.set_scans_host_ptr(ptr, counts[f0 + j], stride)
            else:   # the PointCloud2 `data` array of an unorganised cloud: width = n points, point_step 16, x y z intensity
                gpu.set_scans_pointcloud2_ptr(ptr, counts[f0 + j], stride * 16, 16, 0, 4, 8, 12, False)

        in_flight = [0]

        def step_host(j):
            """Scan j was staged by upload(j).  Enqueue its processing and the read-back of its poses, stage scan j+1
            (its H2D copy overlaps the kernels of scan j: the library double-buffers the input), then collect the poses
            of scan j-1: a consumer one scan behind, like the reference's stage threads behind their Channels."""
            gpu.process_scans()
            slot = pose_host[j & 1]
            gpu.poses_async(slot[0].data_ptr(), slot[1].data_ptr(), slot[2].data_ptr())
            in_flight[0] += 1
            upload(j + 1)
            if in_flight[0] == 2:
                gpu.wait_poses()
                in_flight[0] -= 1

        def drain():
            while in_flight[0]:
                gpu.wait_poses()
                in_flight[0] -= 1

        upload(0)
        for j in range(e2e_warm):
            step_host(j)
        drain()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h2d = 0
        with torch.cuda.stream(stream):
            e0.record(stream)
            fork(e0)
            for j in range(e2e_warm, e2e_warm + steps):
                # what the library copies: per handle one strided copy whose rows carry the longest scan of that sub-batch
                cj = counts[f0 + j].reshape(len(parts), -1)
                h2d += int((cj.max(axis=1) * cj.shape[1]).sum()) * (12 if kind == "xyz" else 16) + B * 4
                step_host(j)
            drain()
            gpu.join_mapping()
            join()
            e1.record(stream)
        torch.cuda.synchronize(dev)
        # scan j+1 of the last step was staged but never processed: consume it so that the frame sequence stays gap-free
        gpu.process_scans()
        gpu.synchronize()
        del host
        return e0.elapsed_time(e1), steps, h2d, f0 + nf

    e2e_res = {}
    for kind in kinds:
        ms, steps, h2d, f = run_e2e(kind, f)
        e2e_res[kind] = (ms, steps, h2d)
    clocks = sampler.stop(t_begin, t_end)
    if clocks.get("samples", 0) == 0:
        clocks = sampler.stop()  # timed region shorter than one sample: report the whole run

    latency = None
    if rank == 0 and not args.skip_latency:
        latency = single_sequence_latency(wl, params, seq_ids[0], dev, torch, 1 + 8 * params.mapping_frequency_divider)

    # ---- max over ranks ----
    times = torch.tensor([dev_ms] + [e2e_res[k][0] for k in kinds], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    per_rank = total_seqs / world
    value, dev_ms = aggregate_throughput(world, per_rank, args.steps, float(times[0]))
    e2e_out = {}
    for i, k in enumerate(kinds):
        v, ms = aggregate_throughput(world, per_rank, e2e_res[k][1], float(times[1 + i]))
        e2e_out[k] = {"value": v, "unit": UNIT, "h2d_bytes_per_step": e2e_res[k][2] // max(1, e2e_res[k][1]),
                      "d2h_bytes_per_step": B * 6 * 4 * 3, "steps": e2e_res[k][1], "ms_per_step": ms / max(1, e2e_res[k][1]),
                      "host_points": {"pc2": "sensor_msgs/PointCloud2 data bytes, point_step 16 (ll_set_scans_pointcloud2_host)",
                                      "xyzi": "16-byte xyzi points (ll_set_scans_host)", "xyz": "packed 12-byte xyz points (ll_set_scans_xyz_host)"}[k]}

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
        try:  # dram bytes per launch from the committed ncu --set full captures (profiles/*traffic.json)
            for name in ("r2_final_traffic.json", "r2_traffic.json", "r1_traffic.json"):
                pth = os.path.join(ROOT, "profiles", name)
                if os.path.exists(pth):
                    tr = json.load(open(pth))
                    if dominant in tr:
                        traffic = tr[dominant]["dram_bytes_per_sequence"] * sub
                        break
        except Exception:
            pass
        achieved = alg / per_launch_s / 1e9 if k_n else 0.0
        shares = {k: round(v[0] / sum(x[0] for x in table.values()), 4) for k, v in sorted(table.items(), key=lambda kv: -kv[1][0])}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong" if args.total_seqs else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": wl.describe(B, "gpu", args),
            "p50_scan_latency_ms": latency["p50_ms"] if latency else None, "latency_single_sequence": latency,
            "step_ms_spread": {"p50": round(float(np.percentile(step_ms, 50)), 3), "p95": round(float(np.percentile(step_ms, 95)), 3),
                               "min": round(float(step_ms.min()), 3), "max": round(float(step_ms.max()), 3),
                               "mapping_steps": mapping_steps,
                               "note": "device time between the completions of consecutive steps (max over the streams)"},
            "e2e": e2e_out[args.e2e_input],
            "e2e_alt": {k: v for k, v in e2e_out.items() if k != args.e2e_input},
            "gpu_launches": int(gpu_launches),
            "roofline": {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak if peak else None, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg, "avg_launch_us": per_launch_s * 1e6, "launches_timed": k_n,
                         "kernel_time_share_profiling_pass": shares},
            "kernel_rooflines": kernel_rooflines(alone, st, sub, peak, alone_steps, len(parts)),
            "clocks": clocks,
            "stats": {**{k: round(v, 1) for k, v in st.items()}, "odom_iters": [float(x) for x in odom_iters],
                      "map_iters_rows": [float(x) for x in map_iters], "launches_per_step": launches_per_step,
                      "dataset_gen_s": round(gen_s, 1), "settle_steps_untimed": settle, **setup_info,
                      **({"keyframe_state_mean": [float(x) for x in kstate.mean(axis=0)]} if wl.kind in ("kf500", "live") else {})},
        }
        if not args.skip_cpu_baseline:
            cores = args.cpu_threads or min(usable_cores(), 64)
            r = cpu_run(wl, cores, 1, args.cpu_steps)
            line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": cores, "kind": "port", "knn": r["knn"],
                                    "sample": f"{cores} sequences x {args.cpu_steps} scans of the same workload on {cores} host threads "
                                              f"({r['timed_wall_s']} s timed, {r['total_wall_s']} s with the set-up of the key frames)",
                                    "stage_ms": r["stage_ms"], "pipeline_3_threads": r.get("pipeline_3_threads")}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
