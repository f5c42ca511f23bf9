"""BASELINE.json configs[3]: 64x2048 with a 500-key-frame local map (way 1: pre-built key frames,
then the sub-map assembly + scan2MapOptimization called directly).

  python tools/map500.py [batch=4] [keyframes=500] [--cpu]

Every sequence first runs 7 frames of the normal pipeline (so that real down-sampled scan clouds exist), then
`keyframes` key frames are stored through ll_map_save_keyframe with forced poses on a 1.5 m lawn-mower grid (every pose
its own 1 m voxel, all within the 50 m search radius of the last one), each holding the clouds of the current scan.
Timed with CUDA events on the handle's stream:
  rebuild   ll_map_extract_surrounding_keyframes with every key frame pending (all per-voxel sums from scratch)
  append    the same call after one more key frame (only that key frame is summed in)
  scan2map  ll_scan_to_map against the assembled map (k-NN structure build + LM iterations)
With --cpu the oracle does the same for ONE sequence on one host core (std::stable_sort VoxelGrid over the
concatenated clouds + nanoflann builds), the reference's cost for the same mapping cycle.
Prints one JSON line."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def grid_pose(i, cols=23, step=1.5):
    r, c = divmod(i, cols)
    c = c if r % 2 == 0 else cols - 1 - c          # lawn-mower: consecutive poses are neighbours
    return np.array([0.0, 0.02 * i, 0.0, 0.3 + step * c, 0.0, 0.3 + step * r], np.float32)


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    B = int(args[0]) if len(args) > 0 else 4
    K = int(args[1]) if len(args) > 1 else 500
    with_cpu = "--cpu" in sys.argv
    import torch
    import bench
    from lego_loam_bor_b200 import config_params
    from lego_loam_bor_b200.capi import LegoLoam
    params = config_params("C")
    N = params.num_vertical_scans * params.num_horizontal_scans
    seqs = list(range(B))
    cfg, scans, counts, _, _ = bench.gen_dataset(params, seqs, 7)
    stream = torch.cuda.Stream()
    gpu = LegoLoam(params, batch=B, max_points=N, device=0, stream=stream.cuda_stream)
    gpu.map_enable_keyframes(max_keyframes=min(1024, K + 8), pool_points=(K + 8) * 16384, max_map_corner=1 << 20, max_map_surf=3 << 20)
    for f in range(7):
        gpu.set_scans_host([scans[(s, f)] for s in seqs])
        gpu.process_scans()
    gpu.synchronize()
    kf0 = int(gpu.download("KEYFRAME_STATE")[0])

    def save(i):
        aft = np.tile(grid_pose(i), (B, 1))
        gpu.map_set_poses(aft, np.zeros_like(aft))
        gpu.map_save_keyframe()

    def timed(fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            fn()
            e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    t_save = timed(lambda: [save(i) for i in range(K - kf0)])
    ms_rebuild = timed(gpu.map_extract_surrounding_keyframes)          # first extract: every key frame pending
    st = gpu.download("KEYFRAME_STATE")
    save(K - kf0)
    ms_append = timed(gpu.map_extract_surrounding_keyframes)
    st2 = gpu.download("KEYFRAME_STATE")
    n_corner, n_surf = len(gpu.download("MAP_CORNER")), len(gpu.download("MAP_SURF"))
    # scan-to-map of the current scan against that map, starting 5 cm / 0.01 rad off the last key pose
    guess = np.tile(grid_pose(K - kf0) + np.array([0.002, 0.01, -0.002, 0.05, 0.02, -0.05], np.float32), (B, 1))
    gpu.map_set_initial_guess(guess)
    ms_s2m = timed(gpu.scan_to_map)
    iters = gpu.download("MAP_ITERS")
    pts_per_kf = (len(gpu.download("SCAN_CORNER_DS")), len(gpu.download("SCAN_SURF_DS")), len(gpu.download("SCAN_OUTLIER_DS")))
    line = {"workload": f"64x2048, {B} sequences, local map assembled from {int(st2[0])} key frames on the device",
            "key_frames": int(st2[0]), "surrounding_key_frames": int(st2[1]), "capacity_error_bits": int(st2[3]),
            "points_per_key_frame": pts_per_kf, "concatenated_points_per_sequence": int(sum(pts_per_kf)) * int(st2[1]),
            "map_corner_points": n_corner, "map_surf_points": n_surf,
            "gpu_ms": {"save_keyframe_avg": t_save / max(1, K - kf0), "extract_rebuild_all": ms_rebuild, "extract_append_one": ms_append,
                       "scan_to_map": ms_s2m},
            "gpu_ms_per_sequence": {"extract_rebuild_all": ms_rebuild / B, "extract_append_one": ms_append / B, "scan_to_map": ms_s2m / B},
            "scan_to_map_iters_rows": [int(x) for x in iters], "state_after_rebuild": [int(x) for x in st]}
    if with_cpu:
        from oracle.oracle_py import Oracle
        o = Oracle(params, libm=False, nanoflann=True)  # portable math: bit-comparable with the device
        for f in range(7):
            o.image_projection(scans[(0, f)])
            if o.feature_association() == 1:
                o.mapping_cycle()
        k0 = int(o.download("KEYFRAME_STATE")[0])
        for i in range(K - k0):
            o.map_set_poses(grid_pose(i), np.zeros(6, np.float32))
            o.map_save_keyframe()
        t0 = time.time(); o.map_extract_surrounding_keyframes(); t_ext = time.time() - t0
        o.map_set_poses(grid_pose(K - k0), np.zeros(6, np.float32))
        o.map_save_keyframe()
        o.map_extract_surrounding_keyframes()   # same two-step history as the GPU side (the list order matters)
        o.map_set_initial_guess(guess[0])
        t0 = time.time(); o.scan_to_map(); t_s2m = time.time() - t0
        line["cpu_ms_one_core"] = {"extract": 1e3 * t_ext, "scan_to_map": 1e3 * t_s2m,
                                   "map_corner_points": len(o.download("MAP_CORNER")), "map_surf_points": len(o.download("MAP_SURF")),
                                   "scan_to_map_iters_rows": [int(x) for x in o.download("MAP_ITERS")]}
        line["parity"] = {"map_sizes_equal": line["cpu_ms_one_core"]["map_corner_points"] == n_corner and line["cpu_ms_one_core"]["map_surf_points"] == n_surf,
                          "map_surf_bits_equal": bool(np.array_equal(o.download("MAP_SURF"), gpu.download("MAP_SURF"))),
                          "map_corner_bits_equal": bool(np.array_equal(o.download("MAP_CORNER"), gpu.download("MAP_CORNER")))}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
