"""Workload set-up shared by bench.py, the tests and the tools: the 500-key-frame local map of BASELINE.json
configs[3]/[4] (way 1).

Key frame i of a sequence is what MapOptimization would have stored had the sensor stood at pose i of the arena's
spiral: the scan taken there goes through ImageProjection and a freshly constructed FeatureAssociation (first frame:
features extracted, clouds swapped to "last", no odometry; featureAssociation.cpp:1414-1417), downsampleCurrentScan,
and saveKeyFramesAndFactor with transformAftMapped = the true pose.  Only entry points of the C ABI are used; the
CPU arm does the same with the oracle (bench.py / tests)."""
import numpy as np

from . import synth


def keyframe_transforms(cfg, seq_ids, i):
    return np.stack([synth.pose_to_transform(synth.arena_pose(cfg, s, synth.KEYFRAME, i)) for s in seq_ids])


def drive_transforms(cfg, seq_ids, f):
    return np.stack([synth.pose_to_transform(synth.arena_pose(cfg, s, synth.DRIVE, f)) for s in seq_ids])


def keyframe_capacities(params, n_keyframes, extra_keyframes=64):
    """(max_keyframes, pool_points, max_map_corner, max_map_surf) for ll_map_enable_keyframes: a key frame of an
    NxN-cell scan holds ~0.16 N (64 beams) to ~0.22 N (16 beams) down-sampled points; the arena's local map of 500 key frames has
    ~1.0 N corner and ~1.2 N surf voxels at 64x2048 (a sparse 16-beam sensor spreads its few points over more voxels per
    point: the floors)."""
    N = params.num_vertical_scans * params.num_horizontal_scans
    kf = n_keyframes + extra_keyframes
    return kf, kf * (N // 4 + 512), max(16384, 2 * N), max(49152, 2 * N)


def prebuild_keyframes(gpu, cfg, seq_ids, n_keyframes, scans_of, sync=None):
    """Stores key frames 0..n_keyframes-1 of every sequence of `gpu` (a LegoLoam or LegoLoamStreams).
    scans_of(i) -> (device pointer of packed float32 [B][stride][4], counts int32 [B], stride) for key frame i;
    sync(): called before the scans are handed over (the generator's stream must have finished writing them)."""
    B = len(seq_ids)
    zero = np.zeros((B, 6), np.float32)
    for i in range(n_keyframes):
        ptr, counts, stride = scans_of(i)
        if sync:
            sync()
        gpu.reset_feature_association()
        gpu.set_scans_device(ptr, counts, stride)
        gpu.image_projection()
        gpu.feature_association()
        gpu.map_downsample_current_scan()
        T = keyframe_transforms(cfg, seq_ids, i)
        gpu.map_set_poses(T, zero)          # synchronises: the scan buffer may be rewritten after this
        gpu.map_set_initial_guess(T)        # the first key frame is stored from transformTobeMapped (mapOptmization.cpp:1362-1376)
        gpu.map_save_keyframe()
    gpu.synchronize()
    gpu.reset_feature_association()


def start_drive(gpu, cfg, seq_ids, frame=0):
    """transformAftMapped = true pose of the first driving frame in the map frame, transformBefMapped = the odometry
    origin: from then on transformAssociateToMap / transformUpdate chain odometry into the map frame on the device."""
    B = len(seq_ids)
    gpu.map_set_poses(drive_transforms(cfg, seq_ids, frame), np.zeros((B, 6), np.float32))
