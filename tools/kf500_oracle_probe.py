"""CPU probe of the kf500 workload (SURVEY 8d way 1) with the oracle only: map sizes, per-cycle cost, and that
scan-to-map against the key-frame map converges to the true pose.  python tools/kf500_oracle_probe.py [K=500] [config=C]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lego_loam_bor_b200 import config_params, synth
from oracle.oracle_py import Oracle

K = int(sys.argv[1]) if len(sys.argv) > 1 else 500
p = config_params(sys.argv[2] if len(sys.argv) > 2 else "C")
cfg = synth.make_arena(p, n_keyframes=K)
o = Oracle(p, libm=False, nanoflann=True)
t0 = time.time()
tg = 0.0
sizes = []
for i in range(K):
    t1 = time.time(); sc = synth.arena_scan(cfg, 0, synth.KEYFRAME, i); tg += time.time() - t1
    o.reset_feature_association()
    o.image_projection(sc)
    o.feature_association()
    o.map_downsample_current_scan()
    T = synth.pose_to_transform(synth.arena_pose(cfg, 0, synth.KEYFRAME, i))
    o.map_set_poses(T, np.zeros(6, np.float32))
    o.map_set_initial_guess(T)
    o.map_save_keyframe()
    if i % 100 == 0:
        sizes.append((len(o.download("SCAN_CORNER_DS")), len(o.download("SCAN_SURF_DS")), len(o.download("SCAN_OUTLIER_DS"))))
print("prebuild", time.time() - t0, "s (generator", tg, ") kf clouds", sizes, "state", o.download("KEYFRAME_STATE"))
# drive
o.reset_feature_association()
T0 = synth.pose_to_transform(synth.arena_pose(cfg, 0, synth.DRIVE, 0))
o.map_set_poses(T0, np.zeros(6, np.float32))
for f in range(16):
    sc = synth.arena_scan(cfg, 0, synth.DRIVE, f)
    o.image_projection(sc)
    if o.feature_association() == 1:
        t1 = time.time(); o.mapping_cycle(); dt = time.time() - t1
        Tt = synth.pose_to_transform(synth.arena_pose(cfg, 0, synth.DRIVE, f))
        aft = o.download("TRANSFORM_AFT_MAPPED")
        print(f"frame {f}: cycle {dt*1e3:.0f} ms map {len(o.download('MAP_CORNER'))}+{len(o.download('MAP_SURF'))} iters {o.download('MAP_ITERS')} "
              f"state {o.download('KEYFRAME_STATE')} err_rot {np.abs(aft[:3]-Tt[:3]).max():.5f} err_t {np.abs(aft[3:]-Tt[3:]).max():.4f}")
print("timers", o.timers(), o.timer_map_assembly())
