"""CPU evidence at BASELINE config-2 scale (no GPU): the reference's own compiled code (oracle/_ref/libsrl_reference.so) against the
oracle on the bench's 10.28M-point map and every keypoint of a 100k-point sweep, one pass each for nb = 1 and nb = 2, plus the
iterated update.  Prints one summary; run from the repo root where /root/reference exists:  python scripts/cpu_pin_full_scale.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle_py as O, reference_py as Rf
from sr_livo_b200 import synth
BIG = 2 ** 31 - 1
t0 = time.time()
pts = synth.sample_map_points(600.0, 60.0, seed=1)
om = O.OracleMap(); om.add_points(pts); del pts
ref = Rf.Reference(); ref.load(*om.snapshot())
print(f"map: {om.num_points} points, {om.num_voxels} voxels ({time.time() - t0:.0f} s); reference holds {ref.num_points()} / {ref.num_voxels()}")
sw = synth.make_sweep(100000, seed=1000, yaw=0.5)
for frame_id, n in ((100, 100000), (5, 20000)):
    prm = O.r3live_params(max_num_residuals=BIG, frame_id=frame_id)
    raw = sw.raw_xyz[:n]
    t = time.time(); r = ref.build_plane_residuals(raw, sw.q_init, sw.t_init, sw.t_last, prm); t_r = time.time() - t
    t = time.time(); o = om.build_plane_residuals(raw, sw.q_init, sw.t_init, sw.t_last, prm, debug=True); t_o = time.time() - t
    rows = o.plane[o.status == 2][:, :15]
    print(f"pass frame_id={frame_id} ({n} keypoints, nb={'2' if frame_id < 20 else '1'}): residuals {r['num_residuals_used']} / {o.num_residuals}, fragile {o.num_fragile}, "
          f"rows bit-identical {np.array_equal(r['rows'], rows)}, transformed keypoints bit-identical {np.array_equal(r['world_xyz'], o.world_xyz)}, "
          f"loss equal {r['loss_sum'] == o.loss_sum}; reference {t_r:.2f} s, oracle {t_o:.2f} s (one thread each)")
prm = O.r3live_params(max_num_residuals=BIG)
e0 = O.Eskf(p=sw.t_init.copy(), q=sw.q_init.copy(), cov=synth.prior_covariance())
t = time.time(); r = ref.update_iekf(sw.raw_xyz, e0, sw.t_last, prm); t_r = time.time() - t
t = time.time(); o = om.update_iekf(sw.raw_xyz, e0, sw.t_last, prm); t_o = time.time() - t
print(f"updateIEKF (100000 keypoints, {o['passes']} passes): residuals {r['num_residuals_used']} / {o['num_residuals_used']}, |dp| {np.abs(r['eskf'].p - o['eskf'].p).max():.2e}, "
      f"|dq| {np.abs(r['eskf'].q - o['eskf'].q).max():.2e}, |dcov| {np.abs(r['eskf'].cov - o['eskf'].cov).max():.2e}, |p - p_true| {np.linalg.norm(r['eskf'].p - sw.t_true):.4f} m; "
      f"reference {t_r:.2f} s, oracle {t_o:.2f} s")
