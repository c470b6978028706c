"""First-contact GPU diagnostics: map insert parity, one-pass parity, iterated update, rough timing.
Usage: python scripts/gpu_diag.py [small|full]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from oracle import oracle_py as O
from sr_livo_b200 import lio, synth

mode = sys.argv[1] if len(sys.argv) > 1 else "small"
extent, n_sweep = (80.0, 4000) if mode == "small" else (240.0, 100000)

t0 = time.time()
pts = synth.sample_map_points(extent, 60.0, seed=1)
sw = synth.make_sweep(n_sweep, seed=1000, yaw=0.5)
print(f"synth: {pts.shape[0]} offered pts, sweep {sw.raw_xyz.shape[0]} in {time.time() - t0:.2f}s", flush=True)

L = lio.LioOptimization(max_voxels=1 << 20, sweep_capacity=max(8192, n_sweep))
t0 = time.time()
added = L.addPointsToMap(pts)
t_gpu_ins = time.time() - t0
om = O.OracleMap()
t0 = time.time()
o_added = om.add_points(pts)
t_cpu_ins = time.time() - t0
nv, npts = L.voxel_map.stats()
print(f"insert: gpu added {added} (voxels {nv}, pts {npts}) in {t_gpu_ins:.3f}s | oracle added {o_added} "
      f"(voxels {om.num_voxels}, pts {om.num_points}) in {t_cpu_ins:.3f}s", flush=True)

# map content parity
gk, gc, gx = L.voxel_map.download()
ok_, oc, ox = om.snapshot()
gd = {tuple(k): (c, x[:c].copy()) for k, c, x in zip(gk.tolist(), gc.tolist(), gx)}
od = {tuple(k): (c, x[:c].copy()) for k, c, x in zip(ok_.tolist(), oc.tolist(), ox)}
bad = 0
for k, (c, x) in od.items():
    if k not in gd or gd[k][0] != c or not np.array_equal(gd[k][1], x):
        bad += 1
        if bad <= 3:
            print("  map mismatch at", k, "oracle count", c, "gpu", gd.get(k, (None,))[0])
print(f"map parity: {len(od)} oracle voxels, {len(gd)} gpu voxels, mismatching voxels = {bad}", flush=True)

# second insert (sweep into existing map): exercises the present-voxel path
sw2 = synth.make_sweep(min(n_sweep, 20000), seed=1001, yaw=0.2, position=(1.0, 3.0, 1.8))
reg = synth.registered_points(sw2)
a2 = L.addPointsToMap(reg)
o2 = om.add_points(reg)
print(f"second insert: gpu {a2} oracle {o2}; sizes {L.mapSize()} {om.num_points}", flush=True)

print("device loop active:", L.ctx.counter("device_loop_active"), flush=True)
prm = lio.r3live_params(max_num_residuals=2 ** 31 - 1)
oprm = O.r3live_params(max_num_residuals=2 ** 31 - 1)
L.setKeypoints(sw.raw_xyz)
g = L.buildPlaneResiduals(prm, sw.q_init, sw.t_init, sw.t_last, debug=True)
t0 = time.time()
o = om.build_plane_residuals(sw.raw_xyz, sw.q_init, sw.t_init, sw.t_last, oprm, debug=True)
t_cpu_pass = time.time() - t0
print(f"pass: gpu residuals {g.num_residuals} full {g.num_full_neighborhoods} scanned {g.num_candidates_scanned} | "
      f"oracle residuals {o.num_residuals} full {o.num_full_neighborhoods} sumC {o.sum_candidates} fragile {o.num_fragile} "
      f"cpu {t_cpu_pass:.3f}s", flush=True)
print("  world_xyz max abs diff", np.abs(g.world_xyz - o.world_xyz).max())
st_bad = np.nonzero(g.status != o.status)[0]
print("  status mismatches", st_bad.size, st_bad[:10], g.status[st_bad[:10]], o.status[st_bad[:10]])
full = o.status >= 1
nb_bad = np.nonzero((g.nbr[full] != o.nbr[full]).any(axis=(1, 2)))[0]
print("  neighbour-list mismatches (full keypoints)", nb_bad.size, "of", int(full.sum()))
if nb_bad.size:
    k = np.nonzero(full)[0][nb_bad[0]]
    print("   first at", k, "\n gpu", g.nbr[k].tolist(), "\n orc", o.nbr[k].tolist(), "\n gd", g.nbr_dist[k], "\n od", o.nbr_dist[k])
print("  nbr_dist max abs diff", np.abs(g.nbr_dist[full] - o.nbr_dist[full]).max())
acc = o.status == 2
pd = np.abs(g.plane[full] - o.plane[full])
names = ["raw3", "n3", "J6", "off", "dist", "w", "a2D"]
sl = [slice(0, 3), slice(3, 6), slice(6, 12), slice(12, 13), slice(13, 14), slice(14, 15), slice(15, 16)]
for nme, s in zip(names, sl):
    den = np.maximum(np.abs(o.plane[full][:, s]).max(), 1e-300)
    print(f"  plane {nme}: max abs diff {pd[:, s].max():.3e} (scale {den:.3e})")
sc = np.abs(o.HTH).max()
print("  HTH rel diff", np.abs(g.HTH - o.HTH).max() / sc, " HTh rel diff", np.abs(g.HTh - o.HTh).max() / np.abs(o.HTh).max(),
      " loss", g.loss_sum, o.loss_sum)

# iterated update
L.eskf_pro = lio.EskfEstimator(p=sw.t_init.copy(), q=sw.q_init.copy(), cov=synth.prior_covariance())
t0 = time.time()
summ, fq, ft = L.updateIEKF(prm, sw.t_last)
t_gpu_iekf = time.time() - t0
oe = O.Eskf(p=sw.t_init.copy(), q=sw.q_init.copy(), cov=synth.prior_covariance())
t0 = time.time()
oo = om.update_iekf(sw.raw_xyz, oe, sw.t_last, oprm)
t_cpu_iekf = time.time() - t0
print(f"iekf: gpu passes {summ.passes_run} conv {summ.converged} {t_gpu_iekf:.3f}s | oracle passes {oo['passes']} {t_cpu_iekf:.3f}s")
print("  p diff", np.abs(L.eskf_pro.p - oo["eskf"].p).max(), "q diff", np.abs(L.eskf_pro.q - oo["eskf"].q).max(),
      "cov rel diff", np.abs(L.eskf_pro.cov - oo["eskf"].cov).max() / np.abs(oo["eskf"].cov).max())
print("  |p - p_true|", np.linalg.norm(L.eskf_pro.p - sw.t_true), " init err", np.linalg.norm(sw.t_init - sw.t_true))

# capped pass (max_num_residuals = 600)
prm_c = lio.r3live_params()
oprm_c = O.r3live_params()
gcap = L.buildPlaneResiduals(prm_c, sw.q_init, sw.t_init, sw.t_last, debug=True)
ocap = om.build_plane_residuals(sw.raw_xyz, sw.q_init, sw.t_init, sw.t_last, oprm_c, debug=True)
print("cap600: residuals", gcap.num_residuals, ocap.num_residuals, "status equal", np.array_equal(gcap.status, ocap.status),
      "HTH rel", np.abs(gcap.HTH - ocap.HTH).max() / np.abs(ocap.HTH).max())

print("exact fallbacks so far:", L.ctx.counter("exact_fallbacks"), "fast ambiguous:", L.ctx.counter("fast_ambiguous"))
# rough timing of the pass
import torch
torch.cuda.synchronize()
for rep in range(3):
    L.buildPlaneResiduals(prm, sw.q_init, sw.t_init, sw.t_last)
t0 = time.time()
R = 10
for rep in range(R):
    L.buildPlaneResiduals(prm, sw.q_init, sw.t_init, sw.t_last)
dt = (time.time() - t0) / R
print(f"timing: {dt * 1e6:.1f} us per pass (wall, incl. sync) -> {n_sweep / dt / 1e6:.2f} M assoc/s; kernels {L.ctx.kernel_launches}")
L.close()
