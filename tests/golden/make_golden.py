"""Generates tests/golden/{scan_matching,map_insert,sweep_prep}.npz with the CPU oracle (the reference ships no vectors of its
own: SURVEY.md §8(c)): inputs plus the ORACLE's outputs incl. the neighbour ids the reference never exposes.  The outputs of the
reference's own compiled code on the same inputs are in reference_outputs.npz (make_reference_golden.py); tests/test_golden.py
checks the oracle and the CUDA path against both.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_py as O          # noqa: E402
from sr_livo_b200 import synth             # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
BIG = 2 ** 31 - 1


def scan_matching_fixture():
    pts = synth.sample_map_points(40.0, 60.0, seed=21, center=(20.0, 20.0))
    sw = synth.make_sweep(512, seed=1021, yaw=0.3, position=(5.0, 5.0, 1.7), max_range=25.0)
    om = O.OracleMap()
    om.add_points(pts)
    keys, counts, xyz = om.snapshot()
    out = dict(map_keys=keys, map_counts=counts, map_xyz=xyz, raw_xyz=sw.raw_xyz, q_init=sw.q_init, t_init=sw.t_init,
               t_last=sw.t_last, q_true=sw.q_true, t_true=sw.t_true, prior_cov=synth.prior_covariance())
    for tag, kw in (("nb1", dict(max_num_residuals=BIG, frame_id=100)), ("nb2", dict(max_num_residuals=BIG, frame_id=5)),
                    ("cap", dict(max_num_residuals=100, frame_id=100))):
        r = om.build_plane_residuals(sw.raw_xyz, sw.q_init, sw.t_init, sw.t_last, O.r3live_params(**kw), debug=True)
        assert r.num_fragile == 0
        out.update({f"{tag}_status": r.status, f"{tag}_nbr": r.nbr, f"{tag}_nbr_dist": r.nbr_dist, f"{tag}_plane": r.plane,
                    f"{tag}_HTH": r.HTH, f"{tag}_HTh": r.HTh, f"{tag}_loss": np.array(r.loss_sum),
                    f"{tag}_num_residuals": np.array(r.num_residuals), f"{tag}_world": r.world_xyz,
                    f"{tag}_sum_candidates": np.array(r.sum_candidates)})
    e0 = O.Eskf(p=sw.t_init.copy(), q=sw.q_init.copy(), cov=synth.prior_covariance())
    res = om.update_iekf(sw.raw_xyz, e0, sw.t_last, O.r3live_params(max_num_residuals=BIG))
    out.update(iekf_passes=np.array(res["passes"]), iekf_trace=res["trace"], iekf_p=res["eskf"].p, iekf_q=res["eskf"].q,
               iekf_v=res["eskf"].v, iekf_g=res["eskf"].g, iekf_cov=res["eskf"].cov)
    np.savez_compressed(os.path.join(HERE, "scan_matching.npz"), **out)
    print("scan_matching.npz:", keys.shape[0], "voxels,", int(counts.sum()), "map points,", sw.raw_xyz.shape[0], "keypoints,",
          int(r.num_residuals), "residuals(cap case), iekf passes", res["passes"])


def map_insert_fixture():
    rng = synth.rng_for(31)
    # a 12 m x 12 m ground patch + one wall, offered in two batches (the second into an already-populated map)
    def patch(n):
        g = np.stack([rng.random(n) * 12 - 3, rng.random(n) * 12 - 3, np.zeros(n)], 1)
        g[:, 2] = synth.ground_h(g[:, 0], g[:, 1])
        w = np.stack([np.full(n // 2, 4.37), rng.random(n // 2) * 12 - 3, rng.random(n // 2) * 5], 1)
        p = np.concatenate([g, w]) + rng.normal(0, 1e-3, (n + n // 2, 3))
        return p[rng.permutation(p.shape[0])]
    a, b = patch(3000), patch(1500)
    om = O.OracleMap()
    n1 = om.add_points(a)
    k1, c1, x1 = om.snapshot()
    n2 = om.add_points(b)
    k2, c2, x2 = om.snapshot()
    np.savez_compressed(os.path.join(HERE, "map_insert.npz"), batch_a=a, batch_b=b, added_a=np.array(n1), added_b=np.array(n2),
                        keys_a=k1, counts_a=c1, xyz_a=x1, keys_b=k2, counts_b=c2, xyz_b=x2)
    print("map_insert.npz:", n1, "+", n2, "points,", k2.shape[0], "voxels")


def sweep_prep_fixture():
    """Rows N2 / N3: gridSampling order and the undistortion / sweep-end transforms on one small sweep."""
    rng = synth.rng_for(41)
    n = 2000
    raw = rng.normal(0, 15, (n, 3))
    rel = synth.make_sweep_times(n, seed=6)
    st = synth.make_imu_states(seed=4)
    R_il = O.quat_to_rot(np.array([0.01, -0.02, 0.015, 0.9996]) / np.linalg.norm([0.01, -0.02, 0.015, 0.9996]))
    t_il = np.array([0.04, 0.02, -0.03])
    t0 = st[0]["timestamp"]
    by_const = O.distort_frame_by_constant(raw, rel, st, t0, R_il, t_il)
    by_imu, n_imu = O.distort_frame_by_imu(raw, rel, st, t0, R_il, t_il)
    rel_cut = rel.copy(); rel_cut[1500] = -5.0                       # the iterator walk stops here
    by_imu_cut, n_cut = O.distort_frame_by_imu(raw, rel_cut, st, t0, R_il, t_il, imu_xyz_in=np.full_like(raw, 9.0))
    end_frame = O.transform_all_imu_point(by_imu, st[-1], R_il, t_il)
    keep = O.grid_sampling(by_imu, 0.8)
    np.savez_compressed(os.path.join(HERE, "sweep_prep.npz"), raw=raw, rel=rel, rel_cut=rel_cut, imu_states=O.imu_states_array(st),
                        R_il=R_il, t_il=t_il, t0=np.array(t0), by_const=by_const, by_imu=by_imu, n_imu=np.array(n_imu),
                        by_imu_cut=by_imu_cut, n_cut=np.array(n_cut), end_frame=end_frame, grid_keep=keep, grid_size=np.array(0.8))
    print("sweep_prep.npz:", n, "points,", len(st), "IMU states,", keep.shape[0], "keypoints after grid sampling, walk cut at", n_cut)


if __name__ == "__main__":
    scan_matching_fixture()
    map_insert_fixture()
    sweep_prep_fixture()
