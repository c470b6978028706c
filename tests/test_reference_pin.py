"""Pins the oracle restatement (oracle/srl_oracle.cpp) against the REFERENCE'S OWN CODE.

oracle/_ref/libsrl_reference.so holds /root/reference/src/{optimize,utility,eskfEstimator,cloudMap,state,lioOptimization,
rgbMapTracker,parameters}.cpp compiled unmodified from where they lie (oracle/Makefile target `reference`), over the stand-in
headers of oracle/shim/ — Eigen, OpenCV, ROS and PCL are not in this image.  So the left-hand side of every comparison below
is produced by the reference's own control flow, containers (tsl::robin_map, std::tr1::unordered_map, std::priority_queue),
member functions, casts and quirks; the right-hand side by the oracle, which the GPU parity tests use as their checker.

What stays unpinned: the arithmetic INSIDE Eigen / OpenCV calls (reduction order of small dot products, SelfAdjointEigenSolver,
PartialPivLU, saturating Vec3b operators) — oracle/shim restates it with the same evaluation orders the oracle assumes, so
equality here says nothing about those.  Everything else that a restatement can get wrong is checked.

CPU only; skipped when the library was not built (no /root/reference on the box and no prebuilt copy).
"""
import os

import numpy as np
import pytest

from oracle import oracle_py as O
from oracle import reference_py as Rf
from sr_livo_b200 import synth

pytestmark = pytest.mark.skipif(not Rf.available(), reason="oracle/_ref/libsrl_reference.so not built (needs /root/reference)")

BIG = 2 ** 31 - 1
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def sm():
    return np.load(os.path.join(HERE, "golden", "scan_matching.npz"))


@pytest.fixture(scope="module")
def sp():
    return np.load(os.path.join(HERE, "golden", "sweep_prep.npz"))


def _pair(sm):
    ref = Rf.Reference(); ref.load(sm["map_keys"], sm["map_counts"], sm["map_xyz"])
    om = O.OracleMap(); om.load(sm["map_keys"], sm["map_counts"], sm["map_xyz"])
    return ref, om


def _tsl():
    return "tsl" in O.backend()


def _as_dict(keys, counts, xyz):
    return {tuple(k): x[:c].copy() for k, c, x in zip(keys.tolist(), counts.tolist(), xyz)}


# ---- the reference object itself ----------------------------------------------------------------------------------
def test_reference_constructor_and_hash():
    ref = Rf.Reference()
    assert Rf.lib().ref_laser_point_cov(ref._h) == 0.001                     # src/lioOptimization.cpp:364, what orc_icp_params carries
    rng = np.random.default_rng(0)
    for x, y, z in [(-1, 2, 3), (0, 0, 0), (-32767, 32767, -1)] + rng.integers(-32767, 32767, (200, 3)).tolist():
        assert Rf.voxel_hash(x, y, z) == O.voxel_hash(x, y, z)              # std::hash<voxel>, include/cloudMap.h:173-184


# ---- A7 addPointsToMap / addPointToMap, N4 removePointsFarFromLocation -------------------------------------------
@pytest.mark.parametrize("case", [dict(voxel_size=1.0, cap=20, min_dist=0.15, min_num=0), dict(voxel_size=0.5, cap=8, min_dist=0.05, min_num=0),
                                  dict(voxel_size=1.0, cap=20, min_dist=0.0, min_num=0)])
def test_add_points_to_map_equals_the_oracle(case):
    rng = np.random.default_rng(11)
    ref = Rf.Reference(); om = O.OracleMap()
    for sweep in range(4):
        pts = np.concatenate([rng.uniform(-6, 6, (6000, 3)), rng.normal(0, 0.3, (3000, 3)) + [2.2, -1.7, 0.4],
                              np.array([[0.99999999999, 0.1, 0.1], [-0.5, -0.5, -0.5], [0.5, 0.5, 0.5]])])
        a = ref.add_points_to_map(pts, case["voxel_size"], case["cap"], case["min_dist"], case["min_num"])
        b = om.add_points(pts, case["voxel_size"], case["cap"], case["min_dist"], case["min_num"])
        assert a == b
    s = ref.snapshot(cap=case["cap"]); k, c, x = om.snapshot(cap=case["cap"])
    assert ref.num_points() == om.num_points and ref.num_voxels() == om.num_voxels
    da, db = _as_dict(s["keys"], s["counts"], s["xyz"]), _as_dict(k, c, x)
    assert da.keys() == db.keys()
    assert all(np.array_equal(da[key], db[key]) for key in da)              # contents AND order inside every voxel
    if _tsl():                                                               # same container, same insertions: same iteration order
        assert np.array_equal(s["keys"], k) and np.array_equal(s["xyz"], x)
    # min_num_points > 0: only voxels that already hold enough points accept more, unknown voxels are not created (:431,:439)
    more = rng.uniform(-7, 7, (5000, 3))
    assert ref.add_points_to_map(more, case["voxel_size"], case["cap"], case["min_dist"], 3) == om.add_points(more, case["voxel_size"], case["cap"], case["min_dist"], 3)
    s = ref.snapshot(cap=case["cap"]); k, c, x = om.snapshot(cap=case["cap"])
    da, db = _as_dict(s["keys"], s["counts"], s["xyz"]), _as_dict(k, c, x)
    assert da.keys() == db.keys() and all(np.array_equal(da[key], db[key]) for key in da)
    # removePointsFarFromLocation (src/lioOptimization.cpp:556-572)
    loc = np.array([1.0, -0.5, 0.2])
    assert ref.remove_far(loc, 4.0) == om.remove_far(loc, 4.0)
    s = ref.snapshot(cap=case["cap"]); k, c, x = om.snapshot(cap=case["cap"])
    da, db = _as_dict(s["keys"], s["counts"], s["xyz"]), _as_dict(k, c, x)
    assert da.keys() == db.keys() and all(np.array_equal(da[key], db[key]) for key in da)


# ---- A3 searchNeighbors ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nb,thr", [(1, 1), (2, 1), (1, 3)])
def test_search_neighbors_equals_the_oracle(sm, nb, thr):
    ref, om = _pair(sm)
    blocks = _as_dict(sm["map_keys"], sm["map_counts"], sm["map_xyz"])
    prm = O.r3live_params(max_num_residuals=BIG, frame_id=100 if nb == 1 else 5, threshold_voxel_occupancy=thr)
    o = om.build_plane_residuals(sm["raw_xyz"], sm["q_init"], sm["t_init"], sm["t_last"], prm, debug=True)
    assert o.num_fragile == 0
    idx = np.flatnonzero(o.status >= 0)[::7]
    n_full = 0
    for k in idx:
        xyz, vox = ref.search_neighbors(o.world_xyz[k], nb=nb, size=1.0, K=20, thr=1 if nb == 2 else thr)
        if o.status[k] == 0:
            assert xyz.shape[0] < 20                                          # :78: fewer than min_number_neighbors, keypoint skipped
            continue
        n_full += 1
        assert xyz.shape[0] == 20
        want = np.array([blocks[tuple(v[:3])][v[3]] for v in o.nbr[k].tolist()], np.float64)
        assert np.array_equal(xyz, want)                                      # same 20 map points, nearest first
        assert np.array_equal(vox, o.nbr[k][:, :3])
        d = np.sqrt(((xyz - o.world_xyz[k]) ** 2)[:, 0] + (((xyz - o.world_xyz[k]) ** 2)[:, 1] + ((xyz - o.world_xyz[k]) ** 2)[:, 2]))
        assert np.array_equal(d, o.nbr_dist[k])
    assert n_full > 50


# ---- A2 / A4 buildPlaneResiduals + computeNeighborhoodDistribution ------------------------------------------------
PASS_CASES = {
    "nb1": dict(max_num_residuals=BIG, frame_id=100),
    "nb2_init_frames": dict(max_num_residuals=BIG, frame_id=5),
    "cap100": dict(max_num_residuals=100, frame_id=100),
    "cap600_yaml": dict(max_num_residuals=600, frame_id=100),
    "cap_default_minus1": dict(max_num_residuals=-1, frame_id=100),
    "occupancy3": dict(max_num_residuals=BIG, frame_id=100, threshold_voxel_occupancy=3),
    "gate_tight": dict(max_num_residuals=BIG, frame_id=100, max_dist_to_plane_icp=0.05),
    "weights": dict(max_num_residuals=BIG, frame_id=100, weight_alpha=-0.7, weight_neighborhood=0.4, power_planarity=1.0),
    "too_few": dict(max_num_residuals=BIG, frame_id=100, min_number_neighbors=20, max_dist_to_plane_icp=-1.0),
}


@pytest.mark.parametrize("tag", list(PASS_CASES))
@pytest.mark.parametrize("ext", [False, True])
def test_build_plane_residuals_equals_the_oracle(sm, tag, ext):
    ref, om = _pair(sm)
    prm = O.r3live_params(**PASS_CASES[tag])
    R_il = synth.quat_to_rot(synth.quat_from_rotvec([0.02, -0.01, 0.03])) if ext else None
    t_il = np.array([0.05, -0.02, 0.1]) if ext else None
    raw = sm["raw_xyz"] if not ext else (sm["raw_xyz"] - t_il) @ R_il     # same body-frame points through a non-trivial extrinsic
    r = ref.build_plane_residuals(raw, sm["q_init"], sm["t_init"], sm["t_last"], prm, R_il, t_il)
    o = om.build_plane_residuals(raw, sm["q_init"], sm["t_init"], sm["t_last"], prm, R_il, t_il, debug=True)
    assert not r["threw"] and not o.nan_planarity
    assert r["success"] == o.success and r["num_residuals_used"] == o.num_residuals
    visited = o.status >= 0
    assert np.array_equal(r["world_xyz"][visited], o.world_xyz[visited])     # transformKeypoints (:31-42), bit for bit
    rows = o.plane[o.status == 2][:, :15]                                     # accepted keypoints in keypoint order = plane_residuals
    assert r["rows"].shape == rows.shape
    assert np.array_equal(r["rows"], rows)                                    # raw_point, normal, Jacobian, offset, distance, weight: bit for bit
    assert r["loss_sum"] == o.loss_sum
    # the normal equations the oracle hands to the GPU tests are those rows' products (src/optimize.cpp:160-170,235,239)
    if rows.shape[0]:
        J = rows[:, 6:12]; h = rows[:, 13] * rows[:, 14]
        assert np.allclose(J.T @ J, o.HTH, rtol=1e-12, atol=1e-18) and np.allclose(J.T @ h, o.HTh, rtol=1e-11, atol=1e-18)


def test_neighborhood_distribution_on_random_and_degenerate_sets(sm):
    ref, om = _pair(sm)
    rng = np.random.default_rng(5)
    for trial in range(200):
        kind = trial % 4
        pts = rng.normal(0, 1, (20, 3)) * ([1, 1, 0.01] if kind == 0 else [1, 0.02, 0.02] if kind == 1 else [1, 1, 1] if kind == 2 else [1, 1, 1e-7])
        pts = pts @ synth.quat_to_rot(synth.quat_from_rotvec(rng.normal(0, 1, 3))).T + rng.uniform(-50, 50, 3)
        rc, nh = ref.neighborhood(pts)
        assert rc == 0
        # against the oracle's eigen restatement (same covariance accumulation order as src/optimize.cpp:327-335)
        bary = np.zeros(3)
        for p in pts:
            bary = bary + p
        bary = bary / 20.0
        cov = np.zeros((3, 3))
        for p in pts:
            for k in range(3):
                for l in range(k, 3):
                    cov[k, l] += (p[k] - bary[k]) * (p[l] - bary[l])
        cov[1, 0], cov[2, 0], cov[2, 1] = cov[0, 1], cov[0, 2], cov[1, 2]
        assert np.array_equal(nh["center"], bary) and np.array_equal(nh["covariance"], cov)
        ev, evec = O.eig3_sym(cov)
        n = evec[:, 0]
        nn = n[0] * n[0] + (n[1] * n[1] + n[2] * n[2])
        n = n / np.sqrt(nn) if nn > 0 else n
        assert np.array_equal(nh["normal"], n)
        s1, s2, s3 = np.sqrt(abs(ev[2])), np.sqrt(abs(ev[1])), np.sqrt(abs(ev[0]))
        assert nh["a2D"] == (s2 - s3) / s1
    # all points identical: 0 / 0 -> NaN planarity -> the reference throws (:348-350)
    rc, _ = ref.neighborhood(np.tile([[1.0, 2.0, 3.0]], (20, 1)))
    assert rc == 1


def test_nan_planarity_throws_in_both(sm):
    # a map whose voxels hold 20 copies of one point each: every neighbourhood is degenerate
    keys = np.array([[0, 0, 0], [1, 0, 0]], np.int16); counts = np.array([20, 20], np.int32)
    xyz = np.zeros((2, 20, 3), np.float32); xyz[0] = [0.5, 0.5, 0.5]; xyz[1] = [1.5, 0.5, 0.5]
    ref = Rf.Reference(); ref.load(keys, counts, xyz)
    om = O.OracleMap(); om.load(keys, counts, xyz)
    raw = np.array([[0.6, 0.5, 0.5]])
    prm = O.r3live_params(max_num_residuals=BIG)
    q, t = np.array([0, 0, 0, 1.0]), np.zeros(3)
    assert ref.build_plane_residuals(raw, q, t, t, prm)["threw"]
    assert om.build_plane_residuals(raw, q, t, t, prm).nan_planarity


# ---- A1 / A8 updateIEKF + observe ---------------------------------------------------------------------------------
IEKF_CASES = {
    "steady": dict(max_num_residuals=BIG),
    "cap600_yaml": dict(max_num_residuals=600),
    "init_frame_15_iterations": dict(max_num_residuals=BIG, frame_id=5),
    "frame_1_never_converges_early": dict(max_num_residuals=BIG, frame_id=1, init_num_frames=0),
    "loose_thresholds_early_exit": dict(max_num_residuals=BIG, threshold_orientation_norm=5.0, threshold_translation_norm=0.5),
    "one_iteration": dict(max_num_residuals=BIG, num_iters_icp=1),
    "too_few_residuals": dict(max_num_residuals=BIG, max_dist_to_plane_icp=-1.0),
}


def _assert_eskf_equal(a: O.Eskf, b: O.Eskf, rtol, atol):
    for name in ("p", "q", "v", "ba", "bg", "g"):
        assert np.allclose(getattr(a, name), getattr(b, name), rtol=rtol, atol=atol), name
    # (P / c)^-1 is ill-conditioned (c = 0.001, prior entries 1e-5 .. 1): small entries move in their 7th digit between two
    # LU loop orders; measured against the matrix's scale the two covariances agree to 1e-11
    assert np.allclose(a.cov, b.cov, rtol=1e-6, atol=1e-11 * np.abs(b.cov).max())


@pytest.mark.parametrize("tag", list(IEKF_CASES))
def test_update_iekf_equals_the_oracle(sm, tag):
    ref, om = _pair(sm)
    prm = O.r3live_params(**IEKF_CASES[tag])
    rng = np.random.default_rng(3)
    e0 = O.Eskf(p=sm["t_init"].copy(), q=sm["q_init"].copy(), v=rng.normal(0, 0.3, 3), ba=rng.normal(0, 0.01, 3), bg=rng.normal(0, 0.001, 3),
                g=np.array([0.03, -0.02, 9.8]), cov=sm["prior_cov"].copy())
    R_il = synth.quat_to_rot(synth.quat_from_rotvec([0.01, 0.02, -0.015])); t_il = np.array([0.04, 0.0, -0.03])
    raw = (sm["raw_xyz"] - t_il) @ R_il
    r = ref.update_iekf(raw, e0, sm["t_last"], prm, R_il=R_il, t_il=t_il)
    o = om.update_iekf(raw, e0, sm["t_last"], prm, R_il=R_il, t_il=t_il)
    assert not r["threw"]
    assert r["success"] == o["success"] and r["num_residuals_used"] == o["num_residuals_used"]
    # the associations are identical pass by pass (previous test), so the two states can only differ by the rounding of the
    # 17-dimensional algebra (Eigen's GEMM / LU loop orders are restated, not pinned): 1e-9
    _assert_eskf_equal(r["eskf"], o["eskf"], rtol=1e-9, atol=1e-11)
    assert np.allclose(r["frame_q"], o["frame_q"], rtol=1e-9, atol=1e-12) and np.allclose(r["frame_t"], o["frame_t"], rtol=1e-9, atol=1e-11)
    if tag == "too_few_residuals":
        assert not r["success"]
        assert np.array_equal(r["eskf"].p, e0.p) and np.array_equal(r["eskf"].cov, e0.cov)   # early return (:155) leaves the filter untouched
    else:
        assert r["success"] and np.linalg.norm(r["eskf"].p - sm["t_true"]) < np.linalg.norm(e0.p - sm["t_true"])


def test_eskf_observe_equals_the_oracle():
    rng = np.random.default_rng(9)
    for trial in range(100):
        q = rng.normal(0, 1, 4); q /= np.linalg.norm(q)
        e = O.Eskf(p=rng.normal(0, 5, 3), q=q, v=rng.normal(0, 1, 3), ba=rng.normal(0, 0.1, 3), bg=rng.normal(0, 0.01, 3),
                   g=np.array([0, 0, 9.81]) + rng.normal(0, 0.2, 3))
        dx = rng.normal(0, 1, 17) * (1e-6 if trial % 3 == 0 else 1e-2 if trial % 3 == 1 else 0.5)   # both branches of so3ToQuat / so3ToRotation
        a, b = Rf.eskf_observe(e, dx), e.observe(dx)
        for name in ("p", "q", "v", "ba", "bg", "g"):
            assert np.allclose(getattr(a, name), getattr(b, name), rtol=1e-14, atol=1e-15), name


# ---- the caller: optimize() = gridSampling -> updateIEKF -> transformPoint (src/optimize.cpp:428-447) -------------
def test_optimize_equals_the_composition_of_oracle_pieces(sm):
    ref, om = _pair(sm)
    prm = O.r3live_params(max_num_residuals=600)
    raw = sm["raw_xyz"]
    R0 = synth.quat_to_rot(sm["q_init"])
    world0 = raw @ R0.T + sm["t_init"]                      # point_frame[i].point as the caller left it (pose prediction)
    e0 = O.Eskf(p=sm["t_init"].copy(), q=sm["q_init"].copy(), cov=sm["prior_cov"].copy())
    size = 1.5
    r = ref.optimize(world0, raw, size, e0, sm["t_last"], prm)
    idx = O.grid_sampling(world0, size)
    assert np.array_equal(idx, Rf.grid_sampling(world0, size))
    o = om.update_iekf(raw[idx], e0, sm["t_last"], prm)
    assert r["success"] == o["success"] and r["num_residuals_used"] == o["num_residuals_used"]
    _assert_eskf_equal(r["eskf"], o["eskf"], rtol=1e-9, atol=1e-11)
    want = Rf.transform_point(raw, o["frame_q"], o["frame_t"])
    assert np.allclose(r["world"], want, rtol=0, atol=1e-8)


# ---- N2 gridSampling: the iteration order of std::tr1::unordered_map ------------------------------------------------
@pytest.mark.parametrize("n,size,spread", [(5000, 1.5, 40.0), (20000, 0.8, 60.0), (300, 1.0, 3.0), (1, 1.0, 1.0)])
def test_grid_sampling_order_equals_the_oracle(n, size, spread):
    rng = np.random.default_rng(n)
    xyz = rng.uniform(-spread, spread, (n, 3))
    xyz[: n // 10] = xyz[n // 10: 2 * (n // 10)][: n // 10] + 1e-3            # several points per cell: the first one wins
    a, b = Rf.grid_sampling(xyz, size), O.grid_sampling(xyz, size)
    assert np.array_equal(a, b)


# ---- N3 undistortion / per-point transforms --------------------------------------------------------------------------
def _states(sp):
    return [dict(timestamp=r[0], quat=r[1:5], trans=r[5:8], vel=r[8:11], un_acc=r[11:14], un_gyr=r[14:17]) for r in sp["imu_states"]]


def test_point_transforms_equal_the_oracle(sp):
    states = _states(sp)
    raw, rel, t0 = sp["raw"], sp["rel"], float(sp["t0"])
    R_il, t_il = sp["R_il"], sp["t_il"]
    a = Rf.distort_frame_by_constant(raw, rel, states, t0, R_il, t_il)
    b = O.distort_frame_by_constant(raw, rel, states, t0, R_il, t_il)
    assert np.array_equal(a, b)
    seed_in = np.full_like(raw, -7.0)                                       # points the iterator never reaches keep their value
    a = Rf.distort_frame_by_imu(raw, rel, states, t0, R_il, t_il, seed_in)
    b, m = O.distort_frame_by_imu(raw, rel, states, t0, R_il, t_il, seed_in)
    assert np.array_equal(a, b) and 0 < m <= raw.shape[0]
    # a point outside every IMU interval stops the one-iterator walk for good (src/utility.cpp:263-308)
    rel2 = rel.copy(); rel2[rel.shape[0] // 3] = 1e6
    a = Rf.distort_frame_by_imu(raw, rel2, states, t0, R_il, t_il, seed_in)
    b, m2 = O.distort_frame_by_imu(raw, rel2, states, t0, R_il, t_il, seed_in)
    assert np.array_equal(a, b) and m2 == rel.shape[0] // 3
    a = Rf.transform_all_imu_point(b, states[-1], R_il, t_il)
    c = O.transform_all_imu_point(b, states[-1], R_il, t_il)
    assert np.array_equal(a, c)


# ---- N4 colour map + renderer -------------------------------------------------------------------------------------------
def test_color_map_and_renderer_equal_the_oracle():
    rng = np.random.default_rng(21)
    ref = Rf.Reference()
    oc = O.OracleColorMap(voxel_size=0.5, max_num_points_in_voxel=8, min_distance_points=0.05)
    om = O.OracleMap()
    rows, cols = 96, 128
    fx = fy = 90.0; cx, cy = cols / 2.0, rows / 2.0
    for sweep in range(3):
        pts = np.concatenate([rng.uniform(-2.5, 2.5, (2500, 2)), rng.uniform(3.0, 6.0, (2500, 1))], axis=1)   # in front of the camera
        t_end, t_proc = 10.0 + sweep, 10.0 + sweep - (0.0 if sweep == 1 else 0.5)       # sweep 1: |dt| < 1e-5 -> no voxel becomes recent
        a = ref.add_points_to_map(pts, 1.0, 20, 0.15, 0, color_voxel_size=0.5, color_max_points=8, color_min_distance=0.05, add_point_step=3,
                                  time_sweep_end=t_end, time_last_process=t_proc, to_rendering=True)
        assert a == om.add_points(pts, 1.0, 20, 0.15, 0)
        oc.add_points(pts, add_point_step=3, time_sweep_end=t_end, time_last_process=t_proc, to_rendering=True)
        ca, cb = ref.color_counts(), oc.counts()
        assert ca == cb
        la, lb = ref.color_lists(), oc.lists()
        assert np.array_equal(la[0], lb[0]) and np.array_equal(la[1], lb[1])  # rgb_points_vec and voxels_recent_visited, in order
        # render twice (first observation, then the fusion branch of updateRgb)
        for k in range(2):
            img = rng.integers(0, 256, (rows, cols, 3), dtype=np.uint8)
            q_cw = synth.quat_from_rotvec(rng.normal(0, 0.02, 3)); t_cw = rng.normal(0, 0.05, 3)
            R = synth.quat_to_rot(q_cw); t_wc = -R.T @ t_cw
            cam = np.concatenate([q_cw, t_cw, t_wc, [fx, fy, cx, cy, 0.005]])
            obs = t_end + 0.01 * (k + 1)
            assert ref.color_render(cam, img, obs) == oc.render(cam, img, obs)
        sa = ref.snapshot(which=1, cap=8, color=True); sb = oc.snapshot()
        da = {tuple(k): i for i, k in enumerate(sa["keys"].tolist())}
        assert da.keys() == {tuple(k) for k in sb["keys"].tolist()}
        for j, key in enumerate(sb["keys"].tolist()):
            i = da[tuple(key)]
            for f in ("counts", "xyz", "rgb", "n_rgb", "cov", "obs_dist", "last_obs", "last_visited"):
                assert np.array_equal(sa[f][i], sb[f][j]), (f, key)
    assert ca["rgb_points"] > 500 and (sa["n_rgb"] > 1).sum() > 100


# ---- a different, larger scene: several sweeps and poses, map built by the reference's own addPointsToMap -------------
def test_randomized_scene_passes_equal_the_oracle():
    pts = synth.sample_map_points(80.0, 40.0, seed=3)
    ref = Rf.Reference(); om = O.OracleMap()
    assert ref.add_points_to_map(pts) == om.add_points(pts)
    total = 0
    for i, kw in enumerate([dict(yaw=0.4, position=(0.0, 3.0, 1.8)), dict(yaw=-2.0, position=(6.0, -4.0, 2.2), pattern="spinning"),
                            dict(yaw=1.3, position=(-9.0, 8.0, 1.5), dp_max=0.3, dth_max_deg=3.0)]):
        sw = synth.make_sweep(2500, seed=3100 + i, **kw)
        for frame_id, cap in [(100, BIG), (3, BIG), (100, 600)]:
            prm = O.r3live_params(max_num_residuals=cap, frame_id=frame_id)
            r = ref.build_plane_residuals(sw.raw_xyz, sw.q_init, sw.t_init, sw.t_last, prm)
            o = om.build_plane_residuals(sw.raw_xyz, sw.q_init, sw.t_init, sw.t_last, prm, debug=True)
            assert o.num_fragile == 0
            assert r["success"] == o.success and r["num_residuals_used"] == o.num_residuals
            assert np.array_equal(r["rows"], o.plane[o.status == 2][:, :15]) and r["loss_sum"] == o.loss_sum
            total += o.num_residuals
    assert total > 5000


# ---- config 4 in miniature: insert, register the next sweep, insert it — all through the reference's member functions --
def test_streaming_insert_then_register_equals_the_oracle():
    ref = Rf.Reference(); om = O.OracleMap()
    prm = O.r3live_params(max_num_residuals=BIG)
    e_r = e_o = None
    for i in range(4):
        sw = synth.make_sweep(6000, seed=3300 + i, yaw=0.5, position=(-4.0 + 1.0 * i, 3.0, 1.8))
        if i >= 2:                                                            # the first sweeps only build the map
            e0 = O.Eskf(p=sw.t_init.copy(), q=sw.q_init.copy(), cov=synth.prior_covariance())
            r = ref.update_iekf(sw.raw_xyz, e0, sw.t_last, prm)
            o = om.update_iekf(sw.raw_xyz, e0, sw.t_last, prm)
            assert r["success"] and o["success"] and r["num_residuals_used"] == o["num_residuals_used"]
            _assert_eskf_equal(r["eskf"], o["eskf"], rtol=1e-9, atol=1e-11)
            assert np.linalg.norm(r["eskf"].p - sw.t_true) < 0.03
        reg = synth.registered_points(sw)
        assert ref.add_points_to_map(reg) == om.add_points(reg)
    s = ref.snapshot(); k, c, x = om.snapshot()
    da, db = _as_dict(s["keys"], s["counts"], s["xyz"]), _as_dict(k, c, x)
    assert da.keys() == db.keys() and all(np.array_equal(da[key], db[key]) for key in da)


def test_many_sweeps_on_host_threads_equal_the_single_calls(sm):
    """ref_update_iekf_many (the throughput form bench.py's CPU arm times): every sweep's result equals the single-sweep call."""
    ref, om = _pair(sm)
    prm = O.r3live_params(max_num_residuals=BIG)
    rng = np.random.default_rng(17)
    raws, eskfs, tls = [], [], []
    for s in range(6):
        keep = rng.permutation(sm["raw_xyz"].shape[0])[:400]
        raws.append(sm["raw_xyz"][keep])
        eskfs.append(O.Eskf(p=sm["t_init"] + rng.normal(0, 0.02, 3), q=sm["q_init"].copy(), cov=sm["prior_cov"].copy()))
        tls.append(sm["t_last"])
    ok, out, fq, ft = ref.update_iekf_many(raws, eskfs, tls, prm, n_threads=3)
    assert ok == 6
    for s in range(6):
        one = ref.update_iekf(raws[s], eskfs[s], tls[s], prm)
        assert np.array_equal(out[s].p, one["eskf"].p) and np.array_equal(out[s].cov, one["eskf"].cov) and np.array_equal(fq[s], one["frame_q"])


# ---- a 1.6M-point map (the bench's world, 240 m side) and a sample of a 100k-point sweep: the scale the GPU id tests run at
def test_mid_scale_sample_equals_the_oracle():
    pts = synth.sample_map_points(240.0, 60.0, seed=1)
    om = O.OracleMap(); om.add_points(pts)
    del pts
    ref = Rf.Reference(); ref.load(*om.snapshot())
    assert ref.num_points() == om.num_points > 1_500_000
    sw = synth.make_sweep(100000, seed=1000, yaw=0.5)
    pick = np.random.default_rng(1).permutation(100000)[:6000]
    raw = sw.raw_xyz[pick]
    prm = O.r3live_params(max_num_residuals=BIG)
    r = ref.build_plane_residuals(raw, sw.q_init, sw.t_init, sw.t_last, prm)
    o = om.build_plane_residuals(raw, sw.q_init, sw.t_init, sw.t_last, prm, debug=True)
    assert o.num_fragile == 0 and r["num_residuals_used"] == o.num_residuals > 4000
    assert np.array_equal(r["rows"], o.plane[o.status == 2][:, :15]) and np.array_equal(r["world_xyz"], o.world_xyz)


# ---- the one unpinned choice that reaches a discrete outcome: the order of a 3-term reduction inside Eigen ------------------
@pytest.mark.skipif(not Rf.available("packet"), reason="oracle/_ref/libsrl_reference_packet.so not built")
def test_no_discrete_outcome_depends_on_the_order_of_three_term_reductions(sm):
    """oracle/_ref/libsrl_reference_packet.so = the same reference sources over the stand-in with dot / norm / product rows of three
    terms evaluated as (c0 + c1) + c2 (what a vectorised Eigen 3.3 most likely does for plain Vector3d operands) instead of
    c0 + (c1 + c2) (what the oracle, the default stand-in and the CUDA kernels do).  Transformed keypoints and distances move by an
    ulp; the accepted keypoints, their 20 neighbours (same map points, same order) and the final pose do not change on any scene."""
    scenes = [(sm["map_keys"], sm["map_counts"], sm["map_xyz"], sm["raw_xyz"], sm["q_init"], sm["t_init"], sm["t_last"])]
    om = O.OracleMap(); om.add_points(synth.sample_map_points(80.0, 40.0, seed=3))
    sw = synth.make_sweep(3000, seed=3100, yaw=0.4)
    scenes.append((*om.snapshot(), sw.raw_xyz, sw.q_init, sw.t_init, sw.t_last))
    moved = 0
    for keys, counts, xyz, raw, q, t, tl in scenes:
        a, b = Rf.Reference(), Rf.Reference("packet")
        a.load(keys, counts, xyz); b.load(keys, counts, xyz)
        for kw in (dict(max_num_residuals=BIG), dict(max_num_residuals=BIG, frame_id=5), dict(max_num_residuals=600)):
            prm = O.r3live_params(**kw)
            ra, rb = a.build_plane_residuals(raw, q, t, tl, prm), b.build_plane_residuals(raw, q, t, tl, prm)
            assert ra["num_residuals_used"] == rb["num_residuals_used"] and ra["rows"].shape == rb["rows"].shape
            assert np.array_equal(ra["rows"][:, 0:3], rb["rows"][:, 0:3])                       # the same keypoints were accepted, in order
            assert np.allclose(ra["rows"], rb["rows"], rtol=1e-11, atol=1e-13)
            assert np.abs(ra["world_xyz"] - rb["world_xyz"]).max() < 1e-13
            moved += int((ra["world_xyz"] != rb["world_xyz"]).sum())
        nb = 0
        for k in range(0, raw.shape[0], 5):                                                       # neighbour lists: identical map points, identical order
            xa, va = a.search_neighbors(ra["world_xyz"][k]); xb, vb = b.search_neighbors(rb["world_xyz"][k])
            assert np.array_equal(xa, xb) and np.array_equal(va, vb)
            nb += xa.shape[0] == 20
        assert nb > 50
        e0 = O.Eskf(p=t.copy(), q=q.copy(), cov=synth.prior_covariance())
        ua, ub = a.update_iekf(raw, e0, tl, O.r3live_params(max_num_residuals=BIG)), b.update_iekf(raw, e0, tl, O.r3live_params(max_num_residuals=BIG))
        assert ua["success"] and ub["success"] and ua["num_residuals_used"] == ub["num_residuals_used"]
        assert np.allclose(ua["eskf"].p, ub["eskf"].p, rtol=0, atol=1e-11) and np.allclose(ua["eskf"].q, ub["eskf"].q, rtol=0, atol=1e-12)
    assert moved > 100          # the two orders do differ in the last place — the test is not vacuous


# ---- exact distance ties and keypoints on cell boundaries: the heap's behaviour among equal distances is the reference's own ----
def test_exact_ties_and_cell_boundaries_equal_the_oracle():
    g = np.arange(-2.0, 4.0, 0.25)
    lattice = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    lattice = lattice[np.random.default_rng(2).permutation(lattice.shape[0])]
    ref = Rf.Reference(); om = O.OracleMap()
    assert ref.add_points_to_map(lattice, 1.0, 20, 0.15, 0) == om.add_points(lattice, 1.0, 20, 0.15, 0)
    s = ref.snapshot(); k, c, x = om.snapshot()
    assert np.array_equal(s["keys"], k) and np.array_equal(s["xyz"], x) if _tsl() else True
    # keypoints on lattice points, on cell faces / edges / corners (exact integers), at cell centres, and on the doubled cell 0
    kp = np.concatenate([lattice[:300], np.stack(np.meshgrid([-1.0, 0.0, 1.0, 2.0], [0.0, 1.0, 0.5], [1.0, 1.5, -0.0], indexing="ij"), -1).reshape(-1, 3),
                         np.array([[0.999999999999, 0.5, 0.5], [-0.999999999999, 0.5, 0.5], [1e-300, -1e-300, 0.5]])])
    q, t = np.array([0.0, 0.0, 0.0, 1.0]), np.zeros(3)
    for kw in (dict(max_num_residuals=BIG), dict(max_num_residuals=BIG, frame_id=5), dict(max_num_residuals=BIG, max_dist_to_plane_icp=10.0)):
        prm = O.r3live_params(**kw)
        r = ref.build_plane_residuals(kp, q, t, np.array([-5.0, 0.3, 0.2]), prm)
        o = om.build_plane_residuals(kp, q, t, np.array([-5.0, 0.3, 0.2]), prm, debug=True)
        assert o.num_fragile > 100                                            # ties and boundaries everywhere: exactly what the other tests exclude
        if r["threw"]:
            assert o.nan_planarity
            continue
        assert r["num_residuals_used"] == o.num_residuals and np.array_equal(r["world_xyz"], o.world_xyz)
        assert np.array_equal(r["rows"], o.plane[o.status == 2][:, :15])
    blocks = _as_dict(k, c, x)
    o = om.build_plane_residuals(kp, q, t, np.array([-5.0, 0.3, 0.2]), O.r3live_params(max_num_residuals=BIG), debug=True)
    for i in range(0, kp.shape[0], 3):                                        # the neighbour lists themselves, ties included
        xyz, vox = ref.search_neighbors(kp[i])
        if o.status[i] >= 1:
            want = np.array([blocks[tuple(v[:3])][v[3]] for v in o.nbr[i].tolist()], np.float64)
            assert np.array_equal(xyz, want) and np.array_equal(vox, o.nbr[i][:, :3])


# ---- the caller of the path: stateEstimation over a stream (src/lioOptimization.cpp:983-1035) ------------------------------------------
def _rel_pose(q0, t0, q, t):
    """pose (q, t) expressed in the frame of pose (q0, t0)"""
    R0 = synth.quat_to_rot(q0)
    q0_inv = np.array([-q0[0], -q0[1], -q0[2], q0[3]])
    return synth._quat_mul_xyzw(q0_inv, q), R0.T @ (t - t0)


def test_state_estimation_stream_equals_the_composition_of_oracle_pieces():
    """Seven sweeps through the reference's own stateEstimation — frame 1 only fills the map, frames 2-3 run with the init-frame
    parameters (nb = 2, >= 15 iterations, init_sample_voxel_size), the rest with the steady ones — against gridSampling -> updateIEKF ->
    transformPoint -> addPointsToMap composed from the oracle's pieces with the same per-frame switches (the Python the GPU tests use)."""
    n_frames, init_num_frames, n = 7, 4, 5000
    sweeps = [synth.make_sweep(n, seed=3500 + i, yaw=0.2, position=(0.05 * i, 3.0, 1.8), dp_max=0.03, dth_max_deg=0.3) for i in range(n_frames)]
    q0, t0 = sweeps[0].q_true, sweeps[0].t_true
    ref = Rf.Reference(); ref.stream_reset()
    om = O.OracleMap()
    P = synth.prior_covariance()
    t_prev = None
    for k, sw in enumerate(sweeps, start=1):
        if k <= 2:
            q_pred, t_pred = np.array([0.0, 0.0, 0.0, 1.0]), np.zeros(3)        # stateInitialization: identity for the first two frames
        else:
            q_pred, t_pred = _rel_pose(q0, t0, sw.q_init, sw.t_init)            # a perturbed prediction in the frame of sweep 0
        e0 = O.Eskf(p=t_pred.copy(), q=q_pred.copy(), cov=P.copy())
        prm = O.r3live_params(frame_id=k, init_num_frames=init_num_frames)      # yaml cap 600
        r = ref.stream_push(sw.raw_xyz, k, q_pred, t_pred, e0, prm, init_num_frames=init_num_frames)
        assert not r["threw"] and r["success"]
        # the same step from the oracle's pieces
        if k > 1:
            world_pred = sw.raw_xyz if k <= 2 else Rf.transform_point(sw.raw_xyz, q_pred, t_pred)
            idx = O.grid_sampling(world_pred, 1.0 if k < init_num_frames else 1.5)
            o = om.update_iekf(sw.raw_xyz[idx], e0, t_prev, prm, frame_q=q_pred, frame_t=t_pred)
            assert o["success"] and o["num_residuals_used"] == r["num_residuals_used"]
            assert o["passes"] >= (2 if k >= init_num_frames else 2)
            _assert_eskf_equal(r["eskf"], o["eskf"], rtol=1e-9, atol=1e-11)
            assert np.allclose(r["frame_q"], o["frame_q"], atol=1e-11) and np.allclose(r["frame_t"], o["frame_t"], atol=1e-11)
            world = Rf.transform_point(sw.raw_xyz, o["frame_q"], o["frame_t"])
            assert np.allclose(r["world"], world, rtol=0, atol=1e-9)
            t_prev = o["frame_t"].copy()
            if k > 2:                                                            # registration pulls the perturbed prediction back to the truth
                q_true, t_true = _rel_pose(q0, t0, sw.q_true, sw.t_true)
                assert np.linalg.norm(o["frame_t"] - t_true) < 0.02
        else:
            world = sw.raw_xyz.copy()                                            # frame 1: identity, no optimisation (:1010-1019)
            assert np.array_equal(r["world"], world) and np.array_equal(r["frame_t"], np.zeros(3))
            t_prev = np.zeros(3)
        om.add_points(r["world"], 1.0, 20, 0.1, 0)                               # odometryOptions::min_distance_points = 0.1
        assert ref.num_points() == om.num_points
    s = ref.snapshot(); kk, cc, xx = om.snapshot()
    da, db = _as_dict(s["keys"], s["counts"], s["xyz"]), _as_dict(kk, cc, xx)
    assert da.keys() == db.keys() and all(np.array_equal(da[key], db[key]) for key in da)


def test_threaded_oracle_port_equals_the_compiled_reference(sm):
    """The form of the oracle bench.py times on all host threads (keypoint ranges over std::threads, private sums, fixed-order combine)
    against the single-threaded reference: same residuals, state to 1e-9."""
    ref, om = _pair(sm)
    prm = O.r3live_params(max_num_residuals=BIG)
    e0 = O.Eskf(p=sm["t_init"].copy(), q=sm["q_init"].copy(), cov=sm["prior_cov"].copy())
    r = ref.update_iekf(sm["raw_xyz"], e0, sm["t_last"], prm)
    for nt in (2, 5, 8):
        o = om.update_iekf(sm["raw_xyz"], e0, sm["t_last"], prm, nthreads=nt)
        assert o["success"] and o["num_residuals_used"] == r["num_residuals_used"]
        _assert_eskf_equal(r["eskf"], o["eskf"], rtol=1e-9, atol=1e-11)


# ---- a seeded sweep over the parameter space (icpOptions the yaml files never set included) --------------------------------------------
def test_random_parameter_sets_equal_the_oracle():
    rng = np.random.default_rng(77)
    pts = synth.sample_map_points(80.0, 50.0, seed=9)
    sw = synth.make_sweep(1500, seed=3700, yaw=-0.7, position=(2.0, -3.0, 1.6))
    maps = {}
    checked = 0
    for trial in range(36):
        size = [0.5, 1.0, 2.0][trial % 3]
        cap = [20, 8, 32][(trial // 3) % 3]
        if (size, cap) not in maps:
            ref = Rf.Reference(); om = O.OracleMap()
            assert ref.add_points_to_map(pts, size, cap, 0.07 * size, 0) == om.add_points(pts, size, cap, 0.07 * size, 0)
            maps[(size, cap)] = (ref, om)
        ref, om = maps[(size, cap)]
        kmax = int(rng.choice([5, 10, 20, 30]))
        kw = dict(size_voxel_map=size, max_number_neighbors=kmax, min_number_neighbors=int(rng.integers(3, kmax + 1)),
                  threshold_voxel_occupancy=int(rng.choice([1, 2, 5])), voxel_neighborhood=int(rng.choice([0, 1, 2])),
                  power_planarity=float(rng.choice([0.5, 1.0, 2.0, 3.0])), max_dist_to_plane_icp=float(rng.choice([0.05, 0.3, 1.0])),
                  weight_alpha=float(rng.uniform(-1, 1)), weight_neighborhood=float(rng.uniform(0.05, 1)),
                  max_num_residuals=int(rng.choice([BIG, 600, 50, -1])), frame_id=int(rng.choice([1, 5, 19, 20, 100])),
                  init_num_frames=int(rng.choice([0, 20])))
        prm = O.r3live_params(**kw)
        dq = synth.quat_from_rotvec(rng.normal(0, 0.01, 3))
        q = synth.quat_mul(sw.q_true, dq); t = sw.t_true + rng.normal(0, 0.05, 3)
        r = ref.build_plane_residuals(sw.raw_xyz, q, t, sw.t_last, prm)
        o = om.build_plane_residuals(sw.raw_xyz, q, t, sw.t_last, prm, debug=True)
        assert r["threw"] == bool(o.nan_planarity), kw
        if r["threw"]:
            continue
        assert r["success"] == o.success and r["num_residuals_used"] == o.num_residuals, kw
        assert np.array_equal(r["rows"], o.plane[o.status == 2][:, :15]), kw
        assert r["loss_sum"] == o.loss_sum
        checked += r["rows"].shape[0]
    assert checked > 5000


def test_random_update_parameter_sets_equal_the_oracle(sm):
    ref, om = _pair(sm)
    rng = np.random.default_rng(78)
    n_ok = 0
    for trial in range(16):
        kw = dict(max_num_residuals=int(rng.choice([BIG, 600, 200])), num_iters_icp=int(rng.choice([1, 2, 5, 8])),
                  threshold_translation_norm=float(rng.choice([0.0, 0.001, 0.01, 0.2])), threshold_orientation_norm=float(rng.choice([0.0, 0.01, 0.1, 2.0])),
                  frame_id=int(rng.choice([1, 2, 5, 100])), init_num_frames=int(rng.choice([0, 20])), laser_point_cov=float(rng.choice([0.001, 0.01, 0.0001])),
                  max_dist_to_plane_icp=float(rng.choice([0.1, 0.3])), power_planarity=float(rng.choice([1.0, 2.0])))
        prm = O.r3live_params(**kw)
        e0 = O.Eskf(p=sm["t_init"] + rng.normal(0, 0.03, 3), q=synth.quat_mul(sm["q_init"], synth.quat_from_rotvec(rng.normal(0, 0.004, 3))),
                    v=rng.normal(0, 0.5, 3), ba=rng.normal(0, 0.02, 3), bg=rng.normal(0, 0.002, 3), g=np.array([0.0, 0.0, 9.81]) + rng.normal(0, 0.05, 3),
                    cov=sm["prior_cov"] * float(rng.choice([0.1, 1.0, 10.0])))
        r = ref.update_iekf(sm["raw_xyz"], e0, sm["t_last"], prm)
        o = om.update_iekf(sm["raw_xyz"], e0, sm["t_last"], prm)
        assert not r["threw"] and r["success"] == o["success"] and r["num_residuals_used"] == o["num_residuals_used"], kw
        _assert_eskf_equal(r["eskf"], o["eskf"], rtol=1e-8, atol=1e-10)
        assert np.allclose(r["frame_q"], o["frame_q"], atol=1e-10) and np.allclose(r["frame_t"], o["frame_t"], atol=1e-10), kw
        n_ok += r["success"]
    assert n_ok >= 12
