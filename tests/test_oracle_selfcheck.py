"""Self-checks that pin the CPU oracle (the reference ships no golden vectors for this path: SURVEY.md §8(c)).

Each test checks one restated piece of the reference against an independent computation (numpy / brute force /
finite differences) or against a reference quirk listed in SURVEY.md §8(a).
"""
import numpy as np
import pytest

from oracle import oracle_py as O
from sr_livo_b200 import synth

BIG = 2 ** 31 - 1


# ---- voxel keys / hash (include/cloudMap.h:124-184, src/optimize.cpp:372-374) -----------------------------
def test_voxel_hash_matches_reference_formula():
    # additive primes in size_t with sign-extended shorts; (−1,2,3) -> 0x0CD58C12 (probed)
    assert O.voxel_hash(-1, 2, 3) & 0xFFFFFFFF == 0x0CD58C12
    for x, y, z in [(0, 0, 0), (5, -7, 100), (-32767, 32767, -1)]:
        expect = (x * 73856093 + y * 19349669 + z * 83492791) % (1 << 64)
        assert O.voxel_hash(x, y, z) == expect


def test_keys_truncate_toward_zero_and_insert_uses_float_rounded_position():
    m = O.OracleMap()
    # cell 0 spans (-1, 1): both points land in voxel (0,0,0); -1.5 lands in voxel -1
    m.add_points(np.array([[0.5, 0.5, 0.5], [-0.5, -0.5, -0.5], [-1.5, 0.2, 0.2]]), min_distance_points=0.0)
    keys, counts, xyz = m.snapshot()
    d = {tuple(k): c for k, c in zip(keys.tolist(), counts.tolist())}
    assert d == {(0, 0, 0): 2, (-1, 0, 0): 1}
    # insertion key comes from the FLOAT-rounded coordinate: 0.99999999999 rounds to 1.0f -> voxel 1, not 0
    m2 = O.OracleMap()
    m2.add_points(np.array([[0.99999999999, 0.1, 0.1]]))
    keys, counts, xyz = m2.snapshot()
    assert keys.tolist() == [[1, 0, 0]] and xyz[0, 0, 0] == np.float32(1.0)


def test_add_point_semantics_cap_min_distance_and_order():
    rng = np.random.default_rng(0)
    m = O.OracleMap()
    base = np.array([10.5, 10.5, 10.5])
    # second point 0.1 m away is rejected (min distance 0.15, strict >), third at 0.2 m accepted
    pts = np.array([base, base + [0.1, 0, 0], base + [0.2, 0, 0]])
    assert m.add_points(pts) == 2
    # exactly at the threshold distance is rejected (sq_dist > min^2 is strict)
    m3 = O.OracleMap()
    assert m3.add_points(np.array([[2.25, 2.5, 2.5], [2.5, 2.5, 2.5]]), min_distance_points=0.25) == 1
    # cap: a voxel never holds more than max_num_points_in_voxel, later points are dropped
    m4 = O.OracleMap()
    grid = np.stack(np.meshgrid(np.arange(5), np.arange(5), np.arange(2), indexing="ij"), -1).reshape(-1, 3) * 0.19 + 20.01
    assert grid.shape[0] == 50
    added = m4.add_points(grid)
    assert added == 20 and m4.num_points == 20 and m4.num_voxels == 1
    keys, counts, xyz = m4.snapshot()
    assert np.array_equal(xyz[0, :20], grid[:20].astype(np.float32))   # insertion order is kept
    # order dependence: reversing the offer order changes which points survive
    m5 = O.OracleMap()
    m5.add_points(grid[::-1].copy())
    _, _, xyz5 = m5.snapshot()
    assert not np.array_equal(np.sort(xyz5[0], axis=0), np.sort(xyz[0], axis=0))
    # min_num_points > 0: absent voxels are never created (src/lioOptimization.cpp:437)
    m6 = O.OracleMap()
    assert m6.add_points(rng.random((10, 3)) + 40, min_num_points=1) == 0 and m6.num_voxels == 0


# ---- searchNeighbors vs brute force (src/optimize.cpp:365-426) ----------------------------------------------
def _brute_force_knn(keys, counts, xyz, p, nb, K, size=1.0):
    kd = {tuple(k): i for i, k in enumerate(keys.tolist())}
    c = np.trunc(p / size).astype(int)
    cand = []
    for dx in range(-nb, nb + 1):
        for dy in range(-nb, nb + 1):
            for dz in range(-nb, nb + 1):
                k = (c[0] + dx, c[1] + dy, c[2] + dz)
                if k in kd:
                    v = kd[k]
                    for i in range(counts[v]):
                        m = xyz[v, i].astype(np.float64)
                        d = m - p
                        d2 = d[0] * d[0] + (d[1] * d[1] + d[2] * d[2])
                        cand.append((np.sqrt(d2), len(cand), k, i))
    cand.sort(key=lambda t: (t[0], t[1]))
    return cand[:K], len(cand)


@pytest.mark.parametrize("frame_id", [100, 5])   # nb=1 steady state, nb=2 init mode
def test_search_neighbors_equals_brute_force(small_world, frame_id):
    om, sw = small_world["omap"], small_world["sweep"]
    prm = O.r3live_params(max_num_residuals=BIG, frame_id=frame_id)
    n = 400
    r = om.build_plane_residuals(sw.raw_xyz[:n], sw.q_init, sw.t_init, sw.t_last, prm, debug=True)
    assert r.num_fragile == 0
    keys, counts, xyz = om.snapshot()
    nb = 2 if frame_id < 20 else 1
    for k in range(0, n, 7):
        best, C = _brute_force_knn(keys, counts, xyz, r.world_xyz[k], nb, 20)
        assert C == r.num_candidates[k]
        if C < 20:
            assert r.status[k] == 0
            continue
        assert r.status[k] >= 1
        for j, (d, _, vk, i) in enumerate(best):
            assert tuple(r.nbr[k, j, :3]) == vk and r.nbr[k, j, 3] == i
            assert r.nbr_dist[k, j] == d


def test_keypoint_transform_and_quaternion(small_world):
    sw = small_world["sweep"]
    q = sw.q_init * 1.7   # un-normalised on purpose
    R = O.quat_to_rot(q)
    # toRotationMatrix of an un-normalised quaternion is not a rotation: R^T R = |q|^4 I for this formula family
    qn = q / np.linalg.norm(q)
    Rn = O.quat_to_rot(qn)
    assert np.allclose(Rn @ Rn.T, np.eye(3), atol=1e-14)
    assert np.allclose(Rn, synth.quat_to_rot(qn), atol=1e-15)
    v = np.array([0.3, -1.2, 2.0])
    # q v q^-1
    def qmul(a, b):
        return synth.quat_mul(a, b)
    qv = qmul(qmul(qn, np.array([*v, 0.0])), np.array([-qn[0], -qn[1], -qn[2], qn[3]]))
    assert np.allclose(Rn @ v, qv[:3], atol=1e-14)
    assert not np.allclose(R, Rn)


# ---- computeNeighborhoodDistribution: eigen solver vs numpy (src/optimize.cpp:316-353) ------------------------
def test_eig3_against_numpy_eigh():
    rng = np.random.default_rng(1)
    for trial in range(300):
        if trial % 3 == 0:      # planar cloud (the common case)
            P = rng.normal(size=(20, 3)) * np.array([0.4, 0.3, 0.005 * (1 + trial % 7)])
            Rm = np.linalg.qr(rng.normal(size=(3, 3)))[0]
            P = P @ Rm.T
        elif trial % 3 == 1:    # generic
            P = rng.normal(size=(20, 3))
        else:                   # axis-aligned (exercises the v1norm2<=tol branch) / tiny scale
            P = rng.normal(size=(20, 3)) * np.array([1.0, 0.1, 0.0]) * 1e-3
        P = P - P.mean(0)
        S = P.T @ P
        ev, evec = O.eig3_sym(S)
        w, V = np.linalg.eigh(S)
        assert np.allclose(ev, w, rtol=1e-10, atol=1e-14 * max(1.0, np.abs(w).max()))
        assert np.all(np.diff(ev) >= 0)
        for i in range(3):
            r = S @ evec[:, i] - ev[i] * evec[:, i]
            assert np.linalg.norm(r) <= 1e-13 * max(np.abs(S).max(), 1e-300) * 10
        assert np.allclose(evec.T @ evec, np.eye(3), atol=1e-13)


def test_plane_params_match_numpy_recomputation(small_world):
    om, sw = small_world["omap"], small_world["sweep"]
    prm = O.r3live_params(max_num_residuals=BIG)
    n = 600
    r = om.build_plane_residuals(sw.raw_xyz[:n], sw.q_init, sw.t_init, sw.t_last, prm, debug=True)
    keys, counts, xyz = om.snapshot()
    kd = {tuple(k): i for i, k in enumerate(keys.tolist())}
    Rq = synth.quat_to_rot(sw.q_init)   # q_init is unit: R == R'
    checked = 0
    for k in np.nonzero(r.status >= 1)[0][::5]:
        nb = np.array([xyz[kd[tuple(r.nbr[k, j, :3])], r.nbr[k, j, 3]] for j in range(20)], np.float64)
        mu = nb.mean(0)
        S = (nb - mu).T @ (nb - mu)
        w, V = np.linalg.eigh(S)
        n_np = V[:, 0]
        b = sw.raw_xyz[k]
        if n_np @ (sw.t_last - b) < 0:
            n_np = -n_np
        a2D = (np.sqrt(abs(w[1])) - np.sqrt(abs(w[0]))) / np.sqrt(abs(w[2]))
        p = r.world_xyz[k]
        wgt = 0.9 * a2D ** 2 + 0.1 * np.exp(-np.linalg.norm(nb[0] - p) / (0.3 * 20))
        dist = n_np @ (Rq @ b + sw.t_init) - n_np @ nb[0]
        pl = r.plane[k]
        assert np.allclose(pl[3:6], n_np, atol=1e-8)
        assert abs(pl[15] - a2D) < 1e-8 and abs(pl[14] - wgt) < 1e-8 and abs(pl[13] - dist) < 1e-8
        assert (r.status[k] == 2) == (dist < 0.3)          # signed gate, no abs (src/optimize.cpp:98)
        if r.status[k] == 2:
            J = np.concatenate([wgt * n_np, -wgt * (n_np @ Rq @ np.array([[0, -b[2], b[1]], [b[2], 0, -b[0]], [-b[1], b[0], 0]]))])
            assert np.allclose(pl[6:12], J, atol=1e-7)
        checked += 1
    assert checked > 50


def test_jacobian_is_the_derivative_of_the_weighted_distance(small_world):
    """Central finite differences of w * distance w.r.t. (dp, dtheta): pose perturbed as t+dp, R*Exp(dtheta) with the
    plane (normal, offset, weight) held fixed — the linearisation the reference uses (src/optimize.cpp:100-101)."""
    om, sw = small_world["omap"], small_world["sweep"]
    prm = O.r3live_params(max_num_residuals=BIG)
    n = 300
    r = om.build_plane_residuals(sw.raw_xyz[:n], sw.q_init, sw.t_init, sw.t_last, prm, debug=True)
    Rq = synth.quat_to_rot(sw.q_init)
    eps = 1e-6
    ks = np.nonzero(r.status == 2)[0][:40]
    assert ks.size >= 20
    for k in ks:
        nrm, off, wgt = r.plane[k, 3:6], r.plane[k, 12], r.plane[k, 14]
        b = sw.raw_xyz[k]

        def f(dp, dth):
            Rp = Rq @ synth.quat_to_rot(synth.quat_from_rotvec(dth))
            return wgt * (nrm @ (Rp @ b + sw.t_init + dp) + off)
        num = np.zeros(6)
        for a in range(3):
            e = np.zeros(3); e[a] = eps
            num[a] = (f(e, np.zeros(3)) - f(-e, np.zeros(3))) / (2 * eps)
            num[3 + a] = (f(np.zeros(3), e) - f(np.zeros(3), -e)) / (2 * eps)
        assert np.allclose(num, r.plane[k, 6:12], rtol=1e-5, atol=1e-6)


def test_normal_equations_are_the_sum_of_rows(small_world):
    om, sw = small_world["omap"], small_world["sweep"]
    prm = O.r3live_params(max_num_residuals=BIG)
    r = om.build_plane_residuals(sw.raw_xyz, sw.q_init, sw.t_init, sw.t_last, prm, debug=True)
    acc = r.status == 2
    J = r.plane[acc, 6:12]
    h = r.plane[acc, 13] * r.plane[acc, 14]
    assert np.allclose(r.HTH, J.T @ J, rtol=1e-12)
    assert np.allclose(r.HTh, J.T @ h, rtol=1e-11, atol=1e-9)
    assert np.isclose(r.loss_sum, (r.plane[acc, 13] ** 2).sum(), rtol=1e-12)    # unweighted (src/optimize.cpp:104)
    assert r.num_residuals == acc.sum() and r.num_full_neighborhoods == (r.status >= 1).sum()
    assert np.allclose(r.HTH, r.HTH.T) and np.all(np.linalg.eigvalsh(r.HTH) > -1e-6 * np.abs(r.HTH).max())
    # threaded variant (private sums, fixed-order combine) agrees to rounding
    r4 = om.build_plane_residuals(sw.raw_xyz, sw.q_init, sw.t_init, sw.t_last, prm, nthreads=4)
    assert np.allclose(r4.HTH, r.HTH, rtol=1e-12) and r4.num_residuals == r.num_residuals
    assert r4.sum_candidates == r.sum_candidates


# ---- the residual cap and its quirks (src/optimize.cpp:99,107; include/parameters.h:40) -----------------------
def test_max_num_residuals_cap_in_keypoint_order(small_world):
    om, sw = small_world["omap"], small_world["sweep"]
    full = om.build_plane_residuals(sw.raw_xyz, sw.q_init, sw.t_init, sw.t_last, O.r3live_params(max_num_residuals=BIG), debug=True)
    cap = om.build_plane_residuals(sw.raw_xyz, sw.q_init, sw.t_init, sw.t_last, O.r3live_params(max_num_residuals=600), debug=True)
    assert cap.num_residuals == 600
    kstar = np.nonzero(full.status == 2)[0][599]
    assert np.array_equal(cap.status[:kstar + 1], full.status[:kstar + 1])
    assert np.all(cap.status[kstar + 1:] == -1) and cap.num_visited == kstar + 1
    # compiled default -1: the loop stops after the first keypoint with a full neighbourhood
    d = om.build_plane_residuals(sw.raw_xyz, sw.q_init, sw.t_init, sw.t_last, O.r3live_params(max_num_residuals=-1), debug=True)
    first_full = np.nonzero(full.status >= 1)[0][0]
    assert d.num_visited == first_full + 1 and d.num_residuals == int(full.status[first_full] == 2)
    assert not d.success   # < min_number_neighbors residuals -> summary.success = false (src/optimize.cpp:110)


def test_threaded_refuses_when_cap_binds(small_world):
    om, sw = small_world["omap"], small_world["sweep"]
    with pytest.raises(RuntimeError):
        om.build_plane_residuals(sw.raw_xyz, sw.q_init, sw.t_init, sw.t_last, O.r3live_params(max_num_residuals=600), nthreads=2)


# ---- ESIKF algebra vs numpy (src/optimize.cpp:172-310, src/eskfEstimator.cpp:219-230) ------------------------
def _so3_exp(w):
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * K @ K


def test_mat17_inverse_and_observe():
    rng = np.random.default_rng(5)
    A = rng.normal(size=(17, 17)) + 17 * np.eye(17)
    ok, Ai = O.mat17_inverse(A)
    assert ok and np.allclose(Ai, np.linalg.inv(A), rtol=1e-10, atol=1e-12)
    e = O.Eskf(p=np.array([1.0, 2, 3]), q=synth.quat_from_rotvec([0.1, -0.2, 0.3]), g=np.array([0.1, -0.2, 9.8]))
    dx = rng.normal(size=17) * 0.01
    out = e.observe(dx)
    assert np.allclose(out.p, e.p + dx[:3]) and np.allclose(out.v, dx[6:9]) and np.allclose(out.ba, dx[9:12])
    R_new = synth.quat_to_rot(out.q)
    assert np.allclose(R_new, synth.quat_to_rot(e.q) @ _so3_exp(dx[3:6]), atol=1e-12)
    assert abs(np.linalg.norm(out.q) - 1) < 1e-15
    assert np.isclose(np.linalg.norm(out.g), np.linalg.norm(e.g), rtol=1e-12)    # gravity stays on its sphere
    # small-angle branch (theta < 1e-4)
    out2 = e.observe(np.concatenate([np.zeros(3), [1e-6, -2e-6, 3e-6], np.zeros(11)]))
    assert np.allclose(synth.quat_to_rot(out2.q), synth.quat_to_rot(e.q) @ _so3_exp(np.array([1e-6, -2e-6, 3e-6])), atol=1e-13)


def test_update_iekf_first_pass_matches_numpy(small_world):
    """First pass: current == predicted, so d_x_new = 0 and the update is d_x = -(P^-1*cov + H^T H)^-1 H^T h."""
    om, sw = small_world["omap"], small_world["sweep"]
    prm = O.r3live_params(max_num_residuals=BIG)
    P = synth.prior_covariance()
    e0 = O.Eskf(p=sw.t_init.copy(), q=sw.q_init.copy(), cov=P.copy())
    res = om.update_iekf(sw.raw_xyz, e0, sw.t_last, prm)
    r = om.build_plane_residuals(sw.raw_xyz, sw.q_init, sw.t_init, sw.t_last, prm)
    T = np.linalg.inv(P / 0.001)
    T[:6, :6] += r.HTH
    Ti = np.linalg.inv(T)
    dx = -Ti[:, :6] @ r.HTh
    assert np.allclose(res["trace"][0, :17], dx, rtol=1e-8, atol=1e-12)
    assert res["success"] and 2 <= res["passes"] <= 6
    # the estimate moves toward the true pose
    assert np.linalg.norm(res["eskf"].p - sw.t_true) < 0.2 * np.linalg.norm(sw.t_init - sw.t_true)
    # posterior covariance is symmetric-ish, smaller than the prior on the pose block
    C = res["eskf"].cov
    assert np.all(np.diag(C)[:6] < np.diag(P)[:6]) and np.all(np.diag(C) > 0)


def test_update_iekf_pass_count_rules(small_world):
    om, sw = small_world["omap"], small_world["sweep"]
    e0 = O.Eskf(p=sw.t_init.copy(), q=sw.q_init.copy(), cov=synth.prior_covariance())
    # frame_id <= 1: never "converged" (src/optimize.cpp:265) -> runs all num_iters_icp + 1 passes... but frame_id < init_num_frames
    # also forces max(15, num_iters) iterations and nb=2 (src/optimize.cpp:21,135)
    r = om.update_iekf(sw.raw_xyz[:800], e0, sw.t_last, O.r3live_params(max_num_residuals=BIG, frame_id=1))
    assert r["passes"] == 16
    r2 = om.update_iekf(sw.raw_xyz[:800], e0, sw.t_last, O.r3live_params(max_num_residuals=BIG, frame_id=100,
                                                                         threshold_translation_norm=0.0))
    assert r2["passes"] == 6     # i = -1..4, posterior covariance on the last
    # too few residuals -> success false after the first pass
    r3 = om.update_iekf(sw.raw_xyz[:5], e0, sw.t_last, O.r3live_params(max_num_residuals=BIG))
    assert not r3["success"] and r3["passes"] == 1


# ---- gridSampling / subSampleFrame (src/utility.cpp:167-201) -------------------------------------------------------
def test_grid_sampling_keeps_the_first_point_of_every_cell(small_world):
    sw = small_world["sweep"]
    w = synth.registered_points(sw)
    for size in (1.5, 0.4):
        idx = O.grid_sampling(w, size)
        cells = np.trunc(w / size).astype(np.int64)           # static_cast<short>: truncation toward zero
        uniq, first = np.unique(cells, axis=0, return_index=True)
        assert len(idx) == len(uniq) == len(set(idx.tolist()))
        assert set(idx.tolist()) == set(first.tolist())       # n.second[0] = first point pushed into the cell
    # the order is the hash container's, not the frame's (which is why it has to be reproduced, not re-invented)
    assert not np.array_equal(idx, np.sort(idx))
    assert O.grid_sampling(np.zeros((0, 3)), 1.5).size == 0


# ---- row N3: undistortion / point transforms (src/utility.cpp:203-332) -------------------------------------------
def _np_quat_to_rot(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _np_quat_mul(a, b):
    ax, ay, az, aw = a; bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz])


make_imu_states, make_sweep_times = synth.make_imu_states, synth.make_sweep_times


def test_slerp_matches_the_closed_form():
    rng = np.random.default_rng(1)
    for _ in range(50):
        a = rng.normal(size=4); a /= np.linalg.norm(a)
        b = rng.normal(size=4); b /= np.linalg.norm(b)
        t = rng.uniform()
        d = float(a @ b)
        th = np.arccos(abs(d))
        s0, s1 = np.sin((1 - t) * th) / np.sin(th), np.sin(t * th) / np.sin(th)
        ref = s0 * a + (s1 if d >= 0 else -s1) * b
        assert np.allclose(O.quat_slerp(a, t, b), ref, rtol=0, atol=1e-14)
    a = np.array([0, 0, 0, 1.0])
    assert np.allclose(O.quat_slerp(a, 0.3, a), a)                      # |dot| >= 1 - eps: linear branch
    assert np.allclose(O.quat_slerp(a, 0.3, -a), 0.7 * a + 0.3 * a)     # opposite hemisphere: scale1 flips sign


def test_distort_frame_by_constant_against_numpy():
    st = make_imu_states()
    rng = np.random.default_rng(9)
    raw = rng.normal(0, 20, (500, 3)); rel = make_sweep_times(500)
    R_il = _np_quat_to_rot(np.array([0.01, 0.02, -0.03, 0.9993]) / np.linalg.norm([0.01, 0.02, -0.03, 0.9993])); t_il = np.array([0.05, -0.02, 0.1])
    got = O.distort_frame_by_constant(raw, rel, st, st[0]["timestamp"], R_il, t_il)
    t0, t1 = st[0]["timestamp"], st[-1]["timestamp"]
    for i in range(500):
        tp = t0 + rel[i] / 1000.0
        if abs(tp - t0) < 1e-6: tp = t0 + 1e-6
        if abs(tp - t1) < 1e-6: tp = t1 - 1e-6
        a = min(max((tp - t0) / (t1 - t0), 0.0), 1.0)
        q = O.quat_slerp(st[0]["quat"], a, st[-1]["quat"])
        ref = _np_quat_to_rot(q) @ (R_il @ raw[i] + t_il) + (1 - a) * st[0]["trans"] + a * st[-1]["trans"]
        assert np.allclose(got[i], ref, rtol=1e-13, atol=1e-12)


def test_distort_frame_by_imu_against_numpy_and_its_iterator_semantics():
    st = make_imu_states()
    rng = np.random.default_rng(10)
    raw = rng.normal(0, 20, (400, 3)); rel = make_sweep_times(400)
    t0 = st[0]["timestamp"]
    got, m = O.distort_frame_by_imu(raw, rel, st, t0)
    assert m == 400
    ts = np.array([s["timestamp"] for s in st])
    for i in range(400):
        tp = t0 + rel[i] / 1000.0
        k = next(k for k in range(len(st) - 1) if tp > ts[k] - 1e-6 and tp < ts[k + 1] + 1e-6)
        if abs(tp - ts[k]) < 1e-6: tp = ts[k] + 1e-6
        if abs(tp - ts[k + 1]) < 1e-6: tp = ts[k + 1] - 1e-6
        dt = tp - ts[k]
        w = st[k + 1]["un_gyr"] * dt; th = np.linalg.norm(w)
        dq = np.r_[w / 2, 1.0] if th < 1e-4 else np.r_[w / th * np.sin(th / 2), np.cos(th / 2)]
        q = _np_quat_mul(st[k]["quat"], dq / np.linalg.norm(dq)); q /= np.linalg.norm(q)
        ref = _np_quat_to_rot(q) @ raw[i] + st[k]["trans"] + st[k]["vel"] * dt + 0.5 * st[k + 1]["un_acc"] * dt * dt
        assert np.allclose(got[i], ref, rtol=1e-12, atol=1e-11), i
    # a point that is earlier than the interval the walk has reached stops it: the rest keep their old values
    rel2 = rel.copy(); rel2[100] = 1.0
    keep = np.full_like(raw, 7.0)
    got2, m2 = O.distort_frame_by_imu(raw, rel2, st, t0, imu_xyz_in=keep)
    assert m2 == 100 and np.array_equal(got2[:100], got[:100]) and np.all(got2[100:] == 7.0)
    # a point after the last IMU stamp: same
    rel3 = rel.copy(); rel3[250:] += 500.0
    got3, m3 = O.distort_frame_by_imu(raw, rel3, st, t0, imu_xyz_in=keep)
    assert m3 == 250 and np.all(got3[250:] == 7.0)
    # fewer than two states: the outer loop never runs
    assert O.distort_frame_by_imu(raw, rel, st[:1], t0)[1] == 0


def test_transform_all_imu_point_inverts_the_end_pose():
    st = make_imu_states()
    rng = np.random.default_rng(11)
    raw = rng.normal(0, 20, (300, 3))
    R_il = _np_quat_to_rot(np.array([0.0, 0.0, 0.0, 1.0])); t_il = np.array([0.05, -0.02, 0.1])
    last = st[-1]
    imu = (_np_quat_to_rot(last["quat"]) @ (R_il @ raw.T + t_il[:, None])).T + last["trans"]   # transformPoint with the end pose
    back = O.transform_all_imu_point(imu, last, R_il, t_il)
    assert np.allclose(back, raw, rtol=1e-12, atol=1e-11)


def test_remove_points_far_from_location_by_first_point():
    pts = synth.sample_map_points(40.0, 40.0, seed=8)
    om = O.OracleMap(); om.add_points(pts)
    keys, counts, xyz = om.snapshot()
    loc = np.array([3.0, -2.0, 1.0]); dist = 12.0
    first = xyz[:, 0, :].astype(np.float64)
    d = first - loc
    far = (d[:, 0] ** 2 + (d[:, 1] ** 2 + d[:, 2] ** 2)) > dist * dist
    removed = om.remove_far(loc, dist)
    assert removed == int(far.sum()) and 0 < removed < keys.shape[0]
    k2, c2, x2 = om.snapshot()
    left = {tuple(k): (c, x) for k, c, x in zip(k2.tolist(), c2, x2)}
    want = {tuple(k): (c, x) for k, c, x, f in zip(keys.tolist(), counts, xyz, far) if not f}
    assert left.keys() == want.keys() and all(left[k][0] == want[k][0] and np.array_equal(left[k][1], want[k][1]) for k in left)
    assert om.num_points == int(counts[~far].sum())
