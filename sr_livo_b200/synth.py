"""Seeded synthetic worlds, maps and sweeps for the LIO scan-matching hot path.

World = union of planes so neighbourhoods are planar:
  * ground height-field  z = 0.02 (sin 0.3x + cos 0.2y)
  * square buildings on a 40 m street grid (footprint [40i+10.37, 40i+29.63]^2, 15 m high walls)
Every sample carries +-1 mm Gaussian jitter (kills lattice ties).  Wall planes sit at non-integer
coordinates so a wall does not straddle a voxel boundary.

Map points are *offered* in a seeded random order and pushed through the map-insert semantics of the
reference (src/lioOptimization.cpp:400-446: 1 m voxels, cap 20, min distance 0.15 m) by whoever builds
the map (GPU `srl_map_insert`, or the CPU oracle in tests).

Sweeps: Livox-style non-repetitive rosette in a 70.4 x 77.2 deg FoV, or a 128-ring spinning pattern,
ray-cast against the analytic world, 2 cm range noise, expressed in the LiDAR frame of a seeded true pose.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

PITCH = 40.0
B_LO = 10.37
B_HI = 29.63
B_H = 15.0


def ground_h(x, y):
    return 0.02 * (np.sin(0.3 * x) + np.cos(0.2 * y))


def rng_for(seed: int) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64(int(seed)))


# ----------------------------------------------------------------------------------------------
# map samples
# ----------------------------------------------------------------------------------------------
def world_surface_area(extent: float) -> float:
    nb = int(np.floor(extent / PITCH))
    side = B_HI - B_LO
    return extent * extent - nb * nb * side * side + nb * nb * 4 * side * B_H


def sample_map_points(extent: float, density: float, seed: int = 1, center=(0.0, 0.0), jitter: float = 1e-3,
                      chunk: int | None = None) -> np.ndarray:
    """Offer `density` samples per m^2 of every surface inside the square of side `extent` centred at
    `center` (rounded to whole 40 m cells), in seeded random order.  Returns (M,3) float64."""
    rng = rng_for(seed)
    nb = max(1, int(np.floor(extent / PITCH)))
    extent = nb * PITCH
    x0 = center[0] - extent / 2.0
    y0 = center[1] - extent / 2.0
    # snap the origin onto the street grid so buildings sit at 40i + [10.37, 29.63]
    x0 = np.floor(x0 / PITCH) * PITCH
    y0 = np.floor(y0 / PITCH) * PITCH
    side = B_HI - B_LO
    # ground (rejection of building footprints)
    n_g = int(extent * extent * density)
    gx = x0 + rng.random(n_g) * extent
    gy = y0 + rng.random(n_g) * extent
    fx = np.mod(gx, PITCH)
    fy = np.mod(gy, PITCH)
    keep = ~((fx > B_LO) & (fx < B_HI) & (fy > B_LO) & (fy < B_HI))
    gx, gy = gx[keep], gy[keep]
    gz = ground_h(gx, gy)
    ground = np.stack([gx, gy, gz], axis=1)
    # walls
    n_w = int(nb * nb * 4 * side * B_H * density)
    bi = rng.integers(0, nb, n_w)
    bj = rng.integers(0, nb, n_w)
    face = rng.integers(0, 4, n_w)
    u = B_LO + rng.random(n_w) * side
    h = rng.random(n_w) * B_H
    ox = x0 + bi * PITCH
    oy = y0 + bj * PITCH
    wx = np.where(face == 0, ox + B_LO, np.where(face == 1, ox + B_HI, ox + u))
    wy = np.where(face == 2, oy + B_LO, np.where(face == 3, oy + B_HI, oy + u))
    wz = ground_h(wx, wy) + h
    walls = np.stack([wx, wy, wz], axis=1)
    pts = np.concatenate([ground, walls], axis=0)
    pts += rng.normal(0.0, jitter, pts.shape)
    perm = rng.permutation(pts.shape[0])
    return np.ascontiguousarray(pts[perm])


# ----------------------------------------------------------------------------------------------
# ray casting against the analytic world
# ----------------------------------------------------------------------------------------------
def raycast(origin: np.ndarray, dirs: np.ndarray, max_range: float = 120.0) -> np.ndarray:
    """Range along each unit direction to the first surface (np.inf if none within max_range)."""
    o = np.asarray(origin, np.float64)
    d = np.asarray(dirs, np.float64)
    n = d.shape[0]
    # ground: fixed-point on the (tiny-amplitude) height field
    t_g = np.full(n, np.inf)
    down = d[:, 2] < -1e-3
    s = np.where(down, (0.0 - o[2]) / np.where(down, d[:, 2], -1.0), np.inf)
    for _ in range(4):
        hx = o[0] + s * d[:, 0]
        hy = o[1] + s * d[:, 1]
        s = np.where(down, (ground_h(np.where(down, hx, 0.0), np.where(down, hy, 0.0)) - o[2]) /
                     np.where(down, d[:, 2], -1.0), np.inf)
    t_g = np.where(down & (s > 0), s, np.inf)
    # buildings: march over 40 m cells, slab-test the cell's box
    t_w = np.full(n, np.inf)
    eps = 1e-9
    dx = np.where(np.abs(d[:, 0]) < eps, eps, d[:, 0])
    dy = np.where(np.abs(d[:, 1]) < eps, eps, d[:, 1])
    t_cur = np.zeros(n)
    active = np.ones(n, bool)
    for _ in range(12):
        px = o[0] + (t_cur + 1e-6) * dx
        py = o[1] + (t_cur + 1e-6) * dy
        ci = np.floor(px / PITCH)
        cj = np.floor(py / PITCH)
        bx0, bx1 = ci * PITCH + B_LO, ci * PITCH + B_HI
        by0, by1 = cj * PITCH + B_LO, cj * PITCH + B_HI
        tx0, tx1 = (bx0 - o[0]) / dx, (bx1 - o[0]) / dx
        ty0, ty1 = (by0 - o[1]) / dy, (by1 - o[1]) / dy
        t_en = np.maximum(np.minimum(tx0, tx1), np.minimum(ty0, ty1))
        t_ex = np.minimum(np.maximum(tx0, tx1), np.maximum(ty0, ty1))
        hit = active & (t_en < t_ex) & (t_en > 0)
        hz = o[2] + t_en * d[:, 2]
        hxx = o[0] + t_en * dx
        hyy = o[1] + t_en * dy
        ok = hit & (hz < ground_h(hxx, hyy) + B_H) & np.isinf(t_w)
        t_w = np.where(ok, t_en, t_w)
        # leave the cell
        cx_next = np.where(dx > 0, (ci + 1) * PITCH, ci * PITCH)
        cy_next = np.where(dy > 0, (cj + 1) * PITCH, cj * PITCH)
        t_leave = np.minimum((cx_next - o[0]) / dx, (cy_next - o[1]) / dy)
        t_cur = np.maximum(t_leave, t_cur + 1e-6)
        active = active & np.isinf(t_w) & (t_cur < max_range)
        if not active.any():
            break
    t = np.minimum(t_g, t_w)
    return np.where(t < max_range, t, np.inf)


def livox_dirs(n_rays: int, rng: np.random.Generator) -> np.ndarray:
    """Non-repetitive rosette in a 70.4 (H) x 77.2 (V) degree FoV, boresight +x (sensor frame)."""
    phase = rng.random() * 2 * np.pi
    tau = phase + np.arange(n_rays) * (2 * np.pi * 0.61803398875 / 97.0)
    r = np.cos(3.0 * np.sqrt(2.0) * tau + 0.3)
    u = r * np.cos(tau)
    v = r * np.sin(tau)
    az = np.deg2rad(35.2) * u + rng.normal(0, 1e-4, n_rays)
    el = np.deg2rad(38.6) * v + rng.normal(0, 1e-4, n_rays)
    ce = np.cos(el)
    return np.stack([ce * np.cos(az), ce * np.sin(az), np.sin(el)], axis=1)


def spinning_dirs(n_rings: int, n_az: int, rng: np.random.Generator) -> np.ndarray:
    """128-ring style spinning LiDAR: 360 x 45 deg (+-22.5), ring-major order per azimuth step."""
    el = np.deg2rad(np.linspace(-22.5, 22.5, n_rings))
    az = np.linspace(0, 2 * np.pi, n_az, endpoint=False) + rng.random() * 1e-3
    A, E = np.meshgrid(az, el, indexing="ij")
    A = A.reshape(-1)
    E = E.reshape(-1)
    ce = np.cos(E)
    return np.stack([ce * np.cos(A), ce * np.sin(A), np.sin(E)], axis=1)


# ----------------------------------------------------------------------------------------------
# rotations (host-side helpers for generators and tests; q = (x, y, z, w))
# ----------------------------------------------------------------------------------------------
def quat_from_rotvec(rv) -> np.ndarray:
    rv = np.asarray(rv, np.float64)
    th = np.linalg.norm(rv)
    if th < 1e-12:
        q = np.array([rv[0] / 2, rv[1] / 2, rv[2] / 2, 1.0])
    else:
        a = rv / th
        q = np.array([a[0] * np.sin(th / 2), a[1] * np.sin(th / 2), a[2] * np.sin(th / 2), np.cos(th / 2)])
    return q / np.linalg.norm(q)


def quat_mul(a, b) -> np.ndarray:
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx, aw * bw - ax * bx - ay * by - az * bz])


def quat_to_rot(q) -> np.ndarray:
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


@dataclass
class Sweep:
    raw_xyz: np.ndarray      # (N,3) float64, LiDAR frame (R_il = I, t_il = 0 => also body frame)
    q_true: np.ndarray       # (4,) x,y,z,w
    t_true: np.ndarray       # (3,)
    q_init: np.ndarray       # perturbed estimate handed to the optimiser
    t_init: np.ndarray
    t_last: np.ndarray       # previous frame translation (normal orientation test, src/optimize.cpp:49)
    seed: int


def make_sweep(n_points: int, seed: int = 1000, pattern: str = "livox", position=(0.0, 3.0, 1.8), yaw: float = 0.0,
               max_range: float = 100.0, range_noise: float = 0.02, pose_seed: int = 7,
               dp_max: float = 0.1, dth_max_deg: float = 1.0) -> Sweep:
    rng = rng_for(seed)
    tilt = rng.normal(0.0, np.deg2rad(1.0), 2)
    q_true = quat_mul(quat_from_rotvec([0, 0, yaw]), quat_from_rotvec([tilt[0], tilt[1], 0.0]))
    t_true = np.asarray(position, np.float64)
    R = quat_to_rot(q_true)
    out = []
    have = 0
    tries = 0
    while have < n_points and tries < 8:
        want = int((n_points - have) * 1.8) + 1024
        if pattern == "livox":
            ds = livox_dirs(want, rng)
        elif pattern == "spinning":
            n_az = int(np.ceil(want / 128))
            ds = spinning_dirs(128, n_az, rng)
        else:
            raise ValueError(pattern)
        dw = ds @ R.T
        rg = raycast(t_true, dw, max_range)
        ok = np.isfinite(rg) & (rg > 0.5)
        rr = rg[ok] + rng.normal(0.0, range_noise, int(ok.sum()))
        out.append(ds[ok] * rr[:, None])
        have += int(ok.sum())
        tries += 1
    raw = np.concatenate(out, axis=0)[:n_points]
    prng = rng_for(pose_seed * 1000003 + seed)
    dp = prng.uniform(-dp_max, dp_max, 3)
    dth = np.deg2rad(prng.uniform(-dth_max_deg, dth_max_deg, 3))
    q_init = quat_mul(q_true, quat_from_rotvec(dth))
    t_init = t_true + dp
    t_last = t_true - R @ np.array([1.0, 0.0, 0.0])   # previous frame 1 m behind along the boresight
    return Sweep(np.ascontiguousarray(raw), q_true, t_true, q_init, t_init, t_last, seed)


def registered_points(sweep: Sweep, q=None, t=None) -> np.ndarray:
    """World-frame points of a sweep under pose (q,t) (what addPointsToMap receives, src/optimize.cpp:441-445)."""
    q = sweep.q_true if q is None else q
    t = sweep.t_true if t is None else t
    return sweep.raw_xyz @ quat_to_rot(np.asarray(q)).T + np.asarray(t)


def prior_covariance() -> np.ndarray:
    """eskfEstimator ctor identity with the tryInit scalings (src/eskfEstimator.cpp:7,74-76)."""
    P = np.eye(17)
    P[9:12, 9:12] *= 0.001
    P[12:15, 12:15] *= 0.0001
    P[15:17, 15:17] *= 0.00001
    return P


# ---- synthetic IMU track and point times for the undistortion functions (row N3) ---------------------------------
def _quat_mul_xyzw(a, b):
    ax, ay, az, aw = a; bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz])


def make_imu_states(n_states=21, t0=100.0, dt=0.005, seed=3):
    """A smooth synthetic IMU track (200 Hz over a 0.1 s sweep) with every field the functions read."""
    rng = np.random.default_rng(seed)
    states = []
    q = np.array([0.02, -0.01, 0.3, 0.95]); q /= np.linalg.norm(q)
    p = np.array([1.0, -2.0, 0.5]); v = np.array([2.0, 0.3, -0.1])
    for k in range(n_states):
        gyr = rng.normal(0, 0.4, 3); acc = rng.normal(0, 1.5, 3)
        states.append(dict(timestamp=t0 + k * dt, quat=q.copy(), trans=p.copy(), vel=v.copy(), un_acc=acc, un_gyr=gyr))
        th = np.linalg.norm(gyr * dt)
        dq = np.r_[gyr * dt / th * np.sin(th / 2), np.cos(th / 2)] if th > 0 else np.array([0, 0, 0, 1.0])
        q = _quat_mul_xyzw(q, dq); q /= np.linalg.norm(q)
        p = p + v * dt + 0.5 * acc * dt * dt; v = v + acc * dt
    return states


def make_sweep_times(n, span_ms=100.0, seed=5):
    rng = np.random.default_rng(seed)
    rel = np.sort(rng.uniform(0.0, span_ms, n))
    rel[0] = 0.0; rel[-1] = span_ms                      # exactly on the sweep's first / last IMU stamp
    rel[n // 2] = 50.0; rel[n // 3] = 35.0 + 4e-4        # on and within 1e-6 s of interior IMU stamps
    return np.sort(rel)
