"""ctypes binding of oracle/_ref/libsrl_reference.so — the REFERENCE's own LIO sources compiled unmodified from where they
lie under /root/reference, over the stand-in headers in oracle/shim/ (oracle/Makefile target `reference`,
oracle/srl_reference_harness.cpp).

TEST INFRASTRUCTURE ONLY: it exists to pin the oracle restatement (tests/test_reference_pin.py) and as the single-thread
"reference as written" CPU timing.  The product package sr_livo_b200 never imports this.  What it pins: the reference's
control flow, containers, casts and quirks.  What it cannot pin: the arithmetic inside Eigen / OpenCV calls, which
oracle/shim restates.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .oracle_py import Eskf, EskfState, IcpParams, _f64, _ptr, imu_states_array

_HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(_HERE, "_ref", "libsrl_reference.so")
# the same reference objects with updateIEKF / addPointsToMap replaced at link time by the product's C++ adapter (the maintainer
# patch of INTEGRATION.md section 2: oracle/srl_reference_gpu_patch.cpp); needs a GPU at run time
PATH_GPU = os.path.join(_HERE, "_ref", "libsrl_reference_gpu.so")
# sensitivity build: 3-term reductions of the Eigen stand-in in the other plausible order, (c0 + c1) + c2
PATH_PACKET = os.path.join(_HERE, "_ref", "libsrl_reference_packet.so")
_libs = {}


def _path(gpu):
    return PATH_PACKET if gpu == "packet" else (PATH_GPU if gpu else PATH)


def available(gpu=False) -> bool:
    return os.path.exists(_path(gpu))


def lib(gpu=False):
    if gpu in _libs:
        return _libs[gpu]
    if not available(gpu):
        raise RuntimeError("oracle/_ref/libsrl_reference*.so missing: run `make -C oracle` where /root/reference exists")
    L = C.CDLL(_path(gpu))
    if gpu is True:
        L.refgpu_map_points.argtypes = [C.c_void_p]
        L.refgpu_map_points.restype = C.c_int64
        L.refgpu_map_is_on_gpu.argtypes = [C.c_void_p]
        L.refgpu_map_is_on_gpu.restype = C.c_int32
        L.refgpu_release.argtypes = [C.c_void_p]
    P, I64, I32, D = C.c_void_p, C.c_int64, C.c_int32, C.c_double
    L.ref_build_info.restype = C.c_char_p
    L.ref_create.restype = P
    L.ref_destroy.argtypes = [P]
    L.ref_voxel_hash.argtypes = [C.c_int16] * 3
    L.ref_voxel_hash.restype = C.c_uint64
    L.ref_laser_point_cov.argtypes = [P]
    L.ref_laser_point_cov.restype = D
    for f in ("ref_map_num_voxels", "ref_map_num_points"):
        getattr(L, f).argtypes = [P, I32]
        getattr(L, f).restype = I64
    L.ref_map_clear.argtypes = [P, I32]
    L.ref_map_load.argtypes = [P, P, P, P, I64, I32]
    L.ref_add_points_to_map.argtypes = [P, P, I64, D, I32, D, I32, D, I32, D, I32, D, D, I32]
    L.ref_add_points_to_map.restype = I64
    L.ref_map_remove_far.argtypes = [P, P, D]
    L.ref_map_remove_far.restype = I64
    L.ref_map_snapshot.argtypes = [P, I32, I32] + [P] * 9
    L.ref_map_snapshot.restype = I64
    L.ref_search_neighbors.argtypes = [P, P, I32, D, I32, I32, P, P]
    L.ref_search_neighbors.restype = I32
    L.ref_neighborhood.argtypes = [P, P, I32, P]
    L.ref_neighborhood.restype = I32
    L.ref_build_plane_residuals.argtypes = [P, P, I64, P, P, P, P, P, C.POINTER(IcpParams), C.POINTER(I32), C.POINTER(I32),
                                            C.POINTER(D), P, P]
    L.ref_build_plane_residuals.restype = I32
    L.ref_update_iekf.argtypes = [P, P, I64, C.POINTER(EskfState), P, P, P, P, P, C.POINTER(IcpParams), C.POINTER(I32), C.POINTER(I32)]
    L.ref_update_iekf.restype = I32
    L.ref_update_iekf_many.argtypes = [P, P, I64, I32, P, P, P, P, P, P, C.POINTER(IcpParams), I32]
    L.ref_update_iekf_many.restype = I32
    L.ref_stream_reset.argtypes = [P]
    L.ref_stream_push.argtypes = [P, P, I64, I32, P, P, C.POINTER(EskfState), P, P, C.POINTER(IcpParams), P, P, C.POINTER(I32), C.POINTER(I32), P, P, P]
    L.ref_stream_push.restype = I32
    L.ref_optimize.argtypes = [P, P, P, I64, D, C.POINTER(EskfState), P, P, P, P, P, C.POINTER(IcpParams), C.POINTER(I32), C.POINTER(I32)]
    L.ref_optimize.restype = I32
    L.ref_eskf_observe.argtypes = [C.POINTER(EskfState), P]
    L.ref_grid_sampling.argtypes = [P, I64, D, P]
    L.ref_grid_sampling.restype = I64
    L.ref_distort_frame_by_constant.argtypes = [P, P, I64, P, I64, D, P, P, P]
    L.ref_distort_frame_by_imu.argtypes = [P, P, I64, P, I64, D, P, P, P]
    L.ref_transform_all_imu_point.argtypes = [P, I64, P, P, P, P]
    L.ref_transform_point.argtypes = [P, I64, P, P, P, P, P]
    for f in ("ref_color_num_rgb_points", "ref_color_num_recent", "ref_color_num_new_recent"):
        getattr(L, f).argtypes = [P]
        getattr(L, f).restype = I64
    L.ref_color_lists.argtypes = [P, P, P]
    L.ref_color_render.argtypes = [P, P, P, I32, I32, D]
    L.ref_color_render.restype = I64
    _libs[gpu] = L
    return L


def _ext(R_il, t_il):
    return _f64(np.eye(3) if R_il is None else R_il).reshape(9), _f64(np.zeros(3) if t_il is None else t_il)


def voxel_hash(x, y, z) -> int:
    return int(lib().ref_voxel_hash(int(x), int(y), int(z)))


class Reference:
    """One lioOptimization object of the reference (its real constructor over the stub ROS NodeHandle)."""

    def __init__(self, gpu=False):
        """gpu=True: the link-time patched library (updateIEKF and addPointsToMap run on the GPU through the product's adapter);
        gpu="packet": the sensitivity build (packet-first 3-term reductions in the Eigen stand-in)."""
        self._L = lib(gpu)
        self._gpu = gpu is True
        self._h = C.c_void_p(self._L.ref_create())

    def __del__(self):
        try:
            if self._h:
                if self._gpu:
                    self._L.refgpu_release(self._h)
                self._L.ref_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def gpu_map_points(self) -> int:
        return int(self._L.refgpu_map_points(self._h))

    def gpu_map_is_on_gpu(self) -> bool:
        return bool(self._L.refgpu_map_is_on_gpu(self._h))

    # ---- maps (which: 0 = voxel_map, 1 = color_voxel_map)
    def num_voxels(self, which=0) -> int:
        return int(self._L.ref_map_num_voxels(self._h, which))

    def num_points(self, which=0) -> int:
        return int(self._L.ref_map_num_points(self._h, which))

    def load(self, keys, counts, xyz):
        keys = np.ascontiguousarray(keys, np.int16); counts = np.ascontiguousarray(counts, np.int32); xyz = np.ascontiguousarray(xyz, np.float32)
        self._L.ref_map_load(self._h, _ptr(keys), _ptr(counts), _ptr(xyz), keys.shape[0], xyz.shape[1])

    def add_points_to_map(self, world_xyz, voxel_size=1.0, max_num_points_in_voxel=20, min_distance_points=0.15, min_num_points=0,
                          color_voxel_size=0.1, color_max_points=20, color_min_distance=0.01, add_point_step=4,
                          time_sweep_end=1.0, time_last_process=-1e5, to_rendering=False) -> int:
        """lioOptimization::addPointsToMap (src/lioOptimization.cpp:520-554): LIO map and colour map in one call."""
        xyz = _f64(world_xyz).reshape(-1, 3)
        return int(self._L.ref_add_points_to_map(self._h, _ptr(xyz), xyz.shape[0], voxel_size, max_num_points_in_voxel, min_distance_points,
                                               min_num_points, color_voxel_size, color_max_points, color_min_distance, add_point_step,
                                               time_sweep_end, time_last_process, 1 if to_rendering else 0))

    def remove_far(self, location, distance) -> int:
        loc = _f64(location).reshape(3)
        return int(self._L.ref_map_remove_far(self._h, _ptr(loc), float(distance)))

    def snapshot(self, which=0, cap=20, color=False):
        nv = self.num_voxels(which)
        out = dict(keys=np.zeros((nv, 3), np.int16), counts=np.zeros(nv, np.int32), xyz=np.zeros((nv, cap, 3), np.float32))
        extra = [None] * 6
        if color:
            out.update(rgb=np.zeros((nv, cap, 3), np.int16), n_rgb=np.zeros((nv, cap), np.int16), cov=np.zeros((nv, cap, 3), np.float32),
                       obs_dist=np.zeros((nv, cap)), last_obs=np.zeros((nv, cap)), last_visited=np.zeros(nv))
            extra = [_ptr(out[k]) for k in ("rgb", "n_rgb", "cov", "obs_dist", "last_obs", "last_visited")]
        got = self._L.ref_map_snapshot(self._h, which, cap, _ptr(out["keys"]), _ptr(out["counts"]), _ptr(out["xyz"]), *extra)
        assert got == nv
        return out

    # ---- the path
    def search_neighbors(self, point, nb=1, size=1.0, K=20, thr=1):
        p = _f64(point).reshape(3)
        xyz = np.zeros((K, 3)); vox = np.zeros((K, 3), np.int16)
        m = self._L.ref_search_neighbors(self._h, _ptr(p), nb, size, K, thr, _ptr(xyz), _ptr(vox))
        return xyz[:m].copy(), vox[:m].copy()

    def neighborhood(self, pts):
        pts = _f64(pts).reshape(-1, 3)
        out = np.zeros(16)
        rc = self._L.ref_neighborhood(self._h, _ptr(pts), pts.shape[0], _ptr(out))
        return rc, dict(center=out[0:3], normal=out[3:6], covariance=out[6:15].reshape(3, 3), a2D=out[15])

    def build_plane_residuals(self, raw_xyz, q_cur, t_cur, t_last, params: IcpParams, R_il=None, t_il=None):
        raw = _f64(raw_xyz).reshape(-1, 3)
        n = raw.shape[0]
        R, ti = _ext(R_il, t_il)
        q, t, tl = _f64(q_cur), _f64(t_cur), _f64(t_last)
        ok, used, loss = C.c_int32(0), C.c_int32(0), C.c_double(0)
        rows = np.zeros((n, 15)); world = np.zeros((n, 3))
        m = self._L.ref_build_plane_residuals(self._h, _ptr(raw), n, _ptr(q), _ptr(t), _ptr(tl), _ptr(R), _ptr(ti), C.byref(params),
                                            C.byref(ok), C.byref(used), C.byref(loss), _ptr(rows), _ptr(world))
        if m < 0:
            return dict(threw=True)
        return dict(threw=False, success=bool(ok.value), num_residuals_used=used.value, loss_sum=loss.value, rows=rows[:m].copy(), world_xyz=world)

    def update_iekf(self, raw_xyz, eskf: Eskf, t_last, params: IcpParams, frame_q=None, frame_t=None, R_il=None, t_il=None):
        raw = _f64(raw_xyz).reshape(-1, 3)
        fq = _f64(eskf.q if frame_q is None else frame_q).copy(); ft = _f64(eskf.p if frame_t is None else frame_t).copy()
        R, ti = _ext(R_il, t_il)
        tl = _f64(t_last)
        st = eskf.to_c()
        ok, used = C.c_int32(0), C.c_int32(0)
        rc = self._L.ref_update_iekf(self._h, _ptr(raw), raw.shape[0], C.byref(st), _ptr(fq), _ptr(ft), _ptr(tl), _ptr(R), _ptr(ti),
                                   C.byref(params), C.byref(ok), C.byref(used))
        return dict(threw=rc < 0, success=bool(ok.value), num_residuals_used=used.value, eskf=Eskf.from_c(st), frame_q=fq, frame_t=ft)

    def update_iekf_many(self, raws, eskfs, t_lasts, params: IcpParams, n_threads: int, R_il=None, t_il=None):
        """`len(raws)` independent sweeps (same keypoint count) against this object's voxel_map on n_threads host threads, each through
        the reference's single-threaded updateIEKF (ref_update_iekf_many).  Returns (sweeps that succeeded, list of Eskf, frame_q, frame_t)."""
        raw = np.ascontiguousarray(np.stack([_f64(r).reshape(-1, 3) for r in raws]))
        ns, n = raw.shape[0], raw.shape[1]
        st = (EskfState * ns)(*[e.to_c() for e in eskfs])
        fq = np.ascontiguousarray(np.stack([_f64(e.q) for e in eskfs])); ft = np.ascontiguousarray(np.stack([_f64(e.p) for e in eskfs]))
        tl = np.ascontiguousarray(np.stack([_f64(t) for t in t_lasts]))
        R, ti = _ext(R_il, t_il)
        ok = self._L.ref_update_iekf_many(self._h, _ptr(raw), n, ns, C.cast(st, C.c_void_p), _ptr(fq), _ptr(ft), _ptr(tl), _ptr(R), _ptr(ti),
                                        C.byref(params), int(n_threads))
        return int(ok), [Eskf.from_c(s) for s in st], fq, ft

    def optimize(self, frame_world, frame_raw, sample_voxel_size, eskf: Eskf, t_last, params: IcpParams, frame_q=None, frame_t=None,
                 R_il=None, t_il=None):
        world = _f64(frame_world).reshape(-1, 3).copy(); raw = _f64(frame_raw).reshape(-1, 3)
        fq = _f64(eskf.q if frame_q is None else frame_q).copy(); ft = _f64(eskf.p if frame_t is None else frame_t).copy()
        R, ti = _ext(R_il, t_il)
        tl = _f64(t_last)
        st = eskf.to_c()
        ok, used = C.c_int32(0), C.c_int32(0)
        rc = self._L.ref_optimize(self._h, _ptr(world), _ptr(raw), raw.shape[0], float(sample_voxel_size), C.byref(st), _ptr(fq), _ptr(ft), _ptr(tl),
                                _ptr(R), _ptr(ti), C.byref(params), C.byref(ok), C.byref(used))
        return dict(threw=rc < 0, success=bool(ok.value), num_residuals_used=used.value, eskf=Eskf.from_c(st), frame_q=fq, frame_t=ft, world=world)

    # ---- the caller: stateEstimation over a stream of sweeps
    def stream_reset(self):
        self._L.ref_stream_reset(self._h)

    def stream_push(self, raw_xyz, index_frame, q_pred, t_pred, eskf: Eskf, params: IcpParams, R_il=None, t_il=None, init_voxel_size=0.2,
                    init_sample_voxel_size=1.0, voxel_size=0.5, sample_voxel_size=1.5, min_distance_points=0.1, init_num_frames=20,
                    max_num_points_in_voxel=20):
        """One sweep through lioOptimization::stateEstimation (src/lioOptimization.cpp:983-1035): optimize() for frames after the
        first, then addPointsToMap; the odometryOptions defaults are include/parameters.h:58-95."""
        raw = _f64(raw_xyz).reshape(-1, 3)
        R, ti = _ext(R_il, t_il)
        qp, tp = _f64(q_pred), _f64(t_pred)
        st = eskf.to_c()
        odo_d = _f64([init_voxel_size, init_sample_voxel_size, voxel_size, sample_voxel_size, min_distance_points])
        odo_i = np.ascontiguousarray([init_num_frames, max_num_points_in_voxel], np.int32)
        ok, used = C.c_int32(0), C.c_int32(0)
        fq, ft, world = np.zeros(4), np.zeros(3), np.zeros_like(raw)
        rc = self._L.ref_stream_push(self._h, _ptr(raw), raw.shape[0], int(index_frame), _ptr(qp), _ptr(tp), C.byref(st), _ptr(R), _ptr(ti),
                                     C.byref(params), _ptr(odo_d), _ptr(odo_i), C.byref(ok), C.byref(used), _ptr(fq), _ptr(ft), _ptr(world))
        return dict(threw=rc < 0, success=bool(ok.value), num_residuals_used=used.value, eskf=Eskf.from_c(st), frame_q=fq, frame_t=ft, world=world)

    # ---- colour map
    def color_counts(self):
        L = self._L
        return dict(voxels=self.num_voxels(1), rgb_points=int(L.ref_color_num_rgb_points(self._h)), recent=int(L.ref_color_num_recent(self._h)),
                    new_recent=int(L.ref_color_num_new_recent(self._h)))

    def color_lists(self):
        c = self.color_counts()
        rgb_points = np.zeros((c["rgb_points"], 4), np.int16); recent = np.zeros((c["recent"], 3), np.int32)
        self._L.ref_color_lists(self._h, _ptr(rgb_points), _ptr(recent))
        return rgb_points, recent

    def color_render(self, cam15, image_bgr, obs_time) -> int:
        cam = _f64(cam15).reshape(15); img = np.ascontiguousarray(image_bgr, np.uint8)
        return int(self._L.ref_color_render(self._h, _ptr(cam), _ptr(img), img.shape[0], img.shape[1], float(obs_time)))


def eskf_observe(e: Eskf, dx) -> Eskf:
    st = e.to_c(); d = _f64(dx)
    lib().ref_eskf_observe(C.byref(st), _ptr(d))
    return Eskf.from_c(st)


def grid_sampling(xyz, size) -> np.ndarray:
    xyz = _f64(xyz).reshape(-1, 3)
    out = np.zeros(xyz.shape[0], np.int32)
    m = lib().ref_grid_sampling(_ptr(xyz), xyz.shape[0], float(size), _ptr(out))
    return out[:m].copy()


def distort_frame_by_constant(raw_xyz, relative_time_ms, states, time_frame_begin, R_il=None, t_il=None):
    raw = _f64(raw_xyz).reshape(-1, 3); rel = _f64(relative_time_ms).reshape(-1); st = imu_states_array(states)
    R, t = _ext(R_il, t_il)
    out = np.zeros_like(raw)
    lib().ref_distort_frame_by_constant(_ptr(raw), _ptr(rel), raw.shape[0], _ptr(st), st.shape[0], float(time_frame_begin), _ptr(R), _ptr(t), _ptr(out))
    return out


def distort_frame_by_imu(raw_xyz, relative_time_ms, states, time_frame_begin, R_il=None, t_il=None, imu_xyz_in=None):
    raw = _f64(raw_xyz).reshape(-1, 3); rel = _f64(relative_time_ms).reshape(-1); st = imu_states_array(states)
    R, t = _ext(R_il, t_il)
    out = np.zeros_like(raw) if imu_xyz_in is None else _f64(imu_xyz_in).reshape(-1, 3).copy()
    lib().ref_distort_frame_by_imu(_ptr(raw), _ptr(rel), raw.shape[0], _ptr(st), st.shape[0], float(time_frame_begin), _ptr(R), _ptr(t), _ptr(out))
    return out


def transform_all_imu_point(imu_xyz, last_state, R_il=None, t_il=None):
    imu = _f64(imu_xyz).reshape(-1, 3); st = imu_states_array([last_state])
    R, t = _ext(R_il, t_il)
    out = np.zeros_like(imu)
    lib().ref_transform_all_imu_point(_ptr(imu), imu.shape[0], _ptr(st), _ptr(R), _ptr(t), _ptr(out))
    return out


def transform_point(raw_xyz, q_end, t_end, R_il=None, t_il=None):
    raw = _f64(raw_xyz).reshape(-1, 3)
    R, t = _ext(R_il, t_il)
    q, te = _f64(q_end), _f64(t_end)
    out = np.zeros_like(raw)
    lib().ref_transform_point(_ptr(raw), raw.shape[0], _ptr(q), _ptr(te), _ptr(R), _ptr(t), _ptr(out))
    return out
