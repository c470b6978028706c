"""Host-side mirror of the reference's interface for the scan-matching path, over the C ABI.

Names and argument meaning follow the reference so the parity tests read like its call sites:

  voxelHashMap                          include/cloudMap.h:171          -> VoxelHashMap
  lioOptimization::addPointsToMap       src/lioOptimization.cpp:520-554 -> LioOptimization.addPointsToMap
  lioOptimization::mapSize              src/lioOptimization.cpp:574-581 -> LioOptimization.mapSize
  lioOptimization::buildPlaneResiduals  src/optimize.cpp:18-131         -> LioOptimization.buildPlaneResiduals
  lioOptimization::updateIEKF           src/optimize.cpp:133-314        -> LioOptimization.updateIEKF
  lioOptimization::optimize             src/optimize.cpp:428-448        -> LioOptimization.optimize (keypoints given)
  eskfEstimator (state + observe)       src/eskfEstimator.cpp           -> EskfEstimator

Error behaviour: optimizeSummary.success=false <-> OptimizeSummary.success False (SRL_TOO_FEW_RESIDUALS);
the reference's `throw std::runtime_error("error")` on NaN planarity <-> RuntimeError; everything else raises
SrlError.  No CPU fallback exists: without the CUDA library/GPU every call fails loudly.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import capi
from .capi import (DebugOut, EskfState, Frame, IcpParams, IekfSummary, NormalEq, SrlError, f64, lib, ptr,
                   r3live_params)

NS = capi.NS


def _check(ctx, rc, ok=(capi.SRL_OK,)):
    if rc not in ok:
        raise SrlError(rc, lib().srl_last_error(ctx).decode() if ctx else "")
    return rc


class Context:
    """srl_ctx: one per host thread / GPU. `stream` may be a raw cudaStream_t (e.g. torch's current stream)."""

    def __init__(self, device: int = 0, stream: int | None = None):
        h = C.c_void_p()
        rc = lib().srl_ctx_create(device, C.c_void_p(stream) if stream else None, C.byref(h))
        if rc != capi.SRL_OK:
            raise SrlError(rc, "srl_ctx_create failed: no usable CUDA device (this path has no CPU fallback)")
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            lib().srl_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        _check(self.h, lib().srl_ctx_synchronize(self.h))

    @property
    def kernel_launches(self) -> int:
        return int(lib().srl_ctx_kernel_launches(self.h))

    def set_option(self, name: str, value: int):
        _check(self.h, lib().srl_ctx_set_option(self.h, name.encode(), int(value)))

    def counter(self, name: str) -> int:
        v = C.c_int64(0)
        _check(self.h, lib().srl_ctx_get_counter(self.h, name.encode(), C.byref(v)))
        return v.value

    def set_timing(self, enable: bool = True):
        _check(self.h, lib().srl_ctx_set_timing(self.h, 1 if enable else 0))

    def pass_time(self, reset: bool = False):
        """(summed k1_assoc device time in ms, launches) measured with CUDA events on the ctx stream."""
        ms, n = C.c_double(0), C.c_int64(0)
        _check(self.h, lib().srl_ctx_pass_time(self.h, C.byref(ms), C.byref(n), 1 if reset else 0))
        return ms.value, n.value


class VoxelHashMap:
    """HBM-resident voxelHashMap (include/cloudMap.h:171)."""

    def __init__(self, ctx: Context, voxel_size: float = 1.0, max_num_points_in_voxel: int = 20,
                 max_voxels: int = 1 << 20):
        self.ctx = ctx
        self.cap = max_num_points_in_voxel
        self.voxel_size = voxel_size
        h = C.c_void_p()
        _check(ctx.h, lib().srl_map_create(ctx.h, voxel_size, max_num_points_in_voxel, max_voxels, C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            lib().srl_map_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def clear(self):
        _check(self.ctx.h, lib().srl_map_clear(self.h))

    def remove_far(self, location, distance: float) -> int:
        """removePointsFarFromLocation (src/lioOptimization.cpp:556-572): number of voxels erased."""
        loc = f64(location).reshape(3)
        n = C.c_int64(0)
        _check(self.ctx.h, lib().srl_map_remove_far(self.h, ptr(loc), float(distance), C.byref(n)))
        return int(n.value)

    def stats(self):
        nv, npts = C.c_int64(0), C.c_int64(0)
        _check(self.ctx.h, lib().srl_map_stats(self.h, C.byref(nv), C.byref(npts)))
        return nv.value, npts.value

    def upload(self, keys, counts, xyz):
        keys = np.ascontiguousarray(keys, np.int16).reshape(-1, 3)
        counts = np.ascontiguousarray(counts, np.int32)
        xyz = np.ascontiguousarray(xyz, np.float32).reshape(keys.shape[0], self.cap, 3)
        _check(self.ctx.h, lib().srl_map_upload(self.h, ptr(keys), ptr(counts), ptr(xyz), keys.shape[0]))

    def download(self):
        nv, _ = self.stats()
        keys = np.zeros((nv, 3), np.int16)
        counts = np.zeros(nv, np.int32)
        xyz = np.zeros((nv, self.cap, 3), np.float32)
        got = C.c_int64(0)
        _check(self.ctx.h, lib().srl_map_download(self.h, ptr(keys), ptr(counts), ptr(xyz), nv, C.byref(got)))
        return keys, counts, xyz

    def insert(self, xyz_world, min_distance_points: float = 0.15, min_num_points: int = 0) -> int:
        xyz = f64(xyz_world).reshape(-1, 3)
        added = C.c_int64(0)
        _check(self.ctx.h, lib().srl_map_insert(self.h, ptr(xyz), xyz.shape[0], min_distance_points, min_num_points,
                                                C.byref(added)))
        return added.value

    def insert_device(self, d_ptr: int, n: int, min_distance_points: float = 0.15, min_num_points: int = 0) -> int:
        added = C.c_int64(0)
        _check(self.ctx.h, lib().srl_map_insert_device(self.h, C.c_void_p(d_ptr), n, min_distance_points,
                                                       min_num_points, C.byref(added)))
        return added.value


class ColorVoxelMap:
    """color_voxel_map + hashmap_3d_points + rgb_points_vec + voxels_recent_visited (include/lioOptimization.h:275-291,
    include/rgbMapTracker.h:38), fed by the colour branch of addPointsToMap and coloured by renderPointsInRecentVoxel."""

    def __init__(self, ctx: Context, voxel_size: float = 1.0, max_num_points_in_voxel: int = 20, max_voxels: int = 1 << 16,
                 min_distance_points: float = 0.15):
        self.ctx, self.cap = ctx, max_num_points_in_voxel
        h = C.c_void_p()
        _check(ctx.h, lib().srl_color_map_create(ctx.h, voxel_size, max_num_points_in_voxel, max_voxels, min_distance_points, C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            lib().srl_color_map_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def stats(self) -> dict:
        v = [C.c_int64(0) for _ in range(5)]
        _check(self.ctx.h, lib().srl_color_map_stats(self.h, *[C.byref(x) for x in v]))
        return dict(zip(("voxels", "points", "rgb_points", "recent", "new_recent"), [x.value for x in v]))

    def addPoints(self, xyz_world, add_point_step: int = 1, time_sweep_end: float = 1.0, time_last_process: float = 0.0,
                  to_rendering: bool = True) -> int:
        """the loop of src/lioOptimization.cpp:533-551 over the registered frame (host array or a device pointer + n)."""
        xyz = f64(xyz_world).reshape(-1, 3)
        stored = C.c_int64(0)
        _check(self.ctx.h, lib().srl_color_map_add_points(self.h, ptr(xyz), xyz.shape[0], add_point_step, time_sweep_end, time_last_process,
                                                          1 if to_rendering else 0, C.byref(stored)))
        return stored.value

    def renderPointsInRecentVoxel(self, camera: "capi.Camera", image_bgr, obs_time: float) -> int:
        img = np.ascontiguousarray(image_bgr, np.uint8)
        assert img.shape == (camera.rows, camera.cols, 3)
        n = C.c_int64(0)
        _check(self.ctx.h, lib().srl_color_map_render_recent(self.h, C.byref(camera), ptr(img), float(obs_time), C.byref(n)))
        return n.value

    def download(self) -> dict:
        """voxel contents + colour state (block order) + the two lists."""
        st = self.stats()
        nv, cap = st["voxels"], self.cap
        out = dict(keys=np.zeros((nv, 3), np.int16), counts=np.zeros(nv, np.int32), xyz=np.zeros((nv, cap, 3), np.float32),
                   rgb=np.zeros((nv, cap, 3), np.int16), n_rgb=np.zeros((nv, cap), np.int16), cov=np.zeros((nv, cap, 3), np.float32),
                   obs_dist=np.zeros((nv, cap)), last_obs=np.zeros((nv, cap)), last_visited=np.zeros(nv),
                   rgb_points=np.zeros((st["rgb_points"], 4), np.int16), recent=np.zeros((st["recent"], 3), np.int16))
        vox = C.c_void_p(lib().srl_color_map_voxels(self.h))
        got = C.c_int64(0)
        _check(self.ctx.h, lib().srl_map_download(vox, ptr(out["keys"]), ptr(out["counts"]), ptr(out["xyz"]), nv, C.byref(got)))
        _check(self.ctx.h, lib().srl_color_map_download_state(self.h, nv, ptr(out["rgb"]), ptr(out["n_rgb"]), ptr(out["cov"]), ptr(out["obs_dist"]),
                                                              ptr(out["last_obs"]), ptr(out["last_visited"])))
        _check(self.ctx.h, lib().srl_color_map_download_lists(self.h, ptr(out["rgb_points"]), ptr(out["recent"])))
        return out


class Sweep:
    """The keypoints of one reconstructed sweep, resident in HBM (raw LiDAR-frame points, FP64)."""

    def __init__(self, ctx: Context, capacity: int):
        self.ctx = ctx
        self.capacity = capacity
        self.n = 0
        h = C.c_void_p()
        _check(ctx.h, lib().srl_sweep_create(ctx.h, capacity, C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            lib().srl_sweep_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, raw_xyz):
        raw = f64(raw_xyz).reshape(-1, 3)
        _check(self.ctx.h, lib().srl_sweep_upload(self.h, ptr(raw), raw.shape[0]))
        self.n = raw.shape[0]

    def set_device(self, d_ptr: int, n: int):
        _check(self.ctx.h, lib().srl_sweep_set_device(self.h, C.c_void_p(d_ptr), n))
        self.n = n

    def set_shard(self, begin: int, end: int):
        _check(self.ctx.h, lib().srl_sweep_set_shard(self.h, begin, end))


@dataclass
class EskfEstimator:
    """eskfEstimator state (src/eskfEstimator.cpp:3-21); q = (x, y, z, w)."""
    p: np.ndarray = field(default_factory=lambda: np.zeros(3))
    q: np.ndarray = field(default_factory=lambda: np.array([0.0, 0.0, 0.0, 1.0]))
    v: np.ndarray = field(default_factory=lambda: np.zeros(3))
    ba: np.ndarray = field(default_factory=lambda: np.zeros(3))
    bg: np.ndarray = field(default_factory=lambda: np.zeros(3))
    g: np.ndarray = field(default_factory=lambda: np.array([0.0, 0.0, 9.81]))
    cov: np.ndarray = field(default_factory=lambda: np.eye(NS))

    def to_c(self) -> EskfState:
        return capi.eskf_to_c(self.p, self.q, self.v, self.ba, self.bg, self.g, self.cov)

    @staticmethod
    def from_c(s: EskfState) -> "EskfEstimator":
        return EskfEstimator(**capi.eskf_from_c(s))

    def observe(self, d_x) -> "EskfEstimator":
        """eskfEstimator::observe (src/eskfEstimator.cpp:219-230)."""
        s = self.to_c()
        d = f64(d_x)
        rc = lib().srl_eskf_observe(C.byref(s), ptr(d))
        if rc != capi.SRL_OK:
            raise SrlError(rc, "srl_eskf_observe")
        return EskfEstimator.from_c(s)


@dataclass
class OptimizeSummary:
    """optimizeSummary (include/lioOptimization.h) + what the GPU pass reports."""
    success: bool
    num_residuals_used: int
    passes_run: int = 1
    converged: bool = False
    trace: np.ndarray | None = None


@dataclass
class PlaneResiduals:
    HTH: np.ndarray
    HTh: np.ndarray
    loss_sum: float
    num_residuals: int
    num_full_neighborhoods: int
    num_candidates_scanned: int
    success: bool
    world_xyz: np.ndarray | None = None
    status: np.ndarray | None = None
    nbr: np.ndarray | None = None
    nbr_dist: np.ndarray | None = None
    plane: np.ndarray | None = None


def make_frame(q_cur, t_cur, t_last, R_il=None, t_il=None) -> Frame:
    fr = Frame()
    for name, val, n in (("q_cur", q_cur, 4), ("t_cur", t_cur, 3), ("t_last", t_last, 3),
                         ("R_il", np.eye(3) if R_il is None else R_il, 9),
                         ("t_il", np.zeros(3) if t_il is None else t_il, 3)):
        a = f64(val).reshape(-1)
        assert a.size == n, name
        for i in range(n):
            getattr(fr, name)[i] = a[i]
    return fr


class LioOptimization:
    """The scan-matching members of `class lioOptimization` (include/lioOptimization.h:334-353) on one GPU."""

    def __init__(self, device: int = 0, stream: int | None = None, max_voxels: int = 1 << 20,
                 sweep_capacity: int = 1 << 17, size_voxel_map: float = 1.0, max_num_points_in_voxel: int = 20,
                 R_imu_lidar=None, t_imu_lidar=None):
        self.ctx = Context(device, stream)
        self.voxel_map = VoxelHashMap(self.ctx, size_voxel_map, max_num_points_in_voxel, max_voxels)
        self.sweep = Sweep(self.ctx, sweep_capacity)
        self.R_imu_lidar = np.eye(3) if R_imu_lidar is None else f64(R_imu_lidar).reshape(3, 3)
        self.t_imu_lidar = np.zeros(3) if t_imu_lidar is None else f64(t_imu_lidar)
        self.eskf_pro = EskfEstimator()

    def close(self):
        self.sweep.close()
        self.voxel_map.close()
        self.ctx.close()

    # ---- src/lioOptimization.cpp:520-554 (voxel_size is the map's own)
    def addPointsToMap(self, points_world, min_distance_points: float = 0.15, min_num_points: int = 0) -> int:
        return self.voxel_map.insert(points_world, min_distance_points, min_num_points)

    def addSweepToMap(self, frame_q, frame_t, min_distance_points: float = 0.15, min_num_points: int = 0) -> int:
        """stateEstimation's tail (src/lioOptimization.cpp:1027): the resident sweep, re-transformed with the final pose
        (src/optimize.cpp:441-445), goes into the map without leaving the device."""
        q, t = f64(frame_q), f64(frame_t)
        R, ti = f64(self.R_imu_lidar).reshape(9), f64(self.t_imu_lidar)
        added = C.c_int64(0)
        _check(self.ctx.h, lib().srl_map_insert_sweep(self.voxel_map.h, self.sweep.h, ptr(q), ptr(t), ptr(R), ptr(ti),
                                                      min_distance_points, min_num_points, C.byref(added)))
        return added.value

    # ---- src/lioOptimization.cpp:574-581
    def mapSize(self) -> int:
        return self.voxel_map.stats()[1]

    # ---- src/lioOptimization.cpp:556-572
    def removePointsFarFromLocation(self, location, distance: float) -> int:
        return self.voxel_map.remove_far(location, distance)

    # ---- src/utility.cpp:188-201, called at src/optimize.cpp:431
    def gridSampling(self, points_world, size_voxel_subsampling: float) -> np.ndarray:
        """Frame indices of the keypoints (first point of every cell), in the reference's order."""
        xyz = f64(points_world).reshape(-1, 3)
        out = np.zeros(xyz.shape[0], np.uint32)
        m = C.c_size_t(0)
        _check(self.ctx.h, lib().srl_grid_sampling(self.ctx.h, ptr(xyz), xyz.shape[0], size_voxel_subsampling, ptr(out), C.byref(m)))
        return out[:m.value].copy()

    # ---- src/utility.cpp:203-332 (row N3): undistortion and the sweep-end re-expression
    @staticmethod
    def _imu_states(states):
        """states: sequence of dicts / objects with timestamp, quat (x,y,z,w), trans, vel, un_acc, un_gyr."""
        from .capi import ImuState
        arr = (ImuState * len(states))()
        for a, s_ in zip(arr, states):
            g = (lambda k: s_[k]) if isinstance(s_, dict) else (lambda k: getattr(s_, k))
            a.timestamp = float(g("timestamp"))
            for name, m in (("quat", 4), ("trans", 3), ("vel", 3), ("un_acc", 3), ("un_gyr", 3)):
                v = np.asarray(g(name), np.float64).reshape(m)
                getattr(a, name)[:] = v.tolist()
        return arr

    def distortFrameByConstant(self, raw_xyz, relative_time_ms, imu_states, time_frame_begin: float) -> np.ndarray:
        raw = f64(raw_xyz).reshape(-1, 3)
        rel = f64(relative_time_ms).reshape(-1)
        st = self._imu_states(imu_states)
        out = np.zeros_like(raw)
        R, t = f64(self.R_imu_lidar).reshape(9), f64(self.t_imu_lidar)
        _check(self.ctx.h, lib().srl_distort_frame_by_constant(self.ctx.h, ptr(raw), ptr(rel), raw.shape[0], C.cast(st, C.c_void_p),
                                                               len(imu_states), float(time_frame_begin), ptr(R), ptr(t), ptr(out)))
        return out

    def distortFrameByImu(self, raw_xyz, relative_time_ms, imu_states, time_frame_begin: float, imu_xyz_in=None):
        """Returns (imu_xyz, n_written); imu_xyz_in supplies the values kept by points the reference's walk never reaches."""
        raw = f64(raw_xyz).reshape(-1, 3)
        rel = f64(relative_time_ms).reshape(-1)
        st = self._imu_states(imu_states)
        out = np.zeros_like(raw) if imu_xyz_in is None else f64(imu_xyz_in).reshape(-1, 3).copy()
        nw = C.c_int64(0)
        R, t = f64(self.R_imu_lidar).reshape(9), f64(self.t_imu_lidar)
        _check(self.ctx.h, lib().srl_distort_frame_by_imu(self.ctx.h, ptr(raw), ptr(rel), raw.shape[0], C.cast(st, C.c_void_p),
                                                          len(imu_states), float(time_frame_begin), ptr(R), ptr(t), ptr(out), C.byref(nw)))
        return out, int(nw.value)

    def transformAllImuPoint(self, imu_xyz, last_imu_state) -> np.ndarray:
        imu = f64(imu_xyz).reshape(-1, 3)
        st = self._imu_states([last_imu_state])
        out = np.zeros_like(imu)
        R, t = f64(self.R_imu_lidar).reshape(9), f64(self.t_imu_lidar)
        _check(self.ctx.h, lib().srl_transform_all_imu_point(self.ctx.h, ptr(imu), imu.shape[0], C.cast(st, C.c_void_p), ptr(R), ptr(t), ptr(out)))
        return out

    def setKeypoints(self, raw_xyz):
        """std::vector<point3D> keypoints (raw_point members), uploaded once per sweep."""
        self.sweep.upload(raw_xyz)

    # ---- src/optimize.cpp:18-131 (+ :160-170,:235,:239)
    def buildPlaneResiduals(self, cur_icp_options: IcpParams, q_cur, t_cur, t_last, debug: bool = False) -> PlaneResiduals:
        fr = make_frame(q_cur, t_cur, t_last, self.R_imu_lidar, self.t_imu_lidar)
        ne = NormalEq()
        n = self.sweep.n
        K = cur_icp_options.max_number_neighbors
        arrs = {}
        dbg = None
        if debug:
            arrs = dict(world_xyz=np.zeros((n, 3)), status=np.zeros(n, np.int32), nbr=np.zeros((n, K, 4), np.int16),
                        nbr_dist=np.zeros((n, K)), plane=np.zeros((n, 16)))
            dbg = DebugOut(*[ptr(arrs[k]) for k in ("world_xyz", "status", "nbr", "nbr_dist", "plane")])
        rc = lib().srl_build_plane_residuals(self.ctx.h, self.voxel_map.h, self.sweep.h, C.byref(fr),
                                             C.byref(cur_icp_options), C.byref(ne),
                                             C.byref(dbg) if dbg is not None else None)
        if rc == capi.SRL_NAN_PLANARITY:
            raise RuntimeError("error")   # src/optimize.cpp:348-350
        _check(self.ctx.h, rc, ok=(capi.SRL_OK, capi.SRL_TOO_FEW_RESIDUALS))
        return PlaneResiduals(HTH=np.array(ne.HTH).reshape(6, 6), HTh=np.array(ne.HTh), loss_sum=ne.loss_sum,
                              num_residuals=ne.num_residuals, num_full_neighborhoods=ne.num_full_neighborhoods,
                              num_candidates_scanned=ne.num_candidates_scanned, success=(rc == capi.SRL_OK), **arrs)

    def _call_buffers(self):
        """Argument marshalling of the per-sweep calls, built once: a persistent srl_eskf_state with a float64 view over it,
        one float64 block for frame_q | frame_t | t_last | R_il | t_il with precomputed pointers, one srl_iekf_summary.
        (Fresh numpy arrays + ctypes pointer objects per call were ~20 us of Python inside every timed sweep.)"""
        b = getattr(self, "_bufs", None)
        if b is None:
            st = capi.EskfState()
            blk = np.zeros(4 + 3 + 3 + 9 + 3, np.float64)
            base = blk.ctypes.data
            b = self._bufs = dict(st=st, st_view=np.frombuffer(st, dtype=np.float64), st_ref=C.byref(st), blk=blk,
                                  p_fq=C.c_void_p(base), p_ft=C.c_void_p(base + 32), p_tl=C.c_void_p(base + 56),
                                  p_R=C.c_void_p(base + 80), p_ti=C.c_void_p(base + 152), summ=IekfSummary())
            b["summ_ref"] = C.byref(b["summ"])
        return b

    def _marshal(self, b, t_last, frame_q, frame_t):
        e, v, blk = self.eskf_pro, b["st_view"], b["blk"]
        v[0:3] = e.p; v[3:7] = e.q; v[7:10] = e.v; v[10:13] = e.ba; v[13:16] = e.bg; v[16:19] = e.g
        v[19:] = np.asarray(e.cov, np.float64).reshape(-1)
        blk[0:4] = e.q if frame_q is None else frame_q
        blk[4:7] = e.p if frame_t is None else frame_t
        blk[7:10] = t_last
        blk[10:19] = np.asarray(self.R_imu_lidar, np.float64).reshape(-1)
        blk[19:22] = self.t_imu_lidar

    def _unmarshal(self, b, rc):
        if rc == capi.SRL_NAN_PLANARITY:
            raise RuntimeError("error")
        _check(self.ctx.h, rc, ok=(capi.SRL_OK, capi.SRL_TOO_FEW_RESIDUALS))
        a = b["st_view"].copy()
        self.eskf_pro = EskfEstimator(p=a[0:3], q=a[3:7], v=a[7:10], ba=a[10:13], bg=a[13:16], g=a[16:19], cov=a[19:].reshape(NS, NS))
        summ = b["summ"]
        blk = b["blk"]
        return OptimizeSummary(success=bool(summ.success) and rc == capi.SRL_OK, num_residuals_used=summ.num_residuals_used,
                               passes_run=summ.passes_run, converged=bool(summ.converged), trace=capi.summary_trace(summ)), \
            blk[0:4].copy(), blk[4:7].copy()

    # ---- src/optimize.cpp:133-314
    def updateIEKF(self, cur_icp_options: IcpParams, t_last, frame_q=None, frame_t=None):
        b = self._call_buffers()
        self._marshal(b, t_last, frame_q, frame_t)
        rc = lib().srl_update_iekf(self.ctx.h, self.voxel_map.h, self.sweep.h, b["st_ref"], b["p_fq"], b["p_ft"], b["p_tl"],
                                   b["p_R"], b["p_ti"], C.byref(cur_icp_options), b["summ_ref"])
        return self._unmarshal(b, rc)

    # ---- src/optimize.cpp:428-448 with the keypoints already selected (gridSampling is a "next" row)
    def optimize(self, raw_xyz, cur_icp_options: IcpParams, t_last, frame_q=None, frame_t=None, want_world: bool = True,
                 world_out=None):
        """world_out: optional (n,3) float64 C-contiguous array to receive the re-transformed frame (pass a pinned
        buffer to avoid staging); raw_xyz may likewise live in pinned memory."""
        raw = f64(raw_xyz).reshape(-1, 3)
        n = raw.shape[0]
        b = self._call_buffers()
        self._marshal(b, t_last, frame_q, frame_t)
        world = (world_out if world_out is not None else np.empty((n, 3))) if want_world else None
        rc = lib().srl_optimize_host(self.ctx.h, self.voxel_map.h, self.sweep.h, ptr(raw), n, b["st_ref"], b["p_fq"], b["p_ft"],
                                     b["p_tl"], b["p_R"], b["p_ti"], C.byref(cur_icp_options), b["summ_ref"],
                                     ptr(world) if want_world else None)
        self.sweep.n = n
        summ, fq, ft = self._unmarshal(b, rc)
        return summ, fq, ft, world


__all__ = ["Context", "VoxelHashMap", "ColorVoxelMap", "Sweep", "EskfEstimator", "LioOptimization", "OptimizeSummary", "PlaneResiduals",
           "IcpParams", "r3live_params", "make_frame", "SrlError"]
