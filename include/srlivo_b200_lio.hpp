// srlivo_b200_lio.hpp — C++ host mirror of the reference's scan-matching interface over the C ABI.
//
// The reference is C++ (class lioOptimization, include/lioOptimization.h:192-385); a maintainer who wants the B200
// path swaps the bodies of four member functions for the calls below (INTEGRATION.md shows the diff).  Header-only,
// Eigen-free by default; define SRL_HAVE_EIGEN before including to get overloads on the reference's own types
// (point3D / icpOptions / Eigen::Quaterniond), which cannot be compiled in this repository's container.
//
//   reference                                                   this header
//   lioOptimization::addPointsToMap   src/lioOptimization.cpp:520   srl::LioBackend::addPointsToMap
//   lioOptimization::mapSize          src/lioOptimization.cpp:574   srl::LioBackend::mapSize
//   lioOptimization::buildPlaneResiduals  src/optimize.cpp:18       srl::LioBackend::buildPlaneResiduals
//   lioOptimization::updateIEKF       src/optimize.cpp:133          srl::LioBackend::updateIEKF
//   lioOptimization::optimize         src/optimize.cpp:428          srl::LioBackend::optimize (keypoints given)
//   lioOptimization::removePointsFarFromLocation  src/lioOptimization.cpp:556   srl::LioBackend::removePointsFarFromLocation
//   gridSampling                      src/utility.cpp:188           srl::LioBackend::gridSampling (keypoint indices)
//   distortFrameByConstant / ByImu    src/utility.cpp:203,238       srl::LioBackend::distortFrameByConstant / distortFrameByImu
//   transformAllImuPoint              src/utility.cpp:320           srl::LioBackend::transformAllImuPoint
//   addPointToColorMap (loop :533-551) src/lioOptimization.cpp:448  srl::LioBackend::addPointsToColorMap
//   rgbMapTracker::renderPointsInRecentVoxel  src/rgbMapTracker.cpp:216  srl::LioBackend::renderPointsInRecentVoxel
#pragma once

#include <array>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "srlivo_b200.h"

namespace srl {

struct optimizeSummary {   // include/lioOptimization.h (same fields the path fills)
    bool success = false;
    int num_residuals_used = 0;
    std::string error_log;
    int passes_run = 0;
    bool converged = false;
};

class LioBackend {
public:
    // device: CUDA ordinal; stream: cudaStream_t or nullptr; max_voxels / sweep_capacity size the HBM pools
    LioBackend(int device, void* stream, size_t max_voxels, size_t sweep_capacity, double size_voxel_map = 1.0,
               int max_num_points_in_voxel = 20) {
        check(srl_ctx_create(device, stream, &ctx_), "srl_ctx_create (no CPU fallback: a CUDA device is required)");
        check(srl_map_create(ctx_, size_voxel_map, max_num_points_in_voxel, max_voxels, &map_), "srl_map_create");
        check(srl_sweep_create(ctx_, sweep_capacity, &sweep_), "srl_sweep_create");
        const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        std::memcpy(R_imu_lidar, I, sizeof(I));
        std::memset(t_imu_lidar, 0, sizeof(t_imu_lidar));
    }
    ~LioBackend() {
        if (color_) srl_color_map_destroy(color_);
        if (sweep_) srl_sweep_destroy(sweep_);
        if (map_) srl_map_destroy(map_);
        if (ctx_) srl_ctx_destroy(ctx_);
    }
    LioBackend(const LioBackend&) = delete;
    LioBackend& operator=(const LioBackend&) = delete;

    double R_imu_lidar[9];   // include/lioOptimization.h:227-228
    double t_imu_lidar[3];
    srl_eskf_state eskf{};   // eskf_pro's state (src/eskfEstimator.cpp:3-21)

    // src/lioOptimization.cpp:520-554: registered frame (world points, sweep order) into the voxel map
    long long addPointsToMap(const double* xyz_world, size_t n, double min_distance_points, int min_num_points = 0) {
        int64_t added = 0;
        check(srl_map_insert(map_, xyz_world, n, min_distance_points, min_num_points, &added), "srl_map_insert");
        return added;
    }
    // src/lioOptimization.cpp:574-581
    long long mapSize() {
        int64_t nv = 0, np = 0;
        check(srl_map_stats(map_, &nv, &np), "srl_map_stats");
        return np;
    }
    // the keypoints vector of optimize() (raw_point members), uploaded once per sweep
    void setKeypoints(const double* raw_xyz, size_t n) { check(srl_sweep_upload(sweep_, raw_xyz, n), "srl_sweep_upload"); }

    // src/optimize.cpp:18-131 + :160-170,:235,:239 — one pass; returns the normal equations instead of plane_residuals
    optimizeSummary buildPlaneResiduals(const srl_icp_params& cur_icp_options, const double q_cur[4], const double t_cur[3],
                                        const double t_last[3], srl_normal_eq& ne, srl_debug_out* dbg = nullptr) {
        srl_frame fr;
        std::memcpy(fr.q_cur, q_cur, sizeof(fr.q_cur));
        std::memcpy(fr.t_cur, t_cur, sizeof(fr.t_cur));
        std::memcpy(fr.t_last, t_last, sizeof(fr.t_last));
        std::memcpy(fr.R_il, R_imu_lidar, sizeof(fr.R_il));
        std::memcpy(fr.t_il, t_imu_lidar, sizeof(fr.t_il));
        const int rc = srl_build_plane_residuals(ctx_, map_, sweep_, &fr, &cur_icp_options, &ne, dbg);
        optimizeSummary s;
        s.num_residuals_used = (int)ne.num_residuals;
        if (rc == SRL_NAN_PLANARITY) throw std::runtime_error("error");   // src/optimize.cpp:348-350
        if (rc == SRL_TOO_FEW_RESIDUALS) { s.success = false; s.error_log = srl_last_error(ctx_); return s; }   // :110-123
        check(rc, "srl_build_plane_residuals");
        s.success = true;
        return s;
    }

    // src/optimize.cpp:133-314 — frame_q/frame_t = p_frame->p_state rotation/translation (in/out)
    optimizeSummary updateIEKF(const srl_icp_params& cur_icp_options, double frame_q[4], double frame_t[3], const double t_last[3]) {
        srl_iekf_summary sm;
        const int rc = srl_update_iekf(ctx_, map_, sweep_, &eskf, frame_q, frame_t, t_last, R_imu_lidar, t_imu_lidar,
                                       &cur_icp_options, &sm);
        return summarise(rc, sm);
    }

    // src/optimize.cpp:428-448 with the keypoints already selected; world_xyz_out (n*3, may be null) receives the
    // re-transformed frame (:441-445)
    optimizeSummary optimize(const double* raw_xyz, size_t n, const srl_icp_params& cur_icp_options, double frame_q[4],
                             double frame_t[3], const double t_last[3], double* world_xyz_out) {
        srl_iekf_summary sm;
        const int rc = srl_optimize_host(ctx_, map_, sweep_, raw_xyz, n, &eskf, frame_q, frame_t, t_last, R_imu_lidar,
                                         t_imu_lidar, &cur_icp_options, &sm, world_xyz_out);
        return summarise(rc, sm);
    }

    // src/lioOptimization.cpp:556-572 — voxels whose first point is farther than `distance` from `location` go
    long long removePointsFarFromLocation(const double location[3], double distance) {
        int64_t removed = 0;
        check(srl_map_remove_far(map_, location, distance, &removed), "srl_map_remove_far");
        return removed;
    }
    // src/utility.cpp:188-201 — indices (into the frame) of the keypoints, in the reference's order
    std::vector<uint32_t> gridSampling(const double* xyz_world, size_t n, double size_voxel_subsampling) {
        std::vector<uint32_t> keep(n);
        size_t m = 0;
        check(srl_grid_sampling(ctx_, xyz_world, n, size_voxel_subsampling, keep.data(), &m), "srl_grid_sampling");
        keep.resize(m);
        return keep;
    }
    // src/utility.cpp:203-236, :238-312, :320-332 — point buffers may be host or device pointers
    void distortFrameByConstant(const double* raw_xyz, const double* relative_time_ms, size_t n, const std::vector<srl_imu_state>& imu_states,
                                double time_frame_begin, double* imu_xyz) {
        check(srl_distort_frame_by_constant(ctx_, raw_xyz, relative_time_ms, n, imu_states.data(), imu_states.size(), time_frame_begin,
                                            R_imu_lidar, t_imu_lidar, imu_xyz), "srl_distort_frame_by_constant");
    }
    long long distortFrameByImu(const double* raw_xyz, const double* relative_time_ms, size_t n, const std::vector<srl_imu_state>& imu_states,
                                double time_frame_begin, double* imu_xyz) {
        int64_t written = 0;
        check(srl_distort_frame_by_imu(ctx_, raw_xyz, relative_time_ms, n, imu_states.data(), imu_states.size(), time_frame_begin,
                                       R_imu_lidar, t_imu_lidar, imu_xyz, &written), "srl_distort_frame_by_imu");
        return written;
    }
    void transformAllImuPoint(const double* imu_xyz, size_t n, const srl_imu_state& last_state, double* raw_xyz_out) {
        check(srl_transform_all_imu_point(ctx_, imu_xyz, n, &last_state, R_imu_lidar, t_imu_lidar, raw_xyz_out), "srl_transform_all_imu_point");
    }

#ifdef SRL_HAVE_EIGEN
    // overloads on the reference's types (cloudMap.h / parameters.h must be included first)
    static srl_icp_params fromIcpOptions(const icpOptions& o, int frame_id, double laser_point_cov) {
        srl_icp_params p;
        p.size_voxel_map = o.size_voxel_map; p.power_planarity = o.power_planarity; p.max_dist_to_plane_icp = o.max_dist_to_plane_icp;
        p.weight_alpha = o.weight_alpha; p.weight_neighborhood = o.weight_neighborhood;
        p.threshold_orientation_norm = o.threshold_orientation_norm; p.threshold_translation_norm = o.threshold_translation_norm;
        p.laser_point_cov = laser_point_cov; p.voxel_neighborhood = o.voxel_neighborhood; p.min_number_neighbors = o.min_number_neighbors;
        p.max_number_neighbors = o.max_number_neighbors; p.threshold_voxel_occupancy = o.threshold_voxel_occupancy;
        p.max_num_residuals = o.max_num_residuals; p.num_iters_icp = o.num_iters_icp; p.init_num_frames = o.init_num_frames;
        p.frame_id = frame_id;
        return p;
    }
    void setKeypoints(const std::vector<point3D>& keypoints) {
        std::vector<double> raw(keypoints.size() * 3);
        for (size_t i = 0; i < keypoints.size(); ++i) for (int a = 0; a < 3; ++a) raw[3 * i + a] = keypoints[i].raw_point[a];
        setKeypoints(raw.data(), keypoints.size());
    }
#endif

    srl_ctx* ctx() { return ctx_; }
    srl_map* map() { return map_; }
    srl_sweep* sweep() { return sweep_; }

    // ---- row N4: color_voxel_map + hashmap_3d_points + rgb_points_vec + voxels_recent_visited (include/lioOptimization.h:275-291)
    // created on first use with the LiDAR map's voxel size and cap (src/lioOptimization.cpp:539 passes map_options' values)
    void enableColorMap(double size_voxel_map, int max_num_points_in_voxel, size_t max_voxels, double min_distance_points) {
        if (!color_) check(srl_color_map_create(ctx_, size_voxel_map, max_num_points_in_voxel, max_voxels, min_distance_points, &color_), "srl_color_map_create");
    }
    // the colour branch of addPointsToMap (src/lioOptimization.cpp:533-551): every add_point_step-th point of the registered frame
    long long addPointsToColorMap(const double* xyz_world, size_t n, int add_point_step, double time_sweep_end, double time_last_process,
                                  bool to_rendering) {
        int64_t stored = 0;
        check(srl_color_map_add_points(color_, xyz_world, n, add_point_step, time_sweep_end, time_last_process, to_rendering ? 1 : 0, &stored),
              "srl_color_map_add_points");
        return stored;
    }
    // rgbMapTracker::renderPointsInRecentVoxel (src/rgbMapTracker.cpp:216-237); returns render_point_count
    long long renderPointsInRecentVoxel(const srl_camera& cam, const uint8_t* image_bgr, double obs_time) {
        int64_t rendered = 0;
        check(srl_color_map_render_recent(color_, &cam, image_bgr, obs_time, &rendered), "srl_color_map_render_recent");
        return rendered;
    }
    srl_color_map* colorMap() { return color_; }

private:
    srl_ctx* ctx_ = nullptr;
    srl_map* map_ = nullptr;
    srl_sweep* sweep_ = nullptr;
    srl_color_map* color_ = nullptr;

    void check(int rc, const char* what) {
        if (rc != SRL_OK) throw std::runtime_error(std::string(what) + ": " + (ctx_ ? srl_last_error(ctx_) : "no context"));
    }
    optimizeSummary summarise(int rc, const srl_iekf_summary& sm) {
        optimizeSummary s;
        s.num_residuals_used = sm.num_residuals_used; s.passes_run = sm.passes_run; s.converged = sm.converged != 0;
        if (rc == SRL_NAN_PLANARITY) throw std::runtime_error("error");
        if (rc == SRL_TOO_FEW_RESIDUALS) { s.success = false; s.error_log = srl_last_error(ctx_); return s; }
        check(rc, "srl_update_iekf");
        s.success = sm.success != 0;
        return s;
    }
};

}  // namespace srl
