/*
 * srl_oracle.h — C interface of the CPU ORACLE (test infrastructure, NOT product code).
 *
 * The oracle is an Eigen-free CPU restatement of SR-LIVO's LIO scan-matching hot path
 * (reference: src/optimize.cpp, include/cloudMap.h, src/cloudMap.cpp,
 * src/lioOptimization.cpp:400-446,520-554, src/eskfEstimator.cpp:219-230,
 * include/utility.h:205-330).  It exists only so that tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs can check and time the
 * reference algorithm.  Nothing under sr_livo_b200/ may include, link or call it.
 *
 * PARITY: the reference ships no test, golden vector or fixture for this path.  Since round 2 the
 * restatement is pinned against the REFERENCE'S OWN CODE: oracle/_ref/libsrl_reference.so holds src/optimize.cpp,
 * utility.cpp, eskfEstimator.cpp, cloudMap.cpp, state.cpp, lioOptimization.cpp, rgbMapTracker.cpp compiled unmodified
 * from where they lie, over stand-in headers for Eigen / OpenCV / ROS / PCL (oracle/shim/, none of them is in this
 * image), and tests/test_reference_pin.py compares every function of this file with it (rows of buildPlaneResiduals bit
 * for bit).  PARITY STILL UNPINNED for the arithmetic INSIDE Eigen / OpenCV calls (reduction order of 3-vector dot
 * products, SelfAdjointEigenSolver, PartialPivLU, saturating Vec3b operators): oracle/shim restates it the same way
 * this file does, so those two cannot check each other; the self-checks (tests/test_oracle_selfcheck.py: brute-force
 * kNN, numpy eigh, finite-difference Jacobians, numpy ESIKF algebra) bound their effect.
 */
#ifndef SRL_ORACLE_H
#define SRL_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* icpOptions fields that the hot path actually reads (include/parameters.h:8-56,
 * config/r3live.yaml:57-69) plus the three lioOptimization members it uses. */
typedef struct orc_icp_params {
    double size_voxel_map;            /* icp_options.size_voxel_map (1.0) */
    double power_planarity;           /* 2.0 */
    double max_dist_to_plane_icp;     /* 0.3 */
    double weight_alpha;              /* 0.9 */
    double weight_neighborhood;       /* 0.1 */
    double threshold_orientation_norm;/* deg, yaml 0.1 */
    double threshold_translation_norm;/* m, yaml 0.01 */
    double laser_point_cov;           /* src/lioOptimization.cpp:364 -> 0.001 */
    int32_t voxel_neighborhood;       /* 1 */
    int32_t min_number_neighbors;     /* 20 */
    int32_t max_number_neighbors;     /* 20 */
    int32_t threshold_voxel_occupancy;/* 1 */
    int32_t max_num_residuals;        /* yaml 600; compiled default -1 */
    int32_t num_iters_icp;            /* 5 */
    int32_t init_num_frames;          /* 20 */
    int32_t frame_id;                 /* p_frame->frame_id: <init_num_frames => nb=2, thr=1, >=15 iters */
} orc_icp_params;

/* eskfEstimator state (src/eskfEstimator.cpp:3-21): p, q(x,y,z,w), v, ba, bg, g, P(17x17 row-major) */
typedef struct orc_eskf_state {
    double p[3];
    double q[4];   /* x, y, z, w  (Eigen coeffs() order) */
    double v[3];
    double ba[3];
    double bg[3];
    double g[3];
    double cov[17 * 17];
} orc_eskf_state;

/* per-pass sums: what src/optimize.cpp:160-170,235,239 produce from the rows */
typedef struct orc_normal_eq {
    double HTH[36];        /* row-major 6x6 */
    double HTh[6];
    double loss_sum;       /* sum distance^2 (unweighted), :104 */
    int64_t num_residuals; /* :99 */
    int64_t num_full_neighborhoods;   /* keypoints passing :78 */
    int64_t sum_candidates;           /* sum C_k, candidates visited at :391 */
    int64_t sum_probes_hit;           /* voxels found at :386 */
    int64_t num_visited;              /* keypoints entered before the :107 break */
    int64_t num_fragile;              /* keypoints with |d20-d21|<=1e-12*d20 or a sqrt tie: selection not robust */
    int32_t success;                  /* :110-130 */
    int32_t nan_planarity;            /* :348 would have thrown */
} orc_normal_eq;

/* optional per-keypoint outputs (any pointer may be NULL) */
typedef struct orc_debug_out {
    double* world_xyz;     /* N*3  keypoint.point after transformKeypoints (:38) */
    int32_t* status;       /* N    -1 not visited, 0 <K neighbours, 1 gated out (:98), 2 accepted */
    int16_t* nbr;          /* N*K*4 (vx,vy,vz,index-in-block) ascending distance; K = max_number_neighbors */
    double* nbr_dist;      /* N*K  distances (sqrt) ascending */
    double* plane;         /* N*16 raw_point3, norm_vector3, jacobians6, norm_offset, distance, weight, a2D */
    int32_t* num_candidates; /* N  C_k */
} orc_debug_out;

void* orc_map_create(void);
void orc_map_destroy(void* map);
const char* orc_map_backend(void);  /* "tsl::robin_map 0.6.3 (reference vendored)" or "std::unordered_map" */
int64_t orc_map_num_voxels(void* map);
int64_t orc_map_num_points(void* map);

/* lioOptimization::addPointsToMap / addPointToMap (src/lioOptimization.cpp:400-446,533-537) */
int64_t orc_map_add_points(void* map, const double* xyz, int64_t n, double voxel_size,
                           int32_t max_num_points_in_voxel, double min_distance_points,
                           int32_t min_num_points);

/* lioOptimization::removePointsFarFromLocation (src/lioOptimization.cpp:556-572); returns the number of voxels erased */
int64_t orc_map_remove_far(void* map, const double location[3], double distance);

/* dump in map iteration order; xyz is n_voxels*cap*3 floats (unused tail zero). returns n_voxels */
int64_t orc_map_snapshot(void* map, int32_t cap, int16_t* keys, int32_t* counts, float* xyz);
/* test-infra shortcut: fill the map from a snapshot (blocks keep the given point order) */
void orc_map_load(void* map, const int16_t* keys, const int32_t* counts, const float* xyz,
                  int64_t n_voxels, int32_t cap);

/* lioOptimization::buildPlaneResiduals (+ row assembly/HTH/HTh), one ESIKF pass.
 * nthreads<=1: the reference as written (single thread, cap semantics exact).
 * nthreads>1 : keypoint ranges over std::thread, private sums, fixed-order combine;
 *              only legal when max_num_residuals >= n (cap never binds). */
int32_t orc_build_plane_residuals(void* map, const double* raw_xyz, int64_t n,
                                  const double q_cur[4], const double t_cur[3],
                                  const double t_last[3], const double R_il[9],
                                  const double t_il[3], const orc_icp_params* prm,
                                  int32_t nthreads, orc_normal_eq* out, orc_debug_out* dbg);

/* lioOptimization::updateIEKF (src/optimize.cpp:133-314) incl. eskfEstimator::observe.
 * frame_q/frame_t = p_frame->p_state rotation/translation (in/out),
 * trace (optional): per pass 17 d_x + 7 pose (t, q) = 24 doubles, up to max_passes rows. */
int32_t orc_update_iekf(void* map, const double* raw_xyz, int64_t n, orc_eskf_state* eskf,
                        double frame_q[4], double frame_t[3], const double t_last[3],
                        const double R_il[9], const double t_il[3], const orc_icp_params* prm,
                        int32_t nthreads, int32_t* passes_run, int32_t* num_residuals_used,
                        double* trace, int32_t max_trace_rows);

/* gridSampling / subSampleFrame (src/utility.cpp:167-201): indices of the kept points in the reference's order (the
 * iteration order of its std::tr1::unordered_map grid). out has capacity n; returns the number of keypoints */
int64_t orc_grid_sampling(const double* xyz, int64_t n, double size_voxel_subsampling, int32_t* out);

/* row N3 — per-sweep point transforms / undistortion (src/utility.cpp:203-332).  imuState fields the functions read
 * (include/utility.h imuState: timestamp, quat, trans, vel, un_acc, un_gyr) */
typedef struct orc_imu_state {
    double timestamp;
    double quat[4];   /* x, y, z, w */
    double trans[3];
    double vel[3];
    double un_acc[3];
    double un_gyr[3];
} orc_imu_state;
/* distortFrameByConstant (:203-236); relative_time in ms as point3D::relative_time */
void orc_distort_frame_by_constant(const double* raw_xyz, const double* relative_time, int64_t n, const orc_imu_state* st,
                                   int64_t n_states, double time_frame_begin, const double R_il[9], const double t_il[3],
                                   double* imu_xyz);
/* distortFrameByImu (:238-312, method 1); imu_xyz is in/out (points the iterator never reaches keep their value);
 * returns how many leading points were written */
int64_t orc_distort_frame_by_imu(const double* raw_xyz, const double* relative_time, int64_t n, const orc_imu_state* st,
                                 int64_t n_states, double time_frame_begin, const double R_il[9], const double t_il[3],
                                 double* imu_xyz);
/* transformAllImuPoint (:320-332) */
void orc_transform_all_imu_point(const double* imu_xyz, int64_t n, const orc_imu_state* last, const double R_il[9],
                                 const double t_il[3], double* raw_out);
void orc_quat_slerp(const double a[4], double t, const double b[4], double out[4]);   /* Eigen slerp */

/* small pieces exported for self-checks */
void orc_quat_to_rot(const double q[4], double R[9]);                 /* Eigen toRotationMatrix */
void orc_eig3_sym(const double S[9], double evals[3], double evecs[9]); /* SelfAdjointEigenSolver<Matrix3d> */
void orc_eskf_observe(orc_eskf_state* s, const double dx[17]);        /* src/eskfEstimator.cpp:219-230 */
int32_t orc_mat17_inverse(const double* A, double* Ainv);             /* PartialPivLU inverse */
uint64_t orc_voxel_hash(int16_t x, int16_t y, int16_t z);             /* include/cloudMap.h:173-184 */

/* ---- row N4: colour map (src/lioOptimization.cpp:448-551 colour branch, src/rgbMapTracker.cpp:181-237, src/cloudMap.cpp:59-101) */
void* orc_color_create(void);
void orc_color_destroy(void* cm);
int64_t orc_color_add_points(void* cm, const double* xyz, int64_t n, double voxel_size, int32_t max_num_points_in_voxel,
                             double min_distance_points, int32_t add_point_step, double time_sweep_end, double time_last_process,
                             int32_t to_rendering);
int64_t orc_color_render(void* cm, const double* cam15, const uint8_t* image_bgr, int32_t rows, int32_t cols, double obs_time);
int64_t orc_color_num_voxels(void* cm);
int64_t orc_color_num_rgb_points(void* cm);
int64_t orc_color_num_recent(void* cm);
int64_t orc_color_num_new_recent(void* cm);
int64_t orc_color_snapshot(void* cm, int32_t cap, int16_t* keys, int32_t* counts, float* xyz, int16_t* rgb, int16_t* n_rgb,
                           float* cov, double* obs_dist, double* last_obs, double* last_visited);
void orc_color_lists(void* cm, int16_t* rgb_points, int32_t* recent);

#ifdef __cplusplus
}
#endif
#endif
