// TEST INFRASTRUCTURE (oracle/) — the maintainer patch of INTEGRATION.md §2, applied to the reference AT LINK TIME.
//
// oracle/_ref/libsrl_reference_gpu.so = the reference's own objects (the same unmodified translation units as
// libsrl_reference.so) with two of its member functions made weak by objcopy and replaced by the definitions below:
//     lioOptimization::updateIEKF      (src/optimize.cpp:133-314)        -> srl::LioBackend::updateIEKF      (CUDA, all passes)
//     lioOptimization::addPointsToMap  (src/lioOptimization.cpp:520-554) -> srl::LioBackend::addPointsToMap  (CUDA insert kernel)
// Everything around them stays the reference's code: lioOptimization::optimize() (gridSampling -> updateIEKF ->
// transformPoint, src/optimize.cpp:428-447) calls the GPU update through its own, unmodified call site; eskfEstimator,
// cloudFrame and state are the reference's classes.  This is what "drop-in" means for this path, exercised by
// tests/test_gpu_parity.py::test_reference_runs_on_the_gpu_backend against the unpatched library.
//
// The adapter used is the product's public header include/srlivo_b200_lio.hpp over the C ABI (include/srlivo_b200.h) — the
// same two files a maintainer would add to the reference's include path.  No product code lives here.
#include "lioOptimization.h"
#include "imageProcessing.h"
#include "rgbMapTracker.h"

#define SRL_HAVE_EIGEN 1
#include "srlivo_b200_lio.hpp"

namespace {

// INTEGRATION.md §2 adds `std::unique_ptr<srl::LioBackend> gpu_lio` to class lioOptimization; the reference's headers are
// not edited here, so the member lives beside the object (one backend per lioOptimization instance)
std::mutex g_mutex;
std::map<const lioOptimization*, std::unique_ptr<srl::LioBackend>> g_backends;
std::map<const lioOptimization*, bool> g_map_on_gpu;        // true once addPointsToMap fed the HBM map directly
std::map<const lioOptimization*, size_t> g_mirrored_points;  // host-map size at the last mirror upload

srl::LioBackend& backend(lioOptimization* self) {
    std::lock_guard<std::mutex> lock(g_mutex);
    auto& slot = g_backends[self];
    if (!slot) {
        slot.reset(new srl::LioBackend(/*device*/0, /*stream*/nullptr, /*max_voxels*/1u << 20, /*sweep_capacity*/1u << 18,
                                       self->odometry_options.optimize_options.size_voxel_map, self->odometry_options.max_num_points_in_voxel));
    }
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) slot->R_imu_lidar[r * 3 + c] = self->R_imu_lidar(r, c);
    for (int a = 0; a < 3; ++a) slot->t_imu_lidar[a] = self->t_imu_lidar(a);
    return *slot;
}

void copy3(double* o, const Eigen::Vector3d& v) { o[0] = v(0); o[1] = v(1); o[2] = v(2); }

}  // namespace

// lioOptimization::addPointsToMap: the registered frame goes to the HBM map in one batched call; the
// colour-map branch and the RViz cloud of the original are outside this demonstration
void lioOptimization::addPointsToMap(voxelHashMap& map, cloudFrame* p_frame, double voxel_size, int max_num_points_in_voxel,
                                     double min_distance_points, int min_num_points, bool to_rendering) {
    (void)map; (void)voxel_size; (void)max_num_points_in_voxel; (void)to_rendering;
    srl::LioBackend& gpu = backend(this);
    std::vector<double> xyz(3 * p_frame->point_frame.size());
    for (size_t i = 0; i < p_frame->point_frame.size(); ++i)
        for (int a = 0; a < 3; ++a) xyz[3 * i + a] = p_frame->point_frame[i].point[a];
    gpu.addPointsToMap(xyz.data(), p_frame->point_frame.size(), min_distance_points, min_num_points);
    std::lock_guard<std::mutex> lock(g_mutex);
    g_map_on_gpu[this] = true;
}

// lioOptimization::updateIEKF: every pass and the 17-dimensional algebra on the GPU
optimizeSummary lioOptimization::updateIEKF(const icpOptions& cur_icp_options, voxelHashMap& voxel_map_temp, std::vector<point3D>& keypoints,
                                            cloudFrame* p_frame) {
    srl::LioBackend& gpu = backend(this);
    bool on_gpu;
    { std::lock_guard<std::mutex> lock(g_mutex); on_gpu = g_map_on_gpu[this]; }
    if (!on_gpu) {
        // the host voxelHashMap is still the master (filled by the harness' ref_map_load): mirror it when it changed
        size_t pts = 0;
        for (auto& kv : voxel_map_temp) pts += (size_t)kv.second.NumPoints();
        bool stale;
        { std::lock_guard<std::mutex> lock(g_mutex); stale = g_mirrored_points.count(this) == 0 || g_mirrored_points[this] != pts; }
        if (stale) {
            const int cap = odometry_options.max_num_points_in_voxel;
            const size_t nv = voxel_map_temp.size();
            std::vector<int16_t> keys(nv * 3); std::vector<int32_t> counts(nv); std::vector<float> xyz(nv * (size_t)cap * 3, 0.f);
            size_t v = 0;
            for (auto it = voxel_map_temp.begin(); it != voxel_map_temp.end(); ++it, ++v) {
                keys[3 * v] = it->first.x; keys[3 * v + 1] = it->first.y; keys[3 * v + 2] = it->first.z;
                voxelBlock& block = it.value();
                counts[v] = std::min(block.NumPoints(), cap);
                for (int i = 0; i < counts[v]; ++i) {
                    const Eigen::Vector3d p = block.points[(size_t)i].getPosition();
                    for (int a = 0; a < 3; ++a) xyz[(v * cap + i) * 3 + a] = (float)p(a);
                }
            }
            if (srl_map_upload(gpu.map(), keys.data(), counts.data(), xyz.data(), nv) != SRL_OK) throw std::runtime_error(srl_last_error(gpu.ctx()));
            std::lock_guard<std::mutex> lock(g_mutex);
            g_mirrored_points[this] = pts;
        }
    }
    const srl_icp_params prm = srl::LioBackend::fromIcpOptions(cur_icp_options, p_frame->frame_id, laser_point_cov);
    gpu.setKeypoints(keypoints);                                              // H2D once per sweep
    srl_eskf_state& e = gpu.eskf;                                             // copyEskfTo
    copy3(e.p, eskf_pro->getTranslation());
    const Eigen::Quaterniond q = eskf_pro->getRotation();
    e.q[0] = q.x(); e.q[1] = q.y(); e.q[2] = q.z(); e.q[3] = q.w();
    copy3(e.v, eskf_pro->getVelocity()); copy3(e.ba, eskf_pro->getBa()); copy3(e.bg, eskf_pro->getBg()); copy3(e.g, eskf_pro->getGravity());
    const Eigen::Matrix<double, 17, 17> P = eskf_pro->getCovariance();
    for (int r = 0; r < 17; ++r) for (int c = 0; c < 17; ++c) e.cov[r * 17 + c] = P(r, c);
    double fq[4] = {p_frame->p_state->rotation.x(), p_frame->p_state->rotation.y(), p_frame->p_state->rotation.z(), p_frame->p_state->rotation.w()};
    double ft[3] = {p_frame->p_state->translation.x(), p_frame->p_state->translation.y(), p_frame->p_state->translation.z()};
    double tl[3];
    copy3(tl, all_cloud_frame[p_frame->id - 1]->p_state->translation);        // last_state (src/optimize.cpp:25)

    const srl::optimizeSummary s = gpu.updateIEKF(prm, fq, ft, tl);           // throws std::runtime_error("error") on NaN planarity like :348-350

    optimizeSummary summary;
    summary.success = s.success;
    summary.num_residuals_used = s.num_residuals_used;
    summary.error_log = s.error_log;
    // the library hands back the filter as the reference would leave it at this point (untouched when the first pass fails, :154-155)
    eskf_pro->setTranslation(Eigen::Vector3d(e.p[0], e.p[1], e.p[2]));        // copyEskfFrom
    eskf_pro->setRotation(Eigen::Quaterniond(e.q[3], e.q[0], e.q[1], e.q[2]));
    eskf_pro->setVelocity(Eigen::Vector3d(e.v[0], e.v[1], e.v[2]));
    eskf_pro->setBa(Eigen::Vector3d(e.ba[0], e.ba[1], e.ba[2]));
    eskf_pro->setBg(Eigen::Vector3d(e.bg[0], e.bg[1], e.bg[2]));
    eskf_pro->setGravity(Eigen::Vector3d(e.g[0], e.g[1], e.g[2]));
    Eigen::Matrix<double, 17, 17> Pn;
    for (int r = 0; r < 17; ++r) for (int c = 0; c < 17; ++c) Pn(r, c) = e.cov[r * 17 + c];
    eskf_pro->setCovariance(Pn);
    if (s.passes_run > 1 || s.success) {                                      // :255-261, written after every accepted observe
        p_frame->p_state->translation = Eigen::Vector3d(ft[0], ft[1], ft[2]);
        p_frame->p_state->rotation = Eigen::Quaterniond(fq[3], fq[0], fq[1], fq[2]);
        p_frame->p_state->velocity = eskf_pro->getVelocity();
        p_frame->p_state->ba = eskf_pro->getBa();
        p_frame->p_state->bg = eskf_pro->getBg();
    }
    G = eskf_pro->getGravity();
    G_norm = G.norm();
    return summary;
}

extern "C" {
// what the test reads back: the HBM map behind a patched lioOptimization (RefCtx's first member is the lioOptimization*)
int64_t refgpu_map_points(void* ctx) {
    lioOptimization* self = *static_cast<lioOptimization**>(ctx);
    return (int64_t)backend(self).mapSize();
}
int32_t refgpu_map_is_on_gpu(void* ctx) {
    lioOptimization* self = *static_cast<lioOptimization**>(ctx);
    std::lock_guard<std::mutex> lock(g_mutex);
    return g_map_on_gpu[self] ? 1 : 0;
}
void refgpu_release(void* ctx) {   // before the lioOptimization object goes away
    lioOptimization* self = *static_cast<lioOptimization**>(ctx);
    std::lock_guard<std::mutex> lock(g_mutex);
    g_backends.erase(self); g_map_on_gpu.erase(self); g_mirrored_points.erase(self);
}
}
