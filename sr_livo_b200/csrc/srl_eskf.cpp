// srl_eskf.cpp — the iterated ESIKF update on the host: the algebra of the host-driven loop (residual cap, profilers, A/B; the
// default loop runs the same algebra on the device, srl_iekf.cu) and eskfEstimator::observe.
//
// Implements what lioOptimization::updateIEKF does around buildPlaneResiduals
// (src/optimize.cpp:138-143 snapshot, :172-232 boxminus + covariance projection, :234-244 gain and d_x,
// :248-251 divergence guard, :253-261 observe + state write-back, :263-310 convergence + posterior covariance)
// and eskfEstimator::observe (src/eskfEstimator.cpp:219-230) with the manifold helpers of
// include/utility.h:194-330 and AngularDistance (src/utility.cpp:146-153).
// The 6x6 normal equations come from the GPU (srl_normal_eq); everything here is 17-dimensional.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "../../include/srlivo_b200.h"
#include "srl_eskf_math.cuh"

namespace {

using namespace srl::ekf;

// Matrix<double,17,17>::inverse(): LU with partial pivoting
bool inverse17(const Mat<N, N>& A, Mat<N, N>& out) {
    double lu[N][N];
    int perm[N];
    for (int r = 0; r < N; ++r) { perm[r] = r; for (int c = 0; c < N; ++c) lu[r][c] = A(r, c); }
    for (int k = 0; k < N; ++k) {
        int piv = k; double best = std::fabs(lu[k][k]);
        for (int r = k + 1; r < N; ++r) if (std::fabs(lu[r][k]) > best) { best = std::fabs(lu[r][k]); piv = r; }
        if (!(best > 0.0)) return false;
        if (piv != k) { for (int c = 0; c < N; ++c) std::swap(lu[k][c], lu[piv][c]); std::swap(perm[k], perm[piv]); }
        for (int r = k + 1; r < N; ++r) {
            const double f = lu[r][k] / lu[k][k];
            lu[r][k] = f;
            for (int c = k + 1; c < N; ++c) lu[r][c] -= f * lu[k][c];
        }
    }
    for (int col = 0; col < N; ++col) {
        double y[N];
        for (int r = 0; r < N; ++r) { double s = (perm[r] == col) ? 1.0 : 0.0; for (int c = 0; c < r; ++c) s -= lu[r][c] * y[c]; y[r] = s; }
        for (int r = N - 1; r >= 0; --r) { double s = y[r]; for (int c = r + 1; c < N; ++c) s -= lu[r][c] * out(c, col); out(r, col) = s / lu[r][r]; }
    }
    return true;
}

// in-place block products used by the covariance projections (:222-232, :281-303)
void rows_so3(Mat<N, N>& dst, const Mat<N, N>& src, const M3& J, int ncols) {   // dst(3:6, j) = J * src(3:6, j)
    for (int j = 0; j < ncols; ++j) {
        double c0 = src(3, j), c1 = src(4, j), c2 = src(5, j);
        for (int r = 0; r < 3; ++r) dst(3 + r, j) = J(r, 0) * c0 + (J(r, 1) * c1 + J(r, 2) * c2);
    }
}
void rows_s2(Mat<N, N>& dst, const Mat<N, N>& src, const Mat<2, 2>& J, int ncols) {
    for (int j = 0; j < ncols; ++j) {
        double c0 = src(15, j), c1 = src(16, j);
        dst(15, j) = J(0, 0) * c0 + J(0, 1) * c1;
        dst(16, j) = J(1, 0) * c0 + J(1, 1) * c1;
    }
}
void cols_so3(Mat<N, N>& dst, const Mat<N, N>& src, const M3& J) {   // dst(j, 3:6) = src(j, 3:6) * J^T
    for (int j = 0; j < N; ++j) {
        double c0 = src(j, 3), c1 = src(j, 4), c2 = src(j, 5);
        for (int r = 0; r < 3; ++r) dst(j, 3 + r) = J(r, 0) * c0 + (J(r, 1) * c1 + J(r, 2) * c2);
    }
}
void cols_s2(Mat<N, N>& dst, const Mat<N, N>& src, const Mat<2, 2>& J) {
    for (int j = 0; j < N; ++j) {
        double c0 = src(j, 15), c1 = src(j, 16);
        dst(j, 15) = J(0, 0) * c0 + J(0, 1) * c1;
        dst(j, 16) = J(1, 0) * c0 + J(1, 1) * c1;
    }
}

}  // namespace

extern "C" {

void srl_icp_params_r3live(srl_icp_params* p) {
    if (!p) return;
    p->size_voxel_map = 1.0; p->power_planarity = 2.0; p->max_dist_to_plane_icp = 0.3; p->weight_alpha = 0.9;
    p->weight_neighborhood = 0.1; p->threshold_orientation_norm = 0.1; p->threshold_translation_norm = 0.01;
    p->laser_point_cov = 0.001; p->voxel_neighborhood = 1; p->min_number_neighbors = 20; p->max_number_neighbors = 20;
    p->threshold_voxel_occupancy = 1; p->max_num_residuals = 600; p->num_iters_icp = 5; p->init_num_frames = 20;
    p->frame_id = 100;
}

int srl_eskf_observe(srl_eskf_state* s, const double d_x[17]) {
    if (!s || !d_x) return SRL_BAD_ARG;
    for (int i = 0; i < 3; ++i) s->p[i] += d_x[i];
    observe_quat(s->q, d_x + 3, s->q);
    for (int i = 0; i < 3; ++i) { s->v[i] += d_x[6 + i]; s->ba[i] += d_x[9 + i]; s->bg[i] += d_x[12 + i]; }
    observe_gravity(s->g, d_x[15], d_x[16], s->g);
    return SRL_OK;
}

int srl_iekf_begin(const srl_eskf_state* eskf, const srl_icp_params* prm, srl_iekf_iter* it) {
    if (!eskf || !prm || !it) return SRL_BAD_ARG;
    it->predict = *eskf;                                                      // :138-143
    it->max_num_iter = prm->frame_id < prm->init_num_frames ? std::max(15, prm->num_iters_icp) : prm->num_iters_icp;   // :135
    it->pass_index = -1;                                                      // :147
    return SRL_OK;
}

int srl_iekf_step(srl_iekf_iter* it, const srl_normal_eq* ne, const srl_icp_params* prm, srl_eskf_state* eskf,
                  double frame_q[4], double frame_t[3], double d_x_out[17], int32_t* done, int32_t* diverged) {
    if (!it || !ne || !prm || !eskf || !frame_q || !frame_t) return SRL_BAD_ARG;
    const srl_eskf_state& pr = it->predict;
    const int i_pass = it->pass_index;
    it->pass_index++;
    if (done) *done = 0;
    if (diverged) *diverged = 0;

    // boxminus of the current state against the prediction (:172-211)
    double dx[N];
    for (int i = 0; i < 3; ++i) {
        dx[i] = eskf->p[i] - pr.p[i];
        dx[6 + i] = eskf->v[i] - pr.v[i];
        dx[9 + i] = eskf->ba[i] - pr.ba[i];
        dx[12 + i] = eskf->bg[i] - pr.bg[i];
    }
    V3 d_so3; M3 J_so3;
    boxminus_so3(pr.q, eskf->q, d_so3, J_so3);                                            // :183-186, :213
    for (int i = 0; i < 3; ++i) dx[3 + i] = d_so3.a[i];
    Mat<2, 1> d_g; Mat<2, 2> J_s2;
    boxminus_s2(pr.g, eskf->g, d_g, J_s2);                                                // :188-211, :214
    dx[15] = d_g.a[0]; dx[16] = d_g.a[1];

    double dx_new[N];
    std::memcpy(dx_new, dx, sizeof(dx));
    { V3 t = J_so3 * d_so3; for (int i = 0; i < 3; ++i) dx_new[3 + i] = t.a[i]; }         // :217
    { Mat<2, 1> t = J_s2 * d_g; dx_new[15] = t.a[0]; dx_new[16] = t.a[1]; }               // :218

    Mat<N, N> P;
    std::memcpy(P.a, eskf->cov, sizeof(P.a));                                             // :220
    rows_so3(P, P, J_so3, N); rows_s2(P, P, J_s2, N); cols_so3(P, P, J_so3); cols_s2(P, P, J_s2);   // :222-232

    Mat<N, N> scaled, temp, temp_inv;
    for (int e = 0; e < N * N; ++e) scaled.a[e] = P.a[e] / prm->laser_point_cov;          // :234
    if (!inverse17(scaled, temp)) return SRL_SINGULAR;
    for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) temp(r, c) += ne->HTH[r * 6 + c];   // :235-236
    if (!inverse17(temp, temp_inv)) return SRL_SINGULAR;                                  // :237

    double K_h[N];                                                                        // :239
    Mat<N, N> K_x = Mat<N, N>::zero();                                                    // :241-242
    for (int r = 0; r < N; ++r) {
        double s = 0.0;
        for (int a = 0; a < 6; ++a) s += temp_inv(r, a) * ne->HTh[a];
        K_h[r] = s;
        for (int c = 0; c < 6; ++c) {
            double t = 0.0;
            for (int a = 0; a < 6; ++a) t += temp_inv(r, a) * ne->HTH[a * 6 + c];
            K_x(r, c) = t;
        }
    }
    double d_x[N];
    for (int r = 0; r < N; ++r) {                                                         // :244
        double s = 0.0;
        for (int c = 0; c < N; ++c) s += (K_x(r, c) - (r == c ? 1.0 : 0.0)) * dx_new[c];
        d_x[r] = -K_h[r] + s;
    }
    if (d_x_out) std::memcpy(d_x_out, d_x, sizeof(d_x));

    V3 g_before = v3(eskf->g);                                                            // :246
    V3 dp = v3(d_x), dth = v3(d_x + 3);
    if (nrm(dp) > 100.0 || angular_distance(dth) > 100.0) {                               // :248-251 `continue`
        if (diverged) *diverged = 1;
        if (done && it->pass_index >= it->max_num_iter) *done = 1;   // loop bound reached without a break
        return SRL_OK;
    }

    srl_eskf_observe(eskf, d_x);                                                          // :253
    std::memcpy(frame_t, eskf->p, 3 * sizeof(double));                                    // :255-256
    std::memcpy(frame_q, eskf->q, 4 * sizeof(double));

    bool converged = prm->frame_id > 1 && nrm(dp) < prm->threshold_translation_norm &&
                     angular_distance(dth) < prm->threshold_orientation_norm;             // :265-270
    if (converged || i_pass == it->max_num_iter - 1) {                                    // :272
        Mat<N, N> P_new = P;
        posterior_jacobians(d_x + 3, g_before.a, d_x[15], d_x[16], J_so3, J_s2);          // :278-279
        rows_so3(P_new, P, J_so3, N);                                                     // :281-282
        rows_s2(P_new, P, J_s2, N);                                                       // :284-285
        cols_so3(P_new, P, J_so3); cols_so3(P, P, J_so3);                                 // :287-291
        cols_s2(P_new, P, J_s2); cols_s2(P, P, J_s2);                                     // :293-297
        rows_so3(K_x, K_x, J_so3, 6);                                                     // :299-300
        rows_s2(K_x, K_x, J_s2, 6);                                                       // :302-303
        for (int r = 0; r < N; ++r)                                                       // :305-307
            for (int c = 0; c < N; ++c) {
                double s = 0.0;
                for (int a = 0; a < 6; ++a) s += K_x(r, a) * P(a, c);
                eskf->cov[r * N + c] = P_new(r, c) - s;
            }
        if (done) *done = converged ? 2 : 1;
        return SRL_OK;
    }
    if (done && it->pass_index >= it->max_num_iter) *done = 1;
    return SRL_OK;
}

}  // extern "C"
