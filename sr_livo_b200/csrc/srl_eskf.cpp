// srl_eskf.cpp — host side of the iterated ESIKF update (O(1) per pass, stays on the CPU).
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
#include "srl_math.cuh"

namespace {

constexpr int N = 17;
constexpr double kTheta = 1e-4;   // THETA_THRESHOLD, include/utility.h:27

template <int R, int C>
struct Mat {
    double a[R * C];
    double& operator()(int r, int c) { return a[r * C + c]; }
    double operator()(int r, int c) const { return a[r * C + c]; }
    static Mat zero() { Mat m; std::memset(m.a, 0, sizeof(m.a)); return m; }
    static Mat identity() { Mat m = zero(); for (int i = 0; i < (R < C ? R : C); ++i) m(i, i) = 1.0; return m; }
};
template <int R, int K, int C>
Mat<R, C> operator*(const Mat<R, K>& A, const Mat<K, C>& B) {
    Mat<R, C> out;
    for (int r = 0; r < R; ++r)
        for (int c = 0; c < C; ++c) {
            double s = 0.0;
            for (int k = 0; k < K; ++k) s += A(r, k) * B(k, c);
            out(r, c) = s;
        }
    return out;
}
template <int R, int C>
Mat<C, R> tr(const Mat<R, C>& A) {
    Mat<C, R> t;
    for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) t(c, r) = A(r, c);
    return t;
}
template <int R, int C>
Mat<R, C> operator+(const Mat<R, C>& A, const Mat<R, C>& B) { Mat<R, C> o; for (int i = 0; i < R * C; ++i) o.a[i] = A.a[i] + B.a[i]; return o; }
template <int R, int C>
Mat<R, C> operator-(const Mat<R, C>& A, const Mat<R, C>& B) { Mat<R, C> o; for (int i = 0; i < R * C; ++i) o.a[i] = A.a[i] - B.a[i]; return o; }
template <int R, int C>
Mat<R, C> operator*(double s, const Mat<R, C>& A) { Mat<R, C> o; for (int i = 0; i < R * C; ++i) o.a[i] = s * A.a[i]; return o; }

typedef Mat<3, 1> V3;
typedef Mat<3, 3> M3;

V3 v3(const double* p) { V3 v; v.a[0] = p[0]; v.a[1] = p[1]; v.a[2] = p[2]; return v; }
double nrm(const V3& v) { return std::sqrt(v.a[0] * v.a[0] + (v.a[1] * v.a[1] + v.a[2] * v.a[2])); }
V3 unit(const V3& v) { double n2 = v.a[0] * v.a[0] + (v.a[1] * v.a[1] + v.a[2] * v.a[2]); if (n2 > 0) { double n = std::sqrt(n2); V3 o; for (int i = 0; i < 3; ++i) o.a[i] = v.a[i] / n; return o; } return v; }
M3 hat(const V3& v) { M3 m = M3::zero(); m(0, 1) = -v.a[2]; m(0, 2) = v.a[1]; m(1, 0) = v.a[2]; m(1, 2) = -v.a[0]; m(2, 0) = -v.a[1]; m(2, 1) = v.a[0]; return m; }

struct Q { double x, y, z, w; };
double qn2(const Q& q) { return (q.x * q.x + q.z * q.z) + (q.y * q.y + q.w * q.w); }
Q qunit(const Q& q) { double n2 = qn2(q); if (n2 > 0) { double n = std::sqrt(n2); return {q.x / n, q.y / n, q.z / n, q.w / n}; } return q; }
Q qmul(const Q& a, const Q& b) {
    return {a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y, a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
            a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x, a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z};
}
Q qinv(const Q& q) { double n2 = qn2(q); if (n2 > 0) return {-q.x / n2, -q.y / n2, -q.z / n2, q.w / n2}; return {0, 0, 0, 0}; }
M3 qrot(const Q& q) { double qq[4] = {q.x, q.y, q.z, q.w}; M3 R; srl::quat_to_rot(qq, R.a); return R; }
Q rot2q(const M3& m) {   // Eigen's matrix -> quaternion
    double q[4];
    double t = m(0, 0) + m(1, 1) + m(2, 2);
    if (t > 0) {
        t = std::sqrt(t + 1.0); q[3] = 0.5 * t; t = 0.5 / t;
        q[0] = (m(2, 1) - m(1, 2)) * t; q[1] = (m(0, 2) - m(2, 0)) * t; q[2] = (m(1, 0) - m(0, 1)) * t;
    } else {
        int i = 0;
        if (m(1, 1) > m(0, 0)) i = 1;
        if (m(2, 2) > m(i, i)) i = 2;
        int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0); q[i] = 0.5 * t; t = 0.5 / t;
        q[3] = (m(k, j) - m(j, k)) * t; q[j] = (m(j, i) + m(i, j)) * t; q[k] = (m(k, i) + m(i, k)) * t;
    }
    return {q[0], q[1], q[2], q[3]};
}
// numType::rotationToSo3 — normalizeR then acos, not clamped (include/utility.h:267-280)
V3 log_so3(const M3& Rin) {
    M3 R = qrot(qunit(rot2q(Rin)));
    double th = std::acos((R(0, 0) + R(1, 1) + R(2, 2) - 1.0) / 2.0);
    V3 a; a.a[0] = R(2, 1) - R(1, 2); a.a[1] = R(0, 2) - R(2, 0); a.a[2] = R(1, 0) - R(0, 1);
    V3 o;
    if (th < kTheta) for (int i = 0; i < 3; ++i) o.a[i] = a.a[i] / 2.0;
    else for (int i = 0; i < 3; ++i) o.a[i] = th * a.a[i] / (2.0 * std::sin(th));
    return o;
}
// numType::so3ToRotation (include/utility.h:282-299)
M3 exp_so3(const V3& w) {
    double th = nrm(w);
    if (th < kTheta) { M3 U = hat(w); return M3::identity() + U + 0.5 * (U * U); }
    M3 U = hat(unit(w));
    return M3::identity() + std::sin(th) * U + (1.0 - std::cos(th)) * (U * U);
}
// numType::so3ToQuat (include/utility.h:301-324)
Q exp_quat(const V3& w) {
    double th = nrm(w);
    if (th < kTheta) return qunit({w.a[0] / 2.0, w.a[1] / 2.0, w.a[2] / 2.0, 1.0});
    V3 u = unit(w);
    double s = std::sin(0.5 * th), c = std::cos(0.5 * th);
    return qunit({u.a[0] * s, u.a[1] * s, u.a[2] * s, c});
}
// numType::derivativeS2 (include/utility.h:215-235)
Mat<3, 2> s2_basis(const V3& gin) {
    V3 g = unit(gin);
    Mat<3, 2> B;
    B(0, 0) = 1.0 - g.a[0] * g.a[0] / (1.0 + g.a[2]);
    B(0, 1) = -g.a[0] * g.a[1] / (1.0 + g.a[2]);
    B(1, 0) = B(0, 1);
    B(1, 1) = 1.0 - g.a[1] * g.a[1] / (1.0 + g.a[2]);
    B(2, 0) = -g.a[0];
    B(2, 1) = -g.a[1];
    return B;
}
// AngularDistance(const Vector3d&) (src/utility.cpp:146-153), degrees, acos not clamped
double angular_distance(const V3& w) {
    M3 R = exp_so3(w);
    return std::acos((R(0, 0) + R(1, 1) + R(2, 2) - 1.0) / 2.0) * 180.0 / M_PI;
}

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
    Q q = qunit(qmul({s->q[0], s->q[1], s->q[2], s->q[3]}, exp_quat(v3(d_x + 3))));
    s->q[0] = q.x; s->q[1] = q.y; s->q[2] = q.z; s->q[3] = q.w;
    for (int i = 0; i < 3; ++i) { s->v[i] += d_x[6 + i]; s->ba[i] += d_x[9 + i]; s->bg[i] += d_x[12 + i]; }
    V3 g = v3(s->g);
    Mat<3, 2> B = s2_basis(g);
    Mat<2, 1> dg; dg.a[0] = d_x[15]; dg.a[1] = d_x[16];
    V3 gn = exp_so3(B * dg) * g;
    for (int i = 0; i < 3; ++i) s->g[i] = gn.a[i];
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
    Q dq = qmul(qinv({pr.q[0], pr.q[1], pr.q[2], pr.q[3]}), {eskf->q[0], eskf->q[1], eskf->q[2], eskf->q[3]});
    V3 d_so3 = log_so3(qrot(dq));
    for (int i = 0; i < 3; ++i) dx[3 + i] = d_so3.a[i];

    V3 gp = unit(v3(pr.g)), gc = unit(v3(eskf->g));
    V3 cr; cr.a[0] = gp.a[1] * gc.a[2] - gp.a[2] * gc.a[1]; cr.a[1] = gp.a[2] * gc.a[0] - gp.a[0] * gc.a[2]; cr.a[2] = gp.a[0] * gc.a[1] - gp.a[1] * gc.a[0];
    double dot = gp.a[0] * gc.a[0] + (gp.a[1] * gc.a[1] + gp.a[2] * gc.a[2]);
    M3 R_dg;
    if (std::fabs(1.0 - dot) < 1e-6) R_dg = M3::identity();
    else {
        M3 sk = hat(cr);
        M3 sk2 = sk * sk;
        const double den = cr.a[0] * cr.a[0] + cr.a[1] * cr.a[1] + cr.a[2] * cr.a[2];
        R_dg = M3::identity() + sk;
        for (int e = 0; e < 9; ++e) R_dg.a[e] += sk2.a[e] * (1.0 - dot) / den;   // :197-198
    }
    V3 so3_dg = log_so3(R_dg);
    Mat<3, 2> Bp = s2_basis(v3(pr.g));
    Mat<2, 1> d_g = tr(Bp) * so3_dg;
    dx[15] = d_g.a[0]; dx[16] = d_g.a[1];

    M3 J_so3 = M3::identity() - 0.5 * hat(d_so3);                                         // :213
    Mat<2, 2> J_s2 = Mat<2, 2>::identity() + 0.5 * (tr(Bp) * (hat(so3_dg) * Bp));         // :214
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
        Mat<3, 2> Bb = s2_basis(g_before);
        Mat<2, 1> dg2; dg2.a[0] = d_x[15]; dg2.a[1] = d_x[16];
        J_so3 = M3::identity() - 0.5 * hat(dth);                                          // :278
        J_s2 = Mat<2, 2>::identity() + 0.5 * (tr(Bb) * (hat(Bb * dg2) * Bb));             // :279
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
