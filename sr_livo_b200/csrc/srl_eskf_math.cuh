// srl_eskf_math.cuh — the small manifold algebra of the iterated ESIKF update, host + device.
//
// What lioOptimization::updateIEKF does around buildPlaneResiduals is 3-, 4- and 17-dimensional algebra
// (src/optimize.cpp:172-310) on top of the numType helpers (include/utility.h:194-330), AngularDistance
// (src/utility.cpp:146-153) and eskfEstimator::observe (src/eskfEstimator.cpp:219-230).  The same source is compiled
// for the host loop (srl_eskf.cpp, srl_iekf_step) and for the device-resident loop (srl_iekf.cu, k_iekf_loop).
// Operation order follows Eigen's fixed-size reductions (a0 + (a1 + a2)) where the reference uses them.
#pragma once

#include <cmath>

#include "srl_math.cuh"

namespace srl {
namespace ekf {

constexpr int N = 17;
constexpr double kTheta = 1e-4;   // THETA_THRESHOLD, include/utility.h:27
constexpr double kPi = 3.14159265358979323846;

template <int R, int C>
struct Mat {
    double a[R * C];
    SRL_HD double& operator()(int r, int c) { return a[r * C + c]; }
    SRL_HD double operator()(int r, int c) const { return a[r * C + c]; }
    SRL_HD static Mat zero() { Mat m; for (int i = 0; i < R * C; ++i) m.a[i] = 0.0; return m; }
    SRL_HD static Mat identity() { Mat m = zero(); for (int i = 0; i < (R < C ? R : C); ++i) m(i, i) = 1.0; return m; }
};
template <int R, int K, int C>
SRL_HD Mat<R, C> operator*(const Mat<R, K>& A, const Mat<K, C>& B) {
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
SRL_HD Mat<C, R> tr(const Mat<R, C>& A) {
    Mat<C, R> t;
    for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) t(c, r) = A(r, c);
    return t;
}
template <int R, int C>
SRL_HD Mat<R, C> operator+(const Mat<R, C>& A, const Mat<R, C>& B) { Mat<R, C> o; for (int i = 0; i < R * C; ++i) o.a[i] = A.a[i] + B.a[i]; return o; }
template <int R, int C>
SRL_HD Mat<R, C> operator-(const Mat<R, C>& A, const Mat<R, C>& B) { Mat<R, C> o; for (int i = 0; i < R * C; ++i) o.a[i] = A.a[i] - B.a[i]; return o; }
template <int R, int C>
SRL_HD Mat<R, C> operator*(double s, const Mat<R, C>& A) { Mat<R, C> o; for (int i = 0; i < R * C; ++i) o.a[i] = s * A.a[i]; return o; }

typedef Mat<3, 1> V3;
typedef Mat<3, 3> M3;

// On the device a FP64 division is a ~35-instruction subroutine and the ESIKF step is one serial chain per warp: vectors
// are normalised with one reciprocal and multiplies there (<= 1 ulp per component from the host's per-component
// division; the parity budget is 1e-5), and sin/cos of the same angle come from one sincos.
#if defined(__CUDA_ARCH__)
#define SRL_EKF_DIV_ALL3(o, v, n) do { const double inv__ = 1.0 / (n); (o)[0] = (v)[0] * inv__; (o)[1] = (v)[1] * inv__; (o)[2] = (v)[2] * inv__; } while (0)
#define SRL_EKF_SINCOS(x, s, c) sincos((x), &(s), &(c))
#else
#define SRL_EKF_DIV_ALL3(o, v, n) do { (o)[0] = (v)[0] / (n); (o)[1] = (v)[1] / (n); (o)[2] = (v)[2] / (n); } while (0)
#define SRL_EKF_SINCOS(x, s, c) do { (s) = sin(x); (c) = cos(x); } while (0)
#endif

SRL_HD V3 v3(const double* p) { V3 v; v.a[0] = p[0]; v.a[1] = p[1]; v.a[2] = p[2]; return v; }
SRL_HD double nrm(const V3& v) { return sqrt(v.a[0] * v.a[0] + (v.a[1] * v.a[1] + v.a[2] * v.a[2])); }
SRL_HD V3 unit(const V3& v) {
    double n2 = v.a[0] * v.a[0] + (v.a[1] * v.a[1] + v.a[2] * v.a[2]);
    if (n2 > 0) { double n = sqrt(n2); V3 o; SRL_EKF_DIV_ALL3(o.a, v.a, n); return o; }
    return v;
}
SRL_HD M3 hat(const V3& v) {
    M3 m = M3::zero();
    m(0, 1) = -v.a[2]; m(0, 2) = v.a[1]; m(1, 0) = v.a[2]; m(1, 2) = -v.a[0]; m(2, 0) = -v.a[1]; m(2, 1) = v.a[0];
    return m;
}

struct Q { double x, y, z, w; };
SRL_HD Q mkq(const double* q) { Q o; o.x = q[0]; o.y = q[1]; o.z = q[2]; o.w = q[3]; return o; }
SRL_HD double qn2(const Q& q) { return (q.x * q.x + q.z * q.z) + (q.y * q.y + q.w * q.w); }
SRL_HD Q qunit(const Q& q) {
    double n2 = qn2(q);
    if (n2 > 0) {
        double n = sqrt(n2); Q o;
#if defined(__CUDA_ARCH__)
        const double inv = 1.0 / n; o.x = q.x * inv; o.y = q.y * inv; o.z = q.z * inv; o.w = q.w * inv;
#else
        o.x = q.x / n; o.y = q.y / n; o.z = q.z / n; o.w = q.w / n;
#endif
        return o;
    }
    return q;
}
SRL_HD Q qmul(const Q& a, const Q& b) {
    Q o;
    o.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    o.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
    o.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
    o.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    return o;
}
SRL_HD Q qinv(const Q& q) {
    double n2 = qn2(q);
    Q o;
#if defined(__CUDA_ARCH__)
    if (n2 > 0) { const double inv = 1.0 / n2; o.x = -q.x * inv; o.y = -q.y * inv; o.z = -q.z * inv; o.w = q.w * inv; }
#else
    if (n2 > 0) { o.x = -q.x / n2; o.y = -q.y / n2; o.z = -q.z / n2; o.w = q.w / n2; }
#endif
    else { o.x = o.y = o.z = o.w = 0.0; }
    return o;
}
SRL_HD M3 qrot(const Q& q) { double qq[4] = {q.x, q.y, q.z, q.w}; M3 R; quat_to_rot(qq, R.a); return R; }
SRL_HD Q rot2q(const M3& m) {   // Eigen's matrix -> quaternion
    double q[4];
    double t = m(0, 0) + m(1, 1) + m(2, 2);
    if (t > 0) {
        t = sqrt(t + 1.0); q[3] = 0.5 * t; t = 0.5 / t;
        q[0] = (m(2, 1) - m(1, 2)) * t; q[1] = (m(0, 2) - m(2, 0)) * t; q[2] = (m(1, 0) - m(0, 1)) * t;
    } else {
        int i = 0;
        if (m(1, 1) > m(0, 0)) i = 1;
        if (m(2, 2) > m(i, i)) i = 2;
        int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0); q[i] = 0.5 * t; t = 0.5 / t;
        q[3] = (m(k, j) - m(j, k)) * t; q[j] = (m(j, i) + m(i, j)) * t; q[k] = (m(k, i) + m(i, k)) * t;
    }
    return mkq(q);
}
// numType::rotationToSo3 — normalizeR then acos, not clamped (include/utility.h:267-280)
SRL_HD V3 log_so3(const M3& Rin) {
    M3 R = qrot(qunit(rot2q(Rin)));
    double th = acos((R(0, 0) + R(1, 1) + R(2, 2) - 1.0) / 2.0);
    V3 a; a.a[0] = R(2, 1) - R(1, 2); a.a[1] = R(0, 2) - R(2, 0); a.a[2] = R(1, 0) - R(0, 1);
    V3 o;
    if (th < kTheta) for (int i = 0; i < 3; ++i) o.a[i] = a.a[i] / 2.0;
    else {
        const double den = 2.0 * sin(th);
#if defined(__CUDA_ARCH__)
        const double f = th / den; for (int i = 0; i < 3; ++i) o.a[i] = a.a[i] * f;
#else
        for (int i = 0; i < 3; ++i) o.a[i] = th * a.a[i] / den;
#endif
    }
    return o;
}
// numType::so3ToRotation (include/utility.h:282-299)
SRL_HD M3 exp_so3(const V3& w) {
    double th = nrm(w);
    if (th < kTheta) { M3 U = hat(w); return M3::identity() + U + 0.5 * (U * U); }
    M3 U = hat(unit(w));
    double sn, cs;
    SRL_EKF_SINCOS(th, sn, cs);
    return M3::identity() + sn * U + (1.0 - cs) * (U * U);
}
// numType::so3ToQuat (include/utility.h:301-324)
SRL_HD Q exp_quat(const V3& w) {
    double th = nrm(w);
    Q o;
    if (th < kTheta) { o.x = w.a[0] / 2.0; o.y = w.a[1] / 2.0; o.z = w.a[2] / 2.0; o.w = 1.0; return qunit(o); }
    V3 u = unit(w);
    double s, c;
    SRL_EKF_SINCOS(0.5 * th, s, c);
    o.x = u.a[0] * s; o.y = u.a[1] * s; o.z = u.a[2] * s; o.w = c;
    return qunit(o);
}
// numType::derivativeS2 (include/utility.h:215-235)
SRL_HD Mat<3, 2> s2_basis(const V3& gin) {
    V3 g = unit(gin);
    Mat<3, 2> B;
#if defined(__CUDA_ARCH__)
    const double inv = 1.0 / (1.0 + g.a[2]);
    B(0, 0) = 1.0 - g.a[0] * g.a[0] * inv;
    B(0, 1) = -g.a[0] * g.a[1] * inv;
    B(1, 0) = B(0, 1);
    B(1, 1) = 1.0 - g.a[1] * g.a[1] * inv;
#else
    B(0, 0) = 1.0 - g.a[0] * g.a[0] / (1.0 + g.a[2]);
    B(0, 1) = -g.a[0] * g.a[1] / (1.0 + g.a[2]);
    B(1, 0) = B(0, 1);
    B(1, 1) = 1.0 - g.a[1] * g.a[1] / (1.0 + g.a[2]);
#endif
    B(2, 0) = -g.a[0];
    B(2, 1) = -g.a[1];
    return B;
}
// AngularDistance(const Vector3d&) (src/utility.cpp:146-153), degrees, acos not clamped
SRL_HD double angular_distance(const V3& w) {
    M3 R = exp_so3(w);
    return acos((R(0, 0) + R(1, 1) + R(2, 2) - 1.0) / 2.0) * 180.0 / kPi;
}

// ---- the two boxminus chains of src/optimize.cpp:172-218 (independent of each other) -----------------------------
// SO(3): d_so3 = Log(q_pred^-1 * q_cur) (:183-186), J_so3 = I - 0.5 [d_so3]x (:213)
SRL_HD void boxminus_so3(const double* q_pred, const double* q_cur, V3& d_so3, M3& J_so3) {
    Q dq = qmul(qinv(mkq(q_pred)), mkq(q_cur));
    d_so3 = log_so3(qrot(dq));
    J_so3 = M3::identity() - 0.5 * hat(d_so3);
}
// S^2: the rotation taking g_pred to g_cur, its log, the tangent-plane coordinates (:188-211), J_s2 (:214)
SRL_HD void boxminus_s2(const double* g_pred, const double* g_cur, Mat<2, 1>& d_g, Mat<2, 2>& J_s2) {
    V3 gp = unit(v3(g_pred)), gc = unit(v3(g_cur));
    V3 cr;
    cr.a[0] = gp.a[1] * gc.a[2] - gp.a[2] * gc.a[1]; cr.a[1] = gp.a[2] * gc.a[0] - gp.a[0] * gc.a[2]; cr.a[2] = gp.a[0] * gc.a[1] - gp.a[1] * gc.a[0];
    double dot = gp.a[0] * gc.a[0] + (gp.a[1] * gc.a[1] + gp.a[2] * gc.a[2]);
    M3 R_dg;
    if (fabs(1.0 - dot) < 1e-6) R_dg = M3::identity();
    else {
        M3 sk = hat(cr);
        M3 sk2 = sk * sk;
        const double den = cr.a[0] * cr.a[0] + cr.a[1] * cr.a[1] + cr.a[2] * cr.a[2];
        R_dg = M3::identity() + sk;
#if defined(__CUDA_ARCH__)
        const double f = (1.0 - dot) / den;
        for (int e = 0; e < 9; ++e) R_dg.a[e] += sk2.a[e] * f;
#else
        for (int e = 0; e < 9; ++e) R_dg.a[e] += sk2.a[e] * (1.0 - dot) / den;   // :197-198
#endif
    }
    V3 so3_dg = log_so3(R_dg);
    Mat<3, 2> Bp = s2_basis(v3(g_pred));
    d_g = tr(Bp) * so3_dg;
    J_s2 = Mat<2, 2>::identity() + 0.5 * (tr(Bp) * (hat(so3_dg) * Bp));
}
// eskfEstimator::observe, gravity part (src/eskfEstimator.cpp:227-229)
SRL_HD void observe_gravity(const double* g, double d0, double d1, double* g_out) {
    V3 gv = v3(g);
    Mat<3, 2> B = s2_basis(gv);
    Mat<2, 1> dg; dg.a[0] = d0; dg.a[1] = d1;
    V3 gn = exp_so3(B * dg) * gv;
    for (int i = 0; i < 3; ++i) g_out[i] = gn.a[i];
}
// eskfEstimator::observe, rotation part (src/eskfEstimator.cpp:222)
SRL_HD void observe_quat(const double* q, const double* dth, double* q_out) {
    Q r = qunit(qmul(mkq(q), exp_quat(v3(dth))));
    q_out[0] = r.x; q_out[1] = r.y; q_out[2] = r.z; q_out[3] = r.w;
}
// Jacobians of the posterior projection (:278-279)
SRL_HD void posterior_jacobians(const double* dth, const double* g_before, double d15, double d16, M3& J_so3, Mat<2, 2>& J_s2) {
    Mat<3, 2> Bb = s2_basis(v3(g_before));
    Mat<2, 1> dg2; dg2.a[0] = d15; dg2.a[1] = d16;
    J_so3 = M3::identity() - 0.5 * hat(v3(dth));
    J_s2 = Mat<2, 2>::identity() + 0.5 * (tr(Bb) * (hat(Bb * dg2) * Bb));
}

}  // namespace ekf
}  // namespace srl
