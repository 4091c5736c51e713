// SE3 / so3 primitives for the IK kernels, templated on the scalar (double or float).
//
// Everything here operates on values held in registers (small fixed-size arrays, fully unrolled); there is
// no memory traffic.  The formulas are the ones Pinocchio applies on the reference's hot path -- log6 / Jlog6
// (reference frame.hpp:54-60,165-166), exp6 + quaternion update inside pinocchio::integrate (dls.cpp:67-68),
// quaternion -> rotation in the free-flyer FK (data.cpp:28-29) -- written out from SURVEY.md 8c.3-6.
// Conventions: rotation row-major R[9]; spatial vectors [linear; angular]; quaternion (x, y, z, w).
//
// The functions are __host__ __device__ so tests/cpu_harness can unit-test the very same source on the CPU
// build box (no GPU there); the product only ever calls them from device code.
#pragma once
#include "fast_math.cuh"

namespace ikb {

template <typename T> struct Num;
template <> struct Num<double> {
    // Pinocchio TaylorSeriesExpansion<double>::precision<3>() = eps^(1/4), precision<2>() = eps^(1/3)
    static IKB_HD double taylor3() { return 1.220703125e-4; }
    static IKB_HD double taylor2() { return 6.0554544523933395e-6; }
    static IKB_HD double pi() { return 3.14159265358979323846; }
    static IKB_HD double tiny() { return 1e-290; }
};
template <> struct Num<float> {
    static IKB_HD float taylor3() { return 1.8581361e-2f; }
    static IKB_HD float taylor2() { return 4.9215666e-3f; }
    static IKB_HD float pi() { return 3.14159265358979323846f; }
    static IKB_HD float tiny() { return 1e-30f; }
};

template <typename T> IKB_HD T dot3(const T *a, const T *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
template <typename T> IKB_HD void cross3(const T *a, const T *b, T *c) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}
// o = R v
template <typename T> IKB_HD void rot_vec(const T *R, const T *v, T *o) {
#pragma unroll
    for (int i = 0; i < 3; ++i) o[i] = R[3 * i] * v[0] + R[3 * i + 1] * v[1] + R[3 * i + 2] * v[2];
}
// o = R^T v
template <typename T> IKB_HD void rotT_vec(const T *R, const T *v, T *o) {
#pragma unroll
    for (int i = 0; i < 3; ++i) o[i] = R[i] * v[0] + R[3 + i] * v[1] + R[6 + i] * v[2];
}
template <typename T> IKB_HD void mat3_mul(const T *A, const T *B, T *C) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
template <typename T> IKB_HD void mat3T_mul(const T *A, const T *B, T *C) {  // A^T B
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) C[3 * i + j] = A[i] * B[j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j];
}

// (Ra,pa)*(Rb,pb)
template <typename T> IKB_HD void se3_mul(const T *Ra, const T *pa, const T *Rb, const T *pb, T *Rc, T *pc) {
    T R[9], p[3];
    mat3_mul(Ra, Rb, R);
    rot_vec(Ra, pb, p);
#pragma unroll
    for (int i = 0; i < 9; ++i) Rc[i] = R[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) pc[i] = pa[i] + p[i];
}
// (Ra,pa)^-1 * (Rb,pb)
template <typename T> IKB_HD void se3_actinv(const T *Ra, const T *pa, const T *Rb, const T *pb, T *Rc, T *pc) {
    T R[9], d[3], p[3];
    mat3T_mul(Ra, Rb, R);
#pragma unroll
    for (int i = 0; i < 3; ++i) d[i] = pb[i] - pa[i];
    rotT_vec(Ra, d, p);
#pragma unroll
    for (int i = 0; i < 9; ++i) Rc[i] = R[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) pc[i] = p[i];
}

// Eigen QuaternionBase::toRotationMatrix (free-flyer joint transform); the quaternion is not re-normalised.
template <typename T> IKB_HD void quat_to_rot(T x, T y, T z, T w, T *R) {
    const T tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const T twx = tx * w, twy = ty * w, twz = tz * w;
    const T txx = tx * x, txy = ty * x, txz = tz * x;
    const T tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
    R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

// rotation matrix -> quaternion (x,y,z,w), Eigen's branch on the trace / largest diagonal entry.
template <typename T> IKB_HD void rot_to_quat(const T *R, T *q) {
    T t = R[0] + R[4] + R[8];
    if (t > T(0)) {
        t = sqrt_(t + T(1));
        q[3] = T(0.5) * t;
        t = T(0.5) * rcp_(t);
        q[0] = (R[7] - R[5]) * t;
        q[1] = (R[2] - R[6]) * t;
        q[2] = (R[3] - R[1]) * t;
    } else if (R[0] >= R[4] && R[0] >= R[8]) {
        t = sqrt_(R[0] - R[4] - R[8] + T(1));
        q[0] = T(0.5) * t;
        t = T(0.5) * rcp_(t);
        q[3] = (R[7] - R[5]) * t;
        q[1] = (R[3] + R[1]) * t;
        q[2] = (R[6] + R[2]) * t;
    } else if (R[4] > R[0] && R[4] >= R[8]) {
        t = sqrt_(R[4] - R[8] - R[0] + T(1));
        q[1] = T(0.5) * t;
        t = T(0.5) * rcp_(t);
        q[3] = (R[2] - R[6]) * t;
        q[2] = (R[7] + R[5]) * t;
        q[0] = (R[1] + R[3]) * t;
    } else {
        t = sqrt_(R[8] - R[0] - R[4] + T(1));
        q[2] = T(0.5) * t;
        t = T(0.5) * rcp_(t);
        q[3] = (R[3] - R[1]) * t;
        q[0] = (R[2] + R[6]) * t;
        q[1] = (R[5] + R[7]) * t;
    }
}

// exp6([v; w]) -> (R, p)   (SURVEY 8c.6).  Branch-free: both the closed form and the small-angle series are
// evaluated and selected (the closed form's denominators are guarded; its value is discarded below the threshold).
template <typename T> IKB_HD void exp6(const T *v, const T *w, T *R, T *p) {
    const T t2 = dot3(w, w);
    const bool small = t2 < Num<T>::taylor3() * Num<T>::taylor3();
    const T t = sqrt_(t2);
    T st, ct;
    sincos_(t, &st, &ct);
    const T inv_t = rcp_(max_(t, Num<T>::tiny()));
    const T inv_t2 = inv_t * inv_t;
    const T g_av = st * inv_t;
    const T a_wxv = small ? T(0.5) - t2 * T(1.0 / 24) : (1 - ct) * inv_t2;
    const T a_v = small ? 1 - t2 * T(1.0 / 6) : g_av;
    const T a_w = small ? T(1.0 / 6) - t2 * T(1.0 / 120) : (1 - g_av) * inv_t2;
    const T diag = small ? 1 - t2 * T(0.5) : ct;
    T wxv[3];
    cross3(w, v, wxv);
    const T wv = a_w * dot3(w, v);
#pragma unroll
    for (int i = 0; i < 3; ++i) p[i] = a_v * v[i] + wv * w[i] + a_wxv * wxv[i];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) R[3 * i + j] = a_wxv * w[i] * w[j];
    R[1] -= a_v * w[2]; R[3] += a_v * w[2];
    R[2] += a_v * w[1]; R[6] -= a_v * w[1];
    R[5] -= a_v * w[0]; R[7] += a_v * w[0];
    R[0] += diag; R[4] += diag; R[8] += diag;
}

// log3(R) -> w, theta, and (sin, cos) of theta   (diagonal formula within 1e-2 of pi).
// theta = acos((tr R - 1)/2) exactly as Pinocchio computes it -- the closed-form coefficients below cancel
// catastrophically for small angles (1/t^2 - ..., -2/t^4 + ...), so two implementations only agree closely if they
// feed those formulas the same way; the acos / sincos used are the branch-free ones of fast_math.cuh.  The hot path
// has no branch; the near-pi formula stays behind a (cold) one.
template <typename T> IKB_HD void log3(const T *R, T *w, T &theta, T &st, T &ct) {
    const T tr = R[0] + R[4] + R[8];
    const T t = acos_(min_(T(1), max_(T(-1), T(0.5) * (tr - 1))));
    sincos_(t, &st, &ct);
    const T s = t > Num<T>::taylor2() ? T(0.5) * t * rcp_(max_(st, Num<T>::tiny())) : T(0.5);
    w[0] = s * (R[7] - R[5]);
    w[1] = s * (R[2] - R[6]);
    w[2] = s * (R[3] - R[1]);
    if (t >= Num<T>::pi() - T(1e-2)) {
        const T cphi = -(tr - 1) / 2;
        const T beta = t * t / (1 + cphi);
        const T d0 = (R[0] + cphi) * beta, d1 = (R[4] + cphi) * beta, d2 = (R[8] + cphi) * beta;
        w[0] = (R[7] > R[5] ? T(1) : T(-1)) * (d0 > 0 ? sqrt_(d0) : T(0));
        w[1] = (R[2] > R[6] ? T(1) : T(-1)) * (d1 > 0 ? sqrt_(d1) : T(0));
        w[2] = (R[3] > R[1] ? T(1) : T(-1)) * (d2 > 0 ? sqrt_(d2) : T(0));
    }
    theta = t;
}
template <typename T> IKB_HD void log3(const T *R, T *w, T &theta) {
    T st, ct;
    log3(R, w, theta, st, ct);
}

// Shared trigonometric coefficients of log6 / Jlog3 / Jlog6 for one rotation angle.
template <typename T> struct LogCoeffs {
    T alpha;     // log6:  t sin t / (2 (1 - cos t))
    T beta;      // log6 / Jlog6: 1/t^2 - sin t / (2 t (1 - cos t))
    T bdot;      // Jlog6: beta_dot_over_theta
    T a3;        // Jlog3 alpha: 1/t^2 - sin t/(1-cos t)/(2 t)   (== beta)
    T diag3;     // Jlog3 diagonal: t sin t /(2 (1 - cos t))     (== alpha)
};
// (st, ct) = (sin t, cos t) as returned by log3.  Branch-free select between closed form and series.
template <typename T> IKB_HD LogCoeffs<T> log_coeffs(T t, T st, T ct) {
    LogCoeffs<T> c;
    const T t2 = t * t;
    const bool small = t < Num<T>::taylor3();
    const T tinv = rcp_(max_(t, Num<T>::tiny())), t2inv = tinv * tinv;
    const T inv_2_2ct = rcp_(max_(2 * (1 - ct), Num<T>::tiny()));
    const T g_alpha = t * st * inv_2_2ct;
    const T g_beta = t2inv - st * tinv * inv_2_2ct;
    const T g_bdot = -2 * t2inv * t2inv + (1 + st * tinv) * t2inv * inv_2_2ct;
    c.alpha = small ? 1 - t2 * T(1.0 / 12) - t2 * t2 * T(1.0 / 720) : g_alpha;
    c.beta = small ? T(1.0 / 12) + t2 * T(1.0 / 720) : g_beta;
    c.bdot = small ? T(1.0 / 360) : g_bdot;
    c.a3 = c.beta;
    c.diag3 = small ? T(0.5) * (2 - t2 * T(1.0 / 6)) : g_alpha;
    return c;
}
template <typename T> IKB_HD LogCoeffs<T> log_coeffs(T t) {
    T st, ct;
    sincos_(t, &st, &ct);
    return log_coeffs(t, st, ct);
}

// log6 given log3 output
template <typename T> IKB_HD void log6_from(const T *w, const LogCoeffs<T> &c, const T *p, T *lin) {
    T wxp[3];
    cross3(w, p, wxp);
    const T bwp = c.beta * dot3(w, p);
#pragma unroll
    for (int i = 0; i < 3; ++i) lin[i] = c.alpha * p[i] - T(0.5) * wxp[i] + bwp * w[i];
}

template <typename T> IKB_HD void add_skew(const T *v, T s, T *M) {
    M[1] -= s * v[2]; M[2] += s * v[1];
    M[3] += s * v[2]; M[5] -= s * v[0];
    M[6] -= s * v[1]; M[7] += s * v[0];
}

// Jlog6(M) = [[A, B], [0, A]] for M = (R, p) whose log3 is (w, t)    
template <typename T> IKB_HD void jlog6_blocks(const T *w, T t, const LogCoeffs<T> &c, const T *p, T *A, T *B) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) A[3 * i + j] = c.a3 * w[i] * w[j];
    A[0] += c.diag3; A[4] += c.diag3; A[8] += c.diag3;
    add_skew(w, T(0.5), A);
    const T wTp = dot3(w, p);
    const T k1 = c.bdot * wTp, k2 = t * t * c.bdot + 2 * c.beta;
    T v3[3], C[9];
#pragma unroll
    for (int i = 0; i < 3; ++i) v3[i] = k1 * w[i] - k2 * p[i];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) C[3 * i + j] = v3[i] * w[j] + c.beta * w[i] * p[j];
    const T d = wTp * c.beta;
    C[0] += d; C[4] += d; C[8] += d;
    add_skew(p, T(0.5), C);
    mat3_mul(C, A, B);
}

// Free-flyer configuration update of pinocchio::integrate (SpecialEuclideanOperation<3>::integrate_impl):
// M1 = M0 * exp6(v); quaternion(M1.R) with the sign continuous with the old one; first-order normalisation.
// R0 is the rotation already computed from the current quaternion by the FK of this iteration.
template <typename T> IKB_HD void integrate_freeflyer(const T *R0, T *pos, T *quat, const T *v /*[6]*/) {
    T Re[9], pe[3], R1[9], p1[3];
    exp6(v, v + 3, Re, pe);
    se3_mul(R0, pos, Re, pe, R1, p1);
    T qn[4];
    rot_to_quat(R1, qn);
    const T d = qn[0] * quat[0] + qn[1] * quat[1] + qn[2] * quat[2] + qn[3] * quat[3];
    const T sgn = d < T(0) ? T(-1) : T(1);
    const T n2 = qn[0] * qn[0] + qn[1] * qn[1] + qn[2] * qn[2] + qn[3] * qn[3];
    const T a = sgn * (3 - n2) * T(0.5);
#pragma unroll
    for (int i = 0; i < 4; ++i) quat[i] = qn[i] * a;
#pragma unroll
    for (int i = 0; i < 3; ++i) pos[i] = p1[i];
}

}  // namespace ikb
