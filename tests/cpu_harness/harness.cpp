// TEST-ONLY: compiles the product's device math header (ik_b200/csrc/se3_math.cuh, __host__ __device__
// templates) with g++ so the formulas can be unit-tested against the oracle on the GPU-less build box.
// Never linked into libikb200.so; the product has no CPU path.
#include "../../ik_b200/csrc/se3_math.cuh"

using namespace ikb;

template <typename T> static void log6_t(const T *M, T *out) {
    T w[3], th, st, ct;
    log3(M, w, th, st, ct);
    LogCoeffs<T> c = log_coeffs(th, st, ct);
    log6_from(w, c, M + 9, out);
    out[3] = w[0]; out[4] = w[1]; out[5] = w[2];
}
template <typename T> static void jlog6_t(const T *M, T *J) {
    T w[3], th, st, ct, A[9], B[9];
    log3(M, w, th, st, ct);
    LogCoeffs<T> c = log_coeffs(th, st, ct);
    jlog6_blocks(w, th, c, M + 9, A, B);
    for (int i = 0; i < 36; ++i) J[i] = 0;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            J[6 * i + j] = A[3 * i + j];
            J[6 * i + 3 + j] = B[3 * i + j];
            J[6 * (3 + i) + 3 + j] = A[3 * i + j];
        }
}
template <typename T> static void integrate_ff_t(const T *q7, const T *v6, T *out7) {
    T R0[9], pos[3] = {q7[0], q7[1], q7[2]}, quat[4] = {q7[3], q7[4], q7[5], q7[6]};
    quat_to_rot(quat[0], quat[1], quat[2], quat[3], R0);
    integrate_freeflyer(R0, pos, quat, v6);
    for (int i = 0; i < 3; ++i) out7[i] = pos[i];
    for (int i = 0; i < 4; ++i) out7[3 + i] = quat[i];
}

extern "C" {
void h_exp6_d(const double *v, double *M) { exp6(v, v + 3, M, M + 9); }
void h_exp6_f(const float *v, float *M) { exp6(v, v + 3, M, M + 9); }
void h_log6_d(const double *M, double *o) { log6_t(M, o); }
void h_log6_f(const float *M, float *o) { log6_t(M, o); }
void h_jlog6_d(const double *M, double *J) { jlog6_t(M, J); }
void h_jlog6_f(const float *M, float *J) { jlog6_t(M, J); }
void h_integrate_ff_d(const double *q, const double *v, double *o) { integrate_ff_t(q, v, o); }
void h_integrate_ff_f(const float *q, const float *v, float *o) { integrate_ff_t(q, v, o); }
void h_rot_to_quat_d(const double *R, double *q) { rot_to_quat(R, q); }
void h_quat_to_rot_d(const double *q, double *R) { quat_to_rot(q[0], q[1], q[2], q[3], R); }
void h_se3_actinv_d(const double *A, const double *B, double *C) { se3_actinv(A, A + 9, B, B + 9, C, C + 9); }
void h_sincos_d(double x, double *s, double *c) { sincos_(x, s, c); }
double h_acos_d(double x) { return acos_(x); }
double h_atan2pos_d(double y, double x) { return atan2pos_(y, x); }
void h_se3_mul_d(const double *A, const double *B, double *C) { se3_mul(A, A + 9, B, B + 9, C, C + 9); }
}
