// Branch-free scalar math for the IK kernels.
//
// The CUDA library versions of sincos / acos / sqrt / division each carry a rarely-taken slow path (huge arguments,
// denormals, ...) behind a branch and a CALL.  In the generated straight-line solver bodies (gen/*.cuh) those
// branches chop the code into ~100 tiny basic blocks, and because ptxas only schedules inside a basic block the
// dependent DFMA chains of e.g. fourteen consecutive joint sincos cannot overlap.  The versions below have NO
// branches: one evaluation path, selects instead of jumps.  Accuracy is ~1 ulp-level (approximation errors are listed
// next to the coefficients, produced by tools/gen_math_coeffs.py); the domain restrictions are stated per function
// and hold for everything the IK path feeds them (joint angles, rotation angles in [0, pi], LDL^T pivots >= damping^2).
//
// double: hand-written (MUFU seed + Newton on the device; the same polynomials on the host for the CPU harness).
// float : the library functions (FP32 is not the parity path; their slow paths are cheap).
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define IKB_HD __host__ __device__ __forceinline__
#else
#define IKB_HD inline
#endif

namespace ikb {

// ---- reciprocal / square root ---------------------------------------------------------------------------
// 1/x for normal, finite x (|x| in [1e-290, 1e290]); two Newton steps on the 20-bit MUFU.RCP64H seed.
IKB_HD double rcp_(double x) {
#if defined(__CUDA_ARCH__)
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
#else
    return 1.0 / x;
#endif
}
IKB_HD float rcp_(float x) { return 1.0f / x; }

// sqrt(x) for x >= 0 (0 for x <= 0); MUFU.RSQ64H seed, two Newton steps on 1/sqrt, one correction on sqrt.
IKB_HD double sqrt_(double x) {
#if defined(__CUDA_ARCH__)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double h = 0.5 * x;
    double e = fma(-h * y, y, 0.5);
    y = fma(y, e, y);
    e = fma(-h * y, y, 0.5);
    y = fma(y, e, y);
    double s = x * y;
    const double r = fma(-s, s, x);
    s = fma(r, 0.5 * y, s);
    return x > 0.0 ? s : 0.0;
#else
    return x > 0.0 ? std::sqrt(x) : 0.0;
#endif
}
IKB_HD float sqrt_(float x) { return x > 0.0f ? sqrtf(x) : 0.0f; }

// ---- sine and cosine --------------------------------------------------------------------------------------
// Valid for |x| < 2^30 (absolute error < 3e-16; joint angles are radians inside URDF limits).  Three-part
// Cody-Waite reduction by pi/2 with FMA, then the two kernels on |r| <= pi/4 (max abs error 1.9e-17 / 9.2e-19).
IKB_HD void sincos_(double x, double *sp, double *cp) {
#if defined(__CUDA_ARCH__)
    const int k = __double2int_rn(x * 6.36619772367581382e-01);
    const double n = (double)k;
#define IKB_FMA fma
#else
    const double n = std::nearbyint(x * 6.36619772367581382e-01);
    const int k = (int)n;
#define IKB_FMA std::fma
#endif
    double r = IKB_FMA(-n, 1.57079632679489656e+00, x);
    r = IKB_FMA(-n, 6.12323399573676604e-17, r);
    r = IKB_FMA(-n, -1.49738490485916983e-33, r);
    const double z = r * r;
    double ps = 1.59145338203463061e-10;
    ps = IKB_FMA(ps, z, -2.50510813003190153e-08);
    ps = IKB_FMA(ps, z, 2.75573158569131213e-06);
    ps = IKB_FMA(ps, z, -1.98412698362788982e-04);
    ps = IKB_FMA(ps, z, 8.33333333333062705e-03);
    ps = IKB_FMA(ps, z, -1.66666666666666630e-01);
    const double sr = IKB_FMA(ps * z, r, r);
    double pc = -1.13803831008215941e-11;
    pc = IKB_FMA(pc, z, 2.08761146479300446e-09);
    pc = IKB_FMA(pc, z, -2.75573171180605570e-07);
    pc = IKB_FMA(pc, z, 2.48015872984656125e-05);
    pc = IKB_FMA(pc, z, -1.38888888888871959e-03);
    pc = IKB_FMA(pc, z, 4.16666666666666644e-02);
    const double cr = IKB_FMA(pc * z, z, IKB_FMA(-0.5, z, 1.0));
    const double s = (k & 1) ? cr : sr;
    const double c = (k & 1) ? sr : cr;
    *sp = (k & 2) ? -s : s;
    *cp = ((k + 1) & 2) ? -c : c;
}
IKB_HD void sincos_(float x, float *s, float *c) {
#if defined(__CUDA_ARCH__)
    sincosf(x, s, c);
#else
    *s = std::sin(x);
    *c = std::cos(x);
#endif
}

// ---- atan2(y, x) for y >= 0: the angle in [0, pi] of a (sin, cos) pair -------------------------------------
// Octant reduction with ONE reciprocal: a = min/max in [0,1]; for a > tan(pi/8) use atan(a) = pi/4 + atan((a-1)/(a+1));
// odd polynomial on |t| <= sqrt(2)-1 (max abs error 3.3e-18).  atan2(0, 0) = 0.
IKB_HD double atan2pos_(double y, double x) {
    const double ax = fabs(x);
    const double mx = fmax(fmax(ax, y), 1e-300), mn = fmin(ax, y);
    const bool big = mn > 0.41421356237309503 * mx;
    const double num = big ? mn - mx : mn;
    const double den = big ? mn + mx : mx;
    const double t = num * rcp_(den);
    const double z = t * t;
    double p = -1.91053729754003358e-02;
    p = IKB_FMA(p, z, 3.91731226082932538e-02);
    p = IKB_FMA(p, z, -5.08341864538292484e-02);
    p = IKB_FMA(p, z, 5.85775983939145484e-02);
    p = IKB_FMA(p, z, -6.66446655080659978e-02);
    p = IKB_FMA(p, z, 7.69217999576784078e-02);
    p = IKB_FMA(p, z, -9.09090444025744543e-02);
    p = IKB_FMA(p, z, 1.11111110118550543e-01);
    p = IKB_FMA(p, z, -1.42857142846241514e-01);
    p = IKB_FMA(p, z, 1.99999999999953160e-01);
    p = IKB_FMA(p, z, -3.33333333333333315e-01);
    double r = IKB_FMA(p * z, t, t);
    r = big ? r + 7.85398163397448279e-01 : r;
    r = y > ax ? 1.57079632679489656e+00 - r : r;
    return x < 0.0 ? 3.14159265358979312e+00 - r : r;
}
IKB_HD float atan2pos_(float y, float x) { return atan2f(y, x); }

// ---- acos(x), x in [-1, 1] ------------------------------------------------------------------------------------
// |x| <= 0.5: pi/2 - asin(|x|); |x| > 0.5: 2 asin(sqrt((1 - |x|)/2)) (1 - |x| is exact there); reflected for x < 0.
// asin(t) = t + t^3 P(t^2) on t <= 0.5 (max abs error 2.6e-18).  Same conditioning as libm's acos, no branch.
IKB_HD double acos_(double x) {
    const double ax = fabs(x);
    const bool big = ax > 0.5;
    const double z = big ? 0.5 * (1.0 - ax) : x * x;
    const double t = big ? sqrt_(z) : ax;
    double p = 2.88609743879985302e-02;
    p = IKB_FMA(p, z, -1.49998557838113486e-02);
    p = IKB_FMA(p, z, 1.74935694762482045e-02);
    p = IKB_FMA(p, z, 5.42423083980400102e-03);
    p = IKB_FMA(p, z, 1.03303702954044060e-02);
    p = IKB_FMA(p, z, 1.14780474725101455e-02);
    p = IKB_FMA(p, z, 1.39713253287287714e-02);
    p = IKB_FMA(p, z, 1.73523853934809472e-02);
    p = IKB_FMA(p, z, 2.23721732438322066e-02);
    p = IKB_FMA(p, z, 3.03819441312371992e-02);
    p = IKB_FMA(p, z, 4.46428571464460508e-02);
    p = IKB_FMA(p, z, 7.49999999999838851e-02);
    p = IKB_FMA(p, z, 1.66666666666666685e-01);
    const double as = IKB_FMA(p * z, t, t);
    const double r = big ? 2.0 * as : (1.57079632679489656e+00 - as) + 6.12323399573676604e-17;
    return x < 0.0 ? (3.14159265358979312e+00 - r) + 1.22464679914735321e-16 : r;
}
IKB_HD float acos_(float x) { return acosf(x); }
#undef IKB_FMA

IKB_HD double abs_(double x) { return fabs(x); }
IKB_HD float abs_(float x) { return fabsf(x); }
IKB_HD double min_(double a, double b) { return fmin(a, b); }
IKB_HD float min_(float a, float b) { return fminf(a, b); }
IKB_HD double max_(double a, double b) { return fmax(a, b); }
IKB_HD float max_(float a, float b) { return fmaxf(a, b); }

}  // namespace ikb
