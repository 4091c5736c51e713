#!/usr/bin/env python3
"""Derive the polynomial coefficients of the branch-free device math in ik_b200/csrc/fast_math.cuh.

Chebyshev-node interpolation in 60-digit arithmetic (mpmath), coefficients rounded to double, maximum error of the
rounded polynomial reported on a dense grid.  Run:  python tools/gen_math_coeffs.py
"""
import mpmath as mp

mp.mp.dps = 60


def cheb_fit(f, a, b, deg):
    n = deg + 1
    xs = [(a + b) / 2 + (b - a) / 2 * mp.cos(mp.pi * (2 * k + 1) / (2 * n)) for k in range(n)]
    A = mp.matrix(n, n)
    y = mp.matrix(n, 1)
    for i, x in enumerate(xs):
        for j in range(n):
            A[i, j] = x ** j
        y[i] = f(x)
    c = mp.lu_solve(A, y)
    return [float(c[j]) for j in range(n)]


def max_err(f, coeffs, a, b, weight=lambda x: 1, npts=4001):
    worst = mp.mpf(0)
    for k in range(npts):
        x = a + (b - a) * mp.mpf(k) / (npts - 1)
        p = mp.mpf(0)
        for c in reversed(coeffs):
            p = p * x + mp.mpf(c)
        worst = max(worst, abs((p - f(x)) * weight(x)))
    return worst


def show(name, coeffs):
    print("// %s" % name)
    print(", ".join("%.17e" % c for c in coeffs))


def main():
    zmax = (mp.pi / 4 + mp.mpf("0.01")) ** 2

    def fs(z):
        if z == 0:
            return -mp.mpf(1) / 6
        r = mp.sqrt(z)
        return (mp.sin(r) / r - 1) / z

    def fc(z):
        if z == 0:
            return mp.mpf(1) / 24
        r = mp.sqrt(z)
        return (mp.cos(r) - 1 + z / 2) / (z * z)

    S = cheb_fit(fs, mp.mpf(0), zmax, 5)
    Cc = cheb_fit(fc, mp.mpf(0), zmax, 5)
    show("sin(r) = r + r^3 * S(r^2), |r| <= pi/4 + 0.01", S)
    print("//   max abs error of sin: %s" % mp.nstr(max_err(fs, S, mp.mpf(0), zmax, lambda z: mp.sqrt(z) ** 3), 5))
    show("cos(r) = 1 - r^2/2 + r^4 * C(r^2)", Cc)
    print("//   max abs error of cos: %s" % mp.nstr(max_err(fc, Cc, mp.mpf(0), zmax, lambda z: z * z), 5))

    tmax = mp.sqrt(2) - 1 + mp.mpf("0.001")

    def fa(z):
        if z == 0:
            return -mp.mpf(1) / 3
        t = mp.sqrt(z)
        return (mp.atan(t) / t - 1) / z

    for deg in (9, 10, 11, 12):
        A = cheb_fit(fa, mp.mpf(0), tmax ** 2, deg)
        e = max_err(fa, A, mp.mpf(0), tmax ** 2, lambda z: mp.sqrt(z) ** 3)
        print("// atan degree %d: max abs error %s" % (deg, mp.nstr(e, 5)))
        if e < mp.mpf("2e-17"):
            show("atan(t) = t + t^3 * A(t^2), |t| <= sqrt(2)-1", A)
            break

    def fasin(z):
        if z == 0:
            return mp.mpf(1) / 6
        t = mp.sqrt(z)
        return (mp.asin(t) / t - 1) / z

    As = cheb_fit(fasin, mp.mpf(0), mp.mpf("0.2505"), 12)
    show("asin(t) = t + t^3 * P(t^2), t <= 0.5", As)
    print("//   max abs error of asin: %s" % mp.nstr(max_err(fasin, As, mp.mpf(0), mp.mpf("0.2505"), lambda z: mp.sqrt(z) ** 3), 5))
    print("// pi = %.17e + %.17e" % (float(mp.pi), float(mp.pi - mp.mpf(float(mp.pi)))))

    # pi/2 in three doubles (Cody-Waite with FMA)
    p = mp.pi / 2
    hi = float(p)
    mid = float(p - mp.mpf(hi))
    lo = float(p - mp.mpf(hi) - mp.mpf(mid))
    print("// pi/2 = %.17e + %.17e + %.17e" % (hi, mid, lo))
    print("// 2/pi = %.17e, pi = %.17e, pi/4 = %.17e" % (float(2 / mp.pi), float(mp.pi), float(mp.pi / 4)))


if __name__ == "__main__":
    main()
