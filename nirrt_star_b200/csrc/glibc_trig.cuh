// glibc_trig.cuh -- math.sin / math.cos / np.sin / np.cos of the reference, bit for bit.
//
// The reference evaluates its trigonometry with the C library: CPython's math.sin/cos and numpy's float64 sin/cos
// loops both call libm (probed: np.sin(array) == [math.sin(x)] on 2e6 samples), which for the pinned runtime is
// glibc 2.39 on x86-64 -- the IBM Accurate Mathematical Library kernels of sysdeps/ieee754/dbl-64/s_sin.c in the
// build that the dynamic linker selects on every FMA-capable CPU (sysdeps/x86_64/fpu/multiarch: __sin_fma /
// __cos_fma, compiled with -mfma -mavx2, i.e. with the compiler's FMA contraction applied).  Those kernels are
// accurate to ~0.52 ulp, NOT correctly rounded (0.15 % of arguments differ from the correctly rounded value), so a
// correctly rounded device evaluation cannot reproduce them: the operation sequence itself is restated here --
// do_sin / do_cos / TAYLOR_SIN / reduce_sincos with the fused operations exactly where that build has them, and
// glibc's own 440-entry table (glibc_sincostab.inc; its low words are not the exactly rounded residuals, the results
// depend on them).  Call sites: irrt_star_3d.py:154-156 (SampleUnitBall), rrt_star_2d.py:77 (steer),
// datasets_3d/point_cloud_mask_utils_3d.py:167-169 (ellipsoid cloud).  tests/test_host_math.py pins both functions
// against the running libm on 4e6 arguments (0 differences); tests/test_gpu_trig.py does the same for the device build.
// Valid for |x| < 105414350 (every planner argument lies in [-pi, 2 pi]); larger arguments fall back to the
// correctly rounded evaluation (glibc's __branred path is not restated).
#pragma once
#include "exact_math.cuh"

namespace nirrt {

#define GT_FNMA(a, b, c) XFMA(-(a), (b), (c))     // -(a*b) + c, one rounding (vfnmadd)

NIRRT_HD void glibc_sincos_entry(int k, double &sn, double &ssn, double &cs, double &ccs) {
    static const double t[440] = {
#include "glibc_sincostab.inc"
    };
    sn = t[4 * k]; ssn = t[4 * k + 1]; cs = t[4 * k + 2]; ccs = t[4 * k + 3];
}

namespace gt {
constexpr double big = 0x1.8p45, toint = 0x1.8p52, hpinv = 0x1.45f306dc9c883p-1;
constexpr double mp1 = 0x1.921fb58p+0, mp2 = -0x1.dde973cp-27, pp3 = -0x1.cb3b398p-55, pp4 = -0x1.d747f23e32ed7p-83;
constexpr double hp0 = 0x1.921fb54442d18p+0, hp1 = 0x1.1a62633145c07p-54;
constexpr double s1 = -0x1.5555555555555p-3, s2 = 0x1.1111111110ecep-7, s3 = -0x1.a01a019db08b8p-13,
                 s4 = 0x1.71de27b9a7ed9p-19, s5 = -0x1.addffc2fcdf59p-26;
constexpr double sn3 = -0x1.5555555555515p-3, sn5 = 0x1.11110e829872fp-7;
constexpr double cs2 = 0.5, cs4 = -0x1.5555555555535p-5, cs6 = 0x1.6c16bedd9e239p-10;
}  // namespace gt

// index of the table entry nearest to |x| and the remainder: u = big + |x| keeps |x| rounded to 2^-7 in its low word
NIRRT_HD int glibc_split(double ax, double *rem) {
    const double u = XADD(gt::big, ax);
    *rem = XSUB(ax, XSUB(u, gt::big));
    return (int)(uint32_t)d2u(u);
}

// TAYLOR_SIN: x + ((POLY(xx) * x - 0.5 * dx) * xx + dx)
NIRRT_HD double glibc_taylor_sin(double x, double dx) {
    const double xx = XMUL(x, x);
    double p = XFMA(xx, gt::s5, gt::s4);
    p = XFMA(xx, p, gt::s3); p = XFMA(xx, p, gt::s2); p = XFMA(xx, p, gt::s1);
    const double t = XFMA(p, x, -XMUL(dx, 0.5));
    return XADD(x, XFMA(xx, t, dx));
}

NIRRT_HD double glibc_do_sin(double x, double dx) {
    const double xold = x;
    if (fabs(x) < 0.126) return glibc_taylor_sin(x, dx);
    if (x <= 0) dx = -dx;
    const int k = glibc_split(fabs(x), &x);
    const double xx = XMUL(x, x);
    const double s = XADD(x, XFMA(XMUL(x, xx), XFMA(xx, gt::sn5, gt::sn3), dx));
    const double c = XFMA(x, dx, XMUL(xx, XFMA(xx, XFMA(xx, gt::cs6, gt::cs4), gt::cs2)));
    double sn, ssn, cs, ccs;
    glibc_sincos_entry(k, sn, ssn, cs, ccs);
    const double cor = XFMA(s, cs, GT_FNMA(c, sn, XFMA(s, ccs, ssn)));
    return copysign(XADD(sn, cor), xold);
}

NIRRT_HD double glibc_do_cos(double x, double dx) {
    if (x < 0) dx = -dx;
    double r;
    const int k = glibc_split(fabs(x), &r);
    x = XADD(r, dx);
    const double xx = XMUL(x, x);
    const double s = XFMA(XMUL(x, xx), XFMA(xx, gt::sn5, gt::sn3), x);
    const double c = XMUL(xx, XFMA(xx, XFMA(xx, gt::cs6, gt::cs4), gt::cs2));
    double sn, ssn, cs, ccs;
    glibc_sincos_entry(k, sn, ssn, cs, ccs);
    const double cor = GT_FNMA(s, sn, GT_FNMA(c, cs, GT_FNMA(s, ssn, ccs)));
    return XADD(cs, cor);
}

// x = n * pi/2 + (a + da), |a| <= pi/4; returns n mod 4
NIRRT_HD int glibc_reduce_sincos(double x, double *a, double *da) {
    const double t = XFMA(x, gt::hpinv, gt::toint);
    const double xn = XSUB(t, gt::toint);
    const int n = (int)((uint32_t)d2u(t) & 3u);
    const double y = GT_FNMA(xn, gt::mp2, GT_FNMA(xn, gt::mp1, x));
    const double t2 = GT_FNMA(xn, gt::pp3, y);
    double db = GT_FNMA(xn, gt::pp3, XSUB(y, t2));
    const double b = GT_FNMA(xn, gt::pp4, t2);
    db = XADD(db, GT_FNMA(xn, gt::pp4, XSUB(t2, b)));
    *a = b; *da = db;
    return n;
}

NIRRT_HD double glibc_do_sincos(double a, double da, int n) {
    const double r = (n & 1) ? glibc_do_cos(a, da) : glibc_do_sin(a, da);
    return (n & 2) ? -r : r;
}

NIRRT_HD double glibc_sin(double x) {
    const uint32_t k = (uint32_t)(d2u(x) >> 32) & 0x7fffffffu;
    if (k < 0x3e500000u) return x;                                   // |x| < 2^-26
    if (k < 0x3feb6000u) return glibc_do_sin(x, 0.0);                // |x| < 0.855469
    if (k < 0x400368fdu) return copysign(glibc_do_cos(XSUB(gt::hp0, fabs(x)), gt::hp1), x);   // |x| < 2.426265
    if (k < 0x419921fbu) {                                           // |x| < 105414350
        double a, da;
        const int n = glibc_reduce_sincos(x, &a, &da);
        return glibc_do_sincos(a, da, n);
    }
    double s, c;
    cr_sincos(x, &s, &c);
    return s;
}

NIRRT_HD double glibc_cos(double x) {
    const uint32_t k = (uint32_t)(d2u(x) >> 32) & 0x7fffffffu;
    if (k < 0x3e400000u) return 1.0;                                 // |x| < 2^-27
    if (k < 0x3feb6000u) return glibc_do_cos(x, 0.0);
    if (k < 0x400368fdu) {
        const double y = XSUB(gt::hp0, fabs(x));
        const double a = XADD(y, gt::hp1);
        const double da = XADD(XSUB(y, a), gt::hp1);
        return glibc_do_sin(a, da);
    }
    if (k < 0x419921fbu) {
        double a, da;
        const int n = glibc_reduce_sincos(x, &a, &da);
        return glibc_do_sincos(a, da, n + 1);
    }
    double s, c;
    cr_sincos(x, &s, &c);
    return c;
}

}  // namespace nirrt
