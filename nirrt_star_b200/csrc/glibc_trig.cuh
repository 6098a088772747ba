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

// ---- math.atan2 (rrt_star_2d.py:74, rrt_base_2d.py:120): glibc 2.39 __ieee754_atan2, FMA variant
// (sysdeps/ieee754/dbl-64/e_atan2.c): u = min(|x|,|y|) / max(|x|,|y|) with its division error du (exact product by
// FMA), then either the odd polynomial (u < 1/16) or the degree-6 expansion around the nearest of 241 table points
// (glibc_atantab.inc), combined with 0 / pi/2 / pi in two-term arithmetic per quadrant.  Max error 0.56 ulp: not
// correctly rounded, restated with the fused operations exactly where that build has them.
NIRRT_HD const double *glibc_atan_row(int i) {
    static const double t[241 * 7] = {
#include "glibc_atantab.inc"
    };
    return t + 7 * i;
}

namespace gt {
constexpr double hpi = 0x1.921fb54442d18p+0, hpi1 = 0x1.1a62633145c07p-54, opi = 0x1.921fb54442d18p+1, opi1 = 0x1.1a62633145c07p-53;
constexpr double d3 = -0x1.5555555555555p-2, d5 = 0x1.99999999997fdp-3, d7 = -0x1.24924923f7603p-3, d9 = 0x1.c71c6e5129a3bp-4,
                 d11 = -0x1.7458022b13c25p-4, d13 = 0x1.375f08b31cbcep-4;
constexpr double two52 = 0x1p52, two500 = 0x1p500, twom500 = 0x1p-500;
}  // namespace gt

NIRRT_HD double glibc_atan_poly(double v) {
    double p = XFMA(v, gt::d13, gt::d11);
    p = XFMA(v, p, gt::d9); p = XFMA(v, p, gt::d7); p = XFMA(v, p, gt::d5);
    return XFMA(v, p, gt::d3);
}
// degree-4 tail c2 + v (c3 + v (c4 + v (c5 + v c6))) of a table row
NIRRT_HD double glibc_atan_tail(const double *c, double v) {
    double p = XFMA(v, c[6], c[5]);
    p = XFMA(v, p, c[4]); p = XFMA(v, p, c[3]);
    return XFMA(v, p, c[2]);
}

NIRRT_HD double glibc_atan2(double y, double x) {
    const uint64_t bx = d2u(x), by = d2u(y);
    const uint32_t ux = (uint32_t)(bx >> 32), uy = (uint32_t)(by >> 32);
    // NaN / infinity: not produced by the planners; defer to the library formula
    if ((ux & 0x7ff00000u) == 0x7ff00000u || (uy & 0x7ff00000u) == 0x7ff00000u) return atan2(y, x);
    if ((by << 1) == 0) {                                   // y = +-0
        const bool neg_x = (bx >> 63) != 0;
        if (by == 0) return neg_x ? gt::opi : 0.0;
        return neg_x ? -gt::opi : -0.0;
    }
    if ((bx << 1) == 0) return y > 0 ? gt::hpi : -gt::hpi;  // x = +-0
    double ax = x < 0 ? -x : x, ay = y < 0 ? -y : y;
    const int de = (int)(uy & 0x7ff00000u) - (int)(ux & 0x7ff00000u);
    if (de >= 59768832) return y > 0 ? gt::hpi : -gt::hpi;
    if (de <= -59768832) {
        if (x > 0) return copysign(XDIV(ay, ax), y);        // (the subnormal-quotient branch differs only in flags)
        return y > 0 ? gt::opi : -gt::opi;
    }
    if (ax < gt::twom500 || ay < gt::twom500) { ax = XMUL(ax, gt::two500); ay = XMUL(ay, gt::two500); }
    if (ax > gt::two500 || ay > gt::two500) { ax = XMUL(ax, gt::twom500); ay = XMUL(ay, gt::twom500); }
    double u, du;
    if (ay < ax) {
        u = XDIV(ay, ax);
        const double v = XMUL(ax, u), vv = XFMA(ax, u, -v);
        du = XDIV(XSUB(XSUB(ay, v), vv), ax);
    } else {
        u = XDIV(ax, ay);
        const double v = XMUL(ay, u), vv = XFMA(ay, u, -v);
        du = XDIV(XSUB(XSUB(ax, v), vv), ay);
    }
    const bool small = u < 0.0625;
    const double *c = nullptr;
    if (!small) c = glibc_atan_row((int)XSUB(XFMA(u, 256.0, gt::two52), gt::two52) - 16);
    double z;
    if (x > 0) {
        if (ay < ax) {                                      // (i) atan(ay/ax)
            if (small) {
                const double v = XMUL(u, u);
                z = XADD(u, XFMA(XMUL(u, v), glibc_atan_poly(v), du));
            } else {
                const double t3 = XSUB(u, c[0]);
                const double v = XADD(du, t3);
                const double dv = fabs(t3) > fabs(du) ? XADD(XSUB(t3, v), du) : XADD(XSUB(du, v), t3);
                double p = XFMA(v, c[6], c[5]);
                p = XFMA(v, p, c[4]); p = XFMA(v, p, c[3]);
                const double zz = XFMA(v, c[2], XFMA(dv, c[2], XMUL(XMUL(v, v), p)));
                z = XADD(zz, c[1]);
            }
        } else {                                            // (ii) pi/2 - atan(ax/ay)
            if (small) {
                const double v = XMUL(u, u);
                const double zz = XMUL(XMUL(u, v), glibc_atan_poly(v));
                const double t2 = XSUB(gt::hpi, u);
                const double cor = XSUB(XSUB(gt::hpi, t2), u);
                z = XADD(XSUB(XSUB(XADD(cor, gt::hpi1), du), zz), t2);
            } else {
                const double v = XADD(XSUB(u, c[0]), du);
                const double zz = GT_FNMA(v, glibc_atan_tail(c, v), gt::hpi1);
                z = XADD(XSUB(gt::hpi, c[1]), zz);
            }
        }
    } else {
        if (ax < ay) {                                      // (iii) pi/2 + atan(ax/ay)
            if (small) {
                const double v = XMUL(u, u);
                const double zz = XMUL(XMUL(u, v), glibc_atan_poly(v));
                const double t2 = XADD(u, gt::hpi);
                const double cor = XADD(XSUB(gt::hpi, t2), u);
                z = XADD(XADD(XADD(XADD(cor, gt::hpi1), du), zz), t2);
            } else {
                const double v = XADD(XSUB(u, c[0]), du);
                const double zz = XFMA(v, glibc_atan_tail(c, v), gt::hpi1);
                z = XADD(XADD(gt::hpi, c[1]), zz);
            }
        } else {                                            // (iv) pi - atan(ay/ax)
            if (small) {
                const double v = XMUL(u, u);
                const double zz = XMUL(XMUL(u, v), glibc_atan_poly(v));
                const double t2 = XSUB(gt::opi, u);
                const double cor = XSUB(XSUB(gt::opi, t2), u);
                z = XADD(XSUB(XSUB(XADD(cor, gt::opi1), du), zz), t2);
            } else {
                const double v = XADD(XSUB(u, c[0]), du);
                const double zz = GT_FNMA(v, glibc_atan_tail(c, v), gt::opi1);
                z = XADD(XSUB(gt::opi, c[1]), zz);
            }
        }
    }
    return copysign(z, y);
}

}  // namespace nirrt
