// exact_math.cuh -- IEEE-exact double arithmetic building blocks for the planner kernels.
//
// Every planner comparison that decides an index (argmin, dist<=r, cost_a<cost_b, SAT tests) must
// reproduce the reference's float64 results bit for bit, so nothing here may be contracted into an
// FMA or re-associated by the compiler.  On the device every operation is an explicit
// round-to-nearest intrinsic; on the host (CPU unit tests of these same functions, built with
// -ffp-contract=off) they are plain operators.
//
// Reference arithmetic being reproduced (upstream tedhuang96/nirrt_star, paths relative to it):
//   rownorm3  == np.linalg.norm(v, axis=-1) on (n,3)         rrt_base_3d.py:111, rrt_star_3d.py:82,94,103,136
//   vecnorm3  == np.linalg.norm(vec3) (1-D, BLAS ddot)       collision_check_utils_3d.py:24,60; rrt_star_3d.py:41
//   hypot3    == math.hypot(dx,dy,dz) (CPython vector_norm)  rrt_base_3d.py:65,121,136
//   cr_sincos == np.sin/np.cos, correctly rounded            irrt_star_3d.py:154-156
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define NIRRT_HD __host__ __device__ __forceinline__
#else
#define NIRRT_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define XADD(a, b) __dadd_rn((a), (b))
#define XSUB(a, b) __dsub_rn((a), (b))
#define XMUL(a, b) __dmul_rn((a), (b))
#define XDIV(a, b) __ddiv_rn((a), (b))
#define XSQRT(a) __dsqrt_rn((a))
#define XFMA(a, b, c) __fma_rn((a), (b), (c))
#define XINF __longlong_as_double(0x7ff0000000000000LL)
#else
#define XADD(a, b) ((a) + (b))
#define XSUB(a, b) ((a) - (b))
#define XMUL(a, b) ((a) * (b))
#define XDIV(a, b) ((a) / (b))
#define XSQRT(a) sqrt((a))
#define XFMA(a, b, c) fma((a), (b), (c))
#define XINF ((double)INFINITY)
#endif

namespace nirrt {

NIRRT_HD double sq3_rows(double dx, double dy, double dz) {  // (dx*dx + dy*dy) + dz*dz, three roundings
    return XADD(XADD(XMUL(dx, dx), XMUL(dy, dy)), XMUL(dz, dz));
}
NIRRT_HD double rownorm3(double dx, double dy, double dz) { return XSQRT(sq3_rows(dx, dy, dz)); }
NIRRT_HD double vecnorm3(double dx, double dy, double dz) {
    return XSQRT(XFMA(dz, dz, XFMA(dy, dy, XMUL(dx, dx))));
}

// bit helpers -------------------------------------------------------------------------------
NIRRT_HD uint64_t d2u(double x) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(x);
#else
    union { double d; uint64_t u; } c; c.d = x; return c.u;
#endif
}
NIRRT_HD double u2d(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    union { double d; uint64_t u; } c; c.u = u; return c.d;
#endif
}
// next representable double above / below a finite non-negative x
NIRRT_HD double next_up_pos(double x) { return u2d(d2u(x) + 1); }
NIRRT_HD double next_down_pos(double x) { return x == 0.0 ? 0.0 : u2d(d2u(x) - 1); }

// Largest double T with sqrt_rn(T) <= r  (r >= 0 finite).  sqrt_rn is monotone, so
// {s2 : sqrt_rn(s2) <= r} == {s2 <= T}: the scans compare squared distances against T and never
// take a square root, yet decide exactly as `np.linalg.norm(...) <= r` does.
NIRRT_HD double sqrt_le_threshold(double r) {
    double t = XMUL(r, r);
    for (int i = 0; i < 8 && XSQRT(t) > r; i++) t = next_down_pos(t);
    for (int i = 0; i < 8; i++) {
        double u = next_up_pos(t);
        if (XSQRT(u) <= r) t = u; else break;
    }
    return t;
}

// CPython 3.12 Modules/mathmodule.c vector_norm() for 3 (or 2, dz = 0 not allowed: use hypot2)
// components.  Scaling by a power of two is exact, so frexp/ldexp are done on the exponent bits.
// Branch-free (special cases are selected at the end) so that several independent evaluations in one
// basic block overlap their dependency chains (walk_to_root in planner3d.cu relies on this).
NIRRT_HD double hypot_n(const double *in, int n) {
    double v[3] = {0.0, 0.0, 0.0}, mx = 0.0;
    for (int i = 0; i < n; i++) { v[i] = fabs(in[i]); mx = v[i] > mx ? v[i] : mx; }
    // frexp: mx = m * 2^e with 0.5 <= m < 1  (normal numbers; planner coordinates never approach
    // the subnormal range -- zero / subnormal / non-finite inputs take the plain formula)
    const int ebits = (int)((d2u(mx) >> 52) & 0x7ff);
    const bool special = ebits == 0 || ebits == 0x7ff;
    const double plain = XSQRT(sq3_rows(v[0], v[1], v[2]));
    const int max_e = special ? 0 : ebits - 1022;
    const double scale = u2d((uint64_t)(1023 - max_e) << 52);      // ldexp(1.0, -max_e)
    double csum = 1.0, frac1 = 0.0, frac2 = 0.0;
    for (int i = 0; i < n; i++) {
        double x = XMUL(v[i], scale);
        double hi = XMUL(x, x), lo = XFMA(x, x, -hi);
        double s = XADD(csum, hi), slo = XADD(XSUB(csum, s), hi);
        csum = s;
        frac1 = XADD(frac1, lo);
        frac2 = XADD(frac2, slo);
    }
    double h = XSQRT(XADD(XSUB(csum, 1.0), XADD(frac1, frac2)));
    double nh = -h;
    double hi = XMUL(nh, h), lo = XFMA(nh, h, -hi);
    double s = XADD(csum, hi), slo = XADD(XSUB(csum, s), hi);
    csum = s;
    frac1 = XADD(frac1, lo);
    frac2 = XADD(frac2, slo);
    double x = XADD(XSUB(csum, 1.0), XADD(frac1, frac2));
    h = XADD(h, XDIV(x, XMUL(2.0, h)));
    // h / scale: scale is a power of two, so the quotient is the (exact, once-rounded) product with 2^max_e
    h = XMUL(h, u2d((uint64_t)(1023 + max_e) << 52));
    return special ? plain : h;
}
NIRRT_HD double hypot3(double dx, double dy, double dz) { double v[3] = {dx, dy, dz}; return hypot_n(v, 3); }
NIRRT_HD double hypot2(double dx, double dy) { double v[2] = {dx, dy}; return hypot_n(v, 2); }

// double-double helpers ---------------------------------------------------------------------
struct dd_t { double hi, lo; };
NIRRT_HD dd_t two_sum(double a, double b) {
    dd_t r; r.hi = XADD(a, b); double bb = XSUB(r.hi, a);
    r.lo = XADD(XSUB(a, XSUB(r.hi, bb)), XSUB(b, bb)); return r;
}
NIRRT_HD dd_t fast_two_sum(double a, double b) { dd_t r; r.hi = XADD(a, b); r.lo = XSUB(b, XSUB(r.hi, a)); return r; }
NIRRT_HD dd_t two_prod(double a, double b) { dd_t r; r.hi = XMUL(a, b); r.lo = XFMA(a, b, -r.hi); return r; }
NIRRT_HD dd_t dd_add(dd_t a, dd_t b) {
    dd_t s = two_sum(a.hi, b.hi), t = two_sum(a.lo, b.lo);
    s.lo = XADD(s.lo, t.hi); s = fast_two_sum(s.hi, s.lo);
    s.lo = XADD(s.lo, t.lo); s = fast_two_sum(s.hi, s.lo);
    return s;
}
NIRRT_HD dd_t dd_mul(dd_t a, dd_t b) {
    dd_t p = two_prod(a.hi, b.hi);
    double c = XADD(XMUL(a.hi, b.lo), XMUL(a.lo, b.hi));
    p.lo = XADD(p.lo, c);
    return fast_two_sum(p.hi, p.lo);
}

#if defined(__CUDA_ARCH__)
#define NIRRT_CONST_TABLE __device__ __constant__
#else
#define NIRRT_CONST_TABLE static const
#endif

// pi/2 in four doubles, and Taylor coefficients (-1)^k/(2k+1)!, (-1)^k/(2k)! as double-doubles
NIRRT_HD double dd_pio2(int i) {
    const double t[4] = {0x1.921fb54442d18p+0, 0x1.1a62633145c07p-54, -0x1.f1976b7ed8fbcp-110, 0x1.4cf98e804177dp-164};
    return t[i];
}
NIRRT_HD dd_t dd_sin_coef(int i) {
    const double t[15][2] = {
        {-0x1.5555555555555p-3, -0x1.5555555555555p-57}, {0x1.1111111111111p-7, 0x1.1111111111111p-63},
        {-0x1.a01a01a01a01ap-13, -0x1.a01a01a01a01ap-73}, {0x1.71de3a556c734p-19, -0x1.c154f8ddc6c00p-73},
        {-0x1.ae64567f544e4p-26, 0x1.c062e06d1f209p-80}, {0x1.6124613a86d09p-33, 0x1.f28e0cc748ebep-87},
        {-0x1.ae7f3e733b81fp-41, -0x1.1d8656b0ee8cbp-97}, {0x1.952c77030ad4ap-49, 0x1.ac981465ddc6cp-103},
        {-0x1.2f49b46814157p-57, -0x1.2650f61dbdcb4p-112}, {0x1.71b8ef6dcf572p-66, -0x1.d043ae40c4647p-120},
        {-0x1.761b41316381ap-75, 0x1.3423c7d91404fp-130}, {0x1.3f3ccdd165fa9p-84, -0x1.58ddadf344487p-139},
        {-0x1.d1ab1c2dccea3p-94, -0x1.054d0c78aea14p-149}, {0x1.259f98b4358adp-103, 0x1.eaf8c39dd9bc5p-157},
        {-0x1.434d2e783f5bcp-113, -0x1.0b87b91be9affp-167}};
    dd_t r; r.hi = t[i][0]; r.lo = t[i][1]; return r;
}
NIRRT_HD dd_t dd_cos_coef(int i) {
    const double t[15][2] = {
        {-0x1.0000000000000p-1, 0x0.0p+0}, {0x1.5555555555555p-5, 0x1.5555555555555p-59},
        {-0x1.6c16c16c16c17p-10, 0x1.f49f49f49f49fp-65}, {0x1.a01a01a01a01ap-16, 0x1.a01a01a01a01ap-76},
        {-0x1.27e4fb7789f5cp-22, -0x1.cbbc05b4fa99ap-76}, {0x1.1eed8eff8d898p-29, -0x1.2aec959e14c06p-83},
        {-0x1.93974a8c07c9dp-37, -0x1.05d6f8a2efd1fp-92}, {0x1.ae7f3e733b81fp-45, 0x1.1d8656b0ee8cbp-101},
        {-0x1.6827863b97d97p-53, -0x1.eec01221a8b0bp-107}, {0x1.e542ba4020225p-62, 0x1.ea72b4afe3c2fp-120},
        {-0x1.0ce396db7f853p-70, 0x1.aebcdbd20331cp-124}, {0x1.f2cf01972f578p-80, -0x1.9ada5fcc1ab14p-135},
        {-0x1.88e85fc6a4e5ap-89, 0x1.71c37ebd16540p-143}, {0x1.0a18a2635085dp-98, 0x1.b9e2e28e1aa54p-153},
        {-0x1.3932c5047d60ep-108, -0x1.832b7b530a627p-162}};
    dd_t r; r.hi = t[i][0]; r.lo = t[i][1]; return r;
}

// sin and cos as double-doubles (error < 2^-100) for 0 <= |x| < ~1e5: Taylor series in double-double
// after an exact four-term reduction by pi/2.  Same operation sequence as the oracle's cr_sincos.
NIRRT_HD void dd_sincos(double x, dd_t *s_out, dd_t *c_out) {
    double k = rint(XMUL(x, 0x1.45f306dc9c883p-1));
    dd_t r, t, m;
    t = two_prod(k, dd_pio2(0));
    r = two_sum(x, -t.hi);
    m.hi = -t.lo; m.lo = 0.0; r = dd_add(r, m);
    t = two_prod(k, dd_pio2(1)); m.hi = -t.hi; m.lo = -t.lo; r = dd_add(r, m);
    t = two_prod(k, dd_pio2(2)); m.hi = -t.hi; m.lo = -t.lo; r = dd_add(r, m);
    m.hi = -XMUL(k, dd_pio2(3)); m.lo = 0.0; r = dd_add(r, m);
    dd_t r2 = dd_mul(r, r);
    dd_t ps = dd_sin_coef(14), pc = dd_cos_coef(14);
    for (int i = 13; i >= 0; i--) {
        ps = dd_add(dd_mul(ps, r2), dd_sin_coef(i));
        pc = dd_add(dd_mul(pc, r2), dd_cos_coef(i));
    }
    dd_t sn = dd_add(r, dd_mul(dd_mul(r2, ps), r));
    m.hi = 1.0; m.lo = 0.0;
    dd_t cs = dd_add(m, dd_mul(r2, pc));
    dd_t ns, nc;
    ns.hi = -sn.hi; ns.lo = -sn.lo; nc.hi = -cs.hi; nc.lo = -cs.lo;
    int q = ((int)k) & 3;
    if (q == 0) { *s_out = sn; *c_out = cs; }
    else if (q == 1) { *s_out = cs; *c_out = ns; }
    else if (q == 2) { *s_out = ns; *c_out = nc; }
    else { *s_out = nc; *c_out = sn; }
}
// Correctly-rounded sin and cos (np.sin / np.cos / math.sin / math.cos up to glibc's own ~0.13 %
// last-bit misroundings).
NIRRT_HD void cr_sincos(double x, double *s_out, double *c_out) {
    dd_t s, c;
    dd_sincos(x, &s, &c);
    *s_out = s.hi; *c_out = c.hi;
}

// Correctly-rounded atan2 (math.atan2 up to glibc's last-bit misroundings): one Newton step in
// double-double on a <= 2-ulp seed, theta = t0 + (y cos t0 - x sin t0) / (x cos t0 + y sin t0).
NIRRT_HD double cr_atan2(double y, double x) {
    if (y == 0.0 && x == 0.0) return 0.0;          // atan2(+0, +0); the planners never produce -0 here
    const double t0 = atan2(y, x);
    dd_t s, c;
    dd_sincos(t0, &s, &c);
    dd_t a = two_prod(y, c.hi); a.lo = XADD(a.lo, XMUL(y, c.lo));
    dd_t b = two_prod(x, s.hi); b.lo = XADD(b.lo, XMUL(x, s.lo));
    b.hi = -b.hi; b.lo = -b.lo;
    const dd_t num = dd_add(a, b);
    const double den = XADD(XMUL(x, c.hi), XMUL(y, s.hi));
    return XADD(t0, XDIV(XADD(num.hi, num.lo), den));
}

// np.hypot(x, y) == glibc 2.35+ hypot (sysdeps/ieee754/dbl-64/e_hypot.c, the non-FMA kernel that
// x86-64 builds use; pinned against libm on 2e7 samples in tests/test_host_math.py).  Not correctly
// rounded, hence restated operation by operation.  Planner coordinates are far from the
// overflow / underflow scaling branches, which fall back to the plain formula here.
NIRRT_HD double np_hypot(double x, double y) {
    x = fabs(x); y = fabs(y);
    const double ax = x < y ? y : x, ay = x < y ? x : y;
    if (!(ax <= 0x1p+511) || (ay != 0.0 && ay < 0x1p-459)) return XSQRT(XADD(XMUL(ax, ax), XMUL(ay, ay)));
    if (ay <= XMUL(ax, 0x1p-54)) return XADD(ax, ay);
    double t1, t2;
    double h = XSQRT(XADD(XMUL(ax, ax), XMUL(ay, ay)));
    if (h <= XMUL(2.0, ay)) {
        const double delta = XSUB(h, ay);
        t1 = XMUL(ax, XSUB(XMUL(2.0, delta), ax));
        t2 = XMUL(XSUB(delta, XMUL(2.0, XSUB(ax, ay))), delta);
    } else {
        const double delta = XSUB(h, ax);
        t1 = XMUL(XMUL(2.0, delta), XSUB(ax, XMUL(2.0, ay)));
        t2 = XADD(XMUL(XSUB(XMUL(4.0, delta), ay), ay), XMUL(delta, delta));
    }
    h = XSUB(h, XDIV(XADD(t1, t2), XMUL(2.0, h)));
    return h;
}
// 2D twins of the norms above (probed against numpy 2.3 in this image):
//   vecnorm2 == np.linalg.norm(vec2) (1-D, BLAS ddot)          collision_check_utils.py:50,55,76; rrt_star_2d.py:41
//   rownorm2 == np.linalg.norm(v, axis=1) on (n,2)             rrt_base_2d.py:85
//   dot2     == np.dot(vec2, vec2)                             collision_check_utils.py:53
NIRRT_HD double vecnorm2(double dx, double dy) { return XSQRT(XFMA(dy, dy, XMUL(dx, dx))); }
NIRRT_HD double rownorm2(double dx, double dy) { return XSQRT(XADD(XMUL(dx, dx), XMUL(dy, dy))); }
NIRRT_HD double dot2(double a0, double a1, double b0, double b1) { return XFMA(a1, b1, XMUL(a0, b0)); }

// numpy pairwise summation of a contiguous f64 vector (np.add.reduce inner loop, PW_BLOCKSIZE 128)
#if defined(__CUDACC__)
__host__ __device__
#endif
inline double pairwise_sum(const double *a, long n) {
    if (n < 8) {
        double res = 0.;
        for (long i = 0; i < n; i++) res = XADD(res, a[i]);
        return res;
    }
    if (n <= 128) {
        double r[8];
        long i;
        for (i = 0; i < 8; i++) r[i] = a[i];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; j++) r[j] = XADD(r[j], a[i + j]);
        double res = XADD(XADD(XADD(r[0], r[1]), XADD(r[2], r[3])), XADD(XADD(r[4], r[5]), XADD(r[6], r[7])));
        for (; i < n; i++) res = XADD(res, a[i]);
        return res;
    }
    long n2 = n / 2;
    n2 -= n2 % 8;
    return XADD(pairwise_sum(a, n2), pairwise_sum(a + n2, n - n2));
}

}  // namespace nirrt
