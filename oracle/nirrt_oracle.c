/*
 * nirrt_oracle.c -- CPU restatement of the NIRRT* per-iteration hot path (3D and 2D planners).
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle: tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it; the product path
 * (nirrt_star_b200/) never does.  It is a plain scalar C restatement of the reference's
 * Python/numpy algorithm, written from the reference's behaviour (file:line cited per
 * function, paths relative to the upstream repo tedhuang96/nirrt_star).  Parity is PINNED:
 * tests/test_oracle_pin.py checks it against golden vectors produced by running the
 * reference's own classes in the build container (tests/golden/make_golden_planner.py).
 *
 * Build: gcc -O2 -ffp-contract=off -mfma -fPIC -shared (see oracle/Makefile).
 * -ffp-contract=off is REQUIRED: every a*b+c below must round twice unless written as fma().
 *
 * Floating-point facts this file relies on (all probed against numpy 2.3.5 / CPython 3.12.3
 * in the build container, see DESIGN.md "exact arithmetic"):
 *   - np.linalg.norm(v, axis=-1) on an (n,3) array  == sqrt((x*x + y*y) + z*z), no FMA;
 *     on an (n,2) array == sqrt(x*x + y*y).
 *   - np.linalg.norm(vec3) on a 1-D array (BLAS ddot) == sqrt(fma(z,z,fma(y,y,x*x))).
 *   - math.hypot(dx,dy[,dz]) == CPython's vector_norm (Modules/mathmodule.c), restated in
 *     orc_hypot().
 *   - np.random.* legacy global generator == MT19937, next_double = (a>>5, b>>6) / 2^53.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ MT19937 (numpy legacy) */
typedef struct {
    uint32_t key[624];
    int pos;
} orc_mt;

static void mt_regen(orc_mt *s) {
    const uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, MAT = 0x9908b0dfu;
    uint32_t y;
    int i;
    for (i = 0; i < 624 - 397; i++) {
        y = (s->key[i] & UPPER) | (s->key[i + 1] & LOWER);
        s->key[i] = s->key[i + 397] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MAT);
    }
    for (; i < 623; i++) {
        y = (s->key[i] & UPPER) | (s->key[i + 1] & LOWER);
        s->key[i] = s->key[i + (397 - 624)] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MAT);
    }
    y = (s->key[623] & UPPER) | (s->key[0] & LOWER);
    s->key[623] = s->key[396] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MAT);
    s->pos = 0;
}

static uint32_t mt_next(orc_mt *s) {
    uint32_t y;
    if (s->pos == 624) mt_regen(s);
    y = s->key[s->pos++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

static double mt_double(orc_mt *s) {
    int32_t a = mt_next(s) >> 5, b = mt_next(s) >> 6;
    return (a * 67108864.0 + b) / 9007199254740992.0;
}

/* np.random.uniform(lo, hi) == lo + (hi - lo) * next_double */
static double mt_uniform(orc_mt *s, double lo, double hi) {
    double range = hi - lo;
    return lo + range * mt_double(s);
}

/* np.random.randint(0, high) (legacy, masked rejection, 32-bit path).  high >= 1. */
static int64_t mt_randint(orc_mt *s, int64_t high) {
    uint64_t rng = (uint64_t)(high - 1);
    uint32_t mask, val;
    if (rng == 0) return 0;
    mask = (uint32_t)rng;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    do { val = mt_next(s) & mask; } while (val > rng);
    return (int64_t)val;
}

/* ------------------------------------------------------------------ exact scalar helpers */

/* CPython 3.12 Modules/mathmodule.c vector_norm(), as used by math.hypot
 * (rrt_base_3d.py:65,121,136; rrt_base_2d.py cost()).  n = 2 or 3. */
static double orc_hypot(int n, const double *in) {
    double vec[3], max = 0.0, x, h, scale, csum = 1.0, frac1 = 0.0, frac2 = 0.0;
    double hi, lo, s, slo;
    int max_e, i;
    for (i = 0; i < n; i++) {
        vec[i] = fabs(in[i]);
        if (vec[i] > max) max = vec[i];
    }
    if (isinf(max)) return max;
    if (max == 0.0 || n <= 1) return max;
    frexp(max, &max_e);
    if (max_e < -1023) {
        for (i = 0; i < n; i++) vec[i] /= DBL_MIN;
        return DBL_MIN * orc_hypot(n, vec);
    }
    scale = ldexp(1.0, -max_e);
    for (i = 0; i < n; i++) {
        x = vec[i] * scale;
        hi = x * x; lo = fma(x, x, -hi);            /* dl_mul */
        s = csum + hi; slo = (csum - s) + hi;        /* dl_fast_sum */
        csum = s;
        frac1 += lo;
        frac2 += slo;
    }
    h = sqrt(csum - 1.0 + (frac1 + frac2));
    hi = -h * h; lo = fma(-h, h, -hi);
    s = csum + hi; slo = (csum - s) + hi;
    csum = s;
    frac1 += lo;
    frac2 += slo;
    x = csum - 1.0 + (frac1 + frac2);
    h += x / (2.0 * h);
    return h / scale;
}

static double hypot3(double dx, double dy, double dz) { double v[3] = {dx, dy, dz}; return orc_hypot(3, v); }
static double hypot2(double dx, double dy) { double v[2] = {dx, dy}; return orc_hypot(2, v); }

/* np.linalg.norm(.., axis=-1) row norm on (n,3) */
static double rownorm3(double dx, double dy, double dz) { return sqrt((dx * dx + dy * dy) + dz * dz); }
/* np.linalg.norm of a 1-D 3-vector (BLAS ddot path) */
static double vecnorm3(double dx, double dy, double dz) { return sqrt(fma(dz, dz, fma(dy, dy, dx * dx))); }

/* numpy pairwise sum of a contiguous f64 vector (np.add.reduce inner loop) */
static double pairwise_sum(const double *a, long n) {
    if (n < 8) {
        double res = 0.;
        for (long i = 0; i < n; i++) res += a[i];
        return res;
    } else if (n <= 128) {
        double r[8], res;
        long i;
        for (i = 0; i < 8; i++) r[i] = a[i];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; j++) r[j] += a[i + j];
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    } else {
        long n2 = n / 2;
        n2 -= n2 % 8;
        return pairwise_sum(a, n2) + pairwise_sum(a + n2, n - n2);
    }
}


/* ------------------------------------------------------------------ correctly-rounded sin/cos
 * np.sin/np.cos in IRRTStar3D.SampleUnitBall (irrt_star_3d.py:147-157) resolve to the platform
 * libm, whose result is not correctly rounded in ~0.13% of calls (probed, glibc 2.39) and whose
 * ifunc variant depends on the CPU.  The oracle therefore uses a self-contained double-double
 * evaluation (error < 2^-100, i.e. correctly rounded for all practical purposes) built only from
 * IEEE + - * fma; the CUDA sampler executes the same operation sequence.  DESIGN.md quantifies
 * the deviation from glibc. */
typedef struct { double hi, lo; } dd_t;
static dd_t two_sum(double a, double b) { dd_t r; double bb; r.hi = a + b; bb = r.hi - a; r.lo = (a - (r.hi - bb)) + (b - bb); return r; }
static dd_t fast_two_sum(double a, double b) { dd_t r; r.hi = a + b; r.lo = b - (r.hi - a); return r; }
static dd_t two_prod(double a, double b) { dd_t r; r.hi = a * b; r.lo = fma(a, b, -r.hi); return r; }
static dd_t dd_add(dd_t a, dd_t b) {
    dd_t s = two_sum(a.hi, b.hi), t = two_sum(a.lo, b.lo);
    s.lo = s.lo + t.hi; s = fast_two_sum(s.hi, s.lo);
    s.lo = s.lo + t.lo; s = fast_two_sum(s.hi, s.lo);
    return s;
}
static dd_t dd_mul(dd_t a, dd_t b) {
    dd_t p = two_prod(a.hi, b.hi);
    double c1 = a.hi * b.lo, c2 = a.lo * b.hi, c = c1 + c2;
    p.lo = p.lo + c;
    return fast_two_sum(p.hi, p.lo);
}
static const double DD_PIO2[4] = {0x1.921fb54442d18p+0, 0x1.1a62633145c07p-54, -0x1.f1976b7ed8fbcp-110, 0x1.4cf98e804177dp-164};
static const double DD_SIN[15][2] = {
  {-0x1.5555555555555p-3, -0x1.5555555555555p-57}, {0x1.1111111111111p-7, 0x1.1111111111111p-63},
  {-0x1.a01a01a01a01ap-13, -0x1.a01a01a01a01ap-73}, {0x1.71de3a556c734p-19, -0x1.c154f8ddc6c00p-73},
  {-0x1.ae64567f544e4p-26, 0x1.c062e06d1f209p-80}, {0x1.6124613a86d09p-33, 0x1.f28e0cc748ebep-87},
  {-0x1.ae7f3e733b81fp-41, -0x1.1d8656b0ee8cbp-97}, {0x1.952c77030ad4ap-49, 0x1.ac981465ddc6cp-103},
  {-0x1.2f49b46814157p-57, -0x1.2650f61dbdcb4p-112}, {0x1.71b8ef6dcf572p-66, -0x1.d043ae40c4647p-120},
  {-0x1.761b41316381ap-75, 0x1.3423c7d91404fp-130}, {0x1.3f3ccdd165fa9p-84, -0x1.58ddadf344487p-139},
  {-0x1.d1ab1c2dccea3p-94, -0x1.054d0c78aea14p-149}, {0x1.259f98b4358adp-103, 0x1.eaf8c39dd9bc5p-157},
  {-0x1.434d2e783f5bcp-113, -0x1.0b87b91be9affp-167}};
static const double DD_COS[15][2] = {
  {-0x1.0000000000000p-1, 0x0.0p+0}, {0x1.5555555555555p-5, 0x1.5555555555555p-59},
  {-0x1.6c16c16c16c17p-10, 0x1.f49f49f49f49fp-65}, {0x1.a01a01a01a01ap-16, 0x1.a01a01a01a01ap-76},
  {-0x1.27e4fb7789f5cp-22, -0x1.cbbc05b4fa99ap-76}, {0x1.1eed8eff8d898p-29, -0x1.2aec959e14c06p-83},
  {-0x1.93974a8c07c9dp-37, -0x1.05d6f8a2efd1fp-92}, {0x1.ae7f3e733b81fp-45, 0x1.1d8656b0ee8cbp-101},
  {-0x1.6827863b97d97p-53, -0x1.eec01221a8b0bp-107}, {0x1.e542ba4020225p-62, 0x1.ea72b4afe3c2fp-120},
  {-0x1.0ce396db7f853p-70, 0x1.aebcdbd20331cp-124}, {0x1.f2cf01972f578p-80, -0x1.9ada5fcc1ab14p-135},
  {-0x1.88e85fc6a4e5ap-89, 0x1.71c37ebd16540p-143}, {0x1.0a18a2635085dp-98, 0x1.b9e2e28e1aa54p-153},
  {-0x1.3932c5047d60ep-108, -0x1.832b7b530a627p-162}};

/* sin and cos of x, 0 <= |x| < ~1e5 */
static void cr_sincos(double x, double *s_out, double *c_out) {
    double k = rint(x * 0x1.45f306dc9c883p-1);
    dd_t r, t, r2, ps, pc, sn, cs, m;
    int q, i;
    t = two_prod(k, DD_PIO2[0]);
    r = two_sum(x, -t.hi);
    m.hi = -t.lo; m.lo = 0; r = dd_add(r, m);
    t = two_prod(k, DD_PIO2[1]); m.hi = -t.hi; m.lo = -t.lo; r = dd_add(r, m);
    t = two_prod(k, DD_PIO2[2]); m.hi = -t.hi; m.lo = -t.lo; r = dd_add(r, m);
    m.hi = -(k * DD_PIO2[3]); m.lo = 0; r = dd_add(r, m);
    r2 = dd_mul(r, r);
    ps.hi = DD_SIN[14][0]; ps.lo = DD_SIN[14][1];
    pc.hi = DD_COS[14][0]; pc.lo = DD_COS[14][1];
    for (i = 13; i >= 0; i--) {
        m.hi = DD_SIN[i][0]; m.lo = DD_SIN[i][1]; ps = dd_add(dd_mul(ps, r2), m);
        m.hi = DD_COS[i][0]; m.lo = DD_COS[i][1]; pc = dd_add(dd_mul(pc, r2), m);
    }
    sn = dd_add(r, dd_mul(dd_mul(r2, ps), r));
    m.hi = 1.0; m.lo = 0; cs = dd_add(m, dd_mul(r2, pc));
    q = ((int)k) & 3;
    switch (q) {
        case 0: *s_out = sn.hi; *c_out = cs.hi; break;
        case 1: *s_out = cs.hi; *c_out = -sn.hi; break;
        case 2: *s_out = -sn.hi; *c_out = -cs.hi; break;
        default: *s_out = -cs.hi; *c_out = sn.hi; break;
    }
}

/* ------------------------------------------------------------------ 3D collision predicates */

typedef struct {
    int n_balls, n_boxes;
    double *balls;      /* [n_balls][4]  x y z r                (Utils.obs_ball, rrt_utils_3d.py:9-12) */
    double *ball_r2pow; /* [n_balls]     (r+clearance)**2 as the numpy *scalar* power gives it */
    double *boxes;      /* [n_boxes][6]  x y z w h d */
    double clearance;
    double range[6];    /* x0 x1 y0 y1 z0 z1 */
} orc_env3;

/* collision_check_utils_3d.py:96-110 */
static int point_in_single_ball(const double *p, const double *c, double radius, double clearance) {
    return vecnorm3(p[0] - c[0], p[1] - c[1], p[2] - c[2]) <= radius + clearance;
}

/* collision_check_utils_3d.py:113-131 */
static int point_in_single_box(const double *p, const double *b, double cl) {
    return b[0] - cl <= p[0] && p[0] <= b[0] + b[3] + cl &&
           b[1] - cl <= p[1] && p[1] <= b[1] + b[4] + cl &&
           b[2] - cl <= p[2] && p[2] <= b[2] + b[5] + cl;
}

/* collision_check_utils_3d.py:3-38 */
static int line_single_ball(const double *p0, const double *p1, const double *ball, double r2, double clearance) {
    const double *c = ball;
    double l0 = p1[0] - p0[0], l1 = p1[1] - p0[1], l2 = p1[2] - p0[2];
    double d0, d1, d2, t;
    if (vecnorm3(l0, l1, l2) == 0) return point_in_single_ball(p0, c, ball[3], clearance);
    d0 = c[0] - p0[0]; d1 = c[1] - p0[1]; d2 = c[2] - p0[2];
    t = (1 / (l0 * l0 + l1 * l1 + l2 * l2)) * (l0 * d0 + l1 * d1 + l2 * d2);
    if (t <= 0) {
        if ((d0 * d0 + d1 * d1 + d2 * d2) <= r2) return 1;
    } else if (t >= 1) {
        double e0 = c[0] - p1[0], e1 = c[1] - p1[1], e2 = c[2] - p1[2];
        if ((e0 * e0 + e1 * e1 + e2 * e2) <= r2) return 1;
    } else if (0 < t && t < 1) {
        double x0 = p0[0] + t * l0, x1 = p0[1] + t * l1, x2 = p0[2] + t * l2;
        double k0 = c[0] - x0, k1 = c[1] - x1, k2 = c[2] - x2;
        if ((k0 * k0 + k1 * k1 + k2 * k2) <= r2) return 1;
    }
    return 0;
}

/* collision_check_utils_3d.py:41-84 */
static int line_single_box(const double *p0, const double *p1, const double *b, double cl) {
    double mid[3], dir[3], I[3], P[3], E[3], T[3], dist, hl, r;
    int i;
    for (i = 0; i < 3; i++) { mid[i] = (p0[i] + p1[i]) / 2; dir[i] = p1[i] - p0[i]; }
    dist = vecnorm3(dir[0], dir[1], dir[2]);
    if (dist == 0) return point_in_single_box(p0, b, cl);
    for (i = 0; i < 3; i++) I[i] = dir[i] / dist;
    hl = dist / 2;
    for (i = 0; i < 3; i++) { P[i] = b[i] + b[3 + i] / 2; E[i] = b[3 + i] / 2 + cl; T[i] = P[i] - mid[i]; }
    if (fabs(T[0]) > (E[0] + hl * fabs(I[0]))) return 0;
    if (fabs(T[1]) > (E[1] + hl * fabs(I[1]))) return 0;
    if (fabs(T[2]) > (E[2] + hl * fabs(I[2]))) return 0;
    r = E[1] * fabs(I[2]) + E[2] * fabs(I[1]);
    if (fabs(T[1] * I[2] - T[2] * I[1]) > r) return 0;
    r = E[0] * fabs(I[2]) + E[2] * fabs(I[0]);
    if (fabs(T[2] * I[0] - T[0] * I[2]) > r) return 0;
    r = E[0] * fabs(I[1]) + E[1] * fabs(I[0]);
    if (fabs(T[0] * I[1] - T[1] * I[0]) > r) return 0;
    return 1;
}

/* Utils.is_collision -> check_collision_line_balls_boxes
 * (rrt_utils_3d.py:22-36, collision_check_utils_3d.py:151-216) */
static int is_collision3(const orc_env3 *e, const double *p0, const double *p1) {
    double lo[3], hi[3], cl = e->clearance;
    int i, k;
    for (i = 0; i < 3; i++) { lo[i] = p0[i] < p1[i] ? p0[i] : p1[i]; hi[i] = p0[i] > p1[i] ? p0[i] : p1[i]; }
    for (k = 0; k < e->n_balls; k++) {
        const double *b = e->balls + 4 * k;
        int hit = 1;
        for (i = 0; i < 3; i++) {
            double a1 = b[i] - b[3] - cl, a2 = b[i] + b[3] + cl;
            hit = hit && (lo[i] <= a2) && (hi[i] >= a1);
        }
        if (hit && line_single_ball(p0, p1, b, e->ball_r2pow[k], cl)) return 1;
    }
    for (k = 0; k < e->n_boxes; k++) {
        const double *b = e->boxes + 6 * k;
        int hit = 1;
        for (i = 0; i < 3; i++) {
            double a1 = b[i] - cl, a2 = b[i] + b[3 + i] + cl;
            hit = hit && (lo[i] <= a2) && (hi[i] >= a1);
        }
        if (hit && line_single_box(p0, p1, b, cl)) return 1;
    }
    return 0;
}

/* points_in_balls (collision_check_utils_3d.py:259-295): strict <, array power == x*x */
static int point_in_balls3(const orc_env3 *e, const double *p, double cl) {
    for (int k = 0; k < e->n_balls; k++) {
        const double *b = e->balls + 4 * k;
        double rc = b[3] + cl;
        double dx = p[0] - b[0], dy = p[1] - b[1], dz = p[2] - b[2];
        if (dx * dx + dy * dy + dz * dz < rc * rc) return 1;
    }
    return 0;
}

/* points_in_boxes (collision_check_utils_3d.py:219-257): inclusive */
static int point_in_boxes3(const orc_env3 *e, const double *p, double cl) {
    for (int k = 0; k < e->n_boxes; k++) {
        const double *b = e->boxes + 6 * k;
        if (b[0] - cl <= p[0] && p[0] <= b[0] + b[3] + cl &&
            b[1] - cl <= p[1] && p[1] <= b[1] + b[4] + cl &&
            b[2] - cl <= p[2] && p[2] <= b[2] + b[5] + cl) return 1;
    }
    return 0;
}

/* Utils.is_inside_obs (rrt_utils_3d.py:39-51) */
static int is_inside_obs3(const orc_env3 *e, const double *p) {
    return point_in_balls3(e, p, e->clearance) || point_in_boxes3(e, p, e->clearance);
}

/* Utils.is_valid -> points_validity_3d (collision_check_utils_3d.py:354-398); the range test is
 * points_in_boxes with clearance = -clearance on the box (x0,y0,z0,x1-x0,...) (:329-352) */
static int is_valid3(const orc_env3 *e, const double *p) {
    double mc = -e->clearance;
    int in_range = 1;
    for (int i = 0; i < 3; i++) {
        double mn = e->range[2 * i], w = e->range[2 * i + 1] - e->range[2 * i];
        in_range = in_range && (mn - mc <= p[i]) && (p[i] <= mn + w + mc);
    }
    return in_range && !point_in_balls3(e, p, e->clearance) && !point_in_boxes3(e, p, e->clearance);
}

/* ------------------------------------------------------------------ 3D planner state */

typedef struct {
    int cap;            /* 1 + iter_max */
    int n;              /* num_vertices */
    double *v;          /* [cap][3]     (rrt_base_3d.py:25) */
    int64_t *parent;    /* [cap]        (rrt_base_3d.py:26) */
    double start[3], goal[3];
    double step_len, search_radius;
    const double *rtab; /* rtab[n] = (math.log(n)/n)**(1/3.) computed by the caller in Python */
    orc_env3 env;
    orc_mt rng;
    /* IRRT*-family state (irrt_star_3d.py:29-36) */
    int64_t *sol; int n_sol, sol_cap;
    double c_min, center[3], C[9];
    /* NIRRT* guidance cloud (nirrt_star_png_3d.py:129-130) */
    const double *pc; int n_pc; double pc_sample_rate;
    /* per-iteration trace of the last iteration (for parity tests) */
    int64_t tr_nearest, tr_new; int tr_inserted; long tr_near_n;
    int64_t *tr_near; long tr_near_cap;
    double tr_rand[3];
} orc_plan3;

ORC_API orc_plan3 *orc3_create(int cap, const double *start, const double *goal, double step_len,
                               double search_radius, double clearance, const double *range6,
                               int n_balls, const double *balls, const double *ball_r2pow,
                               int n_boxes, const double *boxes, const double *rtab,
                               const uint32_t *mt_key, int mt_pos) {
    orc_plan3 *p = (orc_plan3 *)calloc(1, sizeof(orc_plan3));
    p->cap = cap;
    p->v = (double *)calloc((size_t)cap * 3, sizeof(double));
    p->parent = (int64_t *)calloc((size_t)cap, sizeof(int64_t));
    memcpy(p->start, start, 24); memcpy(p->goal, goal, 24);
    memcpy(p->v, start, 24);
    p->n = 1;
    p->step_len = step_len; p->search_radius = search_radius;
    p->rtab = rtab;
    p->env.n_balls = n_balls; p->env.n_boxes = n_boxes; p->env.clearance = clearance;
    memcpy(p->env.range, range6, 48);
    p->env.balls = (double *)malloc(sizeof(double) * 4 * (n_balls + 1));
    p->env.ball_r2pow = (double *)malloc(sizeof(double) * (n_balls + 1));
    p->env.boxes = (double *)malloc(sizeof(double) * 6 * (n_boxes + 1));
    if (n_balls) { memcpy(p->env.balls, balls, sizeof(double) * 4 * n_balls); memcpy(p->env.ball_r2pow, ball_r2pow, sizeof(double) * n_balls); }
    if (n_boxes) memcpy(p->env.boxes, boxes, sizeof(double) * 6 * n_boxes);
    memcpy(p->rng.key, mt_key, sizeof(uint32_t) * 624); p->rng.pos = mt_pos;
    p->sol_cap = cap + 16; p->sol = (int64_t *)malloc(sizeof(int64_t) * p->sol_cap);
    p->tr_near_cap = 1024; p->tr_near = (int64_t *)malloc(sizeof(int64_t) * p->tr_near_cap);
    return p;
}

ORC_API void orc3_destroy(orc_plan3 *p) {
    if (!p) return;
    free(p->v); free(p->parent); free(p->env.balls); free(p->env.ball_r2pow); free(p->env.boxes);
    free(p->sol); free(p->tr_near); free(p);
}

/* IRRTStar3D.init (irrt_star_3d.py:32-36); C is computed by the caller with numpy
 * (RotationToWorldFrame is an SVD, :159-173) */
ORC_API void orc3_set_informed(orc_plan3 *p, const double *C9) {
    p->c_min = hypot3(p->goal[0] - p->start[0], p->goal[1] - p->start[1], p->goal[2] - p->start[2]);
    for (int i = 0; i < 3; i++) p->center[i] = (p->start[i] + p->goal[i]) / 2.;
    memcpy(p->C, C9, 72);
}

ORC_API void orc3_set_cloud(orc_plan3 *p, const double *pc, int n_pc, double rate) {
    p->pc = pc; p->n_pc = n_pc; p->pc_sample_rate = rate;
}

ORC_API void orc3_load_tree(orc_plan3 *p, int n, const double *v, const int64_t *parent) {
    memcpy(p->v, v, sizeof(double) * 3 * n);
    memcpy(p->parent, parent, sizeof(int64_t) * n);
    p->n = n;
}
ORC_API int orc3_num_vertices(const orc_plan3 *p) { return p->n; }
ORC_API void orc3_get_tree(const orc_plan3 *p, double *v, int64_t *parent) {
    memcpy(v, p->v, sizeof(double) * 3 * p->n);
    memcpy(parent, p->parent, sizeof(int64_t) * p->n);
}
ORC_API int orc3_get_solutions(const orc_plan3 *p, int64_t *out) {
    if (out) memcpy(out, p->sol, sizeof(int64_t) * p->n_sol);
    return p->n_sol;
}
ORC_API void orc3_get_rng(const orc_plan3 *p, uint32_t *key, int *pos) {
    memcpy(key, p->rng.key, sizeof(uint32_t) * 624); *pos = p->rng.pos;
}
ORC_API void orc3_get_trace(const orc_plan3 *p, int64_t *nearest, int64_t *newidx, int *inserted,
                            long *near_n, int64_t *near_out, long near_cap, double *rand3) {
    *nearest = p->tr_nearest; *newidx = p->tr_new; *inserted = p->tr_inserted; *near_n = p->tr_near_n;
    for (long i = 0; i < p->tr_near_n && i < near_cap; i++) near_out[i] = p->tr_near[i];
    memcpy(rand3, p->tr_rand, 24);
}

/* RRTBase3D.cost (rrt_base_3d.py:60-67): leaf -> root accumulation of math.hypot edges */
static double cost3(const orc_plan3 *p, int64_t idx) {
    double c = 0.;
    while (idx != 0) {
        int64_t par = p->parent[idx];
        const double *a = p->v + 3 * idx, *b = p->v + 3 * par;
        c += hypot3(a[0] - b[0], a[1] - b[1], a[2] - b[2]);
        idx = par;
    }
    return c;
}

/* RRTBase3D.SampleFree (rrt_base_3d.py:49-58) */
static void sample_free3(orc_plan3 *p, double *out) {
    const double *r = p->env.range; double cl = p->env.clearance;
    do {
        out[0] = mt_uniform(&p->rng, r[0] + cl, r[1] - cl);
        out[1] = mt_uniform(&p->rng, r[2] + cl, r[3] - cl);
        out[2] = mt_uniform(&p->rng, r[4] + cl, r[5] - cl);
    } while (is_inside_obs3(&p->env, out));
}

/* IRRTStar3D.SampleInformedSubset + SampleUnitBall (irrt_star_3d.py:117-157).
 * c_max**2 is a Python float power (libm pow(c,2.0), which deviates from the correctly rounded
 * c*c in ~0.09% of inputs, probed); the oracle uses c*c.  C@L@xball is evaluated as (C@L)@xball;
 * L is diagonal so (C@L)[i][j] = C[i][j]*r[j] (single rounding), and the 3-term row dot products
 * follow the order this container's BLAS dgemv produces: fma(M2,x2, fma(M0,x0, M1*x1))
 * (probed, 60000/60000 rows).  sin/cos: the platform libm, exactly what numpy calls (the CUDA path restates glibc
 * 2.39's kernels operation by operation, nirrt_star_b200/csrc/glibc_trig.cuh; cr_sincos above is kept as the
 * correctly rounded yardstick the tests compare both against). */
static void sample_informed3(orc_plan3 *p, double c_max, double *out) {
    double c2 = c_max * c_max - p->c_min * p->c_min;
    double eps = (c2 < 0) ? 1e-6 : 0;
    double r[3], M[9];
    r[0] = c_max / 2;
    r[1] = r[2] = sqrt(c2 + eps) / 2;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) M[3 * i + j] = p->C[3 * i + j] * r[j];
    for (;;) {
        double rr = mt_uniform(&p->rng, 0.0, 1.0);
        double th = mt_uniform(&p->rng, 0, M_PI);
        double ph = mt_uniform(&p->rng, 0, 2 * M_PI);
        double xb[3], st, ct, sp, cp;
        /* np.sin / np.cos == the C library's sin / cos (probed: np.sin(array) == [math.sin(x)] on 2e6 samples) */
        st = sin(th); ct = cos(th);
        sp = sin(ph); cp = cos(ph);
        xb[0] = rr * st * cp;
        xb[1] = rr * st * sp;
        xb[2] = rr * ct;
        for (int i = 0; i < 3; i++)
            out[i] = fma(M[3 * i + 2], xb[2], fma(M[3 * i], xb[0], M[3 * i + 1] * xb[1])) + p->center[i];
        if (is_valid3(&p->env, out)) break;
    }
}

/* IRRTStar3D.find_best_path_solution (irrt_star_3d.py:80-93) */
static double best_solution3(const orc_plan3 *p, int64_t *x_best) {
    double best = INFINITY; int64_t bi = -1;
    for (int i = 0; i < p->n_sol; i++) {
        int64_t idx = p->sol[i];
        const double *a = p->v + 3 * idx;
        double c = cost3(p, idx) + hypot3(p->goal[0] - a[0], p->goal[1] - a[1], p->goal[2] - a[2]);
        if (bi < 0 || c < best) { best = c; bi = idx; }
    }
    if (x_best) *x_best = bi;
    return best;
}

/* RRTStar3D.search_goal_parent (rrt_star_3d.py:101-117): returns -1 for None */
static int64_t search_goal_parent3(const orc_plan3 *p) {
    double best = 0; int64_t bi = -1; int first = 1;
    for (int64_t i = 0; i < p->n; i++) {
        const double *a = p->v + 3 * i;
        double d = rownorm3(p->goal[0] - a[0], p->goal[1] - a[1], p->goal[2] - a[2]);
        double c;
        if (!(d <= p->step_len)) continue;
        if (!is_collision3(&p->env, a, p->goal)) c = cost3(p, i) + d; else c = INFINITY;
        if (first || c < best) { best = c; bi = i; first = 0; }
    }
    return bi;
}

/* extract_path + get_path_len (rrt_base_3d.py:69-91): sum of row norms of consecutive
 * differences, start -> ... -> goal_parent -> goal, numpy reduce order. */
static double path_len3(const orc_plan3 *p, int64_t goal_parent) {
    long m = 0, cap = 64;
    int64_t *chain = (int64_t *)malloc(sizeof(int64_t) * cap);
    double *seg, res;
    int64_t i = goal_parent;
    while (i != 0) {
        if (m + 2 >= cap) { cap *= 2; chain = (int64_t *)realloc(chain, sizeof(int64_t) * cap); }
        chain[m++] = i; i = p->parent[i];
    }
    chain[m++] = 0;
    /* points in order: chain[m-1]=0 ... chain[0]=goal_parent, then goal; m segments */
    seg = (double *)malloc(sizeof(double) * (m + 1));
    for (long k = 0; k < m; k++) {
        const double *a = p->v + 3 * chain[m - 1 - k];
        const double *b = (k + 1 < m) ? p->v + 3 * chain[m - 2 - k] : p->goal;
        seg[k] = rownorm3(b[0] - a[0], b[1] - a[1], b[2] - a[2]);
    }
    res = (m == 1) ? seg[0] : seg[0] + pairwise_sum(seg + 1, m - 1);
    free(chain); free(seg);
    return res;
}

/* One loop body: rrt_star_3d.py:37-55 (== irrt_star_3d.py:50-71 without the goal append).
 * rand = node_rand.  Returns node_new_index or -1 when the steer edge collides. */
static int64_t expand3(orc_plan3 *p, const double *rnd) {
    int64_t nearest = 0, new_idx;
    double best = 0, dist, dir[3], xnew[3], curr_cost, r;
    const double *xn;
    long n_near = 0;
    memcpy(p->tr_rand, rnd, 24);
    /* nearest_neighbor (rrt_base_3d.py:100-113) */
    for (int64_t i = 0; i < p->n; i++) {
        const double *a = p->v + 3 * i;
        double d = rownorm3(rnd[0] - a[0], rnd[1] - a[1], rnd[2] - a[2]);
        if (i == 0 || d < best) { best = d; nearest = i; }
    }
    xn = p->v + 3 * nearest;
    /* new_state (rrt_star_3d.py:67-78) + get_distance_and_direction (rrt_base_3d.py:116-130) */
    dist = hypot3(rnd[0] - xn[0], rnd[1] - xn[1], rnd[2] - xn[2]);
    if (dist == 0) { dir[0] = dir[1] = dir[2] = 0; }
    else for (int i = 0; i < 3; i++) dir[i] = (rnd[i] - xn[i]) / dist;
    if (dist < p->step_len) { /* min(step_len, dist) */ } else dist = p->step_len;
    for (int i = 0; i < 3; i++) xnew[i] = xn[i] + dist * dir[i];
    p->tr_nearest = nearest; p->tr_new = -1; p->tr_inserted = 0; p->tr_near_n = 0;
    if (is_collision3(&p->env, xn, xnew)) return -1;
    if (vecnorm3(xnew[0] - xn[0], xnew[1] - xn[1], xnew[2] - xn[2]) < 1e-8) {
        memcpy(xnew, xn, 24);
        new_idx = nearest;
        curr_cost = cost3(p, nearest);
    } else {
        new_idx = p->n;
        memcpy(p->v + 3 * new_idx, xnew, 24);
        p->parent[new_idx] = nearest;
        p->n += 1;
        p->tr_inserted = 1;
        curr_cost = cost3(p, nearest) + hypot3(xnew[0] - xn[0], xnew[1] - xn[1], xnew[2] - xn[2]);
    }
    p->tr_new = new_idx;
    /* find_near_neighbors (rrt_star_3d.py:125-145) */
    r = p->search_radius * p->rtab[p->n];
    if (p->step_len < r) r = p->step_len;
    for (int64_t i = 0; i < p->n; i++) {
        const double *a = p->v + 3 * i;
        double d = rownorm3(xnew[0] - a[0], xnew[1] - a[1], xnew[2] - a[2]);
        if (d <= r && !is_collision3(&p->env, xnew, a) && i != new_idx) {
            if (n_near >= p->tr_near_cap) {
                p->tr_near_cap *= 2;
                p->tr_near = (int64_t *)realloc(p->tr_near, sizeof(int64_t) * p->tr_near_cap);
            }
            p->tr_near[n_near++] = i;
        }
    }
    p->tr_near_n = n_near;
    if (n_near > 0) {
        /* choose_parent (rrt_star_3d.py:80-90) */
        double bc = 0, new_cost; long bk = -1;
        for (long k = 0; k < n_near; k++) {
            const double *a = p->v + 3 * p->tr_near[k];
            double c = cost3(p, p->tr_near[k]) + rownorm3(xnew[0] - a[0], xnew[1] - a[1], xnew[2] - a[2]);
            if (bk < 0 || c < bc) { bc = c; bk = k; }
        }
        if (bc < curr_cost) p->parent[new_idx] = p->tr_near[bk];
        /* rewire (rrt_star_3d.py:92-99): sequential, later neighbours see earlier re-parentings */
        new_cost = cost3(p, new_idx);
        for (long k = 0; k < n_near; k++) {
            int64_t j = p->tr_near[k];
            const double *a = p->v + 3 * j;
            double d = rownorm3(a[0] - xnew[0], a[1] - xnew[1], a[2] - xnew[2]);
            if (cost3(p, j) > new_cost + d) p->parent[j] = new_idx;
        }
    }
    return new_idx;
}

/* generate_random_node for the three families:
 * variant 0 RRT*  (rrt_star_3d.py:119-123), 1 IRRT* (irrt_star_3d.py:95-115),
 * 2 NIRRT* with a fixed guidance cloud (nirrt_star_png_3d.py:99-130; the cloud update itself
 * is driven from Python), 3 NRRT* = RRT* driver + guidance cloud (nrrt_star_png_3d.py:52-59) */
static void gen_random3(orc_plan3 *p, int variant, double c_best, double *out) {
    if (variant == 2 || variant == 3) {
        if (mt_double(&p->rng) < p->pc_sample_rate) {
            int64_t k = mt_randint(&p->rng, p->n_pc);
            memcpy(out, p->pc + 3 * k, 24);
            return;
        }
    }
    if ((variant == 1 || variant == 2) && c_best < INFINITY) sample_informed3(p, c_best, out);
    else sample_free3(p, out);
}

/* RRTBase3D.InGoalRegion (rrt_base_3d.py:93-95) */
static int in_goal_region3(const orc_plan3 *p, const double *x) {
    return hypot3(p->goal[0] - x[0], p->goal[1] - x[1], p->goal[2] - x[2]) < p->step_len &&
           !is_collision3(&p->env, x, p->goal);
}

/*
 * Run k loop bodies.
 *   variant 0 (RRT* family): mode 0 = planning() body (rrt_star_3d.py:36-55), no goal work;
 *       mode 1 = planning_random body (rrt_star_3d.py:205-236): pathlen[i] = path length after
 *       iteration i (inf when no goal parent).
 *   variant 1/2 (IRRT-star, NIRRT-star): pathlen[i] = c_best refreshed at the TOP of iteration i
 *       (irrt_star_3d.py:253-256); goal append after the expansion (:281-282).
 * Optional per-iteration traces (arrays of length k, may be NULL): nearest index, new index
 * (-1 = steer edge collided), near count, and the near lists concatenated into near_buf.
 * stop_on_first != 0: stop after the first iteration whose recorded value is finite (phase 1 of
 * planning_random).  Returns the number of iterations executed.
 */
ORC_API long orc3_run(orc_plan3 *p, int variant, int mode, long k, int stop_on_first, double *pathlen,
                      int64_t *t_nearest, int64_t *t_new, int64_t *t_near_cnt,
                      int64_t *near_buf, long near_buf_cap, long *near_buf_used) {
    long used = 0, it;
    for (it = 0; it < k; it++) {
        double rnd[3], c_best = INFINITY;
        int64_t ni;
        if (variant == 1 || variant == 2) {
            if (p->n_sol > 0) c_best = best_solution3(p, NULL);
            if (pathlen) pathlen[it] = c_best;
            if (stop_on_first && c_best < INFINITY) { it++; break; }
        }
        gen_random3(p, variant, c_best, rnd);
        ni = expand3(p, rnd);
        if ((variant == 1 || variant == 2) && ni >= 0 && in_goal_region3(p, p->v + 3 * ni)) {
            if (p->n_sol < p->sol_cap) p->sol[p->n_sol++] = ni;
        }
        if (t_nearest) t_nearest[it] = p->tr_nearest;
        if (t_new) t_new[it] = ni;
        if (t_near_cnt) t_near_cnt[it] = p->tr_near_n;
        if (near_buf) {
            for (long q = 0; q < p->tr_near_n && used < near_buf_cap; q++) near_buf[used++] = p->tr_near[q];
        }
        if ((variant == 0 || variant == 3) && mode == 1) {
            int64_t gp = search_goal_parent3(p);
            double len = gp < 0 ? INFINITY : path_len3(p, gp);
            if (pathlen) pathlen[it] = len;
            if (stop_on_first && len < INFINITY) { it++; break; }
        }
    }
    if (near_buf_used) *near_buf_used = used;
    return it;
}

/* c_best refresh used by the drivers after the loops (irrt_star_3d.py:283-286,327-328) */
ORC_API double orc3_best_cost(const orc_plan3 *p, int64_t *x_best) {
    if (p->n_sol == 0) { if (x_best) *x_best = -1; return INFINITY; }
    return best_solution3(p, x_best);
}
ORC_API int64_t orc3_search_goal_parent(const orc_plan3 *p) { return search_goal_parent3(p); }
ORC_API double orc3_path_len(const orc_plan3 *p, int64_t goal_parent) { return path_len3(p, goal_parent); }
ORC_API double orc3_cost(const orc_plan3 *p, int64_t idx) { return cost3(p, idx); }

/* ------------------------------------------------------------------ stand-alone predicates
 * (function-level parity tests; also used to pin this file against the reference) */
ORC_API double orc_hypot3(double dx, double dy, double dz) { return hypot3(dx, dy, dz); }
ORC_API double orc_hypot2(double dx, double dy) { return hypot2(dx, dy); }
ORC_API double orc_rownorm3(double dx, double dy, double dz) { return rownorm3(dx, dy, dz); }
ORC_API double orc_vecnorm3(double dx, double dy, double dz) { return vecnorm3(dx, dy, dz); }
ORC_API double orc_pairwise_sum(const double *a, long n) { return pairwise_sum(a, n); }

ORC_API void orc3_collide_edges(const orc_plan3 *p, long m, const double *edges /* [m][2][3] */, uint8_t *out) {
    for (long i = 0; i < m; i++) out[i] = (uint8_t)is_collision3(&p->env, edges + 6 * i, edges + 6 * i + 3);
}
ORC_API void orc3_points_inside_obs(const orc_plan3 *p, long m, const double *pts, uint8_t *out) {
    for (long i = 0; i < m; i++) out[i] = (uint8_t)is_inside_obs3(&p->env, pts + 3 * i);
}
ORC_API void orc3_points_valid(const orc_plan3 *p, long m, const double *pts, uint8_t *out) {
    for (long i = 0; i < m; i++) out[i] = (uint8_t)is_valid3(&p->env, pts + 3 * i);
}
ORC_API void orc3_nearest(const orc_plan3 *p, long m, const double *q, int64_t *out) {
    for (long k = 0; k < m; k++) {
        double best = 0; int64_t bi = 0;
        for (int64_t i = 0; i < p->n; i++) {
            const double *a = p->v + 3 * i;
            double d = rownorm3(q[3 * k] - a[0], q[3 * k + 1] - a[1], q[3 * k + 2] - a[2]);
            if (i == 0 || d < best) { best = d; bi = i; }
        }
        out[k] = bi;
    }
}
/* np.where(norm(q - v, axis=-1) <= r)[0] (no collision filter) */
ORC_API long orc3_within(const orc_plan3 *p, const double *q, double r, int64_t *out, long cap) {
    long m = 0;
    for (int64_t i = 0; i < p->n; i++) {
        const double *a = p->v + 3 * i;
        if (rownorm3(q[0] - a[0], q[1] - a[1], q[2] - a[2]) <= r) { if (m < cap) out[m] = i; m++; }
    }
    return m;
}
/* draw helpers for pinning the RNG restatement */
ORC_API void orc3_draw_free(orc_plan3 *p, long m, double *out) { for (long i = 0; i < m; i++) sample_free3(p, out + 3 * i); }
ORC_API void orc3_draw_informed(orc_plan3 *p, double c_max, long m, double *out) { for (long i = 0; i < m; i++) sample_informed3(p, c_max, out + 3 * i); }
ORC_API double orc_mt_double(uint32_t *key, int *pos) {
    orc_mt s; double r; memcpy(s.key, key, 2496); s.pos = *pos; r = mt_double(&s); memcpy(key, s.key, 2496); *pos = s.pos; return r;
}
ORC_API int64_t orc_mt_randint(uint32_t *key, int *pos, int64_t high) {
    orc_mt s; int64_t r; memcpy(s.key, key, 2496); s.pos = *pos; r = mt_randint(&s, high); memcpy(key, s.key, 2496); *pos = s.pos; return r;
}

ORC_API void orc_cr_sincos(double x, double *s, double *c) { cr_sincos(x, s, c); }
