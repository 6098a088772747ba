// geometry2d.cuh -- exact float64 segment/point vs circle/rectangle predicates (2D worlds).
//
// Device restatement of the reference's collision_check_utils.py; every arithmetic step keeps the
// reference's operand order and rounds once per operation (see exact_math.cuh):
//   seg_hits_circle      check_collision_line_single_circle      collision_check_utils.py:33-60
//   seg_hits_rect        check_collision_line_single_rectangle   collision_check_utils.py:98-130
//   lines_intersect      line_intersection (eps = 1e-6)          collision_check_utils.py:8-30
//   seg_collides         check_collision_line_circles_rectangles collision_check_utils.py:158-218
//   point_inside_obs     points_in_circles_rectangles            collision_check_utils.py:298-327
//   point_valid          points_validity                         collision_check_utils.py:353-394
// The AABB pre-filter of the reference (:145-155) is part of the boolean and is kept.
#pragma once
#include "exact_math.cuh"
#include "geometry3d.cuh"

namespace nirrt {

struct Geom2 {
    int n_circles, n_rects;
    double clearance;
    double range[4];             // x0 x1 y0 y1
    double circles[kMaxObs][3];  // x y r
    double rects[kMaxObs][4];    // x y w h
};

NIRRT_HD bool point_in_single_circle(const double *p, const double *c, double cl) {
    return vecnorm2(XSUB(p[0], c[0]), XSUB(p[1], c[1])) <= XADD(c[2], cl);
}
NIRRT_HD bool point_in_single_rect(const double *p, const double *r, double cl) {
    return XSUB(r[0], cl) <= p[0] && p[0] <= XADD(XADD(r[0], r[2]), cl) &&
           XSUB(r[1], cl) <= p[1] && p[1] <= XADD(XADD(r[1], r[3]), cl);
}

NIRRT_HD bool seg_hits_circle(const double *p0, const double *p1, const double *c, double cl) {
    const double rwc = XADD(c[2], cl);
    const double lx = XSUB(p1[0], p0[0]), ly = XSUB(p1[1], p0[1]);
    const double len = vecnorm2(lx, ly);
    if (len == 0.0) return point_in_single_circle(p0, c, cl);
    const double ux = XDIV(lx, len), uy = XDIV(ly, len);
    const double sx = XSUB(c[0], p0[0]), sy = XSUB(c[1], p0[1]);
    double t = dot2(sx, sy, ux, uy);
    t = t < 0.0 ? 0.0 : t;                      // np.clip(projection, 0, line_length)
    t = t > len ? len : t;
    const double qx = XADD(XMUL(t, ux), p0[0]), qy = XADD(XMUL(t, uy), p0[1]);
    return vecnorm2(XSUB(c[0], qx), XSUB(c[1], qy)) <= rwc;
}

NIRRT_HD double det2(double a0, double a1, double b0, double b1) { return XSUB(XMUL(a0, b1), XMUL(a1, b0)); }
NIRRT_HD double dmin(double a, double b) { return b < a ? b : a; }   // Python min/max on floats
NIRRT_HD double dmax(double a, double b) { return b > a ? b : a; }

// line_intersection(line1 = (a, b), line2 = (c, d))
NIRRT_HD bool lines_intersect(const double *a, const double *b, const double *c, const double *d) {
    const double xd0 = XSUB(a[0], b[0]), xd1 = XSUB(c[0], d[0]);
    const double yd0 = XSUB(a[1], b[1]), yd1 = XSUB(c[1], d[1]);
    const double div = det2(xd0, xd1, yd0, yd1);
    if (div == 0.0) return false;
    const double d0 = det2(a[0], a[1], b[0], b[1]), d1 = det2(c[0], c[1], d[0], d[1]);
    const double x = XDIV(det2(d0, d1, xd0, xd1), div);
    const double y = XDIV(det2(d0, d1, yd0, yd1), div);
    const double eps = 1e-6;
    return XSUB(dmin(a[0], b[0]), eps) <= x && x <= XADD(dmax(a[0], b[0]), eps) &&
           XSUB(dmin(a[1], b[1]), eps) <= y && y <= XADD(dmax(a[1], b[1]), eps) &&
           XSUB(dmin(c[0], d[0]), eps) <= x && x <= XADD(dmax(c[0], d[0]), eps) &&
           XSUB(dmin(c[1], d[1]), eps) <= y && y <= XADD(dmax(c[1], d[1]), eps);
}

NIRRT_HD bool seg_hits_rect(const double *p0, const double *p1, const double *r, double cl) {
    if (point_in_single_rect(p0, r, cl) || point_in_single_rect(p1, r, cl)) return true;
    const double x0 = XSUB(r[0], cl), y0 = XSUB(r[1], cl);
    const double x1 = XADD(XADD(r[0], r[2]), cl), y1 = XADD(XADD(r[1], r[3]), cl);
    const double A[2] = {x0, y0}, B[2] = {x1, y0}, Cc[2] = {x1, y1}, Dd[2] = {x0, y1};
    return lines_intersect(p0, p1, A, B) || lines_intersect(p0, p1, B, Cc) || lines_intersect(p0, p1, Cc, Dd) ||
           lines_intersect(p0, p1, Dd, A);
}

// one obstacle (k < n_circles: circle k, else rectangle k - n_circles), AABB filter included
NIRRT_HD bool seg_hits_obstacle(const Geom2 &g, int k, const double *p0, const double *p1) {
    const double cl = g.clearance;
    double lo[2], hi[2];
    for (int i = 0; i < 2; i++) { lo[i] = dmin(p0[i], p1[i]); hi[i] = dmax(p0[i], p1[i]); }
    if (k < g.n_circles) {
        const double *c = g.circles[k];
        for (int i = 0; i < 2; i++) {
            const double a1 = XSUB(XSUB(c[i], c[2]), cl), a2 = XADD(XADD(c[i], c[2]), cl);
            if (!(lo[i] <= a2 && hi[i] >= a1)) return false;
        }
        return seg_hits_circle(p0, p1, c, cl);
    }
    const double *r = g.rects[k - g.n_circles];
    for (int i = 0; i < 2; i++) {
        const double a1 = XSUB(r[i], cl), a2 = XADD(XADD(r[i], r[2 + i]), cl);
        if (!(lo[i] <= a2 && hi[i] >= a1)) return false;
    }
    return seg_hits_rect(p0, p1, r, cl);
}
NIRRT_HD int n_obstacles(const Geom2 &g) { return g.n_circles + g.n_rects; }
NIRRT_HD int n_obstacles(const Geom3 &g) { return g.n_balls + g.n_boxes; }

NIRRT_HD bool seg_collides(const Geom2 &g, const double *p0, const double *p1) {
    const int m = g.n_circles + g.n_rects;
    for (int k = 0; k < m; k++)
        if (seg_hits_obstacle(g, k, p0, p1)) return true;
    return false;
}

NIRRT_HD bool point_in_circles(const Geom2 &g, const double *p) {   // strict <
    for (int k = 0; k < g.n_circles; k++) {
        const double *c = g.circles[k];
        const double rc = XADD(c[2], g.clearance);
        const double dx = XSUB(p[0], c[0]), dy = XSUB(p[1], c[1]);
        if (XADD(XMUL(dx, dx), XMUL(dy, dy)) < XMUL(rc, rc)) return true;
    }
    return false;
}
NIRRT_HD bool point_in_rects(const Geom2 &g, const double *p) {      // inclusive
    for (int k = 0; k < g.n_rects; k++)
        if (point_in_single_rect(p, g.rects[k], g.clearance)) return true;
    return false;
}
// one term of is_inside_obs: obstacle k (circles first, then rectangles), so lanes can split the OR
NIRRT_HD bool point_in_obstacle(const Geom2 &g, int k, const double *p) {
    if (k < g.n_circles) {
        const double *c = g.circles[k];
        const double rc = XADD(c[2], g.clearance);
        const double dx = XSUB(p[0], c[0]), dy = XSUB(p[1], c[1]);
        return XADD(XMUL(dx, dx), XMUL(dy, dy)) < XMUL(rc, rc);
    }
    return point_in_single_rect(p, g.rects[k - g.n_circles], g.clearance);
}
// Utils.is_inside_obs (rrt_utils_2d.py:36-48)
NIRRT_HD bool point_inside_obs(const Geom2 &g, const double *p) { return point_in_circles(g, p) || point_in_rects(g, p); }
// Utils.is_valid (rrt_utils_2d.py:62-79); range test == points_in_rectangles with clearance -c
NIRRT_HD bool point_valid(const Geom2 &g, const double *p) {
    const double mc = -g.clearance;
    for (int i = 0; i < 2; i++) {
        const double mn = g.range[2 * i], w = XSUB(g.range[2 * i + 1], g.range[2 * i]);
        if (!(XSUB(mn, mc) <= p[i] && p[i] <= XADD(XADD(mn, w), mc))) return false;
    }
    return !point_in_circles(g, p) && !point_in_rects(g, p);
}

}  // namespace nirrt
