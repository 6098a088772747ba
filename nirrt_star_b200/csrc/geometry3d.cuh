// geometry3d.cuh -- exact float64 segment/point vs ball/box predicates (3D worlds).
//
// Device restatement of the reference's collision_check_utils_3d.py; every arithmetic step keeps
// the reference's operand order and rounds once per operation (see exact_math.cuh), so a predicate
// returns the same bool as the numpy/Python code for the same float64 inputs:
//   seg_hits_ball        check_collision_line_single_ball      collision_check_utils_3d.py:3-38
//   seg_hits_box         check_collision_line_single_box       collision_check_utils_3d.py:41-84
//   seg_collides         check_collision_line_balls_boxes      collision_check_utils_3d.py:151-216
//   point_inside_obs     points_in_balls_boxes                 collision_check_utils_3d.py:298-327
//   point_valid          points_validity_3d                    collision_check_utils_3d.py:354-398
// The AABB pre-filter of the reference is kept (it is part of the boolean: an obstacle whose
// inflated AABB misses the segment's AABB is never tested exactly).
#pragma once
#include "exact_math.cuh"

namespace nirrt {

constexpr int kMaxObs = 32;  // per obstacle type and problem

// One problem's obstacle table.  Lives in global memory, staged into shared memory per CTA.
struct Geom3 {
    int n_balls, n_boxes;
    double clearance;
    double range[6];           // x0 x1 y0 y1 z0 z1
    double balls[kMaxObs][4];  // x y z r
    double ball_r2[kMaxObs];   // (r + clearance) ** 2 as numpy's *scalar* power returns it (host-computed)
    double boxes[kMaxObs][6];  // x y z w h d
};

NIRRT_HD bool point_in_single_ball(const double *p, const double *ball, double cl) {
    return vecnorm3(XSUB(p[0], ball[0]), XSUB(p[1], ball[1]), XSUB(p[2], ball[2])) <= XADD(ball[3], cl);
}

NIRRT_HD bool point_in_single_box(const double *p, const double *b, double cl) {
    return XSUB(b[0], cl) <= p[0] && p[0] <= XADD(XADD(b[0], b[3]), cl) &&
           XSUB(b[1], cl) <= p[1] && p[1] <= XADD(XADD(b[1], b[4]), cl) &&
           XSUB(b[2], cl) <= p[2] && p[2] <= XADD(XADD(b[2], b[5]), cl);
}

NIRRT_HD double dot3_plain(double a0, double a1, double a2, double b0, double b1, double b2) {
    return XADD(XADD(XMUL(a0, b0), XMUL(a1, b1)), XMUL(a2, b2));
}

NIRRT_HD bool seg_hits_ball(const double *p0, const double *p1, const double *ball, double r2, double cl) {
    const double l0 = XSUB(p1[0], p0[0]), l1 = XSUB(p1[1], p0[1]), l2 = XSUB(p1[2], p0[2]);
    if (vecnorm3(l0, l1, l2) == 0.0) return point_in_single_ball(p0, ball, cl);
    const double d0 = XSUB(ball[0], p0[0]), d1 = XSUB(ball[1], p0[1]), d2 = XSUB(ball[2], p0[2]);
    const double t = XMUL(XDIV(1.0, dot3_plain(l0, l1, l2, l0, l1, l2)), dot3_plain(l0, l1, l2, d0, d1, d2));
    if (t <= 0.0) {
        return dot3_plain(d0, d1, d2, d0, d1, d2) <= r2;
    } else if (t >= 1.0) {
        const double e0 = XSUB(ball[0], p1[0]), e1 = XSUB(ball[1], p1[1]), e2 = XSUB(ball[2], p1[2]);
        return dot3_plain(e0, e1, e2, e0, e1, e2) <= r2;
    } else if (0.0 < t && t < 1.0) {
        const double x0 = XADD(p0[0], XMUL(t, l0)), x1 = XADD(p0[1], XMUL(t, l1)), x2 = XADD(p0[2], XMUL(t, l2));
        const double k0 = XSUB(ball[0], x0), k1 = XSUB(ball[1], x1), k2 = XSUB(ball[2], x2);
        return dot3_plain(k0, k1, k2, k0, k1, k2) <= r2;
    }
    return false;  // t is NaN
}

NIRRT_HD bool seg_hits_box(const double *p0, const double *p1, const double *b, double cl) {
    double mid[3], dir[3], I[3], E[3], T[3];
    for (int i = 0; i < 3; i++) { mid[i] = XDIV(XADD(p0[i], p1[i]), 2.0); dir[i] = XSUB(p1[i], p0[i]); }
    const double dist = vecnorm3(dir[0], dir[1], dir[2]);
    if (dist == 0.0) return point_in_single_box(p0, b, cl);
    const double hl = XDIV(dist, 2.0);
    for (int i = 0; i < 3; i++) {
        I[i] = XDIV(dir[i], dist);
        const double half = XDIV(b[3 + i], 2.0);
        E[i] = XADD(half, cl);
        T[i] = XSUB(XADD(b[i], half), mid[i]);
    }
    if (fabs(T[0]) > XADD(E[0], XMUL(hl, fabs(I[0])))) return false;
    if (fabs(T[1]) > XADD(E[1], XMUL(hl, fabs(I[1])))) return false;
    if (fabs(T[2]) > XADD(E[2], XMUL(hl, fabs(I[2])))) return false;
    double r = XADD(XMUL(E[1], fabs(I[2])), XMUL(E[2], fabs(I[1])));
    if (fabs(XSUB(XMUL(T[1], I[2]), XMUL(T[2], I[1]))) > r) return false;
    r = XADD(XMUL(E[0], fabs(I[2])), XMUL(E[2], fabs(I[0])));
    if (fabs(XSUB(XMUL(T[2], I[0]), XMUL(T[0], I[2]))) > r) return false;
    r = XADD(XMUL(E[0], fabs(I[1])), XMUL(E[1], fabs(I[0])));
    if (fabs(XSUB(XMUL(T[0], I[1]), XMUL(T[1], I[0]))) > r) return false;
    return true;
}

// one obstacle (k < n_balls: ball k, else box k - n_balls) against the segment, AABB filter included
NIRRT_HD bool seg_hits_obstacle(const Geom3 &g, int k, const double *p0, const double *p1) {
    const double cl = g.clearance;
    double lo[3], hi[3];
    for (int i = 0; i < 3; i++) { lo[i] = p0[i] < p1[i] ? p0[i] : p1[i]; hi[i] = p0[i] > p1[i] ? p0[i] : p1[i]; }
    if (k < g.n_balls) {
        const double *b = g.balls[k];
        for (int i = 0; i < 3; i++) {
            const double a1 = XSUB(XSUB(b[i], b[3]), cl), a2 = XADD(XADD(b[i], b[3]), cl);
            if (!(lo[i] <= a2 && hi[i] >= a1)) return false;
        }
        return seg_hits_ball(p0, p1, b, g.ball_r2[k], cl);
    }
    const double *b = g.boxes[k - g.n_balls];
    for (int i = 0; i < 3; i++) {
        const double a1 = XSUB(b[i], cl), a2 = XADD(XADD(b[i], b[3 + i]), cl);
        if (!(lo[i] <= a2 && hi[i] >= a1)) return false;
    }
    return seg_hits_box(p0, p1, b, cl);
}

// Utils.is_collision (rrt_utils_3d.py:22-36): OR over all obstacles (order-independent)
NIRRT_HD bool seg_collides(const Geom3 &g, const double *p0, const double *p1) {
    const int m = g.n_balls + g.n_boxes;
    for (int k = 0; k < m; k++)
        if (seg_hits_obstacle(g, k, p0, p1)) return true;
    return false;
}

NIRRT_HD bool point_in_balls(const Geom3 &g, const double *p) {  // strict <, array power == x*x
    for (int k = 0; k < g.n_balls; k++) {
        const double *b = g.balls[k];
        const double rc = XADD(b[3], g.clearance);
        const double dx = XSUB(p[0], b[0]), dy = XSUB(p[1], b[1]), dz = XSUB(p[2], b[2]);
        if (XADD(XADD(XMUL(dx, dx), XMUL(dy, dy)), XMUL(dz, dz)) < XMUL(rc, rc)) return true;
    }
    return false;
}
NIRRT_HD bool point_in_boxes(const Geom3 &g, const double *p) {  // inclusive
    for (int k = 0; k < g.n_boxes; k++)
        if (point_in_single_box(p, g.boxes[k], g.clearance)) return true;
    return false;
}
// one term of is_inside_obs: obstacle k (balls first, then boxes), so lanes can split the OR
NIRRT_HD bool point_in_obstacle(const Geom3 &g, int k, const double *p) {
    if (k < g.n_balls) {
        const double *b = g.balls[k];
        const double rc = XADD(b[3], g.clearance);
        const double dx = XSUB(p[0], b[0]), dy = XSUB(p[1], b[1]), dz = XSUB(p[2], b[2]);
        return XADD(XADD(XMUL(dx, dx), XMUL(dy, dy)), XMUL(dz, dz)) < XMUL(rc, rc);
    }
    return point_in_single_box(p, g.boxes[k - g.n_balls], g.clearance);
}
// Utils.is_inside_obs (rrt_utils_3d.py:39-51)
NIRRT_HD bool point_inside_obs(const Geom3 &g, const double *p) { return point_in_balls(g, p) || point_in_boxes(g, p); }
// Utils.is_valid (rrt_utils_3d.py:68-86); range test == points_in_boxes with clearance -c
NIRRT_HD bool point_valid(const Geom3 &g, const double *p) {
    const double mc = -g.clearance;
    for (int i = 0; i < 3; i++) {
        const double mn = g.range[2 * i], w = XSUB(g.range[2 * i + 1], g.range[2 * i]);
        if (!(XSUB(mn, mc) <= p[i] && p[i] <= XADD(XADD(mn, w), mc))) return false;
    }
    return !point_in_balls(g, p) && !point_in_boxes(g, p);
}

}  // namespace nirrt
