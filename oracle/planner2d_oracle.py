"""CPU parity oracle for the 2D planner path (TEST INFRASTRUCTURE ONLY).

A numpy/Python restatement of the reference's 2D hot path, one planning problem at a time, written
against the reference's arithmetic (operand order, which libm/numpy routine evaluates what):

    geometry            path_planning_classes/collision_check_utils.py:8-30,33-60,98-130,158-218,221-394
    Utils               path_planning_classes/rrt_utils_2d.py:4-79
    tree + cost         path_planning_classes/rrt_base_2d.py:25-28,54-61,94-107,109-125
    loop body           path_planning_classes/rrt_star_2d.py:36-55 (== irrt_star_2d.py:48-73)
    steer               rrt_star_2d.py:67-78      (math.hypot, math.atan2, math.cos, math.sin)
    near / choose / rewire / goal   rrt_star_2d.py:80-144   (np.hypot everywhere)
    informed sampling   irrt_star_2d.py:84-97,121-161   (CPython `random` stream for the unit disc)
    drivers             rrt_star_2d.py:198-268, irrt_star_2d.py:230-316

Pinned against traces recorded from the reference's own RRTStar2D / IRRTStar2D
(tests/golden/make_golden_planner2d.py -> tests/golden/planner2d_*.npz; tests/test_oracle_pin2d.py).
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.

RNG: a private ``np.random.RandomState`` (numpy legacy stream) and a private ``random.Random``
(CPython stream), so a run is reproducible from (seed_np, seed_py) alone.
"""
import math
import random as _pyrandom

import numpy as np


# ---------------------------------------------------------------------------------------- geometry
def _cross2(a, b):
    return a[0] * b[1] - a[1] * b[0]


def segments_touch(p, q, r, s, eps=1e-6):
    """Intersection of the infinite lines pq and rs, accepted when it lies inside both segments'
    bounding boxes grown by eps; parallel lines never touch (collision_check_utils.py:8-30)."""
    xd = (p[0] - q[0], r[0] - s[0])
    yd = (p[1] - q[1], r[1] - s[1])
    den = _cross2(xd, yd)
    if den == 0:
        return False
    d = (_cross2(p, q), _cross2(r, s))
    x = _cross2(d, xd) / den
    y = _cross2(d, yd) / den
    return (min(p[0], q[0]) - eps <= x <= max(p[0], q[0]) + eps and min(p[1], q[1]) - eps <= y <= max(p[1], q[1]) + eps and
            min(r[0], s[0]) - eps <= x <= max(r[0], s[0]) + eps and min(r[1], s[1]) - eps <= y <= max(r[1], s[1]) + eps)


def segment_hits_circle(seg, centre, radius, clearance):
    """collision_check_utils.py:33-60"""
    reach = radius + clearance
    v = seg[1] - seg[0]
    length = np.linalg.norm(v)
    if length == 0:
        return bool(np.linalg.norm(seg[0] - centre) <= radius + clearance)
    u = v / length
    t = np.dot(centre - seg[0], u)
    foot = np.clip(t, 0, length) * u + seg[0]
    return bool(np.linalg.norm(np.array(centre) - foot) <= reach)


def point_in_rect(pt, rect, clearance):
    x, y, w, h = rect
    return x - clearance <= pt[0] <= x + w + clearance and y - clearance <= pt[1] <= y + h + clearance


def segment_hits_rect(seg, rect, clearance):
    """collision_check_utils.py:98-130"""
    if point_in_rect(seg[0], rect, clearance) or point_in_rect(seg[1], rect, clearance):
        return True
    x, y, w, h = rect
    c = np.array([[x - clearance, y - clearance], [x + w + clearance, y - clearance],
                  [x + w + clearance, y + h + clearance], [x - clearance, y + h + clearance]])
    return any(segments_touch(seg[0], seg[1], c[k], c[(k + 1) % 4]) for k in range(4))


def segment_collides(seg, circles, rects, clearance):
    """Utils.is_collision: OR over obstacles whose inflated AABB overlaps the segment's AABB
    (collision_check_utils.py:158-218)."""
    lo = np.minimum(seg[0], seg[1]); hi = np.maximum(seg[0], seg[1])
    if circles is not None:
        for c in circles:
            if (lo[0] <= c[0] + c[2] + clearance and hi[0] >= c[0] - c[2] - clearance and
                    lo[1] <= c[1] + c[2] + clearance and hi[1] >= c[1] - c[2] - clearance):
                if segment_hits_circle(seg, c[:2], c[2], clearance):
                    return True
    if rects is not None:
        for r in rects:
            if (lo[0] <= r[0] + r[2] + clearance and hi[0] >= r[0] - clearance and
                    lo[1] <= r[1] + r[3] + clearance and hi[1] >= r[1] - clearance):
                if segment_hits_rect(seg, r, clearance):
                    return True
    return False


def points_in_obstacles(pts, circles, rects, clearance):
    """points_in_circles_rectangles on an (n,2) array (collision_check_utils.py:221-327):
    circles strict <, rectangles inclusive."""
    pts = np.asarray(pts, dtype=np.float64).reshape(-1, 2)
    out = np.zeros(len(pts), dtype=bool)
    if circles is not None:
        for c in circles:
            rc = c[2] + clearance
            out |= (pts[:, 0] - c[0]) ** 2 + (pts[:, 1] - c[1]) ** 2 < rc ** 2
    if rects is not None:
        for r in rects:
            out |= ((r[0] - clearance <= pts[:, 0]) & (pts[:, 0] <= r[0] + r[2] + clearance) &
                    (r[1] - clearance <= pts[:, 1]) & (pts[:, 1] <= r[1] + r[3] + clearance))
    return out


def points_valid(pts, circles, rects, x_range, y_range, clearance):
    """points_validity with equal obstacle / range clearance (collision_check_utils.py:329-394)."""
    pts = np.asarray(pts, dtype=np.float64).reshape(-1, 2)
    w, h = x_range[1] - x_range[0], y_range[1] - y_range[0]
    mc = -clearance
    inside = ((x_range[0] - mc <= pts[:, 0]) & (pts[:, 0] <= x_range[0] + w + mc) &
              (y_range[0] - mc <= pts[:, 1]) & (pts[:, 1] <= y_range[0] + h + mc))
    return inside & ~points_in_obstacles(pts, circles, rects, clearance)


# ------------------------------------------------------------------------------------------ planner
def rotation_to_world_frame_2d(x_start, x_goal, length):
    """IRRTStar2D.RotationToWorldFrame (irrt_star_2d.py:153-161)."""
    a1 = np.zeros((3, 1))
    a1[:2, 0] = (x_goal - x_start) / length
    e1 = np.array([[1.0], [0.0], [0.0]])
    U, _, Vt = np.linalg.svd(a1 @ e1.T, True, True)
    return U @ np.diag([1.0, 1.0, np.linalg.det(U) * np.linalg.det(Vt.T)]) @ Vt


class Oracle2D:
    """variant 0 RRT*, 1 IRRT*, 2 NIRRT* (fixed cloud between set_cloud calls), 3 NRRT*."""

    def __init__(self, problem, iter_max, step_len=10, clearance=3, seed=0, py_seed=None):
        ed = problem["env_dict"]
        self.start = np.array(problem["x_start"]).astype(np.float64)
        self.goal = np.array(problem["x_goal"]).astype(np.float64)
        self.step_len, self.clearance = step_len, clearance
        self.gamma = problem["search_radius"]
        self.h, self.w = ed["env_dims"]
        self.x_range, self.y_range = (0, self.w), (0, self.h)
        self.circles = np.array(ed["circle_obstacles"]) if len(ed["circle_obstacles"]) else None
        self.rects = np.array(ed["rectangle_obstacles"]) if len(ed["rectangle_obstacles"]) else None
        self.v = np.zeros((1 + iter_max, 2)); self.v[0] = self.start
        self.parent = np.zeros(1 + iter_max, dtype=np.int64)
        self.n = 1
        self.iter_max = iter_max
        self.rs = np.random.RandomState(seed)
        self.py = _pyrandom.Random(seed if py_seed is None else py_seed)
        self.solutions = []
        self.cloud, self.pc_rate = None, 0.0
        self.trace = None
        self.c_min = math.hypot(*(self.goal - self.start))
        self.center = np.zeros((3, 1)); self.center[:2, 0] = (self.start + self.goal) / 2.
        self.C = rotation_to_world_frame_2d(self.start, self.goal, self.c_min) if self.c_min > 0 else np.eye(3)

    # -- predicates
    def collides(self, a, b):
        return segment_collides(np.array([a, b]).astype(np.float64), self.circles, self.rects, self.clearance)

    def inside_obs(self, p):
        return bool(points_in_obstacles(np.array([[p[0], p[1]]]), self.circles, self.rects, self.clearance)[0])

    def valid(self, p):
        return bool(points_valid(np.array([[p[0], p[1]]]), self.circles, self.rects, self.x_range, self.y_range, self.clearance)[0])

    # -- tree
    def cost(self, i):
        c = 0.
        while i != 0:
            p = self.parent[i]
            dx, dy = self.v[i] - self.v[p]
            c += math.hypot(dx, dy)
            i = p
        return c

    @staticmethod
    def line(a, b):
        dx, dy = b - a
        return math.hypot(dx, dy)

    # -- samplers
    def sample_free(self):
        c = self.clearance
        while True:
            p = (self.rs.uniform(self.x_range[0] + c, self.x_range[1] - c), self.rs.uniform(self.y_range[0] + c, self.y_range[1] - c))
            if not self.inside_obs(p):
                return np.array(p)

    def sample_informed(self, c_max):
        eps = 1e-6 if c_max ** 2 - self.c_min ** 2 < 0 else 0
        r = [c_max / 2.0, math.sqrt(c_max ** 2 - self.c_min ** 2 + eps) / 2.0, math.sqrt(c_max ** 2 - self.c_min ** 2 + eps) / 2.0]
        L = np.diag(r)
        while True:
            while True:
                x, y = self.py.uniform(-1, 1), self.py.uniform(-1, 1)
                if x ** 2 + y ** 2 < 1:
                    break
            q = np.dot(np.dot(self.C, L), np.array([[x], [y], [0.0]])) + self.center
            if self.valid((q[0, 0], q[1, 0])):
                return q[:2, 0]

    def sample(self, variant, c_best):
        if variant in (2, 3) and self.rs.random_sample() < self.pc_rate:
            return self.cloud[self.rs.randint(0, len(self.cloud))]
        if variant in (1, 2) and c_best < np.inf:
            return self.sample_informed(c_best)
        return self.sample_free()

    def set_cloud(self, cloud, rate):
        self.cloud, self.pc_rate = np.asarray(cloud, dtype=np.float64), rate

    # -- loop body
    def expand(self, x_rand):
        V = self.v[:self.n]
        d = x_rand - V
        near_i = int(np.argmin(np.hypot(d[:, 0], d[:, 1])))
        x_near = V[near_i]
        dx, dy = x_rand - x_near
        dist, theta = math.hypot(dx, dy), math.atan2(dy, dx)
        dist = min(self.step_len, dist)
        x_new = x_near + dist * np.array([math.cos(theta), math.sin(theta)])
        rec = {"nearest": near_i, "new": -1, "near": None}
        if self.trace is not None:
            self.trace.append(rec)
        if self.collides(x_near, x_new):
            return -1
        if np.linalg.norm(x_new - x_near) < 1e-8:
            x_new, new_i = x_near, near_i
            cur = self.cost(near_i)
        else:
            new_i = self.n
            self.v[new_i] = x_new; self.parent[new_i] = near_i; self.n += 1
            cur = self.cost(near_i) + self.line(x_near, x_new)
        rec["new"] = new_i
        r = min(self.gamma * math.sqrt(math.log(self.n) / self.n), self.step_len)
        V = self.v[:self.n]
        d = x_new - V
        cand = np.where(np.hypot(d[:, 0], d[:, 1]) <= r)[0]
        near = np.array([i for i in cand if not self.collides(x_new, V[i]) and i != new_i], dtype=np.int64)
        rec["near"] = near
        if len(near):
            d = x_new - V[near]
            dn = np.hypot(d[:, 0], d[:, 1])
            cands = np.array([self.cost(i) for i in near]) + dn
            k = int(np.argmin(cands))
            if cands[k] < cur:
                self.parent[new_i] = near[k]
            d = V[near] - x_new
            dn = np.hypot(d[:, 0], d[:, 1])
            c_new = self.cost(new_i)
            for k, i in enumerate(near):
                if self.cost(i) > c_new + dn[k]:
                    self.parent[i] = new_i
        return new_i

    def best_solution(self):
        if not self.solutions:
            return np.inf, -1
        costs = [self.cost(i) + self.line(self.v[i], self.goal) for i in self.solutions]
        k = int(np.argmin(costs))
        return costs[k], self.solutions[k]

    def search_goal_parent(self):
        V = self.v[:self.n]
        d = self.goal - V
        dg = np.hypot(d[:, 0], d[:, 1])
        idx = np.where(dg <= self.step_len)[0]
        if len(idx) == 0:
            return None
        tot = [self.cost(i) + dg[i] if not self.collides(V[i], self.goal) else np.inf for i in idx]
        return int(idx[int(np.argmin(tot))])

    def path_len(self, gp):
        pts = [self.goal]
        i = gp
        while i != 0:
            pts.append(self.v[i]); i = self.parent[i]
        pts.append(self.v[0])
        p = np.stack(pts[::-1], axis=0)
        return np.linalg.norm(p[1:] - p[:-1], axis=1).sum()

    def run(self, k, variant=0, mode=0, stop_on_first=False):
        """k loop bodies; returns the recorded values (see nirrt_oracle.c:orc3_run for the rules)."""
        out = []
        for _ in range(k):
            c_best = np.inf
            if variant in (1, 2):
                c_best, _ = self.best_solution()
                out.append(c_best)
                if stop_on_first and c_best < np.inf:
                    break
            new_i = self.expand(self.sample(variant, c_best))
            if variant in (1, 2) and new_i >= 0:
                x = self.v[new_i]
                if self.line(x, self.goal) < self.step_len and not self.collides(x, self.goal):
                    self.solutions.append(new_i)
            if variant in (0, 3) and mode == 1:
                gp = self.search_goal_parent()
                out.append(np.inf if gp is None else self.path_len(gp))
                if stop_on_first and out[-1] < np.inf:
                    break
        return out

    def planning_random(self, iter_after_initial, variant=0):
        if variant in (0, 3):
            lst = self.run(self.iter_max, variant, 1, True)
            if lst[-1] == np.inf:
                return lst
            return lst + self.run(iter_after_initial, variant, 1)
        lst = self.run(self.iter_max, variant, 1, True)
        found = lst[-1] < np.inf
        lst = lst[1:]
        if not found:
            lst.append(self.best_solution()[0])
            if lst[-1] == np.inf:
                return lst
        lst = lst[:-1]
        lst += self.run(iter_after_initial, variant, 1)
        lst.append(self.best_solution()[0])
        return lst
