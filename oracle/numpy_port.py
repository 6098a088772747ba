"""numpy/Python port of the reference's RRT* 3D loop body (TEST INFRASTRUCTURE ONLY).

Purpose: the reference is pure Python+numpy and /root/reference does not exist on the GPU box, so
bench.py's ``cpu_baseline`` / ``--impl reference`` legs time THIS port.  It performs the same work
per iteration, in the same style, as the reference's own classes -- one numpy distance pass for
Nearest, one for Near, a Python loop of per-edge collision checks that rebuild the obstacle AABBs on
every call, un-cached leaf->root cost walks with math.hypot -- so its iters/s is representative of
the reference's CPU path (SURVEY.md section 6 measured 60.4 it/s at n=100k for the reference itself).
It produces bit-identical trees to oracle/nirrt_oracle.c (tests/test_numpy_port.py), which is
pinned against the reference's golden traces.

Functions cite the reference lines they follow (paths relative to tedhuang96/nirrt_star).
"""
import math

import numpy as np


class Obstacles3D:
    """Utils (rrt_utils_3d.py:5-36) + check_collision_line_balls_boxes (collision_check_utils_3d.py:151-216)."""

    def __init__(self, env_dict, clearance):
        balls = np.asarray(env_dict["ball_obstacles"], dtype=np.float64).reshape(-1, 4)
        boxes = np.asarray(env_dict["box_obstacles"], dtype=np.float64).reshape(-1, 6)
        self.balls = balls if len(balls) else None
        self.boxes = boxes if len(boxes) else None
        self.clearance = clearance
        h, w, d = env_dict["env_dims"]
        self.x_range, self.y_range, self.z_range = (0, w), (0, h), (0, d)

    # collision_check_utils_3d.py:3-38
    def _segment_ball(self, seg, center, radius):
        cl = self.clearance
        r = radius + cl
        a, b = seg[0], seg[1]
        u = [b[0] - a[0], b[1] - a[1], b[2] - a[2]]
        if np.linalg.norm(u) == 0:
            return np.linalg.norm(a - center) <= radius + cl
        w = [center[0] - a[0], center[1] - a[1], center[2] - a[2]]
        t = (1 / (u[0] * u[0] + u[1] * u[1] + u[2] * u[2])) * (u[0] * w[0] + u[1] * w[1] + u[2] * w[2])
        if t <= 0:
            return bool((w[0] * w[0] + w[1] * w[1] + w[2] * w[2]) <= r ** 2)
        if t >= 1:
            w2 = [center[0] - b[0], center[1] - b[1], center[2] - b[2]]
            return bool((w2[0] * w2[0] + w2[1] * w2[1] + w2[2] * w2[2]) <= r ** 2)
        if 0 < t < 1:
            foot = [a[0] + t * u[0], a[1] + t * u[1], a[2] + t * u[2]]
            k = [center[0] - foot[0], center[1] - foot[1], center[2] - foot[2]]
            return bool((k[0] * k[0] + k[1] * k[1] + k[2] * k[2]) <= r ** 2)
        return False

    # collision_check_utils_3d.py:41-84
    def _segment_box(self, seg, box):
        cl = self.clearance
        centre = (seg[0] + seg[1]) / 2
        delta = seg[1] - seg[0]
        length = np.linalg.norm(delta)
        x, y, z, w, h, d = box
        if length == 0:
            p = seg[0]
            return bool(x - cl <= p[0] <= x + w + cl and y - cl <= p[1] <= y + h + cl and z - cl <= p[2] <= z + d + cl)
        unit = delta / length
        half = length / 2
        mid_box = [x + w / 2, y + h / 2, z + d / 2]
        ext = [w / 2 + cl, h / 2 + cl, d / 2 + cl]
        T = [mid_box[0] - centre[0], mid_box[1] - centre[1], mid_box[2] - centre[2]]
        for i in range(3):
            if abs(T[i]) > (ext[i] + half * abs(unit[i])):
                return False
        if abs(T[1] * unit[2] - T[2] * unit[1]) > ext[1] * abs(unit[2]) + ext[2] * abs(unit[1]):
            return False
        if abs(T[2] * unit[0] - T[0] * unit[2]) > ext[0] * abs(unit[2]) + ext[2] * abs(unit[0]):
            return False
        if abs(T[0] * unit[1] - T[1] * unit[0]) > ext[0] * abs(unit[1]) + ext[1] * abs(unit[0]):
            return False
        return True

    @staticmethod
    def _aabb_overlap(seg_box, lows, highs):
        ok = np.ones(len(lows), dtype=bool)
        for i in range(3):
            ok = ok * (seg_box[0, i] <= highs[:, i]) * (seg_box[1, i] >= lows[:, i])
        return ok

    # rrt_utils_3d.py:22-36 -> collision_check_utils_3d.py:151-216 (AABBs rebuilt on every call)
    def is_collision(self, start, end):
        seg = np.array([start, end]).astype(np.float64)
        cl = self.clearance
        seg_box = np.array([np.min(seg, axis=0), np.max(seg, axis=0)])
        if self.balls is not None:
            B = self.balls
            lows = np.stack([B[:, 0] - B[:, 3] - cl, B[:, 1] - B[:, 3] - cl, B[:, 2] - B[:, 3] - cl], axis=1)
            highs = np.stack([B[:, 0] + B[:, 3] + cl, B[:, 1] + B[:, 3] + cl, B[:, 2] + B[:, 3] + cl], axis=1)
            for ball in B[np.where(self._aabb_overlap(seg_box, lows, highs))]:
                if self._segment_ball(seg, ball[:3], ball[3]):
                    return True
        if self.boxes is not None:
            X = self.boxes
            lows = np.stack([X[:, 0] - cl, X[:, 1] - cl, X[:, 2] - cl], axis=1)
            highs = np.stack([X[:, 0] + X[:, 3] + cl, X[:, 1] + X[:, 4] + cl, X[:, 2] + X[:, 5] + cl], axis=1)
            for box in X[np.where(self._aabb_overlap(seg_box, lows, highs))]:
                if self._segment_box(seg, box):
                    return True
        return False

    # rrt_utils_3d.py:39-51 -> points_in_balls_boxes (collision_check_utils_3d.py:219-327), single point
    def is_inside_obs(self, p):
        cl = self.clearance
        pts = np.array(p)[np.newaxis, :]
        hit = False
        if self.balls is not None:
            B = self.balls
            m = len(B)
            xp, yp, zp = pts[:, 0:1] * np.ones((1, m)), pts[:, 1:2] * np.ones((1, m)), pts[:, 2:3] * np.ones((1, m))
            rc = B[:, 3] + cl
            hit = bool(np.sum((xp - B[:, 0]) ** 2 + (yp - B[:, 1]) ** 2 + (zp - B[:, 2]) ** 2 < rc ** 2, axis=1).astype(bool)[0])
        if self.boxes is not None and not hit:
            X = self.boxes
            m = len(X)
            xp, yp, zp = pts[:, 0:1] * np.ones((1, m)), pts[:, 1:2] * np.ones((1, m)), pts[:, 2:3] * np.ones((1, m))
            lo = X[:, :3] - cl
            hi = X[:, :3] + X[:, 3:] + cl
            inside = (lo[:, 0] <= xp) * (xp <= hi[:, 0]) * (lo[:, 1] <= yp) * (yp <= hi[:, 1]) * (lo[:, 2] <= zp) * (zp <= hi[:, 2])
            hit = bool(np.sum(inside, axis=1).astype(bool)[0])
        return hit


class RRTStar3DPort:
    """RRTBase3D + RRTStar3D (rrt_base_3d.py:7-137, rrt_star_3d.py:7-157), planning() loop body."""

    def __init__(self, problem, iter_max, step_len=10, clearance=2, rng=None):
        self.x_start = np.array(problem["x_start"]).astype(np.float64)
        self.x_goal = np.array(problem["x_goal"]).astype(np.float64)
        self.step_len, self.search_radius, self.iter_max = step_len, problem["search_radius"], iter_max
        self.vertices = np.zeros((1 + iter_max, 3))
        self.vertex_parents = np.zeros(1 + iter_max).astype(int)
        self.vertices[0] = self.x_start
        self.num_vertices = 1
        self.obs = Obstacles3D(problem["env_dict"], clearance)
        self.clearance = clearance
        self.rng = rng if rng is not None else np.random.RandomState(0)   # private legacy MT19937 stream

    def load_tree(self, vertices, parents):
        n = len(vertices)
        self.vertices[:n] = vertices
        self.vertex_parents[:n] = parents
        self.num_vertices = n

    # rrt_base_3d.py:49-58
    def sample_free(self):
        o, c = self.obs, self.clearance
        while True:
            p = (self.rng.uniform(o.x_range[0] + c, o.x_range[1] - c),
                 self.rng.uniform(o.y_range[0] + c, o.y_range[1] - c),
                 self.rng.uniform(o.z_range[0] + c, o.z_range[1] - c))
            if not o.is_inside_obs(p):
                return np.array(p)

    # rrt_base_3d.py:60-67
    def cost(self, i):
        total = 0.
        V = self.vertices[:self.num_vertices]
        while i != 0:
            j = self.vertex_parents[i]
            dx, dy, dz = V[i] - V[j]
            total += math.hypot(dx, dy, dz)
            i = j
        return total

    # one body of the for-loop of rrt_star_3d.py:36-55
    def iterate(self):
        q = self.sample_free()
        V = self.vertices[:self.num_vertices]
        near_i = np.argmin(np.linalg.norm(q - V, axis=1))                                # rrt_base_3d.py:100-113
        x_near = V[near_i]
        dx, dy, dz = q - x_near                                                            # rrt_base_3d.py:116-130
        dist = math.hypot(dx, dy, dz)
        direction = np.zeros(3) if dist == 0 else (q - x_near) / dist
        x_new = x_near + min(self.step_len, dist) * direction                              # rrt_star_3d.py:67-78
        if self.obs.is_collision(x_near, x_new):
            return -1
        if np.linalg.norm(x_new - x_near) < 1e-8:
            x_new, new_i = x_near, near_i
            base_cost = self.cost(near_i)
        else:
            new_i = self.num_vertices
            self.vertices[new_i] = x_new
            self.vertex_parents[new_i] = near_i
            self.num_vertices += 1
            ex, ey, ez = x_new - x_near
            base_cost = self.cost(near_i) + math.hypot(ex, ey, ez)
        # find_near_neighbors (rrt_star_3d.py:125-145)
        n = self.num_vertices
        r = min(self.search_radius * (math.log(n) / n) ** (1 / 3.), self.step_len)
        V = self.vertices[:n]
        cand = np.where(np.linalg.norm(x_new - V, axis=-1) <= r)[0]
        near = []
        for k, vk in zip(cand, V[cand]):
            if not self.obs.is_collision(x_new, vk):
                if k != new_i:
                    near.append(k)
        near = np.array(near)
        if len(near) > 0:
            # choose_parent (rrt_star_3d.py:80-90)
            d = np.linalg.norm(x_new - V[near], axis=-1)
            through = np.array([self.cost(k) for k in near]) + d
            b = np.argmin(through)
            if through[b] < base_cost:
                self.vertex_parents[new_i] = near[b]
            # rewire (rrt_star_3d.py:92-99)
            d = np.linalg.norm(V[near] - x_new, axis=-1)
            c_new = self.cost(new_i)
            for i, k in enumerate(near):
                if self.cost(k) > c_new + d[i]:
                    self.vertex_parents[k] = new_i
        return new_i

    def run(self, iters):
        for _ in range(iters):
            self.iterate()
