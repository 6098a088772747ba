"""ctypes front-end of the CPU parity oracle (TEST INFRASTRUCTURE ONLY -- see nirrt_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package never does.
"""
import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_dp = C.POINTER(C.c_double)
c_i64p = C.POINTER(C.c_int64)
c_u32p = C.POINTER(C.c_uint32)
c_u8p = C.POINTER(C.c_uint8)


def build(force=False):
    so = os.path.join(_HERE, "libnirrt_oracle.so")
    src = os.path.join(_HERE, "nirrt_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libnirrt_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        L = _LIB
        L.orc3_create.restype = C.c_void_p
        L.orc3_create.argtypes = [C.c_int, c_dp, c_dp, C.c_double, C.c_double, C.c_double, c_dp,
                                  C.c_int, c_dp, c_dp, C.c_int, c_dp, c_dp, c_u32p, C.c_int]
        L.orc3_destroy.argtypes = [C.c_void_p]
        L.orc3_set_informed.argtypes = [C.c_void_p, c_dp]
        L.orc3_set_cloud.argtypes = [C.c_void_p, c_dp, C.c_int, C.c_double]
        L.orc3_load_tree.argtypes = [C.c_void_p, C.c_int, c_dp, c_i64p]
        L.orc3_num_vertices.argtypes = [C.c_void_p]
        L.orc3_get_tree.argtypes = [C.c_void_p, c_dp, c_i64p]
        L.orc3_get_solutions.argtypes = [C.c_void_p, c_i64p]
        L.orc3_get_rng.argtypes = [C.c_void_p, c_u32p, C.POINTER(C.c_int)]
        L.orc3_run.restype = C.c_long
        L.orc3_run.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_long, C.c_int, c_dp, c_i64p, c_i64p,
                               c_i64p, c_i64p, C.c_long, C.POINTER(C.c_long)]
        L.orc3_best_cost.restype = C.c_double
        L.orc3_best_cost.argtypes = [C.c_void_p, c_i64p]
        L.orc3_search_goal_parent.restype = C.c_int64
        L.orc3_search_goal_parent.argtypes = [C.c_void_p]
        L.orc3_path_len.restype = C.c_double
        L.orc3_path_len.argtypes = [C.c_void_p, C.c_int64]
        L.orc3_cost.restype = C.c_double
        L.orc3_cost.argtypes = [C.c_void_p, C.c_int64]
        for f in ("orc_hypot3", "orc_rownorm3", "orc_vecnorm3"):
            getattr(L, f).restype = C.c_double
            getattr(L, f).argtypes = [C.c_double] * 3
        L.orc_hypot2.restype = C.c_double
        L.orc_hypot2.argtypes = [C.c_double] * 2
        L.orc_pairwise_sum.restype = C.c_double
        L.orc_pairwise_sum.argtypes = [c_dp, C.c_long]
        L.orc3_collide_edges.argtypes = [C.c_void_p, C.c_long, c_dp, c_u8p]
        L.orc3_points_inside_obs.argtypes = [C.c_void_p, C.c_long, c_dp, c_u8p]
        L.orc3_points_valid.argtypes = [C.c_void_p, C.c_long, c_dp, c_u8p]
        L.orc3_nearest.argtypes = [C.c_void_p, C.c_long, c_dp, c_i64p]
        L.orc3_within.restype = C.c_long
        L.orc3_within.argtypes = [C.c_void_p, c_dp, C.c_double, c_i64p, C.c_long]
        L.orc3_draw_free.argtypes = [C.c_void_p, C.c_long, c_dp]
        L.orc3_draw_informed.argtypes = [C.c_void_p, C.c_double, C.c_long, c_dp]
    return _LIB


def _dp(a):
    return a.ctypes.data_as(c_dp)


def _ip(a):
    return a.ctypes.data_as(c_i64p)


def near_radius_table(cap, dim=3):
    """t[n] = (math.log(n)/n)**(1/dim) with CPython's own libm calls, exactly the expression in
    find_near_neighbors (rrt_star_3d.py:134: ``**(1/3.)``; rrt_star_2d.py:133: ``**0.5`` via
    math.sqrt).  t[0] is unused."""
    t = np.zeros(cap + 2, dtype=np.float64)
    for n in range(1, cap + 2):
        t[n] = (math.log(n) / n) ** (1 / 3.) if dim == 3 else math.sqrt(math.log(n) / n)
    return t


def ball_r2_scalar_pow(balls, clearance):
    """(ball_radius + clearance) ** 2 evaluated as the reference does inside
    check_collision_line_single_ball (collision_check_utils_3d.py:21,31-37): numpy *scalar* power."""
    balls = np.asarray(balls, dtype=np.float64).reshape(-1, 4)
    return np.array([float((b[3] + clearance) ** 2) for b in balls], dtype=np.float64)


def rotation_to_world_frame(x_start, x_goal):
    """IRRTStar3D.RotationToWorldFrame (irrt_star_3d.py:159-173), evaluated with numpy itself."""
    x_start = np.array(x_start).astype(np.float64)
    x_goal = np.array(x_goal).astype(np.float64)
    dx, dy, dz = x_goal - x_start
    L = math.hypot(dx, dy, dz)
    a1 = (x_goal - x_start) / L
    M = np.outer(a1, [1, 0, 0])
    U, S, V = np.linalg.svd(M)
    return U @ np.diag([1, 1, np.linalg.det(U) * np.linalg.det(V)]) @ V.T


class Oracle3D:
    """One 3D planning problem on the CPU oracle.  ``seed`` seeds a numpy legacy MT19937 exactly
    like ``np.random.seed(seed)``; alternatively pass ``rng_state=(key[624] u32, pos)``."""

    def __init__(self, problem, iter_max, step_len=10, clearance=2, seed=0, rng_state=None):
        L = lib()
        ed = problem["env_dict"]
        self.cap = 1 + iter_max
        self.start = np.array(problem["x_start"]).astype(np.float64)
        self.goal = np.array(problem["x_goal"]).astype(np.float64)
        h, w, d = ed["env_dims"]
        self.range6 = np.array([0, w, 0, h, 0, d], dtype=np.float64)
        self.balls = np.ascontiguousarray(np.asarray(ed["ball_obstacles"], dtype=np.float64).reshape(-1, 4))
        self.boxes = np.ascontiguousarray(np.asarray(ed["box_obstacles"], dtype=np.float64).reshape(-1, 6))
        self.r2 = ball_r2_scalar_pow(self.balls, clearance)
        self.rtab = near_radius_table(self.cap, 3)
        if rng_state is None:
            st = np.random.RandomState(seed).get_state()
            key, pos = st[1].astype(np.uint32), int(st[2])
        else:
            key, pos = np.ascontiguousarray(rng_state[0], dtype=np.uint32), int(rng_state[1])
        self._key = key
        self.step_len, self.clearance = step_len, clearance
        self.search_radius = float(problem["search_radius"])
        self.h = L.orc3_create(self.cap, _dp(self.start), _dp(self.goal), float(step_len),
                               self.search_radius, float(clearance), _dp(self.range6),
                               len(self.balls), _dp(self.balls), _dp(self.r2),
                               len(self.boxes), _dp(self.boxes), _dp(self.rtab),
                               key.ctypes.data_as(c_u32p), pos)
        self.Cmat = np.ascontiguousarray(rotation_to_world_frame(self.start, self.goal))
        L.orc3_set_informed(self.h, _dp(self.Cmat))
        self._pc = None

    def __del__(self):
        try:
            lib().orc3_destroy(self.h)
        except Exception:
            pass

    def set_cloud(self, pc, rate):
        self._pc = np.ascontiguousarray(pc, dtype=np.float64)
        lib().orc3_set_cloud(self.h, _dp(self._pc), len(self._pc), float(rate))

    def load_tree(self, vertices, parents):
        v = np.ascontiguousarray(vertices, dtype=np.float64)
        p = np.ascontiguousarray(parents, dtype=np.int64)
        lib().orc3_load_tree(self.h, len(v), _dp(v), _ip(p))

    @property
    def num_vertices(self):
        return lib().orc3_num_vertices(self.h)

    def tree(self):
        n = self.num_vertices
        v = np.zeros((n, 3)); p = np.zeros(n, dtype=np.int64)
        lib().orc3_get_tree(self.h, _dp(v), _ip(p))
        return v, p

    def solutions(self):
        n = lib().orc3_get_solutions(self.h, None)
        out = np.zeros(max(n, 1), dtype=np.int64)
        lib().orc3_get_solutions(self.h, _ip(out))
        return out[:n]

    def rng_state(self):
        key = np.zeros(624, dtype=np.uint32); pos = C.c_int(0)
        lib().orc3_get_rng(self.h, key.ctypes.data_as(c_u32p), C.byref(pos))
        return key, pos.value

    def run(self, k, variant=0, mode=0, stop_on_first=False, trace=False, near_cap=None):
        """Runs up to k loop bodies; returns dict(iters, pathlen[, nearest, new, near_cnt, near])."""
        pathlen = np.full(k, np.nan)
        out = {}
        if trace:
            tn = np.zeros(k, dtype=np.int64); tw = np.zeros(k, dtype=np.int64); tc = np.zeros(k, dtype=np.int64)
            cap = near_cap or (k * 256 + 1024)
            nb = np.zeros(cap, dtype=np.int64); used = C.c_long(0)
            it = lib().orc3_run(self.h, variant, mode, k, int(stop_on_first), _dp(pathlen), _ip(tn), _ip(tw), _ip(tc),
                                _ip(nb), cap, C.byref(used))
            out.update(nearest=tn[:it], new=tw[:it], near_cnt=tc[:it], near=nb[:used.value])
        else:
            it = lib().orc3_run(self.h, variant, mode, k, int(stop_on_first), _dp(pathlen), None, None, None, None, 0, None)
        out.update(iters=it, pathlen=pathlen[:it])
        return out

    # --- driver semantics of the reference (Appendix B of SURVEY.md) -------------------------
    def planning_random(self, iter_after_initial, variant=0):
        """RRTStar3D.planning_random (rrt_star_3d.py:200-270) for variant 0,
        IRRTStar3D.planning_random (irrt_star_3d.py:245-331) for variant 1/2."""
        iter_max = self.cap - 1
        if variant in (0, 3):
            r1 = self.run(iter_max, variant, 1, stop_on_first=True)
            lst = list(r1["pathlen"])
            if lst[-1] == np.inf:
                return lst
            r2 = self.run(iter_after_initial, variant, 1)
            return lst + list(r2["pathlen"])
        r1 = self.run(iter_max, variant, 1, stop_on_first=True)
        lst = list(r1["pathlen"])
        found = lst[-1] < np.inf
        lst = lst[1:]
        if not found:
            lst.append(self.best_cost()[0])
            if lst[-1] == np.inf:
                return lst
        lst = lst[:-1]
        r2 = self.run(iter_after_initial, variant, 1)
        lst += list(r2["pathlen"])
        lst.append(self.best_cost()[0])
        return lst

    def best_cost(self):
        xb = C.c_int64(0)
        c = lib().orc3_best_cost(self.h, C.byref(xb))
        return c, xb.value

    def search_goal_parent(self):
        return lib().orc3_search_goal_parent(self.h)

    def cost(self, idx):
        return lib().orc3_cost(self.h, int(idx))

    # --- function-level entry points ----------------------------------------------------------
    def collide_edges(self, edges):
        e = np.ascontiguousarray(edges, dtype=np.float64).reshape(-1, 6)
        out = np.zeros(len(e), dtype=np.uint8)
        lib().orc3_collide_edges(self.h, len(e), _dp(e), out.ctypes.data_as(c_u8p))
        return out.astype(bool)

    def points_inside_obs(self, pts):
        p = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 3)
        out = np.zeros(len(p), dtype=np.uint8)
        lib().orc3_points_inside_obs(self.h, len(p), _dp(p), out.ctypes.data_as(c_u8p))
        return out.astype(bool)

    def points_valid(self, pts):
        p = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 3)
        out = np.zeros(len(p), dtype=np.uint8)
        lib().orc3_points_valid(self.h, len(p), _dp(p), out.ctypes.data_as(c_u8p))
        return out.astype(bool)

    def nearest(self, queries):
        q = np.ascontiguousarray(queries, dtype=np.float64).reshape(-1, 3)
        out = np.zeros(len(q), dtype=np.int64)
        lib().orc3_nearest(self.h, len(q), _dp(q), _ip(out))
        return out

    def within(self, q, r):
        q = np.ascontiguousarray(q, dtype=np.float64)
        cap = self.num_vertices
        out = np.zeros(cap, dtype=np.int64)
        m = lib().orc3_within(self.h, _dp(q), float(r), _ip(out), cap)
        return out[:m]

    def draw_free(self, m):
        out = np.zeros((m, 3))
        lib().orc3_draw_free(self.h, m, _dp(out))
        return out

    def draw_informed(self, c_max, m):
        out = np.zeros((m, 3))
        lib().orc3_draw_informed(self.h, float(c_max), m, _dp(out))
        return out


def guidance_cloud_3d(oracle, rs, n_points=2048, over_sample_scale=5):
    """generate_rectangle_point_cloud_3d (datasets_3d/point_cloud_mask_utils_3d.py:83-113) restated
    with the oracle's predicates: uniform draws on `rs` (an np.random.RandomState standing for the
    global stream), obstacle filter with clearance 0, farthest-point down-sampling with open3d's
    semantics as stubbed in oracle/ref_shim.py (start index 0, f64, first argmax; parity unpinned)."""
    x1, y1, z1 = oracle.range6[1], oracle.range6[3], oracle.range6[5]
    pts = rs.uniform(low=(0, 0, 0), high=(x1, y1, z1), size=(n_points * over_sample_scale, 3))
    inside = np.zeros(len(pts), dtype=bool)
    for b in oracle.balls:
        inside |= (pts[:, 0] - b[0]) ** 2 + (pts[:, 1] - b[1]) ** 2 + (pts[:, 2] - b[2]) ** 2 < b[3] ** 2
    for b in oracle.boxes:
        inside |= ((b[0] <= pts[:, 0]) & (pts[:, 0] <= b[0] + b[3]) & (b[1] <= pts[:, 1]) & (pts[:, 1] <= b[1] + b[4]) &
                   (b[2] <= pts[:, 2]) & (pts[:, 2] <= b[2] + b[5]))
    pts = pts[~inside]
    if len(pts) > n_points:
        dist = np.full(len(pts), np.inf); far = 0; sel = []
        for _ in range(n_points):
            sel.append(far)
            dist = np.minimum(dist, ((pts - pts[far]) ** 2).sum(axis=1))
            far = int(np.argmax(dist))
        pts = pts[sel]
    return pts
