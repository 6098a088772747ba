"""Host-side driver of the lock-step batched 3D planner (C ABI: include/nirrt_b200.h).

``BatchPlanner3D`` owns one ``nirrt_batch`` handle: E independent planning problems whose trees
live in HBM and advance one loop body per iteration.  It mirrors, for a whole batch, what one
``RRTStar3D`` / ``IRRTStar3D`` / ``NIRRTStarPNG3D`` object does in the reference
(path_planning_classes_3d/*.py); the single-problem drop-in classes are thin views over a batch
of one (nirrt_star_b200/dropin/path_planning_classes_3d).

Host responsibilities (all one-off, per problem): evaluating the three libm-dependent constants
the reference computes with CPython/numpy -- the Near radius table (math.log / **), numpy's scalar
``r ** 2`` of each ball, and the informed-sampling rotation (numpy SVD) -- so the device only
executes IEEE-exact arithmetic.  Everything per-iteration runs in CUDA.
"""
import ctypes as C
import math

import numpy as np

from . import _lib
from ._lib import MAX_OBSTACLES, NirrtError, check, dp, f64, i64p, ip, u8p

VARIANT_RRT_STAR, VARIANT_IRRT_STAR, VARIANT_NIRRT_STAR, VARIANT_NRRT_STAR = 0, 1, 2, 3
INFORMED = (VARIANT_IRRT_STAR, VARIANT_NIRRT_STAR)
MODE_PLANNING, MODE_PLANNING_RANDOM = 0, 1
ST_DONE, ST_PHASE1, ST_PHASE2, ST_WAIT_CLOUD = 0, 1, 2, 3
NEAR_CAPACITY_INFORMED = 8192   # per-iteration Near candidate limit for the informed planners (maximum; default is 1024)


def near_radius_table(capacity, dim=3):
    """t[n] = (math.log(n)/n)**(1/3.) -- the n-dependent factor of find_near_neighbors' radius
    (rrt_star_3d.py:134), evaluated with CPython's libm exactly as the reference does."""
    t = np.zeros(capacity + 2, dtype=np.float64)
    if dim == 3:
        for n in range(1, capacity + 2):
            t[n] = (math.log(n) / n) ** (1 / 3.)
    else:
        for n in range(1, capacity + 2):
            t[n] = math.sqrt(math.log(n) / n)
    return t


def rotation_to_world_frame_3d(x_start, x_goal):
    """IRRTStar3D.RotationToWorldFrame (irrt_star_3d.py:159-173): one numpy SVD per problem."""
    x_start = np.array(x_start).astype(np.float64)
    x_goal = np.array(x_goal).astype(np.float64)
    dx, dy, dz = x_goal - x_start
    length = math.hypot(dx, dy, dz)
    a1 = (x_goal - x_start) / length
    U, _, V = np.linalg.svd(np.outer(a1, [1, 0, 0]))
    return U @ np.diag([1, 1, np.linalg.det(U) * np.linalg.det(V)]) @ V.T


def ellipsoid_params_3d(x_start, x_goal, max_min_ratio):
    """The scalars of ellipsoid_point_cloud_sampling_3d (datasets_3d/point_cloud_mask_utils_3d.py:132-160) that go
    through numpy / libm on the host -- c_min (np.linalg.norm), the rotation (SVD), c_max ** 2 (libm pow), np.sqrt --
    as the 12 numbers the device sampler takes: M = C @ L (row major) and the ellipsoid centre."""
    x_start = np.asarray(x_start, dtype=np.float64); x_goal = np.asarray(x_goal, dtype=np.float64)
    c_min = np.linalg.norm(x_goal - x_start)
    a1 = (x_goal - x_start) / c_min
    U, _, V = np.linalg.svd(np.outer(a1, [1, 0, 0]))
    Crot = U @ np.diag([1, 1, np.linalg.det(U) * np.linalg.det(V)]) @ V.T
    centre = (x_start + x_goal) / 2.
    c_max = c_min * max_min_ratio
    eps = 1e-6 if c_max ** 2 - c_min ** 2 < 0 else 0
    r = np.zeros(3)
    r[0] = c_max / 2
    r[1] = r[2] = np.sqrt(c_max ** 2 - c_min ** 2 + eps) / 2
    return np.concatenate([np.dot(Crot, np.diag(r)).reshape(9), centre])


def ellipsoid_params_2d(x_start, x_goal, max_min_ratio):
    """The host-side scalars of the 2D ellipsoid_point_cloud_sampling (datasets/point_cloud_mask_utils.py:104-133):
    math.hypot, the SVD rotation, c_max ** 2 (libm pow), math.sqrt -> M = C @ L (row major 3x3) and the centre (z = 0)."""
    x_start = np.asarray(x_start, dtype=np.float64); x_goal = np.asarray(x_goal, dtype=np.float64)
    dx, dy = x_goal - x_start
    c_min = math.hypot(dx, dy)
    a1 = np.concatenate([(x_goal - x_start) / c_min, np.array([0.])], axis=0)[:, np.newaxis]
    U, _, V_T = np.linalg.svd(a1 @ np.array([[1.0], [0.0], [0.0]]).T, True, True)
    Crot = U @ np.diag([1.0, 1.0, np.linalg.det(U) * np.linalg.det(V_T.T)]) @ V_T
    centre = np.concatenate([(x_start + x_goal) / 2., np.array([0.])], axis=0)
    c_max = c_min * max_min_ratio
    eps = 1e-6 if c_max ** 2 - c_min ** 2 < 0 else 0
    r = [c_max / 2.0, math.sqrt(c_max ** 2 - c_min ** 2 + eps) / 2.0, math.sqrt(c_max ** 2 - c_min ** 2 + eps) / 2.0]
    return np.concatenate([np.dot(Crot, np.diag(r)).reshape(9), centre])


def seed_state(seed):
    """(key[624] uint32, pos) of ``np.random.seed(seed)``."""
    st = np.random.RandomState(seed).get_state()
    return st[1].astype(np.uint32), int(st[2])


def rotation_to_world_frame_2d(x_start, x_goal):
    """IRRTStar2D.RotationToWorldFrame (irrt_star_2d.py:153-161): one numpy SVD per problem."""
    x_start = np.array(x_start).astype(np.float64)
    x_goal = np.array(x_goal).astype(np.float64)
    dx, dy = x_goal - x_start
    L = math.hypot(dx, dy)
    a1 = np.zeros((3, 1))
    a1[:2, 0] = (x_goal - x_start) / L
    e1 = np.array([[1.0], [0.0], [0.0]])
    M = a1 @ e1.T
    U, _, V_T = np.linalg.svd(M, True, True)
    return U @ np.diag([1.0, 1.0, np.linalg.det(U) * np.linalg.det(V_T.T)]) @ V_T


def py_seed_state(seed):
    """(key[624] uint32, pos) of ``random.seed(seed)`` (CPython's MT19937)."""
    import random
    st = random.Random(seed).getstate()[1]
    return np.array(st[:624], dtype=np.uint32), int(st[624])


class BatchPlanner3D:
    dim = 3

    def __init__(self, problems, iter_max, step_len=10, clearance=None, seeds=None, rng_states=None,
                 record_capacity=None, near_capacity=0, device=0, stream=None, py_rng_states=None):
        _lib.require_device()
        self.L = _lib.lib()
        self.E = len(problems)
        if self.E < 1:
            raise ValueError("at least one problem required")
        self.iter_max = int(iter_max)
        self.capacity = 1 + self.iter_max
        self.record_capacity = int(record_capacity) if record_capacity else self.capacity + 8
        self.stream = C.c_void_p(stream) if stream else None
        self.variant = VARIANT_RRT_STAR
        self.mode = MODE_PLANNING
        if clearance is None:
            clearance = 2 if self.dim == 3 else 3          # eval_planning_3d.py:76-77 / eval_planning_2d.py:80-81
        desc = _lib.BatchDesc(self.dim, self.E, self.capacity, self.record_capacity, int(near_capacity), int(device))
        h = C.c_void_p()
        check(self.L.nirrt_batch_create(C.byref(desc), C.byref(h)))
        self.h = h
        self._upload_problems(problems, step_len, clearance)
        if rng_states is None:
            if seeds is None:
                seeds = list(range(self.E))
            rng_states = [seed_state(s) for s in seeds]
            if self.dim == 2 and py_rng_states is None:
                py_rng_states = [py_seed_state(s) for s in seeds]
        self.set_rng(rng_states)
        if self.dim == 2 and py_rng_states is not None:
            self.set_py_rng(py_rng_states)

    # ------------------------------------------------------------------ setup
    def _upload_problems(self, problems, step_len, clearance):
        E = self.E
        start = np.zeros((E, 3)); goal = np.zeros((E, 3)); rng6 = np.zeros((E, 6))
        sl = np.zeros(E); sr = np.zeros(E); cl = np.zeros(E)
        nb = np.zeros(E, dtype=np.int32); nx = np.zeros(E, dtype=np.int32)
        balls = np.zeros((E, MAX_OBSTACLES, 4)); r2 = np.zeros((E, MAX_OBSTACLES)); boxes = np.zeros((E, MAX_OBSTACLES, 6))
        rot = np.zeros((E, 9))
        step_len = np.broadcast_to(np.asarray(step_len, dtype=np.float64), (E,))
        clearance = np.broadcast_to(np.asarray(clearance, dtype=np.float64), (E,))
        for e, p in enumerate(problems):
            ed = p["env_dict"]
            start[e] = np.array(p["x_start"]).astype(np.float64)
            goal[e] = np.array(p["x_goal"]).astype(np.float64)
            h, w, d = ed["env_dims"]                      # rrt_env_3d.py:6-9
            rng6[e] = [0, w, 0, h, 0, d]
            sl[e], sr[e], cl[e] = step_len[e], float(p["search_radius"]), clearance[e]
            b = np.asarray(ed["ball_obstacles"], dtype=np.float64).reshape(-1, 4)
            x = np.asarray(ed["box_obstacles"], dtype=np.float64).reshape(-1, 6)
            if len(b) > MAX_OBSTACLES or len(x) > MAX_OBSTACLES:
                raise ValueError(f"problem {e}: more than {MAX_OBSTACLES} obstacles of one type")
            nb[e], nx[e] = len(b), len(x)
            balls[e, :len(b)] = b
            boxes[e, :len(x)] = x
            # numpy *scalar* power, as check_collision_line_single_ball evaluates r ** 2
            # (collision_check_utils_3d.py:21,31-37)
            for k in range(len(b)):
                r2[e, k] = float((b[k, 3] + clearance[e]) ** 2)
            if np.any(start[e] != goal[e]):
                rot[e] = rotation_to_world_frame_3d(start[e], goal[e]).reshape(9)
            else:
                rot[e] = np.eye(3).reshape(9)
        self.start, self.goal = start, goal
        table = near_radius_table(self.capacity, 3)
        check(self.L.nirrt_batch_set_problems(self.h, dp(start), dp(goal), dp(sl), dp(sr), dp(cl), dp(rng6),
                                              ip(nb), dp(balls), dp(r2), ip(nx), dp(boxes), dp(table), dp(rot),
                                              self.stream))

    def set_rng(self, rng_states):
        key = np.zeros((self.E, 624), dtype=np.uint32); pos = np.zeros(self.E, dtype=np.int32)
        for e, (k, p) in enumerate(rng_states):
            key[e] = k; pos[e] = p
        check(self.L.nirrt_batch_set_rng(self.h, key.ctypes.data_as(_lib.c_u32p), ip(pos), self.stream))

    def set_py_rng(self, rng_states):
        """CPython ``random`` streams (2D informed sampling): [(key[624], pos)] per problem."""
        key = np.zeros((self.E, 624), dtype=np.uint32); pos = np.zeros(self.E, dtype=np.int32)
        for e, (k, p) in enumerate(rng_states):
            key[e] = k; pos[e] = p
        check(self.L.nirrt_batch_set_py_rng(self.h, key.ctypes.data_as(_lib.c_u32p), ip(pos), self.stream))

    def get_py_rng(self):
        key = np.zeros((self.E, 624), dtype=np.uint32); pos = np.zeros(self.E, dtype=np.int32)
        check(self.L.nirrt_batch_get_py_rng_sync(self.h, key.ctypes.data_as(_lib.c_u32p), ip(pos), self.stream))
        return [(key[e].copy(), int(pos[e])) for e in range(self.E)]

    def get_rng(self):
        key = np.zeros((self.E, 624), dtype=np.uint32); pos = np.zeros(self.E, dtype=np.int32)
        check(self.L.nirrt_batch_get_rng_sync(self.h, key.ctypes.data_as(_lib.c_u32p), ip(pos), self.stream))
        return [(key[e].copy(), int(pos[e])) for e in range(self.E)]

    def set_guidance(self, pc_sample_rate, pc_update_cost_ratio):
        check(self.L.nirrt_batch_set_guidance(self.h, float(pc_sample_rate), float(pc_update_cost_ratio)))

    def set_cloud(self, env, points):
        pts = f64(points).reshape(-1, self.dim)
        check(self.L.nirrt_batch_set_cloud(self.h, int(env), dp(pts), len(pts), self.stream))

    # ------------------------------------------------------------------ guidance clouds on the device (3D)
    def sample_clouds(self, envs, kinds, params, n_points, n_raw, radius, d_pc32, d_start_mask, d_goal_mask):
        """Device-side generate_rectangle_point_cloud_3d (kind 0) / ellipsoid_point_cloud_sampling_3d (kind 1) +
        start / goal masks for the listed problems (include/nirrt_b200.h: nirrt_batch_sample_clouds_sync).
        d_* are device pointers (ints) of f32 buffers [count][n_points][3] / [count][n_points].  Returns counts."""
        envs = np.ascontiguousarray(envs, dtype=np.int32); kinds = np.ascontiguousarray(kinds, dtype=np.int32)
        params = np.ascontiguousarray(params, dtype=np.float64).reshape(len(envs), 12)
        counts = np.zeros(len(envs), dtype=np.int32)
        check(self.L.nirrt_batch_sample_clouds_sync(self.h, ip(envs), len(envs), ip(kinds), dp(params), int(n_points), int(n_raw),
                                                    float(radius), C.c_void_p(d_pc32) if d_pc32 else None,
                                                    C.c_void_p(d_start_mask) if d_start_mask else None,
                                                    C.c_void_p(d_goal_mask) if d_goal_mask else None, ip(counts), self.stream))
        return counts

    def set_free_masks(self, masks):
        """2D: binary_mask of every problem (1 = free), all of one shape: needed by sample_clouds on a 2D batch."""
        m = np.ascontiguousarray(np.stack([np.asarray(x) != 0 for x in masks]), dtype=np.uint8)
        assert m.shape[0] == self.E and m.ndim == 3
        check(self.L.nirrt_batch_set_free_masks(self.h, u8p(m), m.shape[1], m.shape[2], self.stream))

    def read_sampled_clouds(self, first, count, n_points):
        out = np.zeros((count, n_points, 3))
        check(self.L.nirrt_batch_read_sampled_clouds_sync(self.h, int(first), int(count), dp(out), self.stream))
        return out if self.dim == 3 else np.ascontiguousarray(out[:, :, :2])

    def commit_clouds(self, d_pred, sel=None):
        """path_point_cloud_pred = cloud[pred != 0] for the clouds of the last sample_clouds call (all, or positions `sel`)."""
        if sel is None:
            check(self.L.nirrt_batch_commit_clouds(self.h, C.c_void_p(d_pred), None, 0, self.stream))
        else:
            sel = np.ascontiguousarray(sel, dtype=np.int32)
            check(self.L.nirrt_batch_commit_clouds(self.h, C.c_void_p(d_pred), ip(sel), len(sel), self.stream))

    def close(self):
        if getattr(self, "h", None):
            self.L.nirrt_batch_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ trees
    def load_trees(self, vertices, parents, n, env_begin=0):
        """vertices [count][capacity][3] f64, parents [count][capacity] int64, n [count]."""
        n = np.ascontiguousarray(n, dtype=np.int32)
        v = f64(vertices); p = np.ascontiguousarray(parents, dtype=np.int64)
        assert v.shape == (len(n), self.capacity, self.dim) and p.shape == (len(n), self.capacity)
        check(self.L.nirrt_batch_load_trees(self.h, int(env_begin), len(n), ip(n), dp(v), i64p(p), self.stream))

    def read_trees(self, env_begin=0, count=None, out=None):
        count = self.E - env_begin if count is None else count
        if out is None:
            v = np.zeros((count, self.capacity, self.dim)); p = np.zeros((count, self.capacity), dtype=np.int64)
        else:
            v, p = out
        n = np.zeros(count, dtype=np.int32)
        check(self.L.nirrt_batch_read_trees_sync(self.h, int(env_begin), count, ip(n), dp(v), i64p(p), self.stream))
        return v, p, n

    # ------------------------------------------------------------------ drivers
    def begin(self, variant, mode, iter_max=None, iter_after_initial=0):
        self.variant, self.mode = int(variant), int(mode)
        im = self.iter_max if iter_max is None else int(iter_max)
        check(self.L.nirrt_batch_begin(self.h, self.variant, self.mode, im, int(iter_after_initial), self.stream))

    def run(self, iters):
        check(self.L.nirrt_batch_run(self.h, int(iters), self.stream))

    def set_stop_threshold(self, stop_below):
        check(self.L.nirrt_batch_set_stop_threshold(self.h, float(stop_below)))

    def set_vertex_limit(self, limit):
        check(self.L.nirrt_batch_set_vertex_limit(self.h, int(limit)))

    def run_profiled(self, iters):
        """Like run(), synchronous, with the kernels launched unfused and serialised: returns the summed device ms of
        k_top / Nearest scan / k_steer / Near scan (0 with a mirror scan: Near is collected by the Nearest pass) / k_expand."""
        ms = (C.c_float * 5)()
        check(self.L.nirrt_batch_run_profiled_sync(self.h, int(iters), ms, self.stream))
        return dict(zip(("top", "nearest", "steer", "near", "expand"), [float(x) for x in ms]))

    def status(self):
        running = C.c_int(0); need = C.c_int(0)
        check(self.L.nirrt_batch_status_sync(self.h, C.byref(running), C.byref(need), self.stream))
        return running.value, need.value

    def env_state(self):
        st = np.zeros(self.E, dtype=np.int32); nr = np.zeros(self.E, dtype=np.int32); nv = np.zeros(self.E, dtype=np.int32)
        check(self.L.nirrt_batch_env_state_sync(self.h, ip(st), ip(nr), ip(nv), self.stream))
        return st, nr, nv

    def run_to_completion(self, chunk=256, cloud_callback=None):
        """Runs until every problem's driver finished.  ``cloud_callback(batch, env_indices)`` is
        invoked for problems that paused for a guidance-cloud update (NIRRT*)."""
        while True:
            self.run(chunk)
            running, need = self.status()
            if need:
                if cloud_callback is None:
                    raise NirrtError("a problem requested a guidance cloud but no cloud_callback was given")
                st, _, _ = self.env_state()
                cloud_callback(self, np.nonzero(st == ST_WAIT_CLOUD)[0])
                continue
            if running == 0:
                return

    def records(self, env_begin=0, count=None):
        count = self.E - env_begin if count is None else count
        rec = np.zeros((count, self.record_capacity)); n = np.zeros(count, dtype=np.int32)
        check(self.L.nirrt_batch_read_records_sync(self.h, int(env_begin), count, dp(rec), ip(n), self.stream))
        return [rec[k, :n[k]].copy() for k in range(count)]

    def path_len_lists(self):
        """path_len_list of every problem exactly as planning_random returns it: the RRT* family
        records after each iteration; the IRRT* family drops the pre-loop entry
        (irrt_star_3d.py:285)."""
        recs = self.records()
        if self.variant not in INFORMED:
            return [list(r) for r in recs]
        return [list(r[1:]) for r in recs]

    def solutions(self, env):
        n = check(self.L.nirrt_batch_read_solutions_sync(self.h, int(env), None, 0, self.stream))
        out = np.zeros(max(n, 1), dtype=np.int64)
        check(self.L.nirrt_batch_read_solutions_sync(self.h, int(env), i64p(out), len(out), self.stream))
        return out[:n]

    def c_best(self):
        """(c_best[E], c_min[E]) as of the top of the last executed iteration."""
        cb = np.zeros(self.E); cm = np.zeros(self.E)
        check(self.L.nirrt_batch_read_cbest_sync(self.h, dp(cb), dp(cm), self.stream))
        return cb, cm

    def goal_parents(self):
        gp = np.zeros(self.E, dtype=np.int64); cost = np.zeros(self.E)
        check(self.L.nirrt_batch_goal_parent_sync(self.h, i64p(gp), dp(cost), self.stream))
        return gp, cost

    def trace(self, near_stride=2048):
        E = self.E
        nearest = np.zeros(E, dtype=np.int32); new = np.zeros(E, dtype=np.int32); cnt = np.zeros(E, dtype=np.int32)
        near = np.zeros((E, near_stride), dtype=np.int32); xr = np.zeros((E, self.dim))
        check(self.L.nirrt_batch_read_trace_sync(self.h, ip(nearest), ip(new), ip(cnt), ip(near), near_stride, dp(xr), self.stream))
        return nearest, new, cnt, near, xr

    # ------------------------------------------------------------------ stand-alone predicates
    def collide_edges(self, env, edges):
        e = f64(edges).reshape(-1, 2 * self.dim)
        out = np.zeros(len(e), dtype=np.uint8)
        check(self.L.nirrt_collide_edges_sync(self.h, int(env), dp(e), len(e), u8p(out), self.stream))
        return out.astype(bool)

    def points_inside_obs(self, env, pts):
        p = f64(pts).reshape(-1, self.dim); out = np.zeros(len(p), dtype=np.uint8)
        check(self.L.nirrt_points_check_sync(self.h, int(env), 0, dp(p), len(p), u8p(out), self.stream))
        return out.astype(bool)

    def points_valid(self, env, pts):
        p = f64(pts).reshape(-1, self.dim); out = np.zeros(len(p), dtype=np.uint8)
        check(self.L.nirrt_points_check_sync(self.h, int(env), 1, dp(p), len(p), u8p(out), self.stream))
        return out.astype(bool)

    def nearest(self, env, queries):
        q = f64(queries).reshape(-1, self.dim); out = np.zeros(len(q), dtype=np.int64)
        check(self.L.nirrt_nearest_sync(self.h, int(env), dp(q), len(q), i64p(out), self.stream))
        return out

    def within(self, env, q, r, cap=2048):
        q = f64(q).reshape(self.dim); out = np.zeros(cap, dtype=np.int64)
        m = check(self.L.nirrt_within_sync(self.h, int(env), dp(q), float(r), i64p(out), cap, self.stream))
        return out[:min(m, cap)]

    def costs(self, env, idx):
        idx = np.ascontiguousarray(idx, dtype=np.int64); out = np.zeros(len(idx))
        check(self.L.nirrt_costs_sync(self.h, int(env), i64p(idx), len(idx), dp(out), self.stream))
        return out

    def kernel_launches(self):
        n = C.c_int64(0)
        check(self.L.nirrt_batch_counters(self.h, C.byref(n), None))
        return n.value

    def graph_stats(self):
        """{builds, replays, fallbacks} of the CUDA-graph replay path (a fallback = capture failed, plain launches)."""
        a, b, c = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        check(self.L.nirrt_batch_graph_stats(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return {"builds": a.value, "replays": b.value, "fallbacks": c.value}

    def work_stats(self):
        out = np.zeros(8, dtype=np.int64)
        check(self.L.nirrt_batch_work_stats_sync(self.h, i64p(out), self.stream))
        keys = ("expansions", "near_sum", "candidate_sum", "track_rounds", "full_goal_evals", "refreshed_candidates",
                "reparented_roots", "full_eval_list_sum")
        return dict(zip(keys, [int(x) for x in out]))

    def scan_bytes_per_vertex(self):
        """bytes one scan pass reads per vertex: 2*dim with the u16 mirror (default), 4 with the u8 mirror,
        4*dim with the f32 mirror, 8*dim without a mirror"""
        n = C.c_int64(0); b = C.c_int64(0)
        check(self.L.nirrt_batch_counters(self.h, C.byref(n), C.byref(b)))
        return b.value

    def time_scan(self, which, reps=20):
        ms = C.c_float(0); nbytes = C.c_int64(0)
        check(self.L.nirrt_batch_time_scan_sync(self.h, int(which), int(reps), C.byref(ms), C.byref(nbytes), self.stream))
        return ms.value, nbytes.value


class BatchPlanner2D(BatchPlanner3D):
    """E independent 2D planning problems (RRTStar2D / IRRTStar2D / N(I)RRTStarPNG2D loop bodies,
    path_planning_classes/*.py).  ``problems[i]['env_dict']`` follows the reference's 2D json schema
    (rrt_env.py:1-20): env_dims (height, width), circle_obstacles [x,y,r], rectangle_obstacles [x,y,w,h]."""
    dim = 2

    def _upload_problems(self, problems, step_len, clearance):
        E = self.E
        start = np.zeros((E, 2)); goal = np.zeros((E, 2)); rng4 = np.zeros((E, 4))
        sl = np.zeros(E); sr = np.zeros(E); cl = np.zeros(E)
        nc = np.zeros(E, dtype=np.int32); nr = np.zeros(E, dtype=np.int32)
        circles = np.zeros((E, MAX_OBSTACLES, 3)); rects = np.zeros((E, MAX_OBSTACLES, 4))
        rot = np.zeros((E, 9))
        step_len = np.broadcast_to(np.asarray(step_len, dtype=np.float64), (E,))
        clearance = np.broadcast_to(np.asarray(clearance, dtype=np.float64), (E,))
        for e, p in enumerate(problems):
            ed = p["env_dict"]
            start[e] = np.array(p["x_start"]).astype(np.float64)
            goal[e] = np.array(p["x_goal"]).astype(np.float64)
            h, w = ed["env_dims"]                         # rrt_env.py:6-8
            rng4[e] = [0, w, 0, h]
            sl[e], sr[e], cl[e] = step_len[e], float(p["search_radius"]), clearance[e]
            c = np.asarray(ed["circle_obstacles"], dtype=np.float64).reshape(-1, 3)
            r = np.asarray(ed["rectangle_obstacles"], dtype=np.float64).reshape(-1, 4)
            if len(c) > MAX_OBSTACLES or len(r) > MAX_OBSTACLES:
                raise ValueError(f"problem {e}: more than {MAX_OBSTACLES} obstacles of one type")
            nc[e], nr[e] = len(c), len(r)
            circles[e, :len(c)] = c
            rects[e, :len(r)] = r
            rot[e] = (rotation_to_world_frame_2d(start[e], goal[e]) if np.any(start[e] != goal[e]) else np.eye(3)).reshape(9)
        self.start, self.goal = start, goal
        table = near_radius_table(self.capacity, 2)
        check(self.L.nirrt_batch_set_problems_2d(self.h, dp(start), dp(goal), dp(sl), dp(sr), dp(cl), dp(rng4),
                                                 ip(nc), dp(circles), ip(nr), dp(rects), dp(table), dp(rot), self.stream))


def sincos(x, stream=None):
    """(math.sin(x), math.cos(x)) element-wise on the device, bit-identical to the reference runtime's libm."""
    _lib.require_device()
    x = f64(x).reshape(-1)
    s = np.empty_like(x); c = np.empty_like(x)
    check(_lib.lib().nirrt_sincos_sync(dp(x), len(x), dp(s), dp(c), C.c_void_p(stream) if stream else None))
    return s, c


def atan2(y, x, stream=None):
    """math.atan2(y, x) element-wise on the device, bit-identical to the reference runtime's libm."""
    _lib.require_device()
    y = f64(y).reshape(-1); x = f64(x).reshape(-1)
    out = np.empty_like(y)
    check(_lib.lib().nirrt_atan2_sync(dp(y), dp(x), len(y), dp(out), C.c_void_p(stream) if stream else None))
    return out


def fps_f64(points, npoint, start=0, stream=None):
    """Indices of the farthest-point down-sampling of an (n,3) f64 point set (open3d semantics:
    start index 0, squared distances in f64, first argmax)."""
    _lib.require_device()
    pts = f64(points).reshape(-1, 3)
    out = np.zeros(int(npoint), dtype=np.int64)
    check(_lib.lib().nirrt_fps_f64_sync(dp(pts), len(pts), int(npoint), int(start), i64p(out),
                                        C.c_void_p(stream) if stream else None))
    return out
