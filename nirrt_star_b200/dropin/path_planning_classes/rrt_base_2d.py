"""RRTBase2D: tree storage and helpers with the reference's attribute contract
(path_planning_classes/rrt_base_2d.py:7-125).  The tree lives in HBM inside a one-problem
``BatchPlanner2D``; ``vertices`` / ``vertex_parents`` / ``num_vertices`` are host mirrors refreshed
whenever a driver returns.  Both process-global RNG streams the 2D planners consume (numpy's and
CPython's ``random``) are handed to the device and handed back."""
import random

import numpy as np

from nirrt_star_b200 import batch as _B
from path_planning_classes.rrt_utils_2d import Utils


class RRTBase2D:
    _variant = _B.VARIANT_RRT_STAR

    def __init__(self, x_start, x_goal, step_len, search_radius, iter_max, env, clearance, path_planner_name):
        self.x_start = np.array(x_start).astype(np.float64)
        self.x_goal = np.array(x_goal).astype(np.float64)
        self.step_len = step_len
        self.search_radius = search_radius
        self.iter_max = iter_max
        self.vertices = np.zeros((1 + iter_max, 2))
        self.vertex_parents = np.zeros(1 + iter_max).astype(int)
        self.vertices[0] = self.x_start
        self.num_vertices = 1
        self.path = []
        self.env = env
        self.utils = Utils(env, clearance)
        self.clearance = clearance
        self.x_range, self.y_range = env.x_range, env.y_range
        self.path_planner_name = path_planner_name
        self._engine = None

    # ---- engine plumbing -----------------------------------------------------------------------
    def _problem(self):
        return {"x_start": tuple(self.x_start), "x_goal": tuple(self.x_goal), "search_radius": self.search_radius,
                "env_dict": {"env_dims": [self.env.img_height, self.env.img_width],
                             "circle_obstacles": self.env.obs_circle, "rectangle_obstacles": self.env.obs_rectangle}}

    @staticmethod
    def _np_state():
        st = np.random.get_state()
        return (st[1], st[2])

    @staticmethod
    def _py_state():
        st = random.getstate()[1]
        return (np.array(st[:624], dtype=np.uint32), int(st[624]))

    def _start_engine(self, record_capacity):
        if self._engine is not None:
            raise RuntimeError("we can only run planning once per planner object (demo_planning_2d.py:90)")
        self._engine = _B.BatchPlanner2D([self._problem()], self.iter_max, step_len=self.step_len,
                                         clearance=self.clearance, rng_states=[self._np_state()],
                                         py_rng_states=[self._py_state()], record_capacity=record_capacity,
                                         near_capacity=_B.NEAR_CAPACITY_INFORMED if self._variant in _B.INFORMED else 0)
        return self._engine

    def _sync_rng_to_host(self, eng):
        key, pos = eng.get_rng()[0]
        np.random.set_state(("MT19937", key, pos, 0, 0.0))
        key, pos = eng.get_py_rng()[0]
        st = random.getstate()
        random.setstate((st[0], tuple(int(x) for x in key) + (pos,), st[2]))

    def _sync_rng_to_device(self, eng):
        eng.set_rng([self._np_state()])
        eng.set_py_rng([self._py_state()])

    def _finish_engine(self):
        eng = self._engine
        v, p, n = eng.read_trees()
        self.num_vertices = int(n[0])
        self.vertices[:] = v[0]
        self.vertex_parents[:] = p[0]
        self._sync_rng_to_host(eng)

    # ---- helpers with reference semantics ------------------------------------------------------
    def cost(self, vertex_index):
        if self._engine is None:
            return 0.
        return float(self._engine.costs(0, [int(vertex_index)])[0])

    def extract_path(self, goal_parent_index):
        path = [self.x_goal]
        i = goal_parent_index
        while i != 0:
            path.append(self.vertices[:self.num_vertices][i])
            i = self.vertex_parents[i]
        path.append(self.vertices[:self.num_vertices][i])
        path.reverse()
        return np.stack(path, axis=0)

    def check_success(self, path):
        if path is None or len(path) == 0:
            return False
        return np.all(path[0] == self.x_start) and np.all(path[-1] == self.x_goal)

    def get_path_len(self, path):
        if path is None or len(path) == 0:
            return np.inf
        path = np.array(path)
        return np.linalg.norm(path[1:] - path[:-1], axis=1).sum()

    def get_path_planner_name(self):
        return self.path_planner_name

    def visualize(self, *args, **kwargs):
        self.visualizer.animation()
