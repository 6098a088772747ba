"""RRTBase3D: tree storage and helpers with the reference's attribute contract
(path_planning_classes_3d/rrt_base_3d.py:7-137).  The tree itself lives in HBM inside a
one-problem ``BatchPlanner3D``; ``vertices`` / ``vertex_parents`` / ``num_vertices`` are host
mirrors refreshed whenever a driver returns."""
import numpy as np

from nirrt_star_b200 import batch as _B
from path_planning_classes_3d.rrt_utils_3d import Utils


class RRTBase3D:
    _variant = _B.VARIANT_RRT_STAR

    def __init__(self, x_start, x_goal, step_len, search_radius, iter_max, env, clearance, path_planner_name):
        self.x_start = np.array(x_start).astype(np.float64)
        self.x_goal = np.array(x_goal).astype(np.float64)
        self.step_len = step_len
        self.search_radius = search_radius
        self.iter_max = iter_max
        self.vertices = np.zeros((1 + iter_max, 3))
        self.vertex_parents = np.zeros(1 + iter_max).astype(int)
        self.vertices[0] = self.x_start
        self.num_vertices = 1
        self.path = []
        self.env = env
        self.utils = Utils(env, clearance)
        self.clearance = clearance
        self.x_range, self.y_range, self.z_range = env.x_range, env.y_range, env.z_range
        self.path_planner_name = path_planner_name
        self._engine = None

    # ---- engine plumbing -----------------------------------------------------------------------
    def _problem(self):
        return {"x_start": tuple(self.x_start), "x_goal": tuple(self.x_goal), "search_radius": self.search_radius,
                "env_dict": {"env_dims": [self.env.env_height, self.env.env_width, self.env.env_depth],
                             "ball_obstacles": self.env.obs_ball, "box_obstacles": self.env.obs_box}}

    def _start_engine(self, record_capacity):
        """Creates the device tree and hands it the process-global numpy RNG stream the reference
        would consume (np.random.* in SampleFree etc.)."""
        if self._engine is not None:
            raise RuntimeError("we can only run planning once per planner object (demo_planning_3d.py)")
        st = np.random.get_state()
        self._engine = _B.BatchPlanner3D([self._problem()], self.iter_max, step_len=self.step_len,
                                         clearance=self.clearance, rng_states=[(st[1], st[2])],
                                         record_capacity=record_capacity,
                                         near_capacity=_B.NEAR_CAPACITY_INFORMED if self._variant in _B.INFORMED else 0)
        return self._engine

    def _finish_engine(self):
        """Mirrors the device tree into the public numpy attributes and gives the advanced RNG
        stream back to np.random."""
        eng = self._engine
        v, p, n = eng.read_trees()
        self.num_vertices = int(n[0])
        self.vertices[:] = v[0]
        self.vertex_parents[:] = p[0]
        key, pos = eng.get_rng()[0]
        np.random.set_state(("MT19937", key, pos, 0, 0.0))

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
