"""Utils(env, clearance) with the reference's four predicates (rrt_utils_2d.py:4-79), evaluated on
the GPU through the C ABI."""
import numpy as np

from path_planning_classes.collision_check_utils import _context, points_in_range


class Utils:
    def __init__(self, env, clearance):
        self.env = env
        self.clearance = clearance
        self.obs_circle = np.array(env.obs_circle) if len(env.obs_circle) > 0 else None
        self.obs_rectangle = np.array(env.obs_rectangle) if len(env.obs_rectangle) > 0 else None
        self.x_range, self.y_range = env.x_range, env.y_range
        self._ctx = _context(self.obs_circle, self.obs_rectangle, clearance,
                             (self.x_range[0], self.x_range[1], self.y_range[0], self.y_range[1]))

    def is_collision(self, start, end):
        line = np.array([start, end]).astype(np.float64).reshape(1, 2, 2)
        return bool(self._ctx.collide_edges(0, line)[0])

    def is_inside_obs(self, node):
        return bool(self._ctx.points_inside_obs(0, np.array([[node[0], node[1]]], dtype=np.float64))[0])

    def is_in_range(self, node):
        return points_in_range((node[0], node[1]), self.x_range, self.y_range, self.clearance)

    def is_valid(self, node):
        return bool(self._ctx.points_valid(0, np.array([[node[0], node[1]]], dtype=np.float64))[0])
