"""Utils(env, clearance) with the reference's four predicates (rrt_utils_3d.py:5-86), evaluated on
the GPU through the C ABI."""
import numpy as np

from path_planning_classes_3d.collision_check_utils_3d import _context


class Utils:
    def __init__(self, env, clearance):
        self.env = env
        self.clearance = clearance
        self.obs_ball = np.array(env.obs_ball).astype(np.float64) if len(env.obs_ball) > 0 else None
        self.obs_box = np.array(env.obs_box).astype(np.float64) if len(env.obs_box) > 0 else None
        self.x_range, self.y_range, self.z_range = env.x_range, env.y_range, env.z_range
        self._ctx = _context(self.obs_ball, self.obs_box, clearance,
                             (self.x_range[0], self.x_range[1], self.y_range[0], self.y_range[1],
                              self.z_range[0], self.z_range[1]))

    def is_collision(self, start, end):
        line = np.array([start, end]).astype(np.float64).reshape(1, 2, 3)
        return bool(self._ctx.collide_edges(0, line)[0])

    def is_inside_obs(self, node):
        return bool(self._ctx.points_inside_obs(0, np.array([[node[0], node[1], node[2]]], dtype=np.float64))[0])

    def is_in_range(self, node):
        from path_planning_classes_3d.collision_check_utils_3d import points_in_range_3d
        return points_in_range_3d((node[0], node[1], node[2]), self.x_range, self.y_range, self.z_range, self.clearance)

    def is_valid(self, node):
        return bool(self._ctx.points_valid(0, np.array([[node[0], node[1], node[2]]], dtype=np.float64))[0])
