"""Env container with the reference's attribute contract (path_planning_utils_3d/rrt_env_3d.py:1-11)."""


class Env:
    def __init__(self, env_dict):
        dims = env_dict['env_dims']
        self.env_height, self.env_width, self.env_depth = dims
        self.x_range = (0, self.env_width)
        self.y_range = (0, self.env_height)
        self.z_range = (0, self.env_depth)
        self.obs_ball = env_dict['ball_obstacles']
        self.obs_box = env_dict['box_obstacles']

    def as_env_dict(self):
        return {'env_dims': [self.env_height, self.env_width, self.env_depth],
                'ball_obstacles': self.obs_ball, 'box_obstacles': self.obs_box}
