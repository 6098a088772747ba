"""NIRRTStarPNGC3D drop-in (reference: path_planning_classes_3d/nirrt_star_png_c_3d.py): NIRRTStarPNG3D whose guidance cloud comes from
Neural Connect (``png_wrapper.generate_connected_path_points``, up to ``connect_max_trial_attempts``
network calls joined by searches over the predicted points' r-disc graph) instead of a single
network call (path_planning_classes_3d/nirrt_star_png_c_3d.py:50-84)."""
import numpy as np

from path_planning_classes_3d.nirrt_star_png_3d import NIRRTStarPNG3D


class NIRRTStarPNGC3D(NIRRTStarPNG3D):
    def __init__(self, x_start, x_goal, step_len, search_radius, iter_max, env_dict, png_wrapper_connect, clearance,
                 pc_n_points, pc_over_sample_scale, pc_sample_rate, pc_update_cost_ratio, connect_max_trial_attempts):
        NIRRTStarPNG3D.__init__(self, x_start, x_goal, step_len, search_radius, iter_max, env_dict, png_wrapper_connect, clearance,
                 pc_n_points, pc_over_sample_scale, pc_sample_rate, pc_update_cost_ratio)
        self.path_planner_name = "NIRRT*-PNG(C) 3D"
        self.env_dict = env_dict
        self.connect_max_trial_attempts = connect_max_trial_attempts

    def _predict(self, pc):
        _, _, path_pred = self.png_wrapper.generate_connected_path_points(
            pc.astype(np.float32), self.x_start, self.x_goal, self.env_dict,
            neighbor_radius=self.pc_neighbor_radius, max_trial_attempts=self.connect_max_trial_attempts)
        return path_pred


def get_path_planner(args, problem, neural_wrapper):
    return NIRRTStarPNGC3D(problem['x_start'], problem['x_goal'], args.step_len, problem['search_radius'], args.iter_max,
                 problem['env_dict'], neural_wrapper, args.clearance, args.pc_n_points, args.pc_over_sample_scale,
                 args.pc_sample_rate, args.pc_update_cost_ratio, args.connect_max_trial_attempts)
