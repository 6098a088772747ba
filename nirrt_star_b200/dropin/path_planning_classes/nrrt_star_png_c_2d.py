"""NRRTStarPNGC2D drop-in (reference: path_planning_classes/nrrt_star_png_c_2d.py): NRRTStarPNG2D whose guidance cloud comes from
Neural Connect (``png_wrapper.generate_connected_path_points``, up to ``connect_max_trial_attempts``
network calls joined by searches over the predicted points' r-disc graph) instead of a single
network call (path_planning_classes/nrrt_star_png_c_2d.py:60-79)."""
import numpy as np

from path_planning_classes.nrrt_star_png_2d import NRRTStarPNG2D


class NRRTStarPNGC2D(NRRTStarPNG2D):
    def __init__(self, x_start, x_goal, step_len, search_radius, iter_max, env_dict, png_wrapper_connect, binary_mask, clearance,
                 pc_n_points, pc_over_sample_scale, pc_sample_rate, connect_max_trial_attempts):
        NRRTStarPNG2D.__init__(self, x_start, x_goal, step_len, search_radius, iter_max, env_dict, png_wrapper_connect, binary_mask, clearance,
                 pc_n_points, pc_over_sample_scale, pc_sample_rate)
        self.path_planner_name = "NRRT*-PNG(C) 2D"
        self.env_dict = env_dict
        self.connect_max_trial_attempts = connect_max_trial_attempts

    def _predict(self, pc):
        _, _, path_pred = self.png_wrapper.generate_connected_path_points(
            pc.astype(np.float32), self.x_start, self.x_goal, self.env_dict,
            neighbor_radius=self.pc_neighbor_radius, max_trial_attempts=self.connect_max_trial_attempts)
        return path_pred


def get_path_planner(args, problem, neural_wrapper):
    return NRRTStarPNGC2D(problem['x_start'], problem['x_goal'], args.step_len, problem['search_radius'], args.iter_max,
                 problem['env_dict'], neural_wrapper, problem['binary_mask'], args.clearance, args.pc_n_points, args.pc_over_sample_scale,
                 args.pc_sample_rate, args.connect_max_trial_attempts)
