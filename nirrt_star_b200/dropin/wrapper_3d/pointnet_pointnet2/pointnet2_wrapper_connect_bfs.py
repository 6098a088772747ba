"""PNGWrapper with Neural Connect, drop-in for
wrapper_3d/pointnet_pointnet2/pointnet2_wrapper_connect_bfs.py (and, through the subclass in
wrapper/, for the 2D file): ``classify_path_points`` as in pointnet2_wrapper.py plus
``generate_connected_path_points`` (:66-233): up to ``max_trial_attempts`` network calls, each
followed by a start->goal and a goal->start search over the r-disc graph of the predicted points;
the masks of the next call are re-centred on the best boundary point of each side."""
import numpy as np

from datasets.point_cloud_mask_utils import get_point_cloud_mask_around_points
from nirrt_star_b200.pointnet2 import connect_analyse
from wrapper.utils.bfs_connect_heuristic import select_heuristic_boundary_point
from wrapper_3d.pointnet_pointnet2.pointnet2_wrapper import PNGWrapper as _PNGWrapper


class PNGWrapper(_PNGWrapper):
    _banner = "PointNet++ wrapper 3d with connect is initialized."

    def generate_connected_path_points(self, pc, x_start, x_goal, env_dict, neighbor_radius, max_trial_attempts,
                                       visualize=False, vis_folderpath="", token=""):
        """-> (connection_success: bool, num_png_runs: int, path_pred_mask: float32 [n_points])"""
        if visualize:
            raise NotImplementedError("nirrt_star_b200 does not ship the matplotlib visualisers")
        has_path = False
        path_pred_mask = np.zeros(len(pc)).astype(np.float32)
        start_mask = get_point_cloud_mask_around_points(pc, x_start[np.newaxis].astype(np.float32), neighbor_radius)
        goal_mask = get_point_cloud_mask_around_points(pc, x_goal[np.newaxis].astype(np.float32), neighbor_radius)
        xs, xg = x_start.astype(np.float32), x_goal.astype(np.float32)
        trial_i = -1
        for trial_i in range(max_trial_attempts):
            path_pred, path_score = self.classify_path_points(pc, start_mask, goal_mask)
            path_pred_mask = ((path_pred_mask + path_pred) > 0).astype(np.float32)
            has_path, _, boundary_mask = connect_analyse(pc, path_pred_mask, xs, xg, neighbor_radius)
            if has_path:
                break
            _, boundary_point, _ = select_heuristic_boundary_point(pc, boundary_mask, xs, xg)
            next_start_mask = start_mask if boundary_point is None else \
                get_point_cloud_mask_around_points(pc, boundary_point, neighbor_radius)
            has_path, _, boundary_mask = connect_analyse(pc, path_pred_mask, xg, xs, neighbor_radius)
            if has_path:
                break
            _, boundary_point, _ = select_heuristic_boundary_point(pc, boundary_mask, xg, xs)
            next_goal_mask = goal_mask if boundary_point is None else \
                get_point_cloud_mask_around_points(pc, boundary_point, neighbor_radius)
            start_mask, goal_mask = next_start_mask, next_goal_mask
        return has_path, trial_i + 1, path_pred_mask
