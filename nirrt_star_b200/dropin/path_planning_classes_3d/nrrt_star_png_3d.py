"""NRRTStarPNG3D drop-in (reference: path_planning_classes_3d/nrrt_star_png_3d.py): RRT* drivers
with a guidance cloud predicted once before the loop; per-iteration switch between SamplePointCloud
and SampleFree (:52-59) on the device."""
import numpy as np

from nirrt_star_b200 import batch as _B
from path_planning_utils_3d.rrt_env_3d import Env
from path_planning_classes_3d.rrt_base_3d import RRTBase3D
from path_planning_classes_3d.rrt_star_3d import RRTStar3D
from path_planning_classes_3d.rrt_visualizer_3d import NRRTStarPNGVisualizer3D
from datasets.point_cloud_mask_utils import get_point_cloud_mask_around_points
from datasets_3d.point_cloud_mask_utils_3d import generate_rectangle_point_cloud_3d


class NRRTStarPNG3D(RRTStar3D):
    _variant = _B.VARIANT_NRRT_STAR

    def __init__(self, x_start, x_goal, step_len, search_radius, iter_max, env_dict, png_wrapper, clearance,
                 pc_n_points, pc_over_sample_scale, pc_sample_rate):
        RRTBase3D.__init__(self, x_start, x_goal, step_len, search_radius, iter_max, Env(env_dict), clearance,
                           "NRRT*-PNG 3D")
        self.png_wrapper = png_wrapper
        self.pc_n_points = pc_n_points
        self.pc_over_sample_scale = pc_over_sample_scale
        self.pc_sample_rate = pc_sample_rate
        self.pc_neighbor_radius = self.step_len
        self.path_point_cloud_pred = None
        self.visualizer = NRRTStarPNGVisualizer3D(self.x_start, self.x_goal, self.env)

    def _prepare(self, eng):
        eng.set_guidance(self.pc_sample_rate, 0.0)
        self.init_pc()
        st = np.random.get_state()
        eng.set_rng([(st[1], st[2])])
        pc = self.path_point_cloud_pred
        eng.set_cloud(0, np.zeros((0, 3)) if pc is None else pc)

    def init_pc(self):
        self.update_point_cloud()

    def SamplePointCloud(self):
        return self.path_point_cloud_pred[np.random.randint(0, len(self.path_point_cloud_pred))]

    def _predict(self, pc):
        """one network call on start/goal neighbourhood masks; the (C) variants override this"""
        start_mask = get_point_cloud_mask_around_points(pc, self.x_start[np.newaxis, :], self.pc_neighbor_radius)
        goal_mask = get_point_cloud_mask_around_points(pc, self.x_goal[np.newaxis, :], self.pc_neighbor_radius)
        path_pred, path_score = self.png_wrapper.classify_path_points(
            pc.astype(np.float32), start_mask.astype(np.float32), goal_mask.astype(np.float32))
        return path_pred

    def update_point_cloud(self):
        """nrrt_star_png_3d.py:74-100"""
        if self.pc_sample_rate == 0:
            self.path_point_cloud_pred = None
            self.visualizer.set_path_point_cloud_pred(self.path_point_cloud_pred)
            return
        pc = generate_rectangle_point_cloud_3d(self.env, self.pc_n_points, over_sample_scale=self.pc_over_sample_scale)
        path_pred = self._predict(pc)
        self.path_point_cloud_pred = pc[path_pred.nonzero()[0]]
        self.visualizer.set_path_point_cloud_pred(self.path_point_cloud_pred)


def get_path_planner(args, problem, neural_wrapper):
    return NRRTStarPNG3D(problem['x_start'], problem['x_goal'], args.step_len, problem['search_radius'],
                         args.iter_max, problem['env_dict'], neural_wrapper, args.clearance, args.pc_n_points,
                         args.pc_over_sample_scale, args.pc_sample_rate)
