"""NRRTStarPNG2D drop-in (reference: path_planning_classes/nrrt_star_png_2d.py): RRT* drivers with
a guidance cloud predicted once before the loop; per-iteration switch between SamplePointCloud and
SampleFree (:52-56) on the device."""
import numpy as np

from nirrt_star_b200 import batch as _B
from path_planning_utils.rrt_env import Env
from path_planning_classes.rrt_base_2d import RRTBase2D
from path_planning_classes.rrt_star_2d import RRTStar2D
from path_planning_classes.rrt_visualizer_2d import NRRTStarPNGVisualizer
from datasets.point_cloud_mask_utils import get_point_cloud_mask_around_points, generate_rectangle_point_cloud


class NRRTStarPNG2D(RRTStar2D):
    _variant = _B.VARIANT_NRRT_STAR

    def __init__(self, x_start, x_goal, step_len, search_radius, iter_max, env_dict, png_wrapper, binary_mask, clearance,
                 pc_n_points, pc_over_sample_scale, pc_sample_rate):
        RRTBase2D.__init__(self, x_start, x_goal, step_len, search_radius, iter_max, Env(env_dict), clearance,
                           "NRRT*-PNG 2D")
        self.png_wrapper = png_wrapper
        self.binary_mask = binary_mask
        self.pc_n_points = pc_n_points
        self.pc_over_sample_scale = pc_over_sample_scale
        self.pc_sample_rate = pc_sample_rate
        self.pc_neighbor_radius = self.step_len
        self.path_point_cloud_pred = None
        self.visualizer = NRRTStarPNGVisualizer(self.x_start, self.x_goal, self.env)

    def _prepare(self, eng):
        eng.set_guidance(self.pc_sample_rate, 0.0)
        self.init_pc()
        self._sync_rng_to_device(eng)
        pc = self.path_point_cloud_pred
        eng.set_cloud(0, np.zeros((0, 2)) if pc is None else pc)

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
        """nrrt_star_png_2d.py:74-100"""
        if self.pc_sample_rate == 0:
            self.path_point_cloud_pred = None
            self.visualizer.set_path_point_cloud_pred(self.path_point_cloud_pred)
            return
        pc = generate_rectangle_point_cloud(self.binary_mask, self.pc_n_points, self.pc_over_sample_scale)
        path_pred = self._predict(pc)
        self.path_point_cloud_pred = pc[path_pred.nonzero()[0]]
        self.visualizer.set_path_point_cloud_pred(self.path_point_cloud_pred)


def get_path_planner(args, problem, neural_wrapper):
    return NRRTStarPNG2D(problem['x_start'], problem['x_goal'], args.step_len, problem['search_radius'],
                         args.iter_max, problem['env_dict'], neural_wrapper, problem['binary_mask'], args.clearance,
                         args.pc_n_points, args.pc_over_sample_scale, args.pc_sample_rate)
