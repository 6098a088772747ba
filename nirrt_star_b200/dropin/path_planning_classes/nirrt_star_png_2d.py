"""NIRRTStarPNG2D drop-in (reference: path_planning_classes/nirrt_star_png_2d.py): see the 3D twin
(path_planning_classes_3d/nirrt_star_png_3d.py) for the device / host split."""
import numpy as np

from nirrt_star_b200 import batch as _B
from path_planning_utils.rrt_env import Env
from path_planning_classes.irrt_star_2d import IRRTStar2D
from path_planning_classes.rrt_base_2d import RRTBase2D
from path_planning_classes.rrt_visualizer_2d import NIRRTStarVisualizer
from datasets.point_cloud_mask_utils import get_point_cloud_mask_around_points, \
    generate_rectangle_point_cloud, ellipsoid_point_cloud_sampling


class NIRRTStarPNG2D(IRRTStar2D):
    _variant = _B.VARIANT_NIRRT_STAR

    def __init__(self, x_start, x_goal, step_len, search_radius, iter_max, env_dict, png_wrapper, binary_mask, clearance,
                 pc_n_points, pc_over_sample_scale, pc_sample_rate, pc_update_cost_ratio):
        RRTBase2D.__init__(self, x_start, x_goal, step_len, search_radius, iter_max, Env(env_dict), clearance,
                           "NIRRT*-PNG 2D")
        self.png_wrapper = png_wrapper
        self.binary_mask = binary_mask
        self.pc_n_points = pc_n_points
        self.pc_over_sample_scale = pc_over_sample_scale
        self.pc_sample_rate = pc_sample_rate
        self.pc_neighbor_radius = self.step_len
        self.pc_update_cost_ratio = pc_update_cost_ratio
        self.path_solutions = []
        self.path_point_cloud_pred = None
        self.visualizer = NIRRTStarVisualizer(self.x_start, self.x_goal, self.env)

    # ---- engine hooks --------------------------------------------------------------------------
    def _upload_cloud(self, eng):
        pc = self.path_point_cloud_pred
        eng.set_cloud(0, np.zeros((0, 2)) if pc is None else pc)

    def _prepare(self, eng):
        eng.set_guidance(self.pc_sample_rate, self.pc_update_cost_ratio)
        self.init_pc()                      # consumes the numpy stream before the loop (nirrt_star_png_2d.py:50-54)
        self._sync_rng_to_device(eng)
        self._upload_cloud(eng)

    def _cloud_callback(self):
        def cb(eng, env_indices):
            self._sync_rng_to_host(eng)
            c_best, c_min = eng.c_best()
            self.update_point_cloud(float(c_best[0]), float(c_min[0]))
            self._sync_rng_to_device(eng)
            self._upload_cloud(eng)
        return cb

    # ---- reference methods ---------------------------------------------------------------------
    def init_pc(self):
        self.update_point_cloud(cmax=np.inf, cmin=None)

    def SamplePointCloud(self):
        return self.path_point_cloud_pred[np.random.randint(0, len(self.path_point_cloud_pred))]

    def _sample_cloud(self, cmax, cmin):
        if cmax < np.inf:
            return ellipsoid_point_cloud_sampling(self.x_start, self.x_goal, cmax / cmin, self.binary_mask, self.pc_n_points,
                                                  n_raw_samples=self.pc_n_points * self.pc_over_sample_scale)
        return generate_rectangle_point_cloud(self.binary_mask, self.pc_n_points, self.pc_over_sample_scale)

    def _predict(self, pc):
        """one network call on start/goal neighbourhood masks; the (C) variants override this"""
        start_mask = get_point_cloud_mask_around_points(pc, self.x_start[np.newaxis, :], self.pc_neighbor_radius)
        goal_mask = get_point_cloud_mask_around_points(pc, self.x_goal[np.newaxis, :], self.pc_neighbor_radius)
        path_pred, path_score = self.png_wrapper.classify_path_points(
            pc.astype(np.float32), start_mask.astype(np.float32), goal_mask.astype(np.float32))
        return path_pred

    def update_point_cloud(self, cmax, cmin):
        """nirrt_star_png_2d.py:132-174"""
        if self.pc_sample_rate == 0:
            self.path_point_cloud_pred = None
            self.visualizer.set_path_point_cloud_pred(self.path_point_cloud_pred)
            return
        pc = self._sample_cloud(cmax, cmin)
        path_pred = self._predict(pc)
        self.path_point_cloud_pred = pc[path_pred.nonzero()[0]]
        self.visualizer.set_path_point_cloud_pred(self.path_point_cloud_pred)
        self.visualizer.set_path_point_cloud_other(pc[np.nonzero(path_pred == 0)[0]])


def get_path_planner(args, problem, neural_wrapper):
    return NIRRTStarPNG2D(problem['x_start'], problem['x_goal'], args.step_len, problem['search_radius'],
                          args.iter_max, problem['env_dict'], neural_wrapper, problem['binary_mask'], args.clearance,
                          args.pc_n_points, args.pc_over_sample_scale, args.pc_sample_rate, args.pc_update_cost_ratio)
