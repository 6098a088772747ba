"""RRTStar3D drop-in (reference: path_planning_classes_3d/rrt_star_3d.py).  Same constructor,
drivers and public state; the loop body runs as CUDA kernels (nirrt_batch_run)."""
import numpy as np

from nirrt_star_b200 import batch as _B
from path_planning_classes_3d.rrt_base_3d import RRTBase3D
from path_planning_classes_3d.rrt_visualizer_3d import RRTStarVisualizer3D


class RRTStar3D(RRTBase3D):
    _variant = _B.VARIANT_RRT_STAR

    def __init__(self, x_start, x_goal, step_len, search_radius, iter_max, env, clearance):
        super().__init__(x_start, x_goal, step_len, search_radius, iter_max, env, clearance, "RRT* 3D")
        self.visualizer = RRTStarVisualizer3D(self.x_start, self.x_goal, self.env)

    def _prepare(self, eng):
        """Hook for the neural variants (initial guidance cloud)."""

    def _cloud_callback(self):
        return None

    # planning(): rrt_star_3d.py:32-65 / irrt_star_3d.py:38-78
    def planning(self, visualize=False):
        eng = self._start_engine(8)
        self._prepare(eng)
        eng.begin(self._variant, _B.MODE_PLANNING, self.iter_max)
        eng.run_to_completion(chunk=min(512, max(1, self.iter_max)), cloud_callback=self._cloud_callback())
        gp, _ = eng.goal_parents()
        self._finish_engine()
        if self._variant in _B.INFORMED:
            self.path_solutions = [int(i) for i in eng.solutions(0)]
        self.path = self.extract_path(int(gp[0])) if gp[0] >= 0 else []
        if visualize:
            self.visualize()

    def _drive(self, iter_after_initial, stop_below=None):
        eng = self._start_engine(self.iter_max + iter_after_initial + 8)
        self._prepare(eng)
        eng.begin(self._variant, _B.MODE_PLANNING_RANDOM, self.iter_max, iter_after_initial)
        if stop_below is not None:
            eng.set_stop_threshold(stop_below)
        eng.run_to_completion(chunk=min(512, max(1, self.iter_max)), cloud_callback=self._cloud_callback())
        self._finish_engine()
        if self._variant in _B.INFORMED:
            self.path_solutions = [int(i) for i in eng.solutions(0)]
        return eng.path_len_lists()[0]

    # rrt_star_3d.py:200-270 / irrt_star_3d.py:245-331
    def planning_random(self, iter_after_initial):
        return self._drive(iter_after_initial)

    # rrt_star_3d.py:160-198 / irrt_star_3d.py:193-243
    def planning_block_gap(self, path_len_threshold):
        return self._drive(0, stop_below=path_len_threshold)

    # rrt_star_3d.py:101-117, on the mirrored tree
    def search_goal_parent(self):
        if self._engine is None:
            return None
        gp, _ = self._engine.goal_parents()
        return None if gp[0] < 0 else int(gp[0])


def get_path_planner(args, problem, neural_wrapper=None):
    return RRTStar3D(problem['x_start'], problem['x_goal'], args.step_len, problem['search_radius'],
                     args.iter_max, problem['env'], args.clearance)
