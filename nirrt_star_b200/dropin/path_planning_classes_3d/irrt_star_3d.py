"""IRRTStar3D drop-in (reference: path_planning_classes_3d/irrt_star_3d.py): informed sampling,
``path_solutions`` bookkeeping and best-solution refresh run on the device (k_top / k_expand)."""
from nirrt_star_b200 import batch as _B
from path_planning_classes_3d.rrt_base_3d import RRTBase3D
from path_planning_classes_3d.rrt_star_3d import RRTStar3D
from path_planning_classes_3d.rrt_visualizer_3d import IRRTStarVisualizer3D


class IRRTStar3D(RRTStar3D):
    _variant = _B.VARIANT_IRRT_STAR

    def __init__(self, x_start, x_goal, step_len, search_radius, iter_max, env, clearance):
        RRTBase3D.__init__(self, x_start, x_goal, step_len, search_radius, iter_max, env, clearance, "IRRT* 3D")
        self.path_solutions = []
        self.visualizer = IRRTStarVisualizer3D(self.x_start, self.x_goal, self.env)

    def find_best_path_solution(self):
        gp, cost = self._engine.goal_parents()
        return float(cost[0]), int(gp[0])


def get_path_planner(args, problem, neural_wrapper=None):
    return IRRTStar3D(problem['x_start'], problem['x_goal'], args.step_len, problem['search_radius'],
                      args.iter_max, problem['env'], args.clearance)
