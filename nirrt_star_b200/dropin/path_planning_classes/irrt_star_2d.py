"""IRRTStar2D drop-in (reference: path_planning_classes/irrt_star_2d.py): informed sampling (unit
disc from the CPython ``random`` stream), ``path_solutions`` bookkeeping and best-solution refresh
run on the device (k_top / k_expand)."""
from nirrt_star_b200 import batch as _B
from path_planning_classes.rrt_base_2d import RRTBase2D
from path_planning_classes.rrt_star_2d import RRTStar2D
from path_planning_classes.rrt_visualizer_2d import IRRTStarVisualizer


class IRRTStar2D(RRTStar2D):
    _variant = _B.VARIANT_IRRT_STAR

    def __init__(self, x_start, x_goal, step_len, search_radius, iter_max, env, clearance):
        RRTBase2D.__init__(self, x_start, x_goal, step_len, search_radius, iter_max, env, clearance, "IRRT* 2D")
        self.path_solutions = []
        self.visualizer = IRRTStarVisualizer(self.x_start, self.x_goal, self.env)

    def find_best_path_solution(self):
        gp, cost = self._engine.goal_parents()
        return float(cost[0]), int(gp[0])


def get_path_planner(args, problem, neural_wrapper=None):
    return IRRTStar2D(problem['x_start'], problem['x_goal'], args.step_len, problem['search_radius'],
                      args.iter_max, problem['env'], args.clearance)
