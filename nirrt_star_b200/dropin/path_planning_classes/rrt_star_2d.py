"""RRTStar2D drop-in (reference: path_planning_classes/rrt_star_2d.py).  Same constructor, drivers
and public state; the loop body runs as CUDA kernels (nirrt_batch_run on a 2D batch)."""
from nirrt_star_b200 import batch as _B
from nirrt_star_b200._lib import NirrtError
from path_planning_classes.rrt_base_2d import RRTBase2D
from path_planning_classes.rrt_visualizer_2d import RRTStarVisualizer


class RRTStar2D(RRTBase2D):
    _variant = _B.VARIANT_RRT_STAR

    def __init__(self, x_start, x_goal, step_len, search_radius, iter_max, env, clearance):
        super().__init__(x_start, x_goal, step_len, search_radius, iter_max, env, clearance, "RRT* 2D")
        self.visualizer = RRTStarVisualizer(self.x_start, self.x_goal, self.env)

    def _prepare(self, eng):
        """Hook for the neural variants (initial guidance cloud)."""

    def _cloud_callback(self):
        return None

    def _run(self, eng):
        try:
            eng.run_to_completion(chunk=min(512, max(1, self.iter_max)), cloud_callback=self._cloud_callback())
        except NirrtError as exc:
            if "empty predicted cloud" in str(exc):      # np.random.randint(0, 0) in SamplePointCloud
                raise ValueError("low >= high") from exc
            raise

    # planning(): rrt_star_2d.py:32-65 / irrt_star_2d.py:42-82
    def planning(self, visualize=False):
        eng = self._start_engine(8)
        self._prepare(eng)
        eng.begin(self._variant, _B.MODE_PLANNING, self.iter_max)
        self._run(eng)
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
        self._run(eng)
        self._finish_engine()
        if self._variant in _B.INFORMED:
            self.path_solutions = [int(i) for i in eng.solutions(0)]
        return eng.path_len_lists()[0]

    # rrt_star_2d.py:198-268 / irrt_star_2d.py:230-316
    def planning_random(self, iter_after_initial):
        return self._drive(iter_after_initial)

    # rrt_star_2d.py:159-196 / irrt_star_2d.py:180-228
    def planning_block_gap(self, path_len_threshold):
        return self._drive(0, stop_below=path_len_threshold)

    # rrt_star_2d.py:101-117, on the device tree
    def search_goal_parent(self):
        if self._engine is None:
            return None
        gp, _ = self._engine.goal_parents()
        return None if gp[0] < 0 else int(gp[0])


def get_path_planner(args, problem, neural_wrapper=None):
    return RRTStar2D(problem['x_start'], problem['x_goal'], args.step_len, problem['search_radius'],
                     args.iter_max, problem['env'], args.clearance)
