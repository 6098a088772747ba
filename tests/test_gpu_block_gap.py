"""planning_block_gap (rrt_star_2d.py:159-196, irrt_star_2d.py:180-228) of the 2D drop-in planners on the reference's
block / gap problems: against traces recorded from the reference's own classes
(tests/golden/make_golden_block_gap.py) and against the analytic known answers the reference's eval uses as stopping
thresholds (generate_block_gap_env_2d.py:19-20,35 -> eval_planning_2d.py:117-121)."""
import glob
import os
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "blockgap2d_*.npz")))


@pytest.fixture(scope="module", autouse=True)
def dropin():
    from nirrt_star_b200 import dropin
    dropin.install()


def _problem(env_dims, rects, x_start, x_goal, search_radius):
    from path_planning_utils.rrt_env import Env
    env_dict = {"env_dims": [int(env_dims[0]), int(env_dims[1])], "rectangle_obstacles": [list(map(float, r)) for r in rects],
                "circle_obstacles": [], "start": [tuple(x_start)], "goal": [tuple(x_goal)]}
    return {"x_start": tuple(x_start), "x_goal": tuple(x_goal), "env_dict": env_dict, "env": Env(env_dict),
            "search_radius": float(search_radius)}


def _planner(kind, problem, iter_max):
    import importlib
    import types
    mod = importlib.import_module("path_planning_classes." + {"rrt": "rrt_star_2d", "irrt": "irrt_star_2d"}[kind])
    args = types.SimpleNamespace(step_len=10.0, iter_max=iter_max, clearance=0.0)      # eval_planning_2d.py:16-18 defaults
    return mod.get_path_planner(args, problem, None)


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_planning_block_gap_matches_reference_golden(path):
    g = np.load(path)
    kind, iter_max, seed = str(g["kind"]), int(g["iter_max"]), int(g["seed"])
    problem = _problem(g["env_dims"], g["rects"], g["x_start"], g["x_goal"], g["search_radius"])
    np.random.seed(seed); random.seed(seed)
    planner = _planner(kind, problem, iter_max)
    lst = planner.planning_block_gap(float(g["threshold"]))
    want = g["path_len_list"]
    assert isinstance(lst, list) and len(lst) == len(want)           # stops on the same iteration
    assert np.array_equal(np.isinf(lst), np.isinf(want))
    f = np.isfinite(want)
    assert np.allclose(np.array(lst)[f], want[f], rtol=1e-5, atol=0)
    assert (lst[-1] < float(g["threshold"])) == (want[-1] < float(g["threshold"]))
    n = planner.num_vertices
    assert n == int(g["num_vertices"])
    assert np.array_equal(planner.vertex_parents[:n], g["parents"])
    assert np.array_equal(planner.vertices[:n], g["vertices"])
    if kind == "irrt":
        assert list(planner.path_solutions) == list(g["solutions"])
    assert np.random.random() == float(g["next_random"])
    assert random.random() == float(g["next_py_random"])


def _block(w, ratio, d_goal=60):
    """get_block_problem_input (datasets/planning_problem_utils_2d.py:49-90) without the image: the obstacle list,
    start / goal and gamma_RRT* from the free PIXEL count (compute_gamma_rrt_star, :164-172: cv2.rectangle fills
    the closed pixel rectangle [x, x+w] x [y, y+h])."""
    import math
    H = W = d_goal * ratio
    x, y = W // 2 - w // 2, H // 2 - w // 2
    free = H * W - (w + 1) * (w + 1)
    gamma = math.ceil((2 * (1 + 1. / 2)) ** (1. / 2) * (free / np.pi) ** (1. / 2))
    best = w + (((d_goal - w) // 2) ** 2 + (w // 2) ** 2) ** 0.5 + (((d_goal - w) - (d_goal - w) // 2) ** 2 + (w // 2) ** 2) ** 0.5
    return _problem((H, W), [[x, y, w, w]], (W // 2 - d_goal // 2, H // 2), (W // 2 + d_goal // 2, H // 2), gamma), best


@pytest.mark.parametrize("kind,w,ratio,seed", [("irrt", 20, 2, 101), ("irrt", 40, 3, 102), ("rrt", 16, 2, 103)])
def test_block_analytic_optimum(kind, w, ratio, seed):
    """The block world's optimal path length is known in closed form (generate_block_gap_env_2d.py:19-20): with
    clearance 0 every path the planner reports is at least that long, and the eval's stopping rule
    best_path_len * 1.02 (eval_planning_2d.py:117-119) is reached well inside the iteration budget."""
    problem, best = _block(w, ratio)
    np.random.seed(seed); random.seed(seed)
    planner = _planner(kind, problem, 30000)
    lst = np.array(planner.planning_block_gap(best * 1.02))
    f = np.isfinite(lst)
    assert f.any() and lst[f].min() >= best - 1e-9
    assert lst[-1] < best * 1.02 and np.all(lst[:-1] >= best * 1.02)      # stopped at the FIRST entry below the threshold
    assert np.all(np.diff(lst[f]) <= 1e-9)                                 # RRT* costs never increase
    assert len(lst) < 30000


def test_gap_flank_threshold():
    """Gap world (get_gap_problem_input, planning_problem_utils_2d.py:92-143): planning_block_gap(flank_path_len) stops
    as soon as the path through the gap is found -- shorter than the flank path (generate_block_gap_env_2d.py:35),
    never shorter than the straight line."""
    import math
    h, t, h_g, y_g, d_goal, H, W = 90, 20, 6, 40, 60, 224, 224
    flank = t + 2 * (((d_goal - t) / 2) ** 2 + (h / 2) ** 2) ** 0.5
    x0, y0 = W // 2 - t // 2, H // 2 - h // 2
    rects = [[x0, y0, t, h - h_g - y_g], [x0, y0 + (h - y_g), t, y_g]]
    free = H * W - sum((r[2] + 1) * (r[3] + 1) for r in rects)
    gamma = math.ceil((2 * (1 + 1. / 2)) ** (1. / 2) * (free / np.pi) ** (1. / 2))
    problem = _problem((H, W), rects, (W // 2 - d_goal // 2, H // 2), (W // 2 + d_goal // 2, H // 2), gamma)
    np.random.seed(77); random.seed(77)
    planner = _planner("irrt", problem, 30000)
    lst = np.array(planner.planning_block_gap(flank))
    assert lst[-1] < flank and np.all(lst[:-1] >= flank) and lst[-1] >= d_goal
    assert planner.check_success(planner.extract_path(planner.path_solutions[0]))
