"""The 2D drop-in classes reproduce the reference's public behaviour (golden fixtures recorded from
the reference's own RRTStar2D / IRRTStar2D, tests/golden/make_golden_planner2d.py)."""
import glob
import os
import random
import types

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "planner2d_*.npz")))


@pytest.fixture(scope="module", autouse=True)
def dropin():
    from nirrt_star_b200 import dropin
    dropin.install()


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_dropin_planner_matches_reference_golden(path):
    import importlib
    from nirrt_star_b200.synthetic import make_problem_2d
    from path_planning_utils.rrt_env import Env
    g = np.load(path)
    kind, mode = str(g["kind"]), str(g["mode"])
    mod = importlib.import_module("path_planning_classes." + {"rrt": "rrt_star_2d", "irrt": "irrt_star_2d"}[kind])
    problem = make_problem_2d(int(g["env_idx"]))
    problem["env"] = Env(problem["env_dict"])
    args = types.SimpleNamespace(step_len=10, iter_max=int(g["iter_max"]), clearance=3)
    seed = int(g["seed"])
    np.random.seed(seed); random.seed(seed)
    planner = mod.get_path_planner(args, problem, None)
    if mode == "planning":
        planner.planning(False)
        want_path = g["path"]
        if len(want_path):
            assert planner.check_success(planner.path)
            assert np.array_equal(planner.path, want_path)
        else:
            assert len(planner.path) == 0
    else:
        lst = planner.planning_random(int(g["iter_after"]))
        want = g["path_len_list"]
        assert isinstance(lst, list) and len(lst) == len(want)
        assert np.array_equal(np.isinf(lst), np.isinf(want))
        f = np.isfinite(want)
        assert np.allclose(np.array(lst)[f], want[f], rtol=1e-5, atol=0)
    n = planner.num_vertices
    assert n == int(g["num_vertices"])
    assert planner.vertices.shape == (1 + args.iter_max, 2) and planner.vertex_parents.shape == (1 + args.iter_max,)
    assert np.array_equal(planner.vertex_parents[:n], g["parents"])
    assert np.array_equal(planner.vertices[:n], g["vertices"])
    if kind == "irrt":
        assert list(planner.path_solutions) == list(g["solutions"])
    assert np.random.random() == float(g["next_random"])
    assert random.random() == float(g["next_py_random"])


def test_dropin_utils_match_reference_golden():
    from nirrt_star_b200.synthetic import make_problem_2d
    from path_planning_classes.rrt_utils_2d import Utils
    from path_planning_classes import collision_check_utils as ccu
    from path_planning_utils.rrt_env import Env
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "geom2d_e1.npz"))
    env = Env(make_problem_2d(1)["env_dict"])
    u = Utils(env, 3)
    assert [u.is_collision(e[0], e[1]) for e in g["edges"][:300]] == list(g["hit"][:300])
    assert [u.is_inside_obs(p) for p in g["pts"][:300]] == list(g["inside"][:300])
    assert [u.is_valid(tuple(p)) for p in g["pts"][:300]] == list(g["valid"][:300])
    circles, rects = np.array(env.obs_circle), np.array(env.obs_rectangle)
    assert np.array_equal(ccu.points_in_circles_rectangles(g["pts"], circles, rects, 3), g["inside"])
    assert np.array_equal(ccu.points_validity(g["pts"], circles, rects, env.x_range, env.y_range, 3, 3), g["valid"])
    assert ccu.points_in_circles_rectangles(tuple(g["pts"][0]), circles, rects, 3) == bool(g["inside"][0])
