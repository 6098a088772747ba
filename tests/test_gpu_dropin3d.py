"""The drop-in classes reproduce the reference's public behaviour (golden fixtures recorded from the
reference's own RRTStar3D / IRRTStar3D, tests/golden/make_golden_planner.py)."""
import glob
import os
import random
import types

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "planner3d_*.npz")))


@pytest.fixture(scope="module", autouse=True)
def dropin():
    from nirrt_star_b200 import dropin
    dropin.install()


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_dropin_planner_matches_reference_golden(path):
    import importlib
    from nirrt_star_b200.synthetic import make_problem_3d
    from path_planning_utils_3d.rrt_env_3d import Env
    g = np.load(path)
    kind, mode = str(g["kind"]), str(g["mode"])
    mod = importlib.import_module("path_planning_classes_3d." + {"rrt": "rrt_star_3d", "irrt": "irrt_star_3d"}[kind])
    problem = make_problem_3d(int(g["env_idx"]))
    problem["env"] = Env(problem["env_dict"])
    args = types.SimpleNamespace(step_len=10, iter_max=int(g["iter_max"]), clearance=2)
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
    assert planner.vertices.shape == (1 + args.iter_max, 3) and planner.vertex_parents.shape == (1 + args.iter_max,)
    assert np.array_equal(planner.vertex_parents[:n], g["parents"])
    assert np.array_equal(planner.vertices[:n], g["vertices"])     # incl. informed sampling: glibc sin / cos restated bit for bit
    if kind == "irrt":
        assert list(planner.path_solutions) == list(g["solutions"])
    # the global numpy stream advanced exactly as the reference would have advanced it
    assert np.random.random() == float(g["next_random"]) if "next_random" in g.files else True


def test_dropin_utils_match_oracle():
    from nirrt_star_b200.synthetic import make_problem_3d
    from oracle.planner_oracle import Oracle3D
    from path_planning_classes_3d.rrt_utils_3d import Utils
    from path_planning_classes_3d import collision_check_utils_3d as ccu
    from path_planning_utils_3d.rrt_env_3d import Env
    pr = make_problem_3d(7)
    env = Env(pr["env_dict"])
    u = Utils(env, 2)
    o = Oracle3D(pr, 10)
    rng = np.random.default_rng(0)
    pts = rng.uniform(-1, 51, (300, 3))
    edges = np.stack([pts[:150], pts[150:]], 1)
    assert [u.is_collision(a, b) for a, b in edges] == list(o.collide_edges(edges))
    assert [u.is_inside_obs(p) for p in pts] == list(o.points_inside_obs(pts))
    assert [u.is_valid(tuple(p)) for p in pts] == list(o.points_valid(pts))
    balls = np.array(env.obs_ball, dtype=np.float64); boxes = np.array(env.obs_box, dtype=np.float64)
    assert np.array_equal(ccu.points_in_balls_boxes(pts, balls, boxes, 2), o.points_inside_obs(pts))
    assert ccu.points_in_balls_boxes(tuple(pts[0]), balls, boxes, 2) == bool(o.points_inside_obs(pts[:1])[0])
    assert np.array_equal(ccu.points_validity_3d(pts, balls, boxes, env.x_range, env.y_range, env.z_range, 2, 2),
                          o.points_valid(pts))
