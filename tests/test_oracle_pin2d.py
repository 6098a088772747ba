"""Pins oracle/planner2d_oracle.py against traces recorded from the reference's own RRTStar2D /
IRRTStar2D and collision_check_utils (tests/golden/make_golden_planner2d.py).  On the CPU the oracle
uses the same libm/numpy as the reference, so everything is compared bit for bit."""
import glob
import os

import numpy as np
import pytest

from nirrt_star_b200.synthetic import make_problem_2d
from oracle.planner2d_oracle import Oracle2D, points_in_obstacles, points_valid

HERE = os.path.dirname(__file__)
GOLD = sorted(glob.glob(os.path.join(HERE, "golden", "planner2d_*.npz")))
GEOM = sorted(glob.glob(os.path.join(HERE, "golden", "geom2d_*.npz")))


@pytest.mark.parametrize("path", GEOM, ids=[os.path.basename(p) for p in GEOM])
def test_geometry_matches_reference(path):
    g = np.load(path)
    o = Oracle2D(make_problem_2d(int(g["env_idx"])), 10)
    assert np.array_equal([o.collides(e[0], e[1]) for e in g["edges"]], g["hit"])
    assert np.array_equal(points_in_obstacles(g["pts"], o.circles, o.rects, 3), g["inside"])
    assert np.array_equal(points_valid(g["pts"], o.circles, o.rects, o.x_range, o.y_range, 3), g["valid"])


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_planner_trace_matches_reference(path):
    g = np.load(path)
    kind, mode = str(g["kind"]), str(g["mode"])
    variant = {"rrt": 0, "irrt": 1}[kind]
    o = Oracle2D(make_problem_2d(int(g["env_idx"])), int(g["iter_max"]), seed=int(g["seed"]))
    o.trace = []
    if mode == "planning":
        o.run(int(g["iter_max"]), variant, 0)
    else:
        lst = o.planning_random(int(g["iter_after"]), variant)
        assert np.array_equal(np.array(lst), g["path_len_list"])
    n = o.n
    assert n == int(g["num_vertices"])
    assert np.array_equal(o.parent[:n], g["parents"])
    assert np.array_equal(o.v[:n], g["vertices"])
    assert np.array_equal([t["nearest"] for t in o.trace], g["nearest"])
    cnt = np.array([-1 if t["near"] is None else len(t["near"]) for t in o.trace])
    assert np.array_equal(cnt, g["near_cnt"])
    flat = [t["near"] for t in o.trace if t["near"] is not None]
    assert np.array_equal(np.concatenate(flat) if flat else np.zeros(0, dtype=np.int64), g["near"])
    if kind == "irrt":
        assert o.solutions == list(g["solutions"])
    assert o.rs.random_sample() == float(g["next_random"])
    assert o.py.random() == float(g["next_py_random"])
