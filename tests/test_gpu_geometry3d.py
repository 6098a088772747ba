"""3D collision / validity predicates against fixtures labelled by the REFERENCE's own code
(collision_check_utils_3d.py:151-216,298-398 through rrt_utils_3d.Utils; tests/golden/make_golden_geom3d.py):
1.02e6 edges and 3.6e5 points over three worlds, incl. zero-length, tangent, face-touching and integer cases.
Bit-exact (boolean) parity is required."""
import glob
import os

import numpy as np
import pytest

from nirrt_star_b200.synthetic import make_problem_3d
from tests import geom3d_cases as G

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "geom3d_e*.npz")))


def _load(path):
    g = np.load(path)
    env_idx = int(g["env_idx"])
    pr = make_problem_3d(env_idx)
    edges = G.make_edges(pr["env_dict"], 4000 + env_idx, int(g["m_edges"]))
    pts = G.make_points(pr["env_dict"], 5000 + env_idx, int(g["m_points"]))
    assert G.digest(edges) == str(g["edges_sha256"]) and G.digest(pts) == str(g["points_sha256"]), "fixture inputs not reproduced"
    want = {k: np.unpackbits(g[k])[:len(edges) if k == "hit" else len(pts)].astype(bool) for k in ("hit", "inside", "valid")}
    return pr, edges, pts, want


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_cuda_predicates_match_reference_fixture(path):
    from nirrt_star_b200 import batch as B
    pr, edges, pts, want = _load(path)
    bp = B.BatchPlanner3D([pr], 16, seeds=[0])
    assert np.array_equal(bp.collide_edges(0, edges), want["hit"])
    assert np.array_equal(bp.points_inside_obs(0, pts), want["inside"])
    assert np.array_equal(bp.points_valid(0, pts), want["valid"])
    bp.close()


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_oracle_predicates_match_reference_fixture(path):
    """pins oracle/nirrt_oracle.c's geometry against the reference on the same million edges (CPU)"""
    from oracle.planner_oracle import Oracle3D
    pr, edges, pts, want = _load(path)
    o = Oracle3D(pr, 16, seed=0)
    assert np.array_equal(o.collide_edges(edges), want["hit"])
    assert np.array_equal(o.points_inside_obs(pts), want["inside"])
    assert np.array_equal(o.points_valid(pts), want["valid"])
