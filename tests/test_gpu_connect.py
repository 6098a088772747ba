"""Neural Connect: the drop-in PNGWrapper.generate_connected_path_points and the
CUDA graph analysis behind it vs fixtures recorded from the reference's own wrapper
(tests/golden/make_golden_connect.py; the network is replaced on both sides by
tests/golden/fake_connect.py)."""
import glob
import os
import random
import sys
import types

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(__file__)
GOLD = sorted(glob.glob(os.path.join(HERE, "golden", "connect_*.npz")))
sys.path.insert(0, os.path.join(HERE, "golden"))


@pytest.fixture(scope="module", autouse=True)
def dropin():
    from nirrt_star_b200 import dropin
    dropin.install()


def _wrapper(dim, fake):
    if dim == 3:
        from wrapper_3d.pointnet_pointnet2.pointnet2_wrapper_connect_bfs import PNGWrapper
    else:
        from wrapper.pointnet_pointnet2.pointnet2_wrapper_connect_bfs import PNGWrapper
    w = object.__new__(PNGWrapper)
    w.classify_path_points = fake.classify_path_points
    return w


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_generate_connected_path_points_matches_reference(path):
    from fake_connect import FakeClassifier
    g = np.load(path)
    fake = FakeClassifier(float(g["reach"]))
    w = _wrapper(int(g["dim"]), fake)
    ok, runs, mask = w.generate_connected_path_points(g["pc"], g["x_start"], g["x_goal"], {"env_dims": None}, neighbor_radius=10,
                                                      max_trial_attempts=int(g["max_trials"]))
    assert bool(ok) == bool(g["success"]) and runs == int(g["runs"])
    assert mask.dtype == np.float32 and np.array_equal(mask, g["mask"])
    assert fake.calls == list(g["calls"])          # every trial saw the same start / goal masks


def test_graph_analysis_matches_numpy():
    from nirrt_star_b200.pointnet2 import connect_analyse
    rs = np.random.RandomState(1)
    for dim, n, r in ((3, 2048, 6.0), (2, 1500, 9.0), (3, 300, 12.0), (2, 2048, 4.0)):
        pc = rs.uniform(0, 60, (n, dim)).astype(np.float32)
        pm = (rs.rand(n) < 0.4).astype(np.float32)
        src, dst = pc[0] + 0.5, pc[1] - 0.5
        has, vis, bnd = connect_analyse(pc, pm, src, dst, r)
        v = np.concatenate([src[None], dst[None], pc[pm > 0]]).astype(np.float32)
        adj = np.linalg.norm(v[:, None] - v, axis=2) < r
        comp = np.zeros(len(v), bool); comp[0] = True; frontier = [0]
        while frontier:
            nxt = np.nonzero(adj[frontier].any(axis=0) & ~comp)[0]
            comp[nxt] = True; frontier = list(nxt)
        assert has == bool(comp[1])
        if not has:
            want_vis = np.zeros(n, np.float32); want_vis[np.nonzero(pm > 0)[0][comp[2:]]] = 1
            assert np.array_equal(vis, want_vis)
            d = np.linalg.norm(pc[vis > 0][:, None] - pc[pm == 0], axis=2)
            want_b = np.zeros(n, np.float32); want_b[np.nonzero(vis > 0)[0][(d < r).any(axis=1)]] = 1
            assert np.array_equal(bnd, want_b)


@pytest.mark.parametrize("dim", [3, 2])
def test_connect_planners_run_end_to_end(dim, tmp_path):
    """NIRRT*-PNG(C) / NRRT*-PNG(C) with the CUDA network and CUDA Neural Connect."""
    import importlib
    import torch
    from nirrt_star_b200.synthetic import make_pointnet2_state, make_problem_2d, make_problem_3d
    d = tmp_path / f"results/model_training/pointnet2_{dim}d/checkpoints"
    d.mkdir(parents=True)
    torch.save({"model_state_dict": {k: torch.from_numpy(np.asarray(v)) for k, v in make_pointnet2_state(0).items()}},
               str(d / f"best_pointnet2_{dim}d.pth"))
    wmod = importlib.import_module(("wrapper_3d" if dim == 3 else "wrapper") + ".pointnet_pointnet2.pointnet2_wrapper_connect_bfs")
    w = wmod.PNGWrapper(root_dir=str(tmp_path), device="cuda")
    problem = (make_problem_3d if dim == 3 else make_problem_2d)(0 if dim == 3 else 1)
    args = types.SimpleNamespace(step_len=10, iter_max=1500, clearance=2 if dim == 3 else 3, pc_n_points=2048, pc_over_sample_scale=5,
                                 pc_sample_rate=0.5, pc_update_cost_ratio=0.9, connect_max_trial_attempts=5)
    pkg = "path_planning_classes_3d." if dim == 3 else "path_planning_classes."
    for name in ("nirrt_star_png_c_", "nrrt_star_png_c_"):
        np.random.seed(2); random.seed(2); torch.manual_seed(2)
        planner = importlib.import_module(pkg + name + f"{dim}d").get_path_planner(args, problem, w)
        lst = planner.planning_random(100)
        assert "(C)" in planner.get_path_planner_name()
        assert len(lst) >= 100 and len(planner.path_point_cloud_pred) > 0
