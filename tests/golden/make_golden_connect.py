"""Generates tests/golden/connect_*.npz by running the REFERENCE's own
PNGWrapper.generate_connected_path_points (wrapper{,_3d}/pointnet_pointnet2/pointnet2_wrapper_connect_bfs.py)
with the network replaced by tests/golden/fake_connect.py.  Run in the build container only."""
import os
import sys

import numpy as np
import torch  # noqa: F401  (must be imported before the matplotlib stubs are registered)

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import ref_shim  # noqa: E402

ref_shim.install()
from fake_connect import FakeClassifier  # noqa: E402
from nirrt_star_b200.synthetic import make_cloud_3d, make_problem_2d, make_problem_3d  # noqa: E402
from wrapper.pointnet_pointnet2.pointnet2_wrapper_connect_bfs import PNGWrapper as W2  # noqa: E402
from wrapper_3d.pointnet_pointnet2.pointnet2_wrapper_connect_bfs import PNGWrapper as W3  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def run_case(dim, env_idx, reach, max_trials=5):
    if dim == 3:
        pr = make_problem_3d(env_idx)
        pc, _, _ = make_cloud_3d(env_idx)
        w = object.__new__(W3)
    else:
        pr = make_problem_2d(env_idx)
        rs = np.random.RandomState(50 + env_idx)
        pts = rs.uniform(0, 224, (20000, 2))
        pix = pts.astype(int)
        pc = pts[pr["binary_mask"][np.clip(pix[:, 1], 0, 223), np.clip(pix[:, 0], 0, 223)] > 0][:2048].astype(np.float32)
        w = object.__new__(W2)
    fake = FakeClassifier(reach)
    w.classify_path_points = fake.classify_path_points
    xs = np.array(pr["x_start"]).astype(np.float64); xg = np.array(pr["x_goal"]).astype(np.float64)
    ok, runs, mask = w.generate_connected_path_points(pc, xs, xg, pr["env_dict"], neighbor_radius=10, max_trial_attempts=max_trials)
    name = f"connect_{dim}d_e{env_idx}_r{int(reach)}.npz"
    np.savez_compressed(os.path.join(OUT, name), dim=dim, env_idx=env_idx, reach=reach, max_trials=max_trials, pc=pc, x_start=xs,
                        x_goal=xg, success=bool(ok), runs=int(runs), mask=mask.astype(np.float32), calls=np.array(fake.calls))
    print(name, "success", ok, "runs", runs, "path points", int(mask.sum()))


if __name__ == "__main__":
    run_case(3, 0, 13.0); run_case(3, 2, 11.0); run_case(3, 4, 16.0); run_case(3, 5, 9.0, max_trials=3)
    run_case(2, 1, 22.0); run_case(2, 4, 30.0); run_case(2, 6, 40.0); run_case(2, 3, 15.0, max_trials=4)
