"""Generates tests/golden/planner3d_*.npz by running the REFERENCE's own planner classes
(/root/reference, via oracle/ref_shim.py) on synthetic random_3d problems under fixed seeds.

Run in the build container only:  python tests/golden/make_golden_planner.py
The GPU box never runs this (no /root/reference there); it only reads the committed .npz files.

Seeding convention: np.random.seed(s); random.seed(s); torch.manual_seed(s)
before constructing the planner; the problem dict is built once beforehand.
"""
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

ref_shim.install()
import contextlib, io  # noqa: E402

from nirrt_star_b200.synthetic import make_problem_3d  # noqa: E402
from path_planning_utils_3d.rrt_env_3d import Env  # noqa: E402
from path_planning_classes_3d.rrt_star_3d import RRTStar3D  # noqa: E402
from path_planning_classes_3d.irrt_star_3d import IRRTStar3D  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def instrument(planner):
    tr = {"nearest": [], "near": [], "rand": []}
    cls = type(planner)
    orig_nn = cls.nearest_neighbor
    orig_near = planner.find_near_neighbors

    def nn(node_list, n):
        node, idx = orig_nn(node_list, n)
        tr["nearest"].append(int(idx)); tr["rand"].append(np.array(n, dtype=np.float64))
        return node, idx

    def near(node_new, node_new_index=None):
        out = orig_near(node_new, node_new_index)
        tr["near"].append((len(tr["nearest"]) - 1, np.array(out, dtype=np.int64)))
        return out

    planner.nearest_neighbor = nn
    planner.find_near_neighbors = near
    return tr


def run_case(kind, env_idx, seed, iter_max, mode, iter_after=0):
    problem = make_problem_3d(env_idx)
    problem["env"] = Env(problem["env_dict"])
    np.random.seed(seed); random.seed(seed)
    cls = {"rrt": RRTStar3D, "irrt": IRRTStar3D}[kind]
    pl = cls(problem["x_start"], problem["x_goal"], 10, problem["search_radius"], iter_max, problem["env"], 2)
    tr = instrument(pl)
    with contextlib.redirect_stdout(io.StringIO()):
        if mode == "planning":
            pl.planning(False)
            plist = np.zeros(0)
        else:
            plist = np.array(pl.planning_random(iter_after), dtype=np.float64)
    next_random = np.random.random()   # pins how far the reference advanced the global stream
    n = pl.num_vertices
    k = len(tr["nearest"])
    near_cnt = np.full(k, -1, dtype=np.int64)   # -1: steer edge collided (find_near not called)
    near_flat = []
    for it, arr in tr["near"]:
        near_cnt[it] = len(arr); near_flat.append(arr)
    near_flat = np.concatenate(near_flat) if near_flat else np.zeros(0, dtype=np.int64)
    sols = np.array(getattr(pl, "path_solutions", []), dtype=np.int64)
    path = np.array(pl.path, dtype=np.float64) if len(pl.path) else np.zeros((0, 3))
    name = f"planner3d_{kind}_{mode}_e{env_idx}_s{seed}_i{iter_max}.npz"
    np.savez_compressed(os.path.join(OUT, name), kind=kind, mode=mode, env_idx=env_idx, seed=seed,
                        iter_max=iter_max, iter_after=iter_after,
                        nearest=np.array(tr["nearest"], dtype=np.int64), rand=np.array(tr["rand"]),
                        near_cnt=near_cnt, near=near_flat, vertices=pl.vertices[:n].copy(),
                        parents=pl.vertex_parents[:n].astype(np.int64), num_vertices=n,
                        path_len_list=plist, solutions=sols, path=path, next_random=next_random)
    print(name, "iters", k, "n", n, "finite", int(np.isfinite(plist).sum()) if len(plist) else "-")


if __name__ == "__main__":
    run_case("rrt", 0, 7, 600, "planning")
    run_case("rrt", 1, 11, 1500, "random", iter_after=200)
    run_case("rrt", 3, 5, 800, "random", iter_after=100)
    run_case("irrt", 0, 21, 1500, "random", iter_after=300)
    run_case("irrt", 2, 9, 1200, "planning")
    run_case("irrt", 5, 13, 1500, "random", iter_after=200)
