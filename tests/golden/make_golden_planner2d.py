"""Generates tests/golden/planner2d_*.npz by running the REFERENCE's own RRTStar2D / IRRTStar2D
(/root/reference, via oracle/ref_shim.py) on synthetic random_2d problems under fixed seeds, plus
geom2d_*.npz (segment / point predicates of collision_check_utils.py on random inputs).
Run in the build container only:  python tests/golden/make_golden_planner2d.py
Seeding: np.random.seed(s); random.seed(s) before constructing the planner."""
import contextlib
import io
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

ref_shim.install()
from nirrt_star_b200.synthetic import make_problem_2d  # noqa: E402
from path_planning_utils.rrt_env import Env  # noqa: E402
from path_planning_classes.rrt_star_2d import RRTStar2D  # noqa: E402
from path_planning_classes.irrt_star_2d import IRRTStar2D  # noqa: E402
from path_planning_classes.rrt_utils_2d import Utils  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def instrument(planner):
    tr = {"nearest": [], "near": [], "rand": []}
    orig_nn = type(planner).nearest_neighbor
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
    problem = make_problem_2d(env_idx)
    env = Env(problem["env_dict"])
    np.random.seed(seed); random.seed(seed)
    cls = {"rrt": RRTStar2D, "irrt": IRRTStar2D}[kind]
    pl = cls(problem["x_start"], problem["x_goal"], 10, problem["search_radius"], iter_max, env, 3)
    tr = instrument(pl)
    with contextlib.redirect_stdout(io.StringIO()):
        if mode == "planning":
            pl.planning(False); plist = np.zeros(0)
        else:
            plist = np.array(pl.planning_random(iter_after), dtype=np.float64)
    next_np, next_py = np.random.random(), random.random()
    n = pl.num_vertices
    k = len(tr["nearest"])
    near_cnt = np.full(k, -1, dtype=np.int64); near_flat = []
    for it, arr in tr["near"]:
        near_cnt[it] = len(arr); near_flat.append(arr)
    near_flat = np.concatenate(near_flat) if near_flat else np.zeros(0, dtype=np.int64)
    name = f"planner2d_{kind}_{mode}_e{env_idx}_s{seed}_i{iter_max}.npz"
    np.savez_compressed(os.path.join(OUT, name), kind=kind, mode=mode, env_idx=env_idx, seed=seed, iter_max=iter_max,
                        iter_after=iter_after, nearest=np.array(tr["nearest"], dtype=np.int64), rand=np.array(tr["rand"]),
                        near_cnt=near_cnt, near=near_flat, vertices=pl.vertices[:n].copy(),
                        parents=pl.vertex_parents[:n].astype(np.int64), num_vertices=n, path_len_list=plist,
                        solutions=np.array(getattr(pl, "path_solutions", []), dtype=np.int64),
                        path=np.array(pl.path, dtype=np.float64) if len(pl.path) else np.zeros((0, 2)),
                        next_random=next_np, next_py_random=next_py)
    print(name, "iters", k, "n", n, "finite", int(np.isfinite(plist).sum()) if len(plist) else "-")


def geometry_case(env_idx, m=4000):
    problem = make_problem_2d(env_idx)
    u = Utils(Env(problem["env_dict"]), 3)
    rs = np.random.RandomState(900 + env_idx)
    a = rs.uniform(-5, 229, (m, 2)); d = rs.normal(size=(m, 2)); d /= np.linalg.norm(d, axis=1)[:, None]
    b = a + d * rs.uniform(0, 12, (m, 1))
    b[:50] = a[:50]                                  # zero-length edges
    a[50:400] = np.round(a[50:400]); b[50:400] = np.round(b[50:400])   # integer coordinates: touching cases
    b[400:500, 0] = a[400:500, 0]                    # vertical
    b[500:600, 1] = a[500:600, 1]                    # horizontal
    edges = np.stack([a, b], 1)
    hit = np.array([u.is_collision(e[0], e[1]) for e in edges])
    pts = rs.uniform(-5, 229, (m, 2)); pts[:1000] = np.round(pts[:1000])
    inside = np.array([u.is_inside_obs(p) for p in pts])
    valid = np.array([u.is_valid(p) for p in pts])
    np.savez_compressed(os.path.join(OUT, f"geom2d_e{env_idx}.npz"), env_idx=env_idx, edges=edges, hit=hit, pts=pts,
                        inside=inside, valid=valid)
    print(f"geom2d_e{env_idx}.npz hits", int(hit.sum()), "inside", int(inside.sum()), "valid", int(valid.sum()))


if __name__ == "__main__":
    for e in (0, 1, 2):
        geometry_case(e)
    run_case("rrt", 0, 7, 500, "planning")
    run_case("rrt", 1, 11, 1500, "random", iter_after=200)
    run_case("rrt", 3, 5, 1000, "random", iter_after=150)
    run_case("irrt", 0, 21, 1500, "random", iter_after=300)
    run_case("irrt", 2, 9, 1200, "planning")
    run_case("irrt", 4, 13, 1500, "random", iter_after=250)
