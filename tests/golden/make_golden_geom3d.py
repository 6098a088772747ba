"""Generates tests/golden/geom3d_e*.npz: the REFERENCE's own 3D predicates (Utils.is_collision / is_inside_obs /
is_valid of path_planning_classes_3d/rrt_utils_3d.py:22-86 -> collision_check_utils_3d.py:151-216,298-398) on the
deterministic edge / point sets of tests/geom3d_cases.py, >= 1e6 edges in total.  Only the packed answers and a
SHA-256 of the inputs are stored (the GPU test regenerates the inputs and checks the digest first).
Run in the build container only:  python tests/golden/make_golden_geom3d.py   (a few minutes, multi-process)."""
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = os.path.dirname(os.path.abspath(__file__))
M_EDGES, M_POINTS = 340000, 120000
ENVS = (0, 3, 7)


def _label(args):
    env_idx, kind, lo, hi = args
    from oracle import ref_shim
    ref_shim.install()
    from nirrt_star_b200.synthetic import make_problem_3d
    from path_planning_utils_3d.rrt_env_3d import Env
    from path_planning_classes_3d.rrt_utils_3d import Utils
    from tests import geom3d_cases as G
    ed = make_problem_3d(env_idx)["env_dict"]
    u = Utils(Env(ed), 2)
    if kind == "edges":
        e = G.make_edges(ed, 4000 + env_idx, M_EDGES)[lo:hi]
        return np.array([bool(u.is_collision(x[0], x[1])) for x in e])
    p = G.make_points(ed, 5000 + env_idx, M_POINTS)[lo:hi]
    if kind == "inside":
        return np.array([bool(u.is_inside_obs(x)) for x in p])
    return np.array([bool(u.is_valid(x)) for x in p])


if __name__ == "__main__":
    from nirrt_star_b200.synthetic import make_problem_3d
    from tests import geom3d_cases as G
    workers = os.cpu_count() or 1
    with mp.get_context("spawn").Pool(workers) as pool:
        for env_idx in ENVS:
            ed = make_problem_3d(env_idx)["env_dict"]
            out = {}
            for kind, m in (("edges", M_EDGES), ("inside", M_POINTS), ("valid", M_POINTS)):
                step = (m + 4 * workers - 1) // (4 * workers)
                jobs = [(env_idx, kind, lo, min(m, lo + step)) for lo in range(0, m, step)]
                out[kind] = np.concatenate(pool.map(_label, jobs))
            edges = G.make_edges(ed, 4000 + env_idx, M_EDGES); pts = G.make_points(ed, 5000 + env_idx, M_POINTS)
            np.savez_compressed(os.path.join(OUT, f"geom3d_e{env_idx}.npz"), env_idx=env_idx, m_edges=M_EDGES, m_points=M_POINTS,
                                edges_sha256=G.digest(edges), points_sha256=G.digest(pts),
                                hit=np.packbits(out["edges"]), inside=np.packbits(out["inside"]), valid=np.packbits(out["valid"]))
            print(f"geom3d_e{env_idx}.npz edges {M_EDGES} hits {int(out['edges'].sum())} inside {int(out['inside'].sum())} valid {int(out['valid'].sum())}")
