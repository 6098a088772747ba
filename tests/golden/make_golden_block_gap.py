"""Generates tests/golden/blockgap2d_*.npz by running the REFERENCE's own RRTStar2D / IRRTStar2D
``planning_block_gap`` (rrt_star_2d.py:159-196, irrt_star_2d.py:180-228) on block / gap problems built by the
reference's own ``get_block_problem_input`` / ``get_gap_problem_input`` (datasets/planning_problem_utils_2d.py:49-143)
from configs shaped like generate_block_gap_env_2d.py:15-47, exactly as eval_planning_2d.py:117-121 drives them
(clearance 0, step_len 10, threshold = best_path_len * 1.02 for block, flank_path_len for gap).
Run in the build container only:  python tests/golden/make_golden_block_gap.py
The fixtures carry the problem itself (obstacles, start / goal, search radius, analytic lengths), so the GPU
tests need neither cv2 nor /root/reference."""
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
from datasets.planning_problem_utils_2d import get_block_problem_input, get_gap_problem_input  # noqa: E402
from path_planning_classes.rrt_star_2d import RRTStar2D  # noqa: E402
from path_planning_classes.irrt_star_2d import IRRTStar2D  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def block_config(w, ratio, d_goal=60):
    # generate_block_gap_env_2d.py:15-28
    best = w + (((d_goal - w) // 2) ** 2 + (w // 2) ** 2) ** 0.5 + (((d_goal - w) - (d_goal - w) // 2) ** 2 + (w // 2) ** 2) ** 0.5
    return {"w": int(w), "d_goal": d_goal, "img_height": d_goal * ratio, "img_width": d_goal * ratio, "best_path_len": best}


def gap_config(h_g, y_g, h=90, t=20, d_goal=60):
    # generate_block_gap_env_2d.py:30-47
    flank = t + 2 * (((d_goal - t) / 2) ** 2 + (h / 2) ** 2) ** 0.5
    return {"h": h, "t": t, "h_g": h_g, "y_g": int(y_g), "d_goal": d_goal, "img_height": 224, "img_width": 224,
            "flank_path_len": flank}


def run_case(tag, kind, problem_kind, cfg, seed, iter_max):
    problem = get_block_problem_input(cfg) if problem_kind == "block" else get_gap_problem_input(cfg)
    threshold = problem["best_path_len"] * 1.02 if problem_kind == "block" else problem["flank_path_len"]
    np.random.seed(seed); random.seed(seed)
    cls = {"rrt": RRTStar2D, "irrt": IRRTStar2D}[kind]
    pl = cls(problem["x_start"], problem["x_goal"], 10.0, problem["search_radius"], iter_max, problem["env"], 0.0)
    with contextlib.redirect_stdout(io.StringIO()):
        plist = np.array(pl.planning_block_gap(threshold), dtype=np.float64)
    next_np, next_py = np.random.random(), random.random()
    n = pl.num_vertices
    ed = problem["env_dict"]
    name = f"blockgap2d_{problem_kind}_{kind}_{tag}_s{seed}_i{iter_max}.npz"
    np.savez_compressed(os.path.join(OUT, name), kind=kind, problem_kind=problem_kind, seed=seed, iter_max=iter_max,
                        threshold=threshold, analytic=problem.get("best_path_len", problem.get("flank_path_len")),
                        env_dims=np.array(ed["env_dims"]), rects=np.array(ed["rectangle_obstacles"], dtype=np.float64),
                        x_start=np.array(problem["x_start"], dtype=np.float64), x_goal=np.array(problem["x_goal"], dtype=np.float64),
                        search_radius=float(problem["search_radius"]), path_len_list=plist,
                        vertices=pl.vertices[:n].copy(), parents=pl.vertex_parents[:n].astype(np.int64), num_vertices=n,
                        solutions=np.array(getattr(pl, "path_solutions", []), dtype=np.int64),
                        next_random=next_np, next_py_random=next_py)
    print(name, "len", len(plist), "n", n, "last", plist[-1], "threshold", threshold, "reached", bool(plist[-1] < threshold))


if __name__ == "__main__":
    run_case("w20r2", "irrt", "block", block_config(20, 2), 3, 3000)
    run_case("w34r3", "irrt", "block", block_config(34, 3), 4, 2500)
    run_case("w12r2", "rrt", "block", block_config(12, 2), 5, 1500)
    run_case("g7y30", "rrt", "gap", gap_config(7, 30), 6, 2500)
    run_case("g5y55", "irrt", "gap", gap_config(5, 55), 7, 3000)
    run_case("g3y42", "irrt", "gap", gap_config(3, 42), 8, 2000)
