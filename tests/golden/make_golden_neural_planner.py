"""Generates tests/golden/neural3d_*.npz by running the REFERENCE's own NIRRTStarPNG3D / NRRTStarPNG3D
(/root/reference, via oracle/ref_shim.py) with a recording stand-in for the network: a geometric
rule picks the "path" points of each guidance cloud, and every call (cloud hash + prediction) is
recorded.  The GPU test replays the recorded predictions through a stub wrapper, so it pins the
planner loop, the sampler switch, the cloud-update trigger and the guidance-cloud generation
(numpy stream + obstacle filters + farthest-point down-sampling) independently of network
numerics.  Run in the build container only:  python tests/golden/make_golden_neural_planner.py
"""
import contextlib
import hashlib
import io
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

ref_shim.install()
from nirrt_star_b200.synthetic import make_problem_2d, make_problem_3d  # noqa: E402
from path_planning_classes_3d.nirrt_star_png_3d import NIRRTStarPNG3D  # noqa: E402
from path_planning_classes_3d.nrrt_star_png_3d import NRRTStarPNG3D  # noqa: E402
from path_planning_classes.nirrt_star_png_2d import NIRRTStarPNG2D  # noqa: E402
from path_planning_classes.nrrt_star_png_2d import NRRTStarPNG2D  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


class RecordingWrapper:
    """classify_path_points stand-in: path = points within `width` of the start-goal segment."""

    def __init__(self, x_start, x_goal, width=9.0):
        self.a = np.asarray(x_start, dtype=np.float64); self.b = np.asarray(x_goal, dtype=np.float64)
        self.width = width
        self.calls = []

    def classify_path_points(self, pc, start_mask, goal_mask):
        assert pc.dtype == np.float32 and start_mask.dtype == np.float32 and goal_mask.dtype == np.float32
        p = pc.astype(np.float64)
        ab = self.b - self.a
        t = np.clip(((p - self.a) @ ab) / (ab @ ab), 0, 1)
        d = np.linalg.norm(p - (self.a + t[:, None] * ab), axis=1)
        pred = (d < self.width).astype(np.int64)
        self.calls.append((hashlib.sha1(np.ascontiguousarray(pc).tobytes()).hexdigest(), len(pc),
                           hashlib.sha1(start_mask.tobytes() + goal_mask.tobytes()).hexdigest(), np.packbits(pred.astype(np.uint8)),
                           pc.copy()))
        return pred, d.astype(np.float32)


def run_case_2d(kind, env_idx, seed, iter_max, mode, iter_after=0, pc_sample_rate=0.5, ratio=0.9):
    problem = make_problem_2d(env_idx)
    np.random.seed(seed); random.seed(seed)
    w = RecordingWrapper(problem["x_start"], problem["x_goal"], width=14.0)
    if kind == "nirrt":
        pl = NIRRTStarPNG2D(problem["x_start"], problem["x_goal"], 10, problem["search_radius"], iter_max,
                            problem["env_dict"], w, problem["binary_mask"], 3, 2048, 5, pc_sample_rate, ratio)
    else:
        pl = NRRTStarPNG2D(problem["x_start"], problem["x_goal"], 10, problem["search_radius"], iter_max,
                           problem["env_dict"], w, problem["binary_mask"], 3, 2048, 5, pc_sample_rate)
    finish(pl, w, kind, env_idx, seed, iter_max, mode, iter_after, pc_sample_rate, ratio, 2)


def run_case(kind, env_idx, seed, iter_max, mode, iter_after=0, pc_sample_rate=0.5, ratio=0.9):
    problem = make_problem_3d(env_idx)
    np.random.seed(seed); random.seed(seed)
    w = RecordingWrapper(problem["x_start"], problem["x_goal"])
    if kind == "nirrt":
        pl = NIRRTStarPNG3D(problem["x_start"], problem["x_goal"], 10, problem["search_radius"], iter_max,
                            problem["env_dict"], w, 2, 2048, 5, pc_sample_rate, ratio)
    else:
        pl = NRRTStarPNG3D(problem["x_start"], problem["x_goal"], 10, problem["search_radius"], iter_max,
                           problem["env_dict"], w, 2, 2048, 5, pc_sample_rate)
    finish(pl, w, kind, env_idx, seed, iter_max, mode, iter_after, pc_sample_rate, ratio, 3)


def finish(pl, w, kind, env_idx, seed, iter_max, mode, iter_after, pc_sample_rate, ratio, dim):
    with contextlib.redirect_stdout(io.StringIO()):
        if mode == "planning":
            pl.planning(False); plist = np.zeros(0)
        else:
            plist = np.array(pl.planning_random(iter_after), dtype=np.float64)
    next_random, next_py_random = np.random.random(), random.random()
    n = pl.num_vertices
    name = f"neural{dim}d_{kind}_{mode}_e{env_idx}_s{seed}_i{iter_max}.npz"
    np.savez_compressed(os.path.join(OUT, name), kind=kind, mode=mode, env_idx=env_idx, seed=seed, iter_max=iter_max,
                        iter_after=iter_after, pc_sample_rate=pc_sample_rate, ratio=ratio,
                        vertices=pl.vertices[:n].copy(), parents=pl.vertex_parents[:n].astype(np.int64), num_vertices=n,
                        path_len_list=plist, solutions=np.array(getattr(pl, "path_solutions", []), dtype=np.int64),
                        path=np.array(pl.path, dtype=np.float64) if len(pl.path) else np.zeros((0, dim)),
                        next_random=next_random, next_py_random=next_py_random, n_calls=len(w.calls), dim=dim,
                        call_pc_sha1=np.array([c[0] for c in w.calls]), call_n=np.array([c[1] for c in w.calls]),
                        call_mask_sha1=np.array([c[2] for c in w.calls]),
                        call_pred=np.stack([np.pad(c[3], (0, 256 - len(c[3]))) for c in w.calls]),
                        **({f"call_pc{k}": c[4] for k, c in enumerate(w.calls)} if dim == 2 else {}))
    print(name, "n", n, "cloud updates", len(w.calls), "finite", int(np.isfinite(plist).sum()) if len(plist) else "-")


if __name__ == "__main__":
    run_case("nirrt", 0, 31, 1500, "random", iter_after=400)
    run_case("nirrt", 2, 32, 1200, "planning")
    run_case("nirrt", 5, 33, 1500, "random", iter_after=300, pc_sample_rate=0.3, ratio=0.95)
    run_case("nrrt", 1, 34, 1500, "random", iter_after=200)
    run_case("nrrt", 3, 35, 700, "planning")
    run_case_2d("nirrt", 1, 41, 1500, "random", iter_after=400)
    run_case_2d("nirrt", 4, 42, 1000, "planning")
    run_case_2d("nirrt", 6, 43, 1500, "random", iter_after=300, pc_sample_rate=0.3, ratio=0.95)
    run_case_2d("nrrt", 1, 44, 1500, "random", iter_after=200)
    run_case_2d("nrrt", 4, 45, 700, "planning")
