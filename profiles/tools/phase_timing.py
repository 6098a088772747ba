"""k_expand phase times (development build with -DNIRRT_PHASE_TIMING): mean SM clocks per phase and problem-iteration.
Build once on the CPU box (it travels with the snapshot):  python profiles/tools/phase_timing.py --build
Run on the GPU:  NIRRT_LIB_VARIANT=timing python profiles/tools/phase_timing.py [--dim 2|3] [--envs E] [--iters K] [--variant 0|1] [--mode 0|1]"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("--build", action="store_true")
ap.add_argument("--dim", type=int, default=3)
ap.add_argument("--envs", type=int, default=64)
ap.add_argument("--iters", type=int, default=5000)
ap.add_argument("--variant", type=int, default=0)
ap.add_argument("--mode", type=int, default=0)
ap.add_argument("--window", type=int, default=1000)
args = ap.parse_args()
if args.build:
    from nirrt_star_b200 import build
    print(build.build_variant("timing", ["-DNIRRT_PHASE_TIMING"]))
    sys.exit(0)
os.environ["NIRRT_LIB_VARIANT"] = "timing"
import torch  # noqa: E402,F401
from nirrt_star_b200 import _lib, batch as B  # noqa: E402
from nirrt_star_b200.synthetic import make_problem_2d, make_problem_3d  # noqa: E402

L = _lib.lib()
mk = make_problem_3d if args.dim == 3 else make_problem_2d
cls = B.BatchPlanner3D if args.dim == 3 else B.BatchPlanner2D
problems = [mk(300 + i) for i in range(args.envs)]
bp = cls(problems, args.iters, seeds=[900 + i for i in range(args.envs)], record_capacity=args.iters + 8,
         near_capacity=B.NEAR_CAPACITY_INFORMED if args.variant in B.INFORMED else 0)
bp.begin(args.variant, args.mode, args.iters, 1 << 30)
out = (C.c_double * 9)()
names = ("steer", "sort", "filter", "walks+choose", "rewire", "goal", "records", "top(next sample)")
clk_ghz = 1.965
done = 0
while done < args.iters:
    k = min(args.window, args.iters - done)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(); bp.run(k); ev1.record(); torch.cuda.synchronize()
    done += k
    L.nirrt_debug_phase_clocks(out)
    nv = bp.env_state()[2]
    row = {"iterations": done, "mean_vertices": float(nv.mean()), "us_per_lockstep_iteration": 1e3 * ev0.elapsed_time(ev1) / k,
           "expansions": out[8], "phase_us": {n: round(out[i] / clk_ghz / 1e3, 2) for i, n in enumerate(names)}}
    row["phase_sum_us"] = round(sum(row["phase_us"].values()), 2)
    print(json.dumps(row), flush=True)
