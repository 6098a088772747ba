"""One persistent-kernel run in the small-tree regime for ncu: python profiles/tools/run_small.py [dim] [envs] [iters] [variant]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from nirrt_star_b200 import batch as B  # noqa: E402
from nirrt_star_b200.synthetic import make_problem_2d, make_problem_3d  # noqa: E402

dim = int(sys.argv[1]) if len(sys.argv) > 1 else 3
E = int(sys.argv[2]) if len(sys.argv) > 2 else 512
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
variant = int(sys.argv[4]) if len(sys.argv) > 4 else 0
mk = make_problem_3d if dim == 3 else make_problem_2d
cls = B.BatchPlanner3D if dim == 3 else B.BatchPlanner2D
bp = cls([mk(300 + i) for i in range(E)], iters, seeds=[900 + i for i in range(E)],
         near_capacity=B.NEAR_CAPACITY_INFORMED if variant in B.INFORMED else 0)
bp.begin(variant, B.MODE_PLANNING, iters)
bp.run(iters)
print(bp.env_state()[2].mean(), bp.work_stats())
