"""Times the Nearest scan kernel alone (nirrt_batch_time_scan_sync: all problems in one launch, CUDA events) for the
scan variants selected by environment variables, on the same 512 x 100k-vertex trees (grown once, re-loaded per variant).
usage: python profiles/tools/scan_variants.py [envs] [nodes]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from nirrt_star_b200 import batch as B  # noqa: E402
from nirrt_star_b200.synthetic import make_problem_3d  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 512
nodes = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
problems = [make_problem_3d(i) for i in range(E)]
bp = B.BatchPlanner3D(problems, nodes + 64, seeds=[5000 + i for i in range(E)])
bp.begin(0, 0, 1 << 30)
bp.set_vertex_limit(nodes)
while True:
    bp.run(4096)
    if bp.env_state()[2].min() >= nodes:
        break
v, p, n = bp.read_trees()
bp.close()
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6545.6
out = []
variants = [("u16 LDG", {"NIRRT_SCAN": "u16ldg"}), ("u16 LDG pipelined", {"NIRRT_TMA": "4"}), ("u16 TMA 3 stages x 4 CTAs", {"NIRRT_TMA": "1"}), ("u16 TMA 2 x 6", {"NIRRT_TMA": "2"}),
            ("u16 TMA 2 x 8", {"NIRRT_TMA": "3"}), ("f32 LDG", {"NIRRT_SCAN": "f32"}), ("u8 dp4a", {"NIRRT_SCAN": "u8"}), ("u8 SAD", {"NIRRT_SCAN": "s8"})]
for chunks in ("10", "5", "20"):
    for name, env in variants:
        if chunks != "10" and name.startswith(("f32", "u8", "u16 TMA")):
            continue
        for k in ("NIRRT_SCAN", "NIRRT_TMA"):
            os.environ.pop(k, None)
        os.environ.update(env)
        os.environ["NIRRT_CHUNKS"] = chunks
        b = B.BatchPlanner3D(problems, nodes + 64, seeds=list(range(E)))
        b.load_trees(v, p, n)
        b.begin(0, 0, 1 << 30)
        b.run(2)
        ms, nbytes = b.time_scan(0, reps=30)
        row = {"variant": name, "chunks": int(chunks), "us": ms * 1e3, "GB/s": nbytes / ms / 1e6, "frac": nbytes / ms / 1e6 / peak,
               "bytes_per_vertex": b.scan_bytes_per_vertex()}
        print(json.dumps(row), flush=True)
        out.append(row)
        b.close()
