#!/usr/bin/env python
"""Tuning sweep for the planner step (not a bench line): grows the trees ONCE with the CUDA planner,
snapshots them to the host, then re-creates the batch under each (NIRRT_GROUPS, NIRRT_CHUNKS)
setting, re-loads the same trees/RNG states and times K lock-step iterations with CUDA events.

  python profiles/tools/sweep_planner.py --envs 512 --nodes 100000 --steps 300 \
      --configs 1:0,2:0,4:0,8:0,4:5,4:20
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=512)
    ap.add_argument("--nodes", type=int, default=100000)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--configs", default="1:0,2:0,4:0,8:0")
    ap.add_argument("--profiled", action="store_true", help="also print the per-kernel event bracket (serialised launches)")
    args = ap.parse_args()

    import torch
    from nirrt_star_b200 import batch as B
    from nirrt_star_b200.synthetic import make_problem_3d

    E, nodes, K, W = args.envs, args.nodes, args.steps, args.warmup
    problems = [make_problem_3d(i) for i in range(E)]
    seeds = [5000 + i for i in range(E)]
    cap = nodes + 8 * (K + W) + 4096
    bp = B.BatchPlanner3D(problems, cap, seeds=seeds, record_capacity=64)
    bp.begin(B.VARIANT_RRT_STAR, B.MODE_PLANNING, 1 << 30)
    bp.set_vertex_limit(nodes)
    while True:
        bp.run(4000)
        _, _, nv = bp.env_state()
        if nv.min() >= nodes:
            break
    v_pin = torch.empty((E, bp.capacity, 3), dtype=torch.float64, pin_memory=True)
    p_pin = torch.empty((E, bp.capacity), dtype=torch.int64, pin_memory=True)
    v_np, p_np = v_pin.numpy(), p_pin.numpy()
    _, _, n_np = bp.read_trees(out=(v_np, p_np))
    rng = bp.get_rng()
    bp.close()
    del bp

    for cfg in args.configs.split(","):
        g, c = cfg.split(":")
        os.environ["NIRRT_GROUPS"] = g
        if int(c) > 0:
            os.environ["NIRRT_CHUNKS"] = c
        else:
            os.environ.pop("NIRRT_CHUNKS", None)
        bp = B.BatchPlanner3D(problems, cap, rng_states=rng, record_capacity=64)
        bp.load_trees(v_np, p_np, n_np)
        bp.begin(B.VARIANT_RRT_STAR, B.MODE_PLANNING, 1 << 30, 1 << 30)
        bp.run(W)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        import time
        t0 = time.perf_counter()
        bp.run(K)
        host_ms = (time.perf_counter() - t0) * 1e3 / K      # host time to ENQUEUE one step (no sync)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / K
        out = {"groups": int(g), "chunks": int(c), "ms_per_step": ms, "env_iters_per_s": E / (ms * 1e-3), "host_enqueue_ms_per_step": host_ms}
        if args.profiled:
            prof = bp.run_profiled(100)
            out["kernel_us"] = {k: round(1e3 * x / 100, 2) for k, x in prof.items()}
        _, _, nv = bp.env_state()
        out["n_end"] = int(nv.max())
        print(json.dumps(out), flush=True)
        bp.close()
        del bp


if __name__ == "__main__":
    main()
