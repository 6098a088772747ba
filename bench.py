#!/usr/bin/env python
"""bench.py -- RRT* iterations/s at 100k-node trees on random_3d worlds (BASELINE.json metric).

A "step" is a block of `--iters-per-step` (default 64) lock-step iterations of the RRT* loop body (Sample,
Nearest scan, Steer + collision, Near scan, ChooseParent, Rewire -- rrt_star_3d.py:36-55 of the reference)
over a batch of E independent planning problems per GPU whose trees hold `--nodes` vertices when the timed
window starts; one step = one nirrt_batch_run call = k_top + 4 replays of the 16-iteration CUDA graph that
nirrt_batch_begin built (no graph is ever captured inside the timed window).  Every leg starts from the same
host snapshot of the 100k-vertex trees.  Workload at
N=1: BASELINE.json configs[4] scaled to one GPU (4096 envs / 8 GPUs = 512 envs per GPU, 100k-node
trees); more GPUs shard more problems (weak scaling), with one NCCL gather of per-problem results.

  python bench.py [--gpus N] [--steps K] [--warmup W]          our CUDA path
  python bench.py --impl reference ...                          CPU arm: numpy port of the reference
                                                                loop body on all host cores

Prints ONE JSON line (see the keys below).  Trees are grown to size by the CUDA planner itself
before the timed region (parity-tested path); inputs are larger than L2 (E x n x 6 B of u16 mirror
coordinates = 0.31 GB streamed per step at the default size), so no L2 flush is needed between steps.
"""
import argparse
import json
import multiprocessing as mp
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "RRT* iters/sec at 100k nodes (random_3d)"
UNIT = "env-iters/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--iters-per-step", type=int, default=64, help="lock-step iterations per step (both arms)")
    ap.add_argument("--parity-iters", type=int, default=128,
                    help="iterations the CPU baseline and the GPU both continue from the snapshot before their trees are compared")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=512, help="planning problems per GPU")
    ap.add_argument("--nodes", type=int, default=100000, help="tree size at the start of the timed window")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU baseline sample length per core")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-pipeline", type=int, default=1,
                    help="e2e leg: also run the job cut into this many planners on their own streams / host threads so that copies "
                         "overlap iterations (measured with 4: 172 ms against 169 ms serial -- four concurrent lock-step planners "
                         "iterate 1.5x slower in aggregate than one, which eats the hidden copies; off by default)")
    ap.add_argument("--profile-range", action="store_true",
                    help="wrap the timed core region in cudaProfilerStart/Stop (ncu --profile-from-start off)")
    ap.add_argument("--core-only", action="store_true", help="skip roofline/eval/e2e/cpu legs (profiling runs)")
    ap.add_argument("--clouds", type=int, default=256, help="PointNet++ leg: clouds per batch per GPU (2048 points each)")
    ap.add_argument("--pn2-steps", type=int, default=20, help="PointNet++ leg: timed forwards")
    ap.add_argument("--no-pointnet2", action="store_true")
    ap.add_argument("--nirrt-envs", type=int, default=512, help="NIRRT* leg: planning problems per GPU (BASELINE configs[3]/[4] shape)")
    ap.add_argument("--nirrt-envs-2d", type=int, default=256, help="2D NIRRT* leg: planning problems per GPU (BASELINE configs[2])")
    ap.add_argument("--nirrt-iter-max", type=int, default=10000)
    ap.add_argument("--nirrt-iter-after", type=int, default=5000)
    ap.add_argument("--nirrt-reps", type=int, default=2, help="NIRRT* legs: repetitions (the fastest is reported, all are listed)")
    ap.add_argument("--no-nirrt", action="store_true")
    ap.add_argument("--no-small", action="store_true", help="skip the small-tree legs (BASELINE configs[0] and configs[1])")
    ap.add_argument("--small-only", action="store_true", help="run only the small-tree legs (profiling runs)")
    ap.add_argument("--nirrt-only", action="store_true", help="run only the NIRRT* leg (profiling runs)")
    ap.add_argument("--pointnet2-only", action="store_true", help="run only the PointNet++ leg (profiling runs)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# clocks

class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU legs (numpy port of the reference loop body; oracle/ is test infrastructure -- this is one of
# the two places allowed to execute it)

def _cpu_worker_snapshot(args):
    """Continues the numpy port from a tree snapshot: first `parity_iters` iterations (their result is returned for
    the parity check against the GPU continuing from the same snapshot), then on until `seconds` have passed.
    Returns (iterations, elapsed, n, parents, vertices-after-parity_iters)."""
    path, env_idx, seconds, parity_iters = args
    from nirrt_star_b200.synthetic import make_problem_3d
    from oracle.numpy_port import RRTStar3DPort
    d = np.load(path)
    pr = make_problem_3d(env_idx)
    rs = np.random.RandomState(0)
    rs.set_state(("MT19937", d["key"], int(d["pos"]), 0, 0.0))
    n = int(d["n"])
    port = RRTStar3DPort(pr, n + 20000, rng=rs)
    port.load_tree(d["v"][:n], d["p"][:n])
    t0 = time.perf_counter(); it = 0
    for _ in range(parity_iters):
        port.iterate(); it += 1
    pn = int(port.num_vertices)
    pp = np.array(port.vertex_parents[:pn], dtype=np.int64)
    pv = np.array(port.vertices[:pn], dtype=np.float64)
    while time.perf_counter() - t0 < seconds:
        port.iterate(); it += 1
    return it, time.perf_counter() - t0, pn, pp, pv


def cpu_baseline_from_snapshot(E, v, p, n, rng_states, env_base, seconds, parity_iters):
    cores = os.cpu_count() or 1
    workers = min(cores, E)
    tmp = tempfile.mkdtemp(prefix="nirrt_cpu_")
    jobs = []
    for k in range(workers):
        path = os.path.join(tmp, f"env{k}.npz")
        np.savez(path, v=v[k], p=p[k], n=n[k], key=rng_states[k][0], pos=rng_states[k][1])
        jobs.append((path, env_base + k, seconds, parity_iters))
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(workers) as pool:
        res = pool.map(_cpu_worker_snapshot, jobs)
    wall = time.perf_counter() - t0
    rate = sum(r[0] / r[1] for r in res)
    base = {"value": rate, "unit": UNIT, "cores": workers, "kind": "port",
            "sample": f"{workers} problems (one per host core) x ~{seconds:.0f} s of the numpy port of the reference "
                      f"loop body, continuing from the same {int(n[0])}-vertex GPU-grown trees and RNG states; "
                      f"{sum(r[0] for r in res)} iterations in {wall:.0f} s wall",
            "per_core": rate / workers}
    return base, [(r[2], r[3], r[4]) for r in res]


def _ref_worker(conn, env_idx, seed, nodes):
    """--impl reference worker: grows its own tree with the C oracle (untimed), then executes
    `count` numpy-port iterations per request."""
    from nirrt_star_b200.synthetic import make_problem_3d
    from oracle.numpy_port import RRTStar3DPort
    from oracle.planner_oracle import Oracle3D
    pr = make_problem_3d(env_idx)
    cap_iters = int(nodes * 1.6) + 50000
    o = Oracle3D(pr, cap_iters, seed=seed)
    while o.num_vertices < nodes:
        o.run(min(2000, max(1, nodes - o.num_vertices)), 0, 0)
    v, p = o.tree()
    key, pos = o.rng_state()
    rs = np.random.RandomState(0)
    rs.set_state(("MT19937", key, pos, 0, 0.0))
    port = RRTStar3DPort(pr, len(v) + 200000, rng=rs)
    port.load_tree(v, p)
    del o
    t0 = time.perf_counter(); port.run(5); rate = 5 / (time.perf_counter() - t0)
    conn.send(("ready", rate, port.num_vertices))
    while True:
        msg = conn.recv()
        if msg == "stop":
            break
        t0 = time.perf_counter()
        port.run(msg)
        conn.send(time.perf_counter() - t0)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    procs, conns = [], []
    for k in range(cores):
        a, b = ctx.Pipe()
        pr = ctx.Process(target=_ref_worker, args=(b, k, 1000 + k, args.nodes), daemon=True)
        pr.start(); procs.append(pr); conns.append(a)
    rates = []
    for c in conns:
        tag, rate, n = c.recv(); rates.append(rate)
    # One step = the same block of iterations per problem as our arm's step (--iters-per-step), one problem per
    # host core.  Only if that would take more than ~4 minutes in total is the block shortened (and reported).
    per_step = args.iters_per_step
    budget_s = 240.0
    if per_step * (args.steps + args.warmup) / max(1e-9, min(rates)) > budget_s:
        per_step = max(1, int(budget_s * min(rates) / max(1, args.steps + args.warmup)))

    def step():
        for c in conns:
            c.send(per_step)
        for c in conns:
            c.recv()

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    el = time.perf_counter() - t0
    for c in conns:
        c.send("stop")
    value = cores * per_step * args.steps / el
    sample = (f"{cores} processes (one per host core), each a {args.nodes}-vertex RRT* tree grown by the C oracle; "
              f"one step = {per_step} numpy-port iterations per process")
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": f"rrt_star 3D random_3d, {args.nodes}-node trees, CPU numpy port of the reference loop body",
                      "nodes": args.nodes, "iters_per_step": per_step, "problems": cores,
                      "iters_per_step_requested": args.iters_per_step},
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------
# PointNet++ leg: clouds/s at 2048 points (second half of BASELINE.json's metric)

PN2_GMAC_SA = 1.040e9      # MACs per 2048-point cloud in the SA MLPs
PN2_GMAC_ALL = 1.380e9     # whole network


def _pn2_cpu_baseline(seconds):
    """The torch-fp32 oracle (== the reference's forward, see oracle/pointnet2_oracle.py) on all host
    cores, batch 1 as the reference's planner calls it."""
    import torch
    from nirrt_star_b200.synthetic import make_cloud_3d, make_pointnet2_state
    from oracle import pointnet2_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in make_pointnet2_state(0).items()}
    pc, sm, gm = make_cloud_3d(0)
    O.classify_path_points(sd, pc, sm, gm, [0, 0, 0, 0])
    t0 = time.perf_counter(); n = 0
    while time.perf_counter() - t0 < seconds:
        O.classify_path_points(sd, pc, sm, gm, [n % 2048, n % 1024, n % 256, n % 64]); n += 1
    el = time.perf_counter() - t0
    return {"value": n / el, "unit": "clouds/s", "cores": cores, "kind": "port",
            "sample": f"{n} single-cloud forwards of the torch-fp32 oracle (the reference's forward restated) in {el:.1f} s, "
                      f"torch.set_num_threads({cores})"}


def bench_pointnet2(args, world, rank, local, peaks, cpu=True):
    import torch
    import torch.distributed as dist
    from nirrt_star_b200.pointnet2 import PointNet2Engine
    from nirrt_star_b200.synthetic import make_cloud_3d, make_pointnet2_state
    Bc, N, K, W = args.clouds, 2048, args.pn2_steps, 3
    eng = PointNet2Engine(make_pointnet2_state(0), n_points=N, max_batch=Bc, device=local)
    base = [make_cloud_3d(i) for i in range(16)]
    rs = np.random.RandomState(rank)
    pc = np.stack([base[i % 16][0] + rs.uniform(-0.01, 0.01, (N, 3)).astype(np.float32) for i in range(Bc)])
    sm = np.stack([base[i % 16][1] for i in range(Bc)]); gm = np.stack([base[i % 16][2] for i in range(Bc)])
    fs = np.stack([rs.randint(0, n, Bc) for n in (2048, 1024, 256, 64)], 1).astype(np.int32)
    d_pc, d_sm, d_gm = (torch.from_numpy(a).cuda() for a in (pc, sm, gm))
    d_fs = torch.from_numpy(fs).cuda()
    d_pred = torch.empty((Bc, N), dtype=torch.int64, device="cuda")
    d_score = torch.empty((Bc, N), dtype=torch.float32, device="cuda")

    def fwd():
        eng.classify_device(Bc, 3, d_pc.data_ptr(), d_sm.data_ptr(), d_gm.data_ptr(), d_fs.data_ptr(),
                            d_pred.data_ptr(), d_score.data_ptr())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        fwd()
    barrier()
    l0 = eng.launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if args.profile_range:
        torch.cuda.profiler.start()
    ev0.record()
    for _ in range(K):
        fwd()
    ev1.record()
    barrier()
    if args.profile_range:
        torch.cuda.profiler.stop()
    ms = ev0.elapsed_time(ev1)
    launches = eng.launches() - l0
    # end to end through the host-buffer entry point (pinned host -> HBM -> pinned host inside)
    pin = [torch.from_numpy(a).pin_memory().numpy() for a in (pc, sm, gm)]
    h_out = (torch.empty((Bc, N), dtype=torch.int64, pin_memory=True).numpy(),
             torch.empty((Bc, N), dtype=torch.float32, pin_memory=True).numpy())
    eng.classify(pin[0], pin[1], pin[2], fps_start=fs, out=h_out)
    barrier()
    t0 = time.perf_counter()
    reps = max(2, K // 4)
    for _ in range(reps):
        eng.classify(pin[0], pin[1], pin[2], fps_start=fs, out=h_out)
    barrier()
    e2e_s = (time.perf_counter() - t0) / reps
    if world > 1:
        t = torch.tensor([ms, e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s = float(t[0].item()), float(t[1].item())
    # stage attribution (event bracket + sync around each stage group)
    eng.set_profiling(True)
    fwd(); torch.cuda.synchronize()             # the serial path launches kernel variants the overlapped path never uses:
    fwd(); torch.cuda.synchronize()             # their first launch loads the module lazily (milliseconds), so time the second
    stages = eng.stage_ms()
    eng.set_profiling(False)
    peak = float(peaks.get("bf16_tflops", 1590.0))
    sa_tflops = 2 * PN2_GMAC_SA * Bc / (stages["sa_mlp"] * 1e-3) / 1e12
    out = {"metric": "PointNet++ clouds/sec @2048pts", "value": world * Bc * K / (ms / 1e3), "unit": "clouds/s",
           "ms_per_step": ms / K, "steps": K, "warmup": W, "dtype": "f16 operands, f32 accumulate",
           "config": {"workload": f"pointnet2 MSG sem-seg forward (classify_path_points), {Bc} clouds x {N} points per GPU "
                                  "(BASELINE configs[2] batch), synthetic checkpoint", "clouds_per_gpu": Bc, "n_points": N},
           "e2e": {"value": world * Bc / e2e_s, "unit": "clouds/s", "h2d_bytes_per_step": float(pc.nbytes + sm.nbytes + gm.nbytes + fs.nbytes),
                   "d2h_bytes_per_step": float(Bc * N * 12), "what": "nirrt_pn2_classify_sync with pinned host buffers"},
           "gpu_launches": int(launches), "stage_ms": stages,
           "roofline": {"bound": "tensor", "kernel": "SA MLPs: safused::k_sa_fused x4 (sa1, sa2: gather + 3 layers + pool per radius) + 12 umma::k_gemm launches (sa3, sa4), tcgen05 kind::f16", "achieved": sa_tflops,
                        "peak": peak, "unit": "TFLOP/s", "frac": sa_tflops / peak, "traffic": None,
                        "peak_source": "MEASURED_PEAKS.json bf16_tflops (of measured)" if "bf16_tflops" in peaks else "fallback 1590 TFLOP/s (of fallback)",
                        "algorithmic_flops_per_step": 2 * PN2_GMAC_SA * Bc}}
    if cpu and rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = _pn2_cpu_baseline(min(10.0, args.cpu_seconds))
    eng.close()
    return out


# ------------------------------------------------------------------------------------------------
# small-tree legs: BASELINE configs[0] (rrt_star 2D, 1 problem, iter_max = 500: the reference's demo_planning_2d.py path)
# and configs[1] (irrt_star 2D, 64 problems batched, iter_max = 5000).  Here the scan is a few microseconds; what is
# measured is the latency of one lock-step iteration (launch + dependency chain of k_expand), the regime every real
# eval configuration of the reference (iter_max <= 10000) lives in.

def _small_cpu_worker(args):
    """numpy oracle of the 2D planners (restates rrt_star_2d.py / irrt_star_2d.py) for `seconds` on one problem"""
    env_idx, seed, iter_max, variant, seconds = args
    from nirrt_star_b200.synthetic import make_problem_2d
    from oracle.planner2d_oracle import Oracle2D
    o = Oracle2D(make_problem_2d(env_idx), iter_max, seed=seed)
    t0 = time.perf_counter(); it = 0
    while it < iter_max and time.perf_counter() - t0 < seconds:
        o.run(min(50, iter_max - it), variant, 0); it += min(50, iter_max - it)
    return it, time.perf_counter() - t0


def bench_small(args, local, cpu=True):
    import torch
    from nirrt_star_b200 import batch as B
    from nirrt_star_b200.synthetic import make_problem_2d
    out = {}
    for name, E, iter_max, variant, reps in (("config0_rrt_star_2d_1x500", 1, 500, B.VARIANT_RRT_STAR, 20),
                                              ("config1_irrt_star_2d_64x5000", 64, 5000, B.VARIANT_IRRT_STAR, 3)):
        problems = [make_problem_2d(300 + i) for i in range(E)]
        best = None
        nv = None
        for r in range(reps + 1):                      # first repetition = warm-up (library, graph build)
            bp = B.BatchPlanner2D(problems, iter_max, seeds=[900 + i for i in range(E)], device=local)
            bp.begin(variant, B.MODE_PLANNING, iter_max)
            torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            bp.run(iter_max)
            ev1.record()
            torch.cuda.synchronize()
            ms = ev0.elapsed_time(ev1)
            if r > 0 and (best is None or ms < best):
                best = ms
            _, _, nv = bp.env_state()
            bp.close()
        leg = {"value": E * iter_max / (best / 1e3), "unit": UNIT, "problems": E, "iter_max": iter_max,
               "us_per_lockstep_iteration": 1e3 * best / iter_max, "ms_total": best, "mean_vertices_at_end": float(nv.mean()),
               "what": "nirrt_batch_run(iter_max) of the planning() loop body, device-resident, CUDA events, best of %d" % reps}
        if cpu:
            cores = os.cpu_count() or 1
            workers = min(cores, E)
            ctx = mp.get_context("spawn")
            with ctx.Pool(workers) as pool:
                res = pool.map(_small_cpu_worker, [(300 + i, 900 + i, iter_max, variant, 8.0) for i in range(workers)])
            rate = sum(it / el for it, el in res)
            leg["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": workers, "kind": "port", "per_core": rate / workers,
                                   "sample": f"{workers} problem(s), one per host core, numpy oracle of the same loop body for up to 8 s or "
                                             f"{iter_max} iterations each ({sum(it for it, _ in res)} iterations)"}
        out[name] = leg
    return out


# ------------------------------------------------------------------------------------------------
# NIRRT* leg: the whole planner of BASELINE configs[3]/[4] -- informed + guidance-cloud sampling, cloud updates by
# batched PointNet++ forwards -- through nirrt_star_b200.eval.plan_batch (what eval_planning_3d.py -p nirrt_star
# -n pointnet2 does one problem at a time, eval_planning_3d.py:101-126)

def bench_nirrt(args, world, rank, local, connect="none", dim=3):
    import torch
    import torch.distributed as dist
    from nirrt_star_b200.eval import default_args, plan_batch
    from nirrt_star_b200.synthetic import make_pointnet2_state, make_problem_2d, make_problem_3d
    E = args.nirrt_envs if dim == 3 else args.nirrt_envs_2d
    sd = make_pointnet2_state(0)
    a = default_args(dim, iter_max=args.nirrt_iter_max, iter_after_initial=args.nirrt_iter_after)
    base = 100000 + rank * E
    problems = [(make_problem_3d if dim == 3 else make_problem_2d)(base + i) for i in range(E)]
    seeds = [base + i for i in range(E)]
    # warm-up on a slice (library load, engine + graph construction paths), untimed
    plan_batch(problems[:max(8, E // 16)], "nirrt_star", dim, default_args(dim, iter_max=300, iter_after_initial=50),
               seeds=seeds[:max(8, E // 16)], state_dict=sd, device=local, connect=connect)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    # the whole call is timed (device buffers, engines, planning, cloud updates, teardown); the host side of it (driver
    # allocations, numpy) is at the mercy of the box's other tenants, so the leg is run args.nirrt_reps times and the
    # fastest repetition is reported, all of them listed
    rep_seconds = []
    el, stats, lists = None, {}, None
    for _ in range(max(1, args.nirrt_reps)):
        st_r = {}
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        lists_r = plan_batch(problems, "nirrt_star", dim, a, seeds=seeds, state_dict=sd, device=local, stats_out=st_r, connect=connect)
        torch.cuda.synchronize()
        el_r = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([el_r], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            el_r = float(t[0])
        rep_seconds.append(el_r)
        if el is None or el_r < el:
            el, stats, lists = el_r, st_r, lists_r
    iters = float(sum(len(x) for x in lists))
    solved = int(sum(1 for x in lists if len(x) and np.isfinite(x[-1])))
    agg = torch.tensor([el, iters, solved, stats.get("cloud_updates", 0), stats.get("forward_calls", 0), stats.get("update_seconds", 0.0)],
                       device="cuda", dtype=torch.float64)
    if world > 1:
        mx = agg.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = agg.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        el = float(mx[0]); iters = float(sm[1]); solved = int(sm[2])
        upd, fwd, upd_s = float(sm[3]), float(sm[4]), float(mx[5])
    else:
        upd, fwd, upd_s = float(agg[3]), float(agg[4]), float(agg[5])
    return {"metric": f"NIRRT* env-iters/sec, whole planner incl. batched cloud updates (random_{dim}d)", "value": iters / el, "unit": UNIT,
            "seconds": el, "repetition_seconds": rep_seconds, "env_iterations": iters,
            "config": {"workload": f"nirrt_star -n pointnet2{' -c bfs' if connect == 'bfs' else ''} {dim}D random_{dim}d (BASELINE {'configs[3]/[4]' if dim == 3 else 'configs[2]'} shape): {E} problems/GPU, "
                                   f"iter_max={a.iter_max}, iter_after_initial={a.iter_after_initial}, 2048-pt clouds (10240 raw samples), "
                                   "synthetic checkpoint", "envs_per_gpu": E},
            "problems_solved": solved, "problems_total": world * E,
            "cloud_updates": upd, "pointnet2_forward_calls": fwd, "clouds_per_forward": upd / max(1.0, fwd),
            "cloud_update_seconds": upd_s, "cloud_update_share_of_time": upd_s / el, "work_rank0": stats.get("work"),
            "update_breakdown_rank0": {k: stats.get(k) for k in ("rounds", "t_params", "t_sample", "t_forward", "t_short", "short_clouds", "t_connect", "heuristic_ties")},
            "time_breakdown_rank0": {k: stats.get(k) for k in ("create_seconds", "setup_seconds", "plan_seconds", "loop_seconds", "run_seconds", "update_seconds", "close_seconds")},
            "what": "wall clock of plan_batch (max over ranks, fastest of the listed repetitions): device-side cloud sampling (MT19937 draws, filters, FPS) + masks + "
                    "ONE PointNet++ forward per lock-step round + commit, interleaved with the lock-step planner iterations"}


# ------------------------------------------------------------------------------------------------
# our arm

def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    from nirrt_star_b200 import batch as B
    from nirrt_star_b200.synthetic import make_problem_3d

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    if args.pointnet2_only:
        pn2 = bench_pointnet2(args, world, rank, local, peaks, cpu=False)
        if rank == 0:
            print(json.dumps(pn2), flush=True)
        return
    if args.small_only:
        sm = bench_small(args, local, cpu=not args.no_cpu_baseline) if rank == 0 else None
        if rank == 0:
            print(json.dumps(sm), flush=True)
        return
    if args.nirrt_only:
        nr = bench_nirrt(args, world, rank, local)
        nc = bench_nirrt(args, world, rank, local, connect="bfs")
        n2 = bench_nirrt(args, world, rank, local, dim=2)
        if rank == 0:
            print(json.dumps({"nirrt_star": nr, "nirrt_star_connect_bfs": nc, "nirrt_star_2d": n2}), flush=True)
        return

    E, nodes, K, W, ips = args.envs, args.nodes, args.steps, args.warmup, args.iters_per_step
    env_base = rank * E
    problems = [make_problem_3d(env_base + i) for i in range(E)]
    seeds = [5000 + env_base + i for i in range(E)]
    slack = (W + K) * ips + 4096
    bp = B.BatchPlanner3D(problems, nodes + slack, seeds=seeds, device=local, record_capacity=(K + W) * ips + 64)

    # ---- untimed: grow every tree to exactly `nodes` vertices with the CUDA planner itself
    t_grow0 = time.perf_counter()
    bp.begin(B.VARIANT_RRT_STAR, B.MODE_PLANNING, 1 << 30)
    bp.set_vertex_limit(nodes)
    while True:
        bp.run(4096)
        _, _, nv = bp.env_state()
        if nv.min() >= nodes:
            break
    bp.set_vertex_limit(0)
    grow_s = time.perf_counter() - t_grow0

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # host snapshot (pinned) of the grown trees + RNG streams: every leg below starts from it
    v_pin = torch.empty((E, bp.capacity, 3), dtype=torch.float64, pin_memory=True)
    p_pin = torch.empty((E, bp.capacity), dtype=torch.int64, pin_memory=True)
    v_np, p_np = v_pin.numpy(), p_pin.numpy()
    _, _, n_np = bp.read_trees(out=(v_np, p_np))
    rng_states = bp.get_rng()

    def restore():
        bp.load_trees(v_np, p_np, n_np)
        bp.set_rng(rng_states)

    # ---- timed: K steps of `ips` iterations each (device-resident)
    def timed_region(variant_mode):
        restore()
        bp.begin(B.VARIANT_RRT_STAR, variant_mode, 1 << 30, 1 << 30)     # builds (or finds) the iteration graph: untimed
        for _ in range(W):
            bp.run(ips)
        _, _, n0 = bp.env_state()
        g0 = bp.graph_stats()
        barrier()
        sampler = ClockSampler(local); sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = bp.kernel_launches()
        if args.profile_range:
            torch.cuda.profiler.start()
        ev0.record()
        for _ in range(K):
            bp.run(ips)
        ev1.record()
        if args.profile_range:
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
        barrier()
        clocks = sampler.stop()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        _, _, n1 = bp.env_state()
        g1 = bp.graph_stats()
        graph = {"built_inside_timed_region": g1["builds"] - g0["builds"], "replays": g1["replays"] - g0["replays"],
                 "fallbacks": g1["fallbacks"]}
        return ms, clocks, bp.kernel_launches() - launches0, n0, n1, graph

    ms_core, clocks, launches, n0, n1, graph_core = timed_region(B.MODE_PLANNING)
    value = world * E * K * ips / (ms_core / 1e3)
    if args.core_only:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "ms_per_step": ms_core / K,
                              "ms_per_iteration": ms_core / K / ips, "graph": graph_core, "core_only": True}))
        return

    # ---- roofline attribution: same iterations with an event bracket around every launch
    prof_iters = min(K * ips, 200)
    prof = bp.run_profiled(prof_iters)
    _, _, n2 = bp.env_state()
    bpv = bp.scan_bytes_per_vertex()      # layout actually scanned: 6 B (u16 mirror), 12 B (f32 mirror) or 24 B (f64) per vertex
    scan_bytes = float(0.5 * (n1.astype(np.float64).sum() + n2.astype(np.float64).sum()) * bpv)
    peak = float(peaks.get("hbm_gbs", 6650.0))
    t_near_ms = prof["nearest"] / prof_iters
    achieved = scan_bytes / (t_near_ms * 1e-3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(f"k_nearest_dram_bytes_per_launch_{bpv}B")
    except Exception:
        pass
    step_ms = sum(prof.values()) / prof_iters
    ms_iter = ms_core / K / ips
    kernel_names = {4: "k_nearest_p<3>: ONE pass over the 10-10-10 packed mirror (4 B / vertex) does the Nearest argmin filter and collects "
                       "the speculative Near ball around x_rand; exact f64 re-check of the few candidates in k_expand",
                    6: "k_nearest_m<3,u16>: ONE pass over the u16 fixed-point mirror does the Nearest argmin filter and collects the "
                       "speculative Near ball around x_rand; exact f64 re-check of the few candidates in k_expand",
                    12: "k_nearest_m<3,f32> (Nearest + speculative Near over the f32 SoA mirror, exact f64 re-check of the candidates)",
                    24: "k_nearest (Nearest argmin scan, f64 SoA)"}
    fused_near = prof["near"] < 0.1 * prof["nearest"]     # mirror modes: no second scan (k_steer proves x_new == x_rand)
    roofline = {"bound": "hbm", "kernel": kernel_names.get(bpv, "k_nearest"), "scan_bytes_per_vertex": bpv, "achieved": achieved,
                "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (of fallback)",
                "algorithmic_bytes_per_launch": scan_bytes,
                "scans_per_iteration": 1 if fused_near else 2,
                "near_scan": None if fused_near else {"achieved": scan_bytes / (prof["near"] / prof_iters * 1e-3) / 1e9,
                                                      "frac": scan_bytes / (prof["near"] / prof_iters * 1e-3) / 1e9 / peak},
                "whole_step_hbm_frac": (1 if fused_near else 2) * scan_bytes / (ms_iter * 1e-3) / 1e9 / peak,
                "kernel_ms_per_iteration": {k: v / prof_iters for k, v in prof.items()},
                "kernel_share_of_iteration": {k: v / prof_iters / step_ms for k, v in prof.items()},
                "attribution": "event bracket around each of k_top / scan / k_steer / k_expand launched UNFUSED and serialised over all "
                               "problems, one launch per iteration; the timed region runs them as 2 fused kernels per iteration on "
                               "overlapped groups of problems, replayed from a CUDA graph"}

    # ---- eval variant (planning_random body: + search_goal_parent / path length every iteration)
    ms_eval, _, _, _, _, graph_eval = timed_region(B.MODE_PLANNING_RANDOM)
    value_eval = world * E * K * ips / (ms_eval / 1e3)

    # ---- e2e through the C ABI with host buffers: one job = trees + RNG streams H2D, K steps each followed by the
    # D2H read of the per-problem status rows a caller polls, then trees + goal parents D2H
    e2e = None
    if not args.no_e2e:
        out_v = torch.empty((E, bp.capacity, 3), dtype=torch.float64, pin_memory=True).numpy()
        out_p = torch.empty((E, bp.capacity), dtype=torch.int64, pin_memory=True).numpy()
        reps = 2
        best = None
        parts = None
        for _ in range(reps):
            barrier()
            t0 = time.perf_counter()
            bp.load_trees(v_np, p_np, n_np)                      # H2D: trees (reference layout)
            bp.set_rng(rng_states)                               # H2D: RNG streams
            t1 = time.perf_counter()
            bp.begin(B.VARIANT_RRT_STAR, B.MODE_PLANNING, 1 << 30)
            for _ in range(K):
                bp.run(ips)
                bp.env_state()                                   # D2H: per-problem state / record count / vertex count
            t2 = time.perf_counter()
            _, _, n_out = bp.read_trees(out=(out_v, out_p))      # D2H: vertices / parents / num_vertices
            gp, cost = bp.goal_parents()                         # D2H: goal parent + path cost per problem
            barrier()
            el = time.perf_counter() - t0
            if best is None or el < best:
                best = el
                parts = {"h2d_trees_s": t1 - t0, "iterations_s": t2 - t1, "d2h_trees_s": t0 + el - t2}
        if world > 1:
            t = torch.tensor([best], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = float(t.item())
        h2d = float((n_np.astype(np.float64) * 32).sum() + E * 625 * 4)
        d2h = float(E * bp.capacity * 32 + E * 16 + K * E * 12)
        serial = {"value": world * E * K * ips / best, "seconds": best, "breakdown": parts}
        # The same job with the copies overlapped with the iterations: the problems are independent, so the batch is cut
        # into P planners (own streams, one host thread each); planner i uploads while planners < i iterate and the
        # finished ones download.  Same calls, same bytes, same results (compared below with the serial job's trees).
        P = max(1, min(args.e2e_pipeline, E))
        piped = None
        if P > 1 and E % P == 0:
            Es = E // P
            streams = [torch.cuda.Stream(device=local) for _ in range(P)]
            subs = [B.BatchPlanner3D(problems[i * Es:(i + 1) * Es], nodes + slack, seeds=seeds[i * Es:(i + 1) * Es], device=local,
                                     record_capacity=(K + W) * ips + 64, stream=streams[i].cuda_stream) for i in range(P)]
            out_v2 = torch.empty((E, bp.capacity, 3), dtype=torch.float64, pin_memory=True).numpy()
            out_p2 = torch.empty((E, bp.capacity), dtype=torch.int64, pin_memory=True).numpy()
            gp2, n2 = [None] * P, [None] * P

            marks = [None] * P

            def job(i, up_done, t_start):
                sl = slice(i * Es, (i + 1) * Es)
                if i > 0:
                    up_done[i - 1].wait()                        # uploads take turns on the link
                m = [time.perf_counter() - t_start]
                subs[i].load_trees(v_np[sl], p_np[sl], n_np[sl])
                subs[i].set_rng(rng_states[sl])
                up_done[i].set()
                m.append(time.perf_counter() - t_start)
                subs[i].begin(B.VARIANT_RRT_STAR, B.MODE_PLANNING, 1 << 30)
                for _ in range(K):
                    subs[i].run(ips)
                    subs[i].env_state()
                m.append(time.perf_counter() - t_start)
                n2[i] = subs[i].read_trees(out=(out_v2[sl], out_p2[sl]))[2]
                gp2[i] = subs[i].goal_parents()
                m.append(time.perf_counter() - t_start)
                marks[i] = [round(x * 1e3, 1) for x in m]

            best2 = None
            for _ in range(reps + 1):                            # the first repetition also builds each planner's graphs
                up_done = [threading.Event() for _ in range(P)]
                barrier()
                t0 = time.perf_counter()
                th = [threading.Thread(target=job, args=(i, up_done, t0)) for i in range(P)]
                for t in th:
                    t.start()
                for t in th:
                    t.join()
                barrier()
                el = time.perf_counter() - t0
                if best2 is None or el < best2:
                    best2, timeline = el, list(marks)
            live = np.arange(bp.capacity)[None, :] < n_out[:, None]          # rows beyond a tree's size are never written
            same = bool(np.array_equal(np.concatenate(n2), n_out) and np.array_equal(out_p2[live], out_p[live])
                        and np.array_equal(out_v2[live], out_v[live]) and np.array_equal(np.concatenate([g[0] for g in gp2]), gp))
            for sb in subs:
                sb.close()
            if world > 1:
                t = torch.tensor([best2], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                best2 = float(t.item())
            piped = {"value": world * E * K * ips / best2, "seconds": best2, "planners": P, "identical_to_serial_job": same,
                     "timeline_ms_upload_start_end_iterations_end_download_end": timeline}
        head = piped if piped is not None and piped["identical_to_serial_job"] and piped["value"] > serial["value"] else serial
        e2e = {"value": head["value"], "unit": UNIT, "h2d_bytes_per_step": h2d / K, "d2h_bytes_per_step": d2h / K,
               "seconds": head["seconds"], "serial": serial, "pipelined": piped,
               "what": "nirrt_batch_load_trees + set_rng (pinned host -> HBM), K x (nirrt_batch_run(iters_per_step) + "
                       "nirrt_batch_env_state_sync), read_trees + goal_parents (HBM -> pinned host); `pipelined`: the batch cut "
                       "into `planners` independent BatchPlanner3D objects on their own streams so that uploads, iterations "
                       "and downloads of different planners overlap (the headline when it reproduces the serial job's trees)"}

    # ---- the one collective: gather per-problem result rows on every rank (NCCL all_gather,
    # nirrt_star_b200/shard.py -- the same code path the gloo CPU tests exercise)
    from nirrt_star_b200.shard import gather_lists, shard_bounds
    assert shard_bounds(world * E, world, rank) == (env_base, env_base + E)
    gp, cost = bp.goal_parents()
    _, _, nv = bp.env_state()
    rows = gather_lists([[float(cost[i]), float(nv[i])] for i in range(E)], world * E,
                        device=torch.device("cuda", local) if world > 1 else None)
    solved = int(sum(1 for r in rows if np.isfinite(r[0])))

    cpu = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        M = args.parity_iters
        cpu, cpu_trees = cpu_baseline_from_snapshot(E, v_np, p_np, n_np, rng_states, env_base, args.cpu_seconds, M)
        # parity at the benchmark size: the GPU continues from the SAME snapshot for the same M iterations
        restore()
        bp.begin(B.VARIANT_RRT_STAR, B.MODE_PLANNING, 1 << 30)
        bp.run(M)
        gv, gpar, gn = bp.read_trees(env_begin=0, count=len(cpu_trees))
        same = 0
        for k, (cn, cp, cv) in enumerate(cpu_trees):
            if gn[k] == cn and np.array_equal(gpar[k, :cn], cp) and np.array_equal(gv[k, :cn], cv):
                same += 1
        parity = {"problems": len(cpu_trees), "iterations": M, "vertices_at_start": int(n_np[0]),
                  "identical": same == len(cpu_trees), "identical_problems": same,
                  "what": "parents and vertices (bit-exact) of the CUDA planner vs the numpy port, both continued from the same "
                          "GPU-grown snapshot and RNG states"}

    pn2 = None
    nirrt = None
    bp.close()
    if not args.no_pointnet2:
        pn2 = bench_pointnet2(args, world, rank, local, peaks)
    nirrt_c = nirrt_2d = None
    if not args.no_nirrt:
        nirrt = bench_nirrt(args, world, rank, local)
        nirrt_c = bench_nirrt(args, world, rank, local, connect="bfs")
        nirrt_2d = bench_nirrt(args, world, rank, local, dim=2)
    small = None
    if not args.no_small and rank == 0 and world == 1:
        small = bench_small(args, local, cpu=not args.no_cpu_baseline)

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
               "ms_per_step": ms_core / K, "ms_per_iteration": ms_core / K / ips, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "f64", "data": "synthetic",
               "config": {"workload": f"rrt_star 3D random_3d (BASELINE configs[4] per-GPU shard): {E} problems/GPU in lock step, "
                                      f"{nodes}-node trees, planning() loop body, {ips} iterations per step",
                          "envs_per_gpu": E, "iters_per_step": ips, "nodes_at_window_start": int(n0.min()), "nodes_at_window_end": int(n1.max()),
                          "l2": "inputs larger than L2: every step streams %.2f GB of mirror coordinates (all problems, all groups) "
                                "through the 126 MB L2, no flush needed" % (scan_bytes / 1e9),
                          "tree_growth": "grown 1 -> %d vertices by the same CUDA planner, untimed (%.0f s)" % (nodes, grow_s),
                          "timing": "CUDA events on the launch stream, barrier + synchronize both sides, max over ranks"},
               "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "graph": graph_core, "roofline": roofline,
               "cpu_baseline": cpu, "parity_at_100k": parity,
               "eval_variant": {"value": value_eval, "unit": UNIT, "ms_per_step": ms_eval / K, "ms_per_iteration": ms_eval / K / ips,
                                "graph": graph_eval,
                                "what": "planning_random loop body (adds goal-parent search + path length per iteration)"},
               "problems_with_solution": solved, "problems_total": world * E, "pointnet2": pn2, "nirrt_star": nirrt, "nirrt_star_connect_bfs": nirrt_c, "nirrt_star_2d": nirrt_2d, "small_trees": small}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
