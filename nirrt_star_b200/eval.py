"""Batched evaluation front-end: what eval_planning_{2d,3d}.py does one problem at a time
(eval_planning_3d.py:101-126: build planner, `planning_random(iter_after_initial)`, store
`path_len_list`), done for a whole list of problems in lock step on the GPU and sharded over ranks.

``plan_batch`` returns, for every problem, exactly the list the reference's
``get_path_planner(args, problem, wrapper).planning_random(iter_after_initial)`` returns under the
per-problem seeding convention of SURVEY.md 8c (`np.random.seed(s); random.seed(s);
torch.manual_seed(s)` with ``s = seeds[i]`` right before the planner is constructed) -- the drop-in
classes are the single-problem special case of this driver and the GPU tests compare the two.

Neural planners (nrrt_star / nirrt_star): guidance clouds are generated per problem on that problem's
own numpy stream (host, same code as the drop-in classes), classified in ONE batched PointNet++
forward for all problems that wait for a cloud at the same lock-step iteration, and uploaded.
"""
import time
import types

import numpy as np

from . import batch as _B
from .shard import gather_lists, shard_bounds

VARIANTS = {"rrt_star": _B.VARIANT_RRT_STAR, "irrt_star": _B.VARIANT_IRRT_STAR,
            "nrrt_star": _B.VARIANT_NRRT_STAR, "nirrt_star": _B.VARIANT_NIRRT_STAR}


def default_args(dim, **kw):
    """The argparse defaults of eval_planning_{2d,3d}.py:10-31 that reach the planners."""
    a = dict(step_len=10, iter_max=30000, clearance=3 if dim == 2 else 2, pc_n_points=2048, pc_over_sample_scale=5,
             pc_sample_rate=0.5, pc_update_cost_ratio=0.9, iter_after_initial=5000, connect_max_trial_attempts=5)
    a.update(kw)
    return types.SimpleNamespace(**a)


class _CloudMaker:
    """update_point_cloud (nirrt_star_png_{2d,3d}.py:132-174) for one problem, split in two halves
    around the network call so that the forwards of many problems can be batched."""

    def __init__(self, dim, problem, args):
        from . import dropin
        dropin.install()
        self.dim, self.args = dim, args
        self.x_start = np.array(problem["x_start"]).astype(np.float64)
        self.x_goal = np.array(problem["x_goal"]).astype(np.float64)
        if dim == 3:
            from path_planning_utils_3d.rrt_env_3d import Env
            self.env = Env(problem["env_dict"])
        else:
            self.mask = problem["binary_mask"]

    def sample(self, cmax, cmin):
        a = self.args
        if self.dim == 3:
            from datasets_3d.point_cloud_mask_utils_3d import ellipsoid_point_cloud_sampling_3d, generate_rectangle_point_cloud_3d
            if cmax < np.inf:
                pc = ellipsoid_point_cloud_sampling_3d(self.x_start, self.x_goal, cmax / cmin, self.env, a.pc_n_points,
                                                       n_raw_samples=a.pc_n_points * a.pc_over_sample_scale)
            else:
                pc = generate_rectangle_point_cloud_3d(self.env, a.pc_n_points, over_sample_scale=a.pc_over_sample_scale)
        else:
            from datasets.point_cloud_mask_utils import ellipsoid_point_cloud_sampling, generate_rectangle_point_cloud
            if cmax < np.inf:
                pc = ellipsoid_point_cloud_sampling(self.x_start, self.x_goal, cmax / cmin, self.mask, a.pc_n_points,
                                                    n_raw_samples=a.pc_n_points * a.pc_over_sample_scale)
            else:
                pc = generate_rectangle_point_cloud(self.mask, a.pc_n_points, a.pc_over_sample_scale)
        from datasets.point_cloud_mask_utils import get_point_cloud_mask_around_points
        sm = get_point_cloud_mask_around_points(pc, self.x_start[np.newaxis, :], a.step_len)
        gm = get_point_cloud_mask_around_points(pc, self.x_goal[np.newaxis, :], a.step_len)
        return pc, sm.astype(np.float32), gm.astype(np.float32)


def plan_batch(problems, planner, dim, args=None, seeds=None, state_dict=None, classify=None, device=0, chunk=256,
               distributed=False, return_planner=False, host_clouds=False, stats_out=None, connect="none"):
    """path_len_list of every problem (global order on every rank when ``distributed``).

    planner: 'rrt_star' | 'irrt_star' | 'nrrt_star' | 'nirrt_star'
    state_dict: PointNet++ ``model_state_dict`` for the neural planners (the sm_100a engine is built
        from it), or pass ``classify(list_of_(pc, start_mask, goal_mask)) -> list_of_path_pred`` to
        supply predictions some other way (tests replay recorded ones).
    connect: 'none' | 'bfs' -- Neural Connect (nirrt_star_png_c_3d.py:50-84, pointnet2_wrapper_connect_bfs.py:66-233): up to
        args.connect_max_trial_attempts network calls per cloud update, the calls of one trial batched over all waiting
        problems, each followed by the start->goal / goal->start searches over the predicted points' r-disc graph (CUDA)
        and the boundary-point heuristic.  Needs device-sampled clouds.
    host_clouds: run the per-problem drop-in samplers one cloud at a time instead of the batched device update.
    stats_out: dict that receives {iterations, cloud_updates, forward_calls, clouds_classified, short_clouds,
        update_seconds, plan_seconds} of this rank's shard.
    """
    import torch
    t_enter = time.perf_counter()
    args = default_args(dim) if args is None else args
    variant = VARIANTS[planner]
    n_total = len(problems)
    seeds = list(range(n_total)) if seeds is None else list(seeds)
    rank, world = 0, 1
    if distributed:
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
    b, e = shard_bounds(n_total, world, rank)
    local, lseeds = problems[b:e], seeds[b:e]
    lists = []
    bp = None
    engines = []
    if local:
        E = len(local)
        cls = _B.BatchPlanner3D if dim == 3 else _B.BatchPlanner2D
        # informed planners concentrate the tree in a thin ellipsoid: one Near ball can hold thousands of vertices there
        bp = cls(local, args.iter_max, step_len=args.step_len, clearance=args.clearance, seeds=lseeds, device=device,
                 record_capacity=args.iter_max + args.iter_after_initial + 8,
                 near_capacity=_B.NEAR_CAPACITY_INFORMED if variant in _B.INFORMED else 0)
        t_created = time.perf_counter()
        neural = variant in (_B.VARIANT_NIRRT_STAR, _B.VARIANT_NRRT_STAR)
        if neural:
            makers = [_CloudMaker(dim, p, args) for p in local]
            gens = [torch.Generator().manual_seed(int(s)) for s in lseeds]
            engine = None
            from .pointnet2 import NPOINTS, PointNet2Engine
            from .dropin import install as _install
            _install()
            from datasets.point_cloud_mask_utils import get_point_cloud_mask_around_points as get_mask
            short_engines = {}
            if classify is None:
                if state_dict is None:
                    raise ValueError("neural planners need state_dict= or classify=")
                engine = PointNet2Engine(state_dict, n_points=args.pc_n_points, max_batch=E, device=device)
                engines.append(engine)

                def classify(items, envs):
                    # torch.randint on each problem's own generator, in problem order (pointnet2_utils.py:77)
                    fs = np.stack([[int(torch.randint(0, n, (1,), generator=gens[env], dtype=torch.long)) for n in (len(it[0]),) + NPOINTS]
                                   for it, env in zip(items, envs)]).astype(np.int32)
                    preds = [None] * len(items)
                    full = [k for k, it in enumerate(items) if len(it[0]) == args.pc_n_points]
                    if full:            # every full-size cloud of this lock-step round in ONE forward
                        pred, _ = engine.classify(np.stack([items[k][0].astype(np.float32) for k in full]),
                                                  np.stack([items[k][1] for k in full]), np.stack([items[k][2] for k in full]),
                                                  fps_start=fs[full])
                        for k, pr in zip(full, pred):
                            preds[k] = pr
                    # clouds with fewer free points than pc_n_points (large informed ellipsoid at the world border, heavy
                    # clutter): the reference only down-samples `if len(point_cloud) > n_points` and classifies whatever
                    # is left (point_cloud_mask_utils_3d.py:104-112) -- one engine per distinct size, kept for the run
                    for k, it in enumerate(items):
                        if preds[k] is None:
                            if "e" not in short_engines:      # one spare engine, re-targeted to each short cloud's size
                                short_engines["e"] = PointNet2Engine(state_dict, n_points=args.pc_n_points, max_batch=1, device=device)
                                engines.append(short_engines["e"])
                            short_engines["e"].set_n_points(len(it[0]))
                            pred, _ = short_engines["e"].classify(it[0].astype(np.float32)[None], it[1][None], it[2][None], fps_start=fs[k:k + 1])
                            preds[k] = pred[0]
                    return preds
            else:
                user = classify

                def classify(items, envs):
                    return user([(it[0].astype(np.float32), it[1], it[2]) for it in items], envs)

            bp.set_guidance(args.pc_sample_rate, args.pc_update_cost_ratio if variant == _B.VARIANT_NIRRT_STAR else 0.0)
            stats = {"cloud_updates": 0, "forward_calls": 0, "clouds_classified": 0, "short_clouds": 0, "update_seconds": 0.0,
                     "rounds": 0, "t_params": 0.0, "t_sample": 0.0, "t_forward": 0.0, "t_short": 0.0}

            def update_host(envs, cbest, cmin):
                """host half of update_point_cloud for the listed problems, one batched forward"""
                states = bp.get_rng()
                items = []
                for env in envs:
                    np.random.set_state(("MT19937", states[env][0], states[env][1], 0, 0.0))
                    items.append(makers[env].sample(cbest[env], cmin[env]))
                    st = np.random.get_state()
                    states[env] = (st[1], st[2])
                preds = classify(items, list(envs))
                bp.set_rng(states)
                for env, it, pred in zip(envs, items, preds):
                    bp.set_cloud(int(env), it[0][np.asarray(pred).nonzero()[0]])
                stats["forward_calls"] += 1; stats["clouds_classified"] += len(envs)

            if connect not in ("none", "bfs"):
                raise ValueError("connect must be 'none' or 'bfs'")
            if connect == "bfs" and host_clouds:
                raise NotImplementedError("batched Neural Connect needs device-sampled clouds")
            dev = None
            if not host_clouds:
                # the whole update stays in HBM -- draws from each problem's device MT19937 stream, filters, farthest
                # point down-sampling, masks, ONE PointNet++ forward for every waiting problem, and the predicted points go
                # straight into the planner's guidance-cloud buffer.  (With a caller-supplied classify= the clouds are read
                # back for it; sampling still runs on the device.)
                n_pts = args.pc_n_points
                if dim == 2:
                    bp.set_free_masks([p["binary_mask"] for p in local])
                dev = {"pc": torch.empty((E, n_pts, dim), dtype=torch.float32, device=f"cuda:{device}"),
                       "sm": torch.empty((E, n_pts), dtype=torch.float32, device=f"cuda:{device}"),
                       "gm": torch.empty((E, n_pts), dtype=torch.float32, device=f"cuda:{device}"),
                       "pred": torch.empty((E, n_pts), dtype=torch.int64, device=f"cuda:{device}"),
                       "score": torch.empty((E, n_pts), dtype=torch.float32, device=f"cuda:{device}")}

            def connect_round_host(envs, pts, counts, ks):
                """generate_connected_path_points for the listed positions of this round, trial by trial, host side of the
                heuristic (clouds with fewer points than pc_n_points: classified at their own size)"""
                from wrapper.utils.bfs_connect_heuristic import select_heuristic_boundary_point
                from .pointnet2 import connect_analyse_batch
                r = args.step_len
                st = []
                for k in ks:
                    env = envs[k]
                    pc = pts[k, :counts[k]].astype(np.float32)
                    xs, xg = makers[env].x_start.astype(np.float32), makers[env].x_goal.astype(np.float32)
                    st.append({"pc": pc, "xs": xs, "xg": xg, "sm": get_mask(pc, xs[np.newaxis], r), "gm": get_mask(pc, xg[np.newaxis], r),
                               "mask": np.zeros(len(pc), dtype=np.float32), "active": True, "env": env})
                for _ in range(args.connect_max_trial_attempts):
                    act = [q for q in st if q["active"]]
                    if not act:
                        break
                    preds = classify([(q["pc"], q["sm"].astype(np.float32), q["gm"].astype(np.float32)) for q in act], [q["env"] for q in act])
                    stats["forward_calls"] += 1; stats["clouds_classified"] += len(act)
                    for q, pred in zip(act, preds):
                        q["mask"] = ((q["mask"] + pred) > 0).astype(np.float32)
                    # both search directions of every active problem in ONE launch (a path start -> goal exists iff one
                    # goal -> start does, so the second search never depends on the first)
                    hp, _, bnd = connect_analyse_batch([q["pc"] for q in act] * 2, [q["mask"] for q in act] * 2,
                                                       [q["xs"] for q in act] + [q["xg"] for q in act],
                                                       [q["xg"] for q in act] + [q["xs"] for q in act], r)
                    A = len(act)
                    for j, q in enumerate(act):
                        if hp[j] or hp[A + j]:
                            q["active"] = False
                            continue
                        _, bpt, _ = select_heuristic_boundary_point(q["pc"], bnd[j], q["xs"], q["xg"])
                        nsm = q["sm"] if bpt is None else get_mask(q["pc"], bpt, r)
                        _, bpt, _ = select_heuristic_boundary_point(q["pc"], bnd[A + j], q["xg"], q["xs"])
                        ngm = q["gm"] if bpt is None else get_mask(q["pc"], bpt, r)
                        q["sm"], q["gm"] = nsm, ngm
                for k, q in zip(ks, st):
                    bp.set_cloud(int(q["env"]), pts[k, :counts[k]][q["mask"].nonzero()[0]])

            def connect_round(envs, counts):
                """generate_connected_path_points (pointnet2_wrapper_connect_bfs.py:76-240) for every problem of this round.
                Full-size clouds never leave HBM: per trial ONE forward over the round's clouds (predictions of problems
                that are already connected are ignored) and ONE nirrt_connect_trial_device (mask union, both searches,
                boundary heuristic, new neighbourhood masks); the host sees 2 flags per problem and trial."""
                from wrapper.utils.bfs_connect_heuristic import select_heuristic_boundary_point
                from .pointnet2 import connect_masks_device, connect_trial_device
                n, r, R = args.pc_n_points, args.step_len, len(envs)
                if engine is None:      # caller-supplied classifier: it wants host arrays
                    connect_round_host(envs, bp.read_sampled_clouds(0, R, n), counts, list(range(R)))
                    return
                ddev = dev["pc"].device
                full = [k for k in range(R) if counts[k] == n]
                short = [k for k in range(R) if counts[k] != n]
                pts = None
                if short:
                    pts = bp.read_sampled_clouds(0, R, n)
                    connect_round_host(envs, pts, counts, short)
                    stats["short_clouds"] += len(short)
                if not full:
                    return
                src = np.zeros((2 * R, 3), dtype=np.float32); dst = np.zeros((2 * R, 3), dtype=np.float32)
                for k, env in enumerate(envs):
                    xs, xg = makers[env].x_start.astype(np.float32), makers[env].x_goal.astype(np.float32)
                    src[k, :dim] = xs; src[R + k, :dim] = xg; dst[k, :dim] = xg; dst[R + k, :dim] = xs
                d_src, d_dst = torch.from_numpy(src).to(ddev), torch.from_numpy(dst).to(ddev)
                connect_masks_device(dev["pc"].data_ptr(), n, dim, R, d_src.data_ptr(), r, dev["sm"].data_ptr(), dev["gm"].data_ptr())
                acc = torch.zeros((R, n), dtype=torch.uint8, device=ddev)
                active = np.zeros(R, dtype=np.uint8); active[full] = 1
                fs = np.zeros((R, 4), dtype=np.int32)
                for _ in range(args.connect_max_trial_attempts):
                    act = np.nonzero(active)[0]
                    if len(act) == 0:
                        break
                    for k in act:       # torch.randint on each problem's own generator (pointnet2_utils.py:77), active problems only
                        fs[k] = [int(torch.randint(0, m, (1,), generator=gens[envs[k]], dtype=torch.long)) for m in (n,) + NPOINTS]
                    d_fs = torch.from_numpy(fs).to(ddev)
                    t0 = time.perf_counter()
                    engine.classify_device(R, dim, dev["pc"].data_ptr(), dev["sm"].data_ptr(), dev["gm"].data_ptr(), d_fs.data_ptr(),
                                           dev["pred"].data_ptr(), dev["score"].data_ptr())
                    torch.cuda.synchronize()
                    t1 = time.perf_counter()
                    stats["t_forward"] += t1 - t0
                    stats["forward_calls"] += 1; stats["clouds_classified"] += len(act)
                    d_active = torch.from_numpy(active).to(ddev)
                    hp, ties, tb = connect_trial_device(dev["pc"].data_ptr(), n, dim, R, d_active.data_ptr(), acc.data_ptr(),
                                                        dev["pred"].data_ptr(), d_src.data_ptr(), d_dst.data_ptr(), r,
                                                        dev["sm"].data_ptr(), dev["gm"].data_ptr())
                    stats["t_connect"] = stats.get("t_connect", 0.0) + time.perf_counter() - t1
                    for k in act:
                        if hp[k] or hp[R + k]:
                            active[k] = 0
                            continue
                        for j, buf in ((k, dev["sm"]), (R + k, dev["gm"])):
                            if ties[j]:          # equal keys: numpy's argsort decides, as in the reference
                                if pts is None:
                                    pts = bp.read_sampled_clouds(0, R, n)
                                pc = pts[k].astype(np.float32)
                                _, bpt, _ = select_heuristic_boundary_point(pc, tb[j].astype(np.float32), src[j, :dim], dst[j, :dim])
                                buf[k] = torch.from_numpy(get_mask(pc, bpt, r).astype(np.float32)).to(ddev)
                                stats["heuristic_ties"] = stats.get("heuristic_ties", 0) + 1
                dev["pred"][:R] = acc.to(torch.int64)
                bp.commit_clouds(dev["pred"].data_ptr(), None if not short else full)

            def update_device(envs, cbest, cmin):
                n_pts, n_raw = args.pc_n_points, args.pc_n_points * args.pc_over_sample_scale
                ta = time.perf_counter()
                kinds = [1 if cbest[env] < np.inf else 0 for env in envs]
                params = np.zeros((len(envs), 12))
                for k, env in enumerate(envs):
                    if kinds[k]:
                        params[k] = (_B.ellipsoid_params_3d if dim == 3 else _B.ellipsoid_params_2d)(
                            makers[env].x_start, makers[env].x_goal, cbest[env] / cmin[env])
                tb = time.perf_counter()
                counts = bp.sample_clouds(envs, kinds, params, n_pts, n_raw, args.step_len, dev["pc"].data_ptr(),
                                          dev["sm"].data_ptr(), dev["gm"].data_ptr())
                tc = time.perf_counter()
                stats["rounds"] += 1; stats["t_params"] += tb - ta; stats["t_sample"] += tc - tb
                if connect == "bfs":
                    connect_round(envs, counts)
                    return
                if engine is None:          # caller-supplied classifier: hand it the clouds, upload what it predicts
                    pts = bp.read_sampled_clouds(0, len(envs), n_pts)
                    items = []
                    for k, env in enumerate(envs):
                        pc = pts[k, :counts[k]]
                        items.append((pc, get_mask(pc, makers[env].x_start[np.newaxis, :], args.step_len).astype(np.float32),
                                      get_mask(pc, makers[env].x_goal[np.newaxis, :], args.step_len).astype(np.float32)))
                    preds = classify(items, list(envs))
                    for env, it, pred in zip(envs, items, preds):
                        bp.set_cloud(int(env), it[0][np.asarray(pred).nonzero()[0]])
                    stats["forward_calls"] += 1; stats["clouds_classified"] += len(envs)
                    return
                # torch.randint on each problem's own generator, in problem order (pointnet2_utils.py:77)
                fs = np.stack([[int(torch.randint(0, n, (1,), generator=gens[env], dtype=torch.long)) for n in (int(counts[k]),) + NPOINTS]
                               for k, env in enumerate(envs)]).astype(np.int32)
                d_fs = torch.from_numpy(fs).to(dev["pc"].device)
                engine.classify_device(len(envs), dim, dev["pc"].data_ptr(), dev["sm"].data_ptr(), dev["gm"].data_ptr(),
                                       d_fs.data_ptr(), dev["pred"].data_ptr(), dev["score"].data_ptr())
                full = [k for k in range(len(envs)) if counts[k] == n_pts]
                bp.commit_clouds(dev["pred"].data_ptr(), None if len(full) == len(envs) else full)
                torch.cuda.synchronize()
                td = time.perf_counter()
                stats["t_forward"] += td - tc
                stats["forward_calls"] += 1; stats["clouds_classified"] += len(envs)
                for k, env in enumerate(envs):          # short clouds: classified at their own size, like the reference
                    if counts[k] != n_pts:
                        pts = bp.read_sampled_clouds(k, 1, n_pts)[0, :counts[k]]
                        sm = get_mask(pts, makers[env].x_start[np.newaxis, :], args.step_len).astype(np.float32)
                        gm = get_mask(pts, makers[env].x_goal[np.newaxis, :], args.step_len).astype(np.float32)
                        if "e" not in short_engines:
                            short_engines["e"] = PointNet2Engine(state_dict, n_points=n_pts, max_batch=1, device=device)
                            engines.append(short_engines["e"])
                        short_engines["e"].set_n_points(len(pts))
                        pred, _ = short_engines["e"].classify(pts.astype(np.float32)[None], sm[None], gm[None], fps_start=fs[k:k + 1])
                        bp.set_cloud(int(env), pts[pred[0].nonzero()[0]])
                        stats["short_clouds"] += 1; stats["forward_calls"] += 1
                stats["t_short"] += time.perf_counter() - td

            def update(envs, cbest, cmin):
                if args.pc_sample_rate == 0:          # nirrt_star_png_3d.py:137-140: no cloud, no draws
                    for env in envs:
                        bp.set_cloud(int(env), np.zeros((0, dim)))
                    return
                t0 = time.perf_counter()
                (update_device if dev is not None else update_host)(list(envs), cbest, cmin)
                stats["cloud_updates"] += len(envs)
                stats["update_seconds"] += time.perf_counter() - t0

            keep = np.random.get_state()
            if args.pc_sample_rate != 0:
                update(list(range(E)), np.full(E, np.inf), np.full(E, np.nan))        # init_pc()
        t_plan = time.perf_counter()
        bp.begin(variant, _B.MODE_PLANNING_RANDOM, args.iter_max, args.iter_after_initial)
        t_run = 0.0
        while True:
            tr = time.perf_counter()
            bp.run(chunk)
            running, need = bp.status()
            t_run += time.perf_counter() - tr
            if need:
                st, _, _ = bp.env_state()
                cb, cm = bp.c_best()
                update(list(np.nonzero(st == _B.ST_WAIT_CLOUD)[0]), cb, cm)
                continue
            if running == 0:
                break
        if neural:
            np.random.set_state(keep)
        t_loop = time.perf_counter() - t_plan
        lists = bp.path_len_lists()
        if stats_out is not None:
            stats_out.update(stats if neural else {})
            stats_out["plan_seconds"] = time.perf_counter() - t_plan
            stats_out["setup_seconds"] = t_plan - t_enter
            stats_out["create_seconds"] = t_created - t_enter
            stats_out["run_seconds"] = t_run            # nirrt_batch_run + the status read that waits for it
            stats_out["loop_seconds"] = t_loop
            stats_out["iterations"] = int(sum(len(x) for x in lists))
            stats_out["work"] = bp.work_stats()
    out = gather_lists(lists, n_total, device=torch.device("cuda", device)) if distributed else lists
    t_close = time.perf_counter()
    for eng in engines:         # the forward engines of this call (device buffers, tensor maps)
        eng.close()
    if return_planner:
        return out, bp
    if bp is not None:
        bp.close()
    if stats_out is not None:
        stats_out["close_seconds"] = time.perf_counter() - t_close
    return out


def run_eval(env_configs, problems, planner, dim, args=None, seeds=None, state_dict=None, problem_name=None, neural_net="none",
             connect="none", result_root="results/evaluation", batch_size=256, device=0, distributed=False):
    """Batched twin of the eval_planning_{2d,3d}.py main loop (:79-125 of the 3D file): writes
    ``<result_root>/<2d|3d>/<problem>-<planner><-c-connect>-<net>-<N>.pickle`` holding the same
    list of ``env_config`` copies with a ``'result'`` key (``path_len_list``), appended in problem order,
    and RESUMES from an existing pickle by skipping the problems it already holds -- so
    result_analysis_*.py read the files unchanged.  Problems are planned ``batch_size`` at a time in
    lock step (and sharded over ranks when ``distributed``); the pickle is rewritten after every batch
    by rank 0."""
    import os
    import pickle
    from copy import copy
    rank = 0
    if distributed:
        import torch.distributed as dist
        rank = dist.get_rank()
    n = len(env_configs)
    problem_name = problem_name or f"random_{dim}d"
    folder = os.path.join(result_root, f"{dim}d")
    os.makedirs(folder, exist_ok=True)
    connect_str = "" if connect == "none" else "-c-" + connect
    path = os.path.join(folder, f"{problem_name}-{planner}{connect_str}-{neural_net}-{n}.pickle")
    done = []
    if os.path.exists(path):
        with open(path, "rb") as f:
            done = pickle.load(f)
    seeds = list(range(n)) if seeds is None else list(seeds)
    for b in range(len(done), n, batch_size):
        e = min(n, b + batch_size)
        lists = plan_batch(problems[b:e], planner, dim, args, seeds=seeds[b:e], state_dict=state_dict, device=device,
                           distributed=distributed, connect=connect)
        for cfg, lst in zip(env_configs[b:e], lists):
            row = copy(cfg)
            row["result"] = [float(x) for x in lst]
            done.append(row)
        if rank == 0:
            with open(path, "wb") as f:
                pickle.dump(done, f)
    return path, done
