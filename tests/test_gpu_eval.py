"""Batched evaluation front-end (nirrt_star_b200/eval.py): a batch of problems planned in lock step
returns, per problem, what the single-problem drop-in planner (and hence the reference, see
test_gpu_dropin*.py) returns under the per-problem seeding convention."""
import glob
import os
import random
import types

import numpy as np
import pytest

from nirrt_star_b200.synthetic import make_pointnet2_state, make_problem_2d, make_problem_3d

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("dim,kind", [(3, "rrt"), (3, "irrt"), (2, "rrt"), (2, "irrt")])
def test_classical_planners_match_reference_golden(dim, kind):
    from nirrt_star_b200.eval import default_args, plan_batch
    files = sorted(glob.glob(os.path.join(GOLD, f"planner{dim}d_{kind}_random_*.npz")))
    gs = [np.load(f) for f in files]
    # one batch per (iter_max, iter_after) pair
    for g in gs:
        mk = make_problem_3d if dim == 3 else make_problem_2d
        others = [mk(20 + k) for k in range(3)]                     # neighbours in the same batch must not matter
        problems = [others[0], mk(int(g["env_idx"])), others[1], others[2]]
        args = default_args(dim, iter_max=int(g["iter_max"]), iter_after_initial=int(g["iter_after"]))
        out = plan_batch(problems, f"{kind}_star", dim, args, seeds=[901, int(g["seed"]), 902, 903])
        lst, want = np.array(out[1]), g["path_len_list"]
        assert len(lst) == len(want) and np.array_equal(np.isinf(lst), np.isinf(want))
        f = np.isfinite(want)
        assert np.allclose(lst[f], want[f], rtol=1e-5, atol=0)


@pytest.mark.parametrize("dim", [3, 2])
def test_batched_nirrt_equals_single_problem_dropin(dim, tmp_path):
    import torch
    from nirrt_star_b200 import dropin
    from nirrt_star_b200.eval import default_args, plan_batch
    dropin.install()
    sd = make_pointnet2_state(0)
    d = tmp_path / f"results/model_training/pointnet2_{dim}d/checkpoints"
    d.mkdir(parents=True)
    torch.save({"model_state_dict": {k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}}, str(d / f"best_pointnet2_{dim}d.pth"))
    if dim == 3:
        from path_planning_classes_3d.nirrt_star_png_3d import get_path_planner
        from wrapper_3d.pointnet_pointnet2.pointnet2_wrapper import PNGWrapper
        problems = [make_problem_3d(i) for i in (0, 2, 5)]
    else:
        from path_planning_classes.nirrt_star_png_2d import get_path_planner
        from wrapper.pointnet_pointnet2.pointnet2_wrapper import PNGWrapper
        problems = [make_problem_2d(i) for i in (1, 4, 6)]
    seeds = [61, 62, 63]
    args = default_args(dim, iter_max=1200, iter_after_initial=150)
    batch = plan_batch(problems, "nirrt_star", dim, args, seeds=seeds, state_dict=sd)
    w = PNGWrapper(root_dir=str(tmp_path), device="cuda")
    for pr, s, got in zip(problems, seeds, batch):
        np.random.seed(s); random.seed(s); torch.manual_seed(s)
        want = get_path_planner(args, pr, w).planning_random(args.iter_after_initial)
        assert len(got) == len(want)
        assert np.array_equal(np.isinf(got), np.isinf(want))
        f = np.isfinite(want)
        assert np.array_equal(np.array(got)[f], np.array(want)[f])


def test_run_eval_writes_reference_format_and_resumes(tmp_path):
    import pickle
    from nirrt_star_b200.eval import default_args, run_eval
    problems = [make_problem_3d(i) for i in range(5)]
    cfgs = [{"env_idx": i, "env_dict": p["env_dict"]} for i, p in enumerate(problems)]
    args = default_args(3, iter_max=400, iter_after_initial=50)
    path, rows = run_eval(cfgs[:5], problems[:5], "irrt_star", 3, args, seeds=[7, 8, 9, 10, 11], result_root=str(tmp_path), batch_size=2)
    assert path.endswith("3d/random_3d-irrt_star-none-5.pickle")
    got = pickle.load(open(path, "rb"))
    assert [r["env_idx"] for r in got] == [0, 1, 2, 3, 4] and all(isinstance(r["result"], list) for r in got)
    # resume: drop the last two problems, rerun -> identical file content
    pickle.dump(got[:3], open(path, "wb"))
    _, rows2 = run_eval(cfgs[:5], problems[:5], "irrt_star", 3, args, seeds=[7, 8, 9, 10, 11], result_root=str(tmp_path), batch_size=2)
    assert [r["result"] for r in rows2] == [r["result"] for r in got]


def test_nirrt_golden_inside_a_batch_device_clouds():
    """The reference's own NIRRT*-PNG 3D trace (tests/golden/neural3d_nirrt_random_e0_s31: recorded predictions,
    SHA-1 of every guidance cloud and mask) reproduced INSIDE a lock-step batch of 8 problems whose clouds are sampled
    on the device (nirrt_batch_sample_clouds_sync: MT19937 draws, filters, farthest point down-sampling): same clouds,
    same tree, same path_len_list -- the neighbours in the batch must not matter."""
    import hashlib
    from nirrt_star_b200.eval import default_args, plan_batch
    g = np.load(os.path.join(GOLD, "neural3d_nirrt_random_e0_s31_i1500.npz"))
    slot = 5
    problems = [make_problem_3d(40 + k) for k in range(8)]
    problems[slot] = make_problem_3d(int(g["env_idx"]))
    seeds = [700 + k for k in range(8)]
    seeds[slot] = int(g["seed"])
    args = default_args(3, iter_max=int(g["iter_max"]), iter_after_initial=int(g["iter_after"]),
                        pc_sample_rate=float(g["pc_sample_rate"]), pc_update_cost_ratio=float(g["ratio"]))
    calls = {"k": 0}

    def classify(items, envs):
        preds = []
        for (pc, sm, gm), env in zip(items, envs):
            if env == slot:
                k = calls["k"]
                assert pc.dtype == np.float32 and len(pc) == int(g["call_n"][k])
                assert hashlib.sha1(np.ascontiguousarray(pc).tobytes()).hexdigest() == str(g["call_pc_sha1"][k]), f"cloud {k} differs"
                assert hashlib.sha1(sm.tobytes() + gm.tobytes()).hexdigest() == str(g["call_mask_sha1"][k])
                preds.append(np.unpackbits(g["call_pred"][k])[:len(pc)].astype(np.int64))
                calls["k"] += 1
            else:       # some deterministic guidance for the neighbours: points near the start-goal segment
                a = np.asarray(problems[env]["x_start"], dtype=np.float64); b = np.asarray(problems[env]["x_goal"], dtype=np.float64)
                t = np.clip(((pc - a) @ (b - a)) / ((b - a) @ (b - a)), 0, 1)
                preds.append((np.linalg.norm(pc - (a + t[:, None] * (b - a)), axis=1) < 12).astype(np.int64))
        return preds

    out, bp = plan_batch(problems, "nirrt_star", 3, args, seeds=seeds, classify=classify, return_planner=True)
    assert calls["k"] == int(g["n_calls"])
    lst, want = np.array(out[slot]), g["path_len_list"]
    assert len(lst) == len(want) and np.array_equal(np.isinf(lst), np.isinf(want))
    f = np.isfinite(want)
    assert np.allclose(lst[f], want[f], rtol=1e-5, atol=0)
    v, p, n = bp.read_trees()
    n = int(n[slot])
    assert n == int(g["num_vertices"]) and np.array_equal(p[slot, :n], g["parents"]) and np.array_equal(v[slot, :n], g["vertices"])
    assert list(bp.solutions(slot)) == list(g["solutions"])
    bp.close()


def test_nirrt_batch_of_64_equals_single_problem_dropin(tmp_path):
    """BASELINE configs[3] shape scaled to a test: 64 NIRRT* problems in lock step with device-side cloud updates and
    batched PointNet++ forwards == the single-problem drop-in planner (reference API, one problem, B = 1 forwards) on 8
    of them, bit for bit; the host-numpy cloud path gives the same answer too."""
    import torch
    from nirrt_star_b200 import dropin
    from nirrt_star_b200.eval import default_args, plan_batch
    dropin.install()
    from path_planning_classes_3d.nirrt_star_png_3d import get_path_planner
    from wrapper_3d.pointnet_pointnet2.pointnet2_wrapper import PNGWrapper
    sd = make_pointnet2_state(0)
    d = tmp_path / "results/model_training/pointnet2_3d/checkpoints"
    d.mkdir(parents=True)
    torch.save({"model_state_dict": {k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}}, str(d / "best_pointnet2_3d.pth"))
    E = 64
    problems = [make_problem_3d(300 + i) for i in range(E)]
    seeds = [4000 + i for i in range(E)]
    args = default_args(3, iter_max=1000, iter_after_initial=200)
    stats = {}
    batch = plan_batch(problems, "nirrt_star", 3, args, seeds=seeds, state_dict=sd, stats_out=stats)
    assert stats["cloud_updates"] >= E and stats["forward_calls"] < stats["cloud_updates"]     # updates really are batched
    w = PNGWrapper(root_dir=str(tmp_path), device="cuda")
    for e in (0, 7, 13, 22, 31, 40, 55, 63):
        s = seeds[e]
        np.random.seed(s); random.seed(s); torch.manual_seed(s)
        want = get_path_planner(args, problems[e], w).planning_random(args.iter_after_initial)
        got = batch[e]
        assert len(got) == len(want), e
        assert np.array_equal(np.isinf(got), np.isinf(want)), e
        f = np.isfinite(want)
        assert np.array_equal(np.array(got)[f], np.array(want)[f]), e
    sub = [0, 7, 13, 22]
    host = plan_batch([problems[e] for e in sub], "nirrt_star", 3, args, seeds=[seeds[e] for e in sub], state_dict=sd, host_clouds=True)
    for k, e in enumerate(sub):
        assert np.array_equal(np.array(host[k]), np.array(batch[e]))


@pytest.mark.parametrize("dim", [3, 2])
def test_batched_neural_connect_equals_single_problem_dropin(dim, tmp_path):
    """-c bfs (BASELINE configs[3]): NIRRT* with Neural Connect in a lock-step batch -- one network call per trial over the
    round's clouds, mask union, both r-disc graph searches, the boundary-point heuristic and the new neighbourhood masks on
    the GPU (nirrt_connect_trial_device) -- equals the single-problem drop-in NIRRTStarPNGC{2,3}D + connect PNGWrapper
    (reference API: nirrt_star_png_c_3d.py:50-84, pointnet2_wrapper_connect_bfs.py:66-233; host numpy heuristic) bit for bit."""
    import torch
    from nirrt_star_b200 import dropin
    from nirrt_star_b200.eval import default_args, plan_batch
    dropin.install()
    if dim == 3:
        from path_planning_classes_3d.nirrt_star_png_c_3d import get_path_planner
        from wrapper_3d.pointnet_pointnet2.pointnet2_wrapper_connect_bfs import PNGWrapper
    else:
        from path_planning_classes.nirrt_star_png_c_2d import get_path_planner
        from wrapper.pointnet_pointnet2.pointnet2_wrapper_connect_bfs import PNGWrapper
    sd = make_pointnet2_state(0)
    d = tmp_path / f"results/model_training/pointnet2_{dim}d/checkpoints"
    d.mkdir(parents=True)
    torch.save({"model_state_dict": {k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}}, str(d / f"best_pointnet2_{dim}d.pth"))
    E = 12
    problems = [(make_problem_3d if dim == 3 else make_problem_2d)(350 + i) for i in range(E)]
    seeds = [5200 + i for i in range(E)]
    args = default_args(dim, iter_max=900, iter_after_initial=150)
    stats = {}
    batch = plan_batch(problems, "nirrt_star", dim, args, seeds=seeds, state_dict=sd, connect="bfs", stats_out=stats)
    assert stats["forward_calls"] >= 2              # several trials happened
    w = PNGWrapper(root_dir=str(tmp_path), device="cuda")
    for e in (0, 3, 5, 8, 11):
        s = seeds[e]
        np.random.seed(s); random.seed(s); torch.manual_seed(s)
        want = get_path_planner(args, problems[e], w).planning_random(args.iter_after_initial)
        got = batch[e]
        assert len(got) == len(want), e
        assert np.array_equal(np.isinf(got), np.isinf(want)), e
        f = np.isfinite(want)
        assert np.array_equal(np.array(got)[f], np.array(want)[f]), e
