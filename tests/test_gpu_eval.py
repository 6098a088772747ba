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
