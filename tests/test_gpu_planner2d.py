"""GPU parity tests of the 2D planner path: CUDA (through the C ABI) vs fixtures recorded from the
reference's own RRTStar2D / IRRTStar2D / collision_check_utils (tests/golden/make_golden_planner2d.py).

Everything must be exact: index work (nearest, near lists, parents, solutions, RNG consumption) AND vertex
coordinates.  The 2D steer goes through libm's atan2 / cos / sin (rrt_star_2d.py:67-78), which glibc does not round
correctly in ~0.15 % of calls; the device restates glibc's kernels operation by operation
(csrc/glibc_trig.cuh), so vertices are bit-identical and the systematic tie |x_new - x_nearest| == step_len == r in
find_near_neighbors falls the same way as in the reference."""
import glob
import os

import numpy as np
import pytest

from nirrt_star_b200.synthetic import make_problem_2d

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(__file__)
GOLD = sorted(glob.glob(os.path.join(HERE, "golden", "planner2d_*.npz")))
GEOM = sorted(glob.glob(os.path.join(HERE, "golden", "geom2d_*.npz")))


@pytest.fixture(scope="module")
def B():
    from nirrt_star_b200 import batch
    return batch


@pytest.mark.parametrize("path", GEOM, ids=[os.path.basename(p) for p in GEOM])
def test_predicates_match_reference(B, path):
    g = np.load(path)
    bp = B.BatchPlanner2D([make_problem_2d(int(g["env_idx"]))], 10, seeds=[0])
    assert np.array_equal(bp.collide_edges(0, g["edges"]), g["hit"])
    assert np.array_equal(bp.points_inside_obs(0, g["pts"]), g["inside"])
    assert np.array_equal(bp.points_valid(0, g["pts"]), g["valid"])
    bp.close()


def _check_final(B, bp, g, variant):
    v, p, n = bp.read_trees()
    n = int(n[0])
    assert n == int(g["num_vertices"])
    assert np.array_equal(p[0, :n], g["parents"])
    assert np.array_equal(v[0, :n], g["vertices"])
    if variant == B.VARIANT_IRRT_STAR:
        assert list(bp.solutions(0)) == list(g["solutions"])
    key, pos = bp.get_rng()[0]
    rs = np.random.RandomState(0); rs.set_state(("MT19937", key, pos, 0, 0.0))
    assert rs.random_sample() == float(g["next_random"])
    import random
    key, pos = bp.get_py_rng()[0]
    r = random.Random(0); r.setstate((3, tuple(int(x) for x in key) + (pos,), None))
    assert r.random() == float(g["next_py_random"])


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_batch_planner_matches_reference_golden(B, path):
    g = np.load(path)
    kind, mode = str(g["kind"]), str(g["mode"])
    variant = {"rrt": B.VARIANT_RRT_STAR, "irrt": B.VARIANT_IRRT_STAR}[kind]
    iter_max, iter_after = int(g["iter_max"]), int(g["iter_after"])
    bp = B.BatchPlanner2D([make_problem_2d(int(g["env_idx"]))], iter_max, seeds=[int(g["seed"])],
                          record_capacity=iter_max + iter_after + 8)
    if mode == "planning":
        bp.begin(variant, B.MODE_PLANNING, iter_max)
        bp.run_to_completion()
    else:
        bp.begin(variant, B.MODE_PLANNING_RANDOM, iter_max, iter_after)
        bp.run_to_completion()
        lst = np.array(bp.path_len_lists()[0]); want = g["path_len_list"]
        assert len(lst) == len(want)
        assert np.array_equal(np.isinf(lst), np.isinf(want))
        f = np.isfinite(want)
        assert np.allclose(lst[f], want[f], rtol=1e-5, atol=0)
    _check_final(B, bp, g, variant)
    bp.close()


@pytest.mark.parametrize("path", [p for p in GOLD if "planning" in p], ids=lambda p: os.path.basename(p))
def test_per_iteration_trace(B, path):
    g = np.load(path)
    variant = {"rrt": B.VARIANT_RRT_STAR, "irrt": B.VARIANT_IRRT_STAR}[str(g["kind"])]
    iter_max = int(g["iter_max"])
    bp = B.BatchPlanner2D([make_problem_2d(int(g["env_idx"]))], iter_max, seeds=[int(g["seed"])])
    bp.begin(variant, B.MODE_PLANNING, iter_max)
    off = 0
    for it in range(min(iter_max, 400)):
        bp.run(1)
        nearest, new, cnt, near, xr = bp.trace()
        assert nearest[0] == g["nearest"][it], it
        assert np.array_equal(xr[0], g["rand"][it])
        want_cnt = int(g["near_cnt"][it])
        if want_cnt < 0:
            assert new[0] == -1
            continue
        want = g["near"][off:off + want_cnt]; off += want_cnt
        got = near[0, :cnt[0]]
        assert np.array_equal(got, want), (it, got, want)
    bp.close()


def test_batch_of_problems_equals_singles(B):
    problems = [make_problem_2d(i) for i in range(6)]
    seeds = [40 + i for i in range(6)]
    bp = B.BatchPlanner2D(problems, 600, seeds=seeds)
    bp.begin(B.VARIANT_IRRT_STAR, B.MODE_PLANNING, 600)
    bp.run_to_completion()
    v, p, n = bp.read_trees()
    for e in (0, 3, 5):
        one = B.BatchPlanner2D([problems[e]], 600, seeds=[seeds[e]])
        one.begin(B.VARIANT_IRRT_STAR, B.MODE_PLANNING, 600)
        one.run_to_completion()
        v1, p1, n1 = one.read_trees()
        assert n1[0] == n[e] and np.array_equal(p1[0], p[e]) and np.array_equal(v1[0], v[e])
        one.close()
    bp.close()


def test_config1_shape_64_problems_5000_iterations(B):
    """BASELINE configs[1]: irrt_star 2D random_2d, 64 problems batched, iter_max = 5000 (planning_random with
    iter_after_initial = 500).  The batch (2 pipelined groups, CUDA-graph replay, u16 mirror scans) must
    equal one-problem batches bit for bit, and the trees keep their invariants."""
    E, iter_max, iter_after = 64, 5000, 500
    problems = [make_problem_2d(100 + i) for i in range(E)]
    seeds = [700 + i for i in range(E)]
    bp = B.BatchPlanner2D(problems, iter_max, seeds=seeds, record_capacity=iter_max + iter_after + 8)
    bp.begin(B.VARIANT_IRRT_STAR, B.MODE_PLANNING_RANDOM, iter_max, iter_after)
    bp.run_to_completion(chunk=512)
    lists = bp.path_len_lists()
    v, p, n = bp.read_trees()
    solved = 0
    for e in range(E):
        ne = int(n[e]); pe = p[e, :ne]; ve = v[e, :ne]
        assert pe[0] == 0 and pe.min() >= 0 and pe.max() < ne
        anc = pe.copy()
        for _ in range(int(np.ceil(np.log2(max(ne, 2)))) + 1):
            anc = anc[anc]
        assert not anc.any()
        seg = np.hypot(*(ve - ve[pe]).T)
        assert seg[1:].max() <= 10.0 + 1e-9
        got = np.array(lists[e])
        assert len(got) <= iter_max + iter_after + 1
        f = np.isfinite(got)
        if f.any():
            solved += 1
            assert np.all(np.diff(got[f]) <= 1e-9)               # the best cost never increases
            assert not f[:int(np.argmax(f))].any() and f[int(np.argmax(f)):].all()   # inf ... inf, then finite to the end
    assert solved >= E // 2
    # against the CPU oracle (numpy restatement of irrt_star_2d.py:230-316, pinned to the reference's golden traces
    # by tests/test_oracle_pin2d.py) for problems spread over both groups of the batch
    from oracle.planner2d_oracle import Oracle2D
    for e in (0, 13, 31, 32, 47, 63):
        o = Oracle2D(problems[e], iter_max, seed=seeds[e])
        want = np.array(o.planning_random(iter_after, 1))
        got = np.array(lists[e])
        assert len(got) == len(want), (e, len(got), len(want))
        assert np.array_equal(np.isinf(got), np.isinf(want)), e
        f = np.isfinite(want)
        assert np.allclose(got[f], want[f], rtol=1e-5, atol=0), e
        assert n[e] == o.n and np.array_equal(p[e, :o.n], o.parent[:o.n]), e
        assert np.array_equal(v[e, :o.n], o.v[:o.n]), e
    for e in (0, 21, 63):
        one = B.BatchPlanner2D([problems[e]], iter_max, seeds=[seeds[e]], record_capacity=iter_max + iter_after + 8)
        one.begin(B.VARIANT_IRRT_STAR, B.MODE_PLANNING_RANDOM, iter_max, iter_after)
        one.run_to_completion(chunk=512)
        v1, p1, n1 = one.read_trees()
        assert n1[0] == n[e] and np.array_equal(p1[0], p[e]) and np.array_equal(v1[0], v[e])
        assert np.array_equal(np.array(one.path_len_lists()[0]), np.array(lists[e]))
        one.close()
    bp.close()
