"""GPU parity tests of the 3D planner path: CUDA (through the C ABI) vs the CPU oracle and vs the
committed reference golden traces.  Bit-exact for indices and (RRT*) coordinates."""
import glob
import os

import numpy as np
import pytest

from nirrt_star_b200.synthetic import make_problem_3d

pytestmark = pytest.mark.gpu

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "planner3d_*.npz")))


@pytest.fixture(scope="module")
def B():
    from nirrt_star_b200 import batch
    return batch


def _oracle(problem, iter_max, seed):
    from oracle.planner_oracle import Oracle3D
    return Oracle3D(problem, iter_max, seed=seed)


def test_predicates_match_oracle(B):
    problems = [make_problem_3d(i) for i in range(3)]
    bp = B.BatchPlanner3D(problems, 50, seeds=[0, 1, 2])
    rng = np.random.default_rng(0)
    for e, pr in enumerate(problems):
        o = _oracle(pr, 50, 0)
        a = rng.uniform(-2, 52, (50000, 3)); d = rng.normal(size=(50000, 3)); d /= np.linalg.norm(d, axis=1)[:, None]
        b = a + d * rng.uniform(0, 12, (50000, 1))
        b[:300] = a[:300]
        edges = np.stack([a, b], 1)
        assert np.array_equal(bp.collide_edges(e, edges), o.collide_edges(edges))
        pts = rng.uniform(-3, 53, (50000, 3)); pts[:2000] = np.round(pts[:2000])
        assert np.array_equal(bp.points_inside_obs(e, pts), o.points_inside_obs(pts))
        assert np.array_equal(bp.points_valid(e, pts), o.points_valid(pts))
    # empty inputs are accepted
    assert bp.collide_edges(0, np.zeros((0, 2, 3))).shape == (0,)


def _grow_oracle(problem, iters, seed):
    o = _oracle(problem, iters, seed)
    o.run(iters, 0, 0)
    return o


def test_nearest_within_costs_on_loaded_tree(B):
    pr = make_problem_3d(4)
    iters = 3000
    o = _grow_oracle(pr, iters, 3)
    v, p = o.tree()
    n = len(v)
    bp = B.BatchPlanner3D([pr], iters, seeds=[3])
    V = np.zeros((1, bp.capacity, 3)); P = np.zeros((1, bp.capacity), dtype=np.int64)
    V[0, :n] = v; P[0, :n] = p
    bp.load_trees(V, P, [n])
    v2, p2, n2 = bp.read_trees()
    assert n2[0] == n and np.array_equal(v2[0, :n], v) and np.array_equal(p2[0, :n], p)
    assert not v2[0, n:].any() and not p2[0, n:].any()
    rng = np.random.default_rng(1)
    q = rng.uniform(0, 50, (64, 3))
    q[:8] = v[rng.integers(0, n, 8)]                       # exact hits
    assert np.array_equal(bp.nearest(0, q), o.nearest(q))
    for k in range(16):
        r = float(rng.uniform(0.5, 10))
        assert np.array_equal(bp.within(0, q[k], r), o.within(q[k], r))
    # radius exactly equal to an existing distance (the <= boundary)
    d = np.linalg.norm(q[20] - v, axis=-1)
    r = float(np.sort(d)[5])
    assert np.array_equal(bp.within(0, q[20], r), o.within(q[20], r))
    idx = rng.integers(0, n, 200)
    want = np.array([o.cost(i) for i in idx])
    assert np.array_equal(bp.costs(0, idx), want)


@pytest.mark.parametrize("variant", [0, 1])
def test_per_iteration_trace_matches_oracle(B, variant):
    """One iteration at a time: nearest index, new index, Near list and sample must be identical."""
    E, iters = 6, 400
    problems = [make_problem_3d(10 + i) for i in range(E)]
    seeds = [100 + i for i in range(E)]
    bp = B.BatchPlanner3D(problems, iters, seeds=seeds)
    bp.begin(variant, B.MODE_PLANNING, iters)
    oracles = [_oracle(pr, iters, s) for pr, s in zip(problems, seeds)]
    for it in range(iters):
        bp.run(1)
        nearest, new, cnt, near, xr = bp.trace()
        for e, o in enumerate(oracles):
            r = o.run(1, variant, 0, trace=True)
            assert nearest[e] == r["nearest"][0], (it, e)
            assert new[e] == r["new"][0], (it, e)
            if r["new"][0] >= 0:
                assert cnt[e] == r["near_cnt"][0], (it, e)
                assert np.array_equal(near[e, :cnt[e]], r["near"]), (it, e)
    running, need = bp.status()
    assert running == 0 and need == 0
    v, p, n = bp.read_trees()
    for e, o in enumerate(oracles):
        ov, op = o.tree()
        assert n[e] == len(ov)
        assert np.array_equal(p[e, :n[e]], op)
        assert np.array_equal(v[e, :n[e]], ov)              # same op sequence on both sides -> bit exact
        if variant == 1:
            assert np.array_equal(bp.solutions(e), o.solutions())


@pytest.mark.parametrize("variant", [0, 1])
def test_planning_random_matches_oracle(B, variant):
    E, iter_max, iter_after = 8, 1500, 250
    problems = [make_problem_3d(30 + i) for i in range(E)]
    seeds = [7 + 3 * i for i in range(E)]
    bp = B.BatchPlanner3D(problems, iter_max, seeds=seeds, record_capacity=iter_max + iter_after + 8)
    bp.begin(variant, B.MODE_PLANNING_RANDOM, iter_max, iter_after)
    bp.run_to_completion(chunk=128)
    lists = bp.path_len_lists()
    v, p, n = bp.read_trees()
    for e in range(E):
        o = _oracle(problems[e], iter_max, seeds[e])
        want = np.array(o.planning_random(iter_after, variant))
        got = np.array(lists[e])
        assert len(got) == len(want), (e, len(got), len(want))
        assert np.array_equal(np.isinf(got), np.isinf(want))
        f = np.isfinite(want)
        assert np.allclose(got[f], want[f], rtol=1e-12, atol=0)
        ov, op = o.tree()
        assert n[e] == len(ov) and np.array_equal(p[e, :n[e]], op) and np.array_equal(v[e, :n[e]], ov)


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_matches_reference_golden(B, path):
    """CUDA path vs traces recorded from the REFERENCE's own classes (tests/golden)."""
    g = np.load(path)
    kind, mode = str(g["kind"]), str(g["mode"])
    variant = {"rrt": 0, "irrt": 1}[kind]
    iter_max, iter_after = int(g["iter_max"]), int(g["iter_after"])
    pr = make_problem_3d(int(g["env_idx"]))
    bp = B.BatchPlanner3D([pr], iter_max, seeds=[int(g["seed"])], record_capacity=iter_max + iter_after + 8)
    if mode == "planning":
        bp.begin(variant, B.MODE_PLANNING, iter_max)
    else:
        bp.begin(variant, B.MODE_PLANNING_RANDOM, iter_max, iter_after)
    bp.run_to_completion(chunk=200)
    v, p, n = bp.read_trees()
    assert n[0] == int(g["num_vertices"])
    assert np.array_equal(p[0, :n[0]], g["parents"])
    if variant == 0:
        assert np.array_equal(v[0, :n[0]], g["vertices"])
    else:
        assert np.array_equal(v[0, :n[0]], g["vertices"])      # informed sampling too: glibc sin / cos restated bit for bit
        assert np.array_equal(bp.solutions(0), g["solutions"])
    if mode == "random":
        got = np.array(bp.path_len_lists()[0]); want = g["path_len_list"]
        assert len(got) == len(want)
        assert np.array_equal(np.isinf(got), np.isinf(want))
        f = np.isfinite(want)
        assert np.allclose(got[f], want[f], rtol=1e-5, atol=0)
    else:
        gp, cost = bp.goal_parents()
        if len(g["path"]):
            # extract_path(goal_parent)[-2] is the goal parent vertex
            assert np.array_equal(v[0, gp[0]], g["path"][-2])
        else:
            assert gp[0] == -1


def test_large_tree_window_matches_oracle(B):
    """Steady-state window on a big tree (the benchmark regime, scaled to what the oracle grows in
    seconds): grow on the GPU, continue on both sides from the same snapshot + RNG state."""
    E, grow, window = 4, 20000, 200
    problems = [make_problem_3d(50 + i) for i in range(E)]
    seeds = [900 + i for i in range(E)]
    cap_iters = grow + window
    bp = B.BatchPlanner3D(problems, cap_iters, seeds=seeds)
    bp.begin(0, B.MODE_PLANNING, cap_iters)
    bp.run(grow)
    v, p, n = bp.read_trees()
    states = bp.get_rng()
    assert n.min() > grow // 4
    from oracle.planner_oracle import Oracle3D
    oracles = []
    for e in range(E):
        o = Oracle3D(problems[e], cap_iters, rng_state=states[e])
        o.load_tree(v[e, :n[e]], p[e, :n[e]])
        oracles.append(o)
    bp.run(window)
    v2, p2, n2 = bp.read_trees()
    for e, o in enumerate(oracles):
        o.run(window, 0, 0)
        ov, op = o.tree()
        assert n2[e] == len(ov)
        assert np.array_equal(p2[e, :n2[e]], op)
        assert np.array_equal(v2[e, :n2[e]], ov)


def _run_batch(B, problems, seeds, iters, variant=0, env=None):
    """Trees after `iters` lock-step iterations with the given environment overrides (read at batch creation)."""
    old = {}
    for k, val in (env or {}).items():
        old[k] = os.environ.get(k)
        os.environ[k] = val
    try:
        bp = B.BatchPlanner3D(problems, iters, seeds=seeds)
    finally:
        for k, val in old.items():
            if val is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = val
    bp.begin(variant, B.MODE_PLANNING, iters)
    bp.run(iters)
    v, p, n = bp.read_trees()
    bpv = bp.scan_bytes_per_vertex()
    bp.close()
    return v, p, n, bpv


@pytest.mark.parametrize("variant", [0, 1])
def test_scan_layouts_and_pipelines_agree(B, variant):
    """The u16 fixed-point mirror (default: TMA-staged k_nearest_t; u16ldg: the LDG kernel), the f32 mirror and the plain f64 scans are filters in
    front of the same exact arithmetic, and groups / programmatic dependent launch only reorder
    independent problems: every configuration must produce bit-identical trees (and match the oracle)."""
    E, iters = 12, 1200
    problems = [make_problem_3d(70 + i) for i in range(E)]
    seeds = [300 + i for i in range(E)]
    base = _run_batch(B, problems, seeds, iters, variant)
    assert base[3] == 6                                           # default layout: 2 B per coordinate
    # the default run above is the persistent one-CTA-per-problem path (small trees); NIRRT_PERSIST=0 and every explicit
    # NIRRT_SCAN select the two-kernels-per-iteration pipeline
    for env, bpv in (({"NIRRT_SCAN": "f32"}, 12), ({"NIRRT_SCAN": "f64"}, 24), ({"NIRRT_GROUPS": "5", "NIRRT_PERSIST": "0"}, 6),
                     ({"NIRRT_SCAN": "s8"}, 4), ({"NIRRT_SCAN": "s8", "NIRRT_GROUPS": "4", "NIRRT_CHUNKS": "3"}, 4),
                     ({"NIRRT_SCAN": "s8", "NIRRT_GROUPS": "2", "NIRRT_CHUNKS": "16", "NIRRT_GRAPH": "0"}, 4),
                     ({"NIRRT_PERSIST": "0", "NIRRT_TMA": "4"}, 6), ({"NIRRT_PERSIST": "0", "NIRRT_TMA": "4", "NIRRT_CHUNKS": "7", "NIRRT_GROUPS": "3"}, 6),
                     ({"NIRRT_PERSIST": "0", "NIRRT_TMA": "3"}, 6),
                     ({"NIRRT_PERSIST": "0"}, 6), ({"NIRRT_PERSIST": "0", "NIRRT_GROUPS": "3", "NIRRT_CHUNKS": "7"}, 6),
                     ({"NIRRT_PERSIST": "0", "NIRRT_TMA": "1"}, 6),
                     ({"NIRRT_CHUNKS": "1", "NIRRT_PERSIST": "0"}, 6), ({"NIRRT_CHUNKS": "13", "NIRRT_GROUPS": "2", "NIRRT_PERSIST": "0"}, 6),
                     ({"NIRRT_SCAN": "u8"}, 4), ({"NIRRT_SCAN": "u8", "NIRRT_GROUPS": "4", "NIRRT_CHUNKS": "3"}, 4),
                     ({"NIRRT_SCAN": "u8", "NIRRT_GROUPS": "3", "NIRRT_PIPELINE": "1"}, 4),
                     ({"NIRRT_GROUPS": "5", "NIRRT_GRAPH": "0", "NIRRT_PERSIST": "0"}, 6), ({"NIRRT_SCAN": "f64", "NIRRT_GROUPS": "3"}, 24),
                     ({"NIRRT_GROUPS": "4", "NIRRT_PIPELINE": "1", "NIRRT_PERSIST": "0"}, 6),
                     ({"NIRRT_GROUPS": "5", "NIRRT_PIPELINE": "1", "NIRRT_GRAPH": "0", "NIRRT_PERSIST": "0"}, 6),
                     ({"NIRRT_GROUPS": "3", "NIRRT_SCAN": "f32", "NIRRT_GRAPH": "6", "NIRRT_PIPELINE": "1"}, 12),
                     ({"NIRRT_GROUPS": "1", "NIRRT_PDL": "0", "NIRRT_PERSIST": "0"}, 6), ({"NIRRT_CHUNKS": "3", "NIRRT_PERSIST": "0"}, 6)):
        v, p, n, got_bpv = _run_batch(B, problems, seeds, iters, variant, env)
        assert got_bpv == bpv, env
        assert np.array_equal(n, base[2]), env
        for e in range(E):
            assert np.array_equal(p[e, :n[e]], base[1][e, :n[e]]), (env, e)
            assert np.array_equal(v[e, :n[e]], base[0][e, :n[e]]), (env, e)
    for e in range(0, E, 4):
        o = _oracle(problems[e], iters, seeds[e])
        o.run(iters, variant, 0)
        ov, op = o.tree()
        assert base[2][e] == len(ov) and np.array_equal(base[1][e, :len(ov)], op)
        if variant == 0:
            assert np.array_equal(base[0][e, :len(ov)], ov)


def test_loaded_trees_continue_identically(B):
    """load_trees rebuilds the walk records (cached edge lengths, ancestor hints, mirrors) from the
    reference layout: a run continued on a fresh batch from a snapshot equals the uninterrupted run,
    split runs equal one run, and both equal the oracle continued from the same snapshot."""
    from oracle.planner_oracle import Oracle3D
    E, first, second = 6, 2500, 700
    problems = [make_problem_3d(90 + i) for i in range(E)]
    seeds = [40 + i for i in range(E)]
    total = first + second
    a = B.BatchPlanner3D(problems, total, seeds=seeds)
    a.begin(0, B.MODE_PLANNING, total)
    a.run(first)
    v, p, n = a.read_trees()
    rng = a.get_rng()
    a.run(second)
    va, pa, na = a.read_trees()
    b = B.BatchPlanner3D(problems, total, rng_states=rng)
    b.load_trees(v, p, n)
    b.begin(0, B.MODE_PLANNING, total)
    b.run(second // 3); b.run(second - second // 3)
    vb, pb, nb = b.read_trees()
    assert np.array_equal(na, nb)
    for e in range(E):
        assert np.array_equal(pa[e, :na[e]], pb[e, :na[e]]) and np.array_equal(va[e, :na[e]], vb[e, :na[e]])
        o = Oracle3D(problems[e], total, rng_state=rng[e])
        o.load_tree(v[e, :n[e]], p[e, :n[e]])
        o.run(second, 0, 0)
        ov, op = o.tree()
        assert nb[e] == len(ov) and np.array_equal(pb[e, :nb[e]], op) and np.array_equal(vb[e, :nb[e]], ov)
    a.close(); b.close()


def test_tree_invariants_at_scale(B):
    """Size-independent properties on bigger trees than the oracle grows in test time: rooted and acyclic,
    edges no longer than step_len and collision free, costs strictly increasing along every edge, and the
    device scans agree with a brute-force numpy search on the final vertex arrays."""
    E, iters = 32, 12000
    problems = [make_problem_3d(200 + i) for i in range(E)]
    bp = B.BatchPlanner3D(problems, iters, seeds=[77 + i for i in range(E)])
    bp.begin(0, B.MODE_PLANNING, iters)
    bp.run(iters)
    v, p, n = bp.read_trees()
    rng = np.random.default_rng(5)
    for e in range(E):
        ne = int(n[e]); ve = v[e, :ne]; pe = p[e, :ne]
        assert ne > iters // 4 and pe[0] == 0 and pe.min() >= 0 and pe.max() < ne
        # pointer jumping: every vertex reaches the root within ne hops (no cycles)
        anc = pe.copy()
        for _ in range(int(np.ceil(np.log2(ne))) + 1):
            anc = anc[anc]
        assert not anc.any()
        seg = np.linalg.norm(ve - ve[pe], axis=1)
        assert seg[1:].max() <= 10.0 + 1e-9 and seg[1:].min() > 1e-8
        if e % 8 == 0:
            edges = np.stack([ve[pe[1:]], ve[1:]], 1)
            assert not bp.collide_edges(e, edges).any()
            idx = rng.integers(1, ne, 300)
            c = bp.costs(e, idx); cp = bp.costs(e, pe[idx])
            assert np.all(c > cp)
            q = rng.uniform(2, 48, (40, 3))
            d = np.sqrt(((q[:, None, :] - ve[None, :, :]) ** 2).sum(-1))
            assert np.array_equal(bp.nearest(e, q), d.argmin(1))
            for k in range(4):
                assert np.array_equal(bp.within(e, q[k], 3.0), np.nonzero(np.linalg.norm(q[k] - ve, axis=-1) <= 3.0)[0])
    bp.close()


def test_parity_at_100k_vertices(B):
    """The benchmark regime itself (missing #3): trees of >= 100 000 vertices grown by the CUDA planner,
    then BOTH sides continue from the same snapshot + RNG state -- the C oracle (pinned against the reference's
    golden traces) and the GPU -- for the same window; parents and vertices must be bit-identical
    (rrt_star_3d.py:36-55).  At n = 1e5 this exercises the u16 mirror margins, the speculative Near ball, the
    candidate-band Nearest finish and the hint walks at the depth / density the bench runs them at."""
    from oracle.planner_oracle import Oracle3D
    E, nodes, window = 3, 100000, 384
    problems = [make_problem_3d(400 + i) for i in range(E)]
    seeds = [6100 + i for i in range(E)]
    cap_iters = nodes + window + 64
    bp = B.BatchPlanner3D(problems, cap_iters, seeds=seeds)
    bp.begin(0, B.MODE_PLANNING, 1 << 30)
    bp.set_vertex_limit(nodes)
    for _ in range(200):
        bp.run(4096)
        _, _, nv = bp.env_state()
        if nv.min() >= nodes:
            break
    bp.set_vertex_limit(0)
    v, p, n = bp.read_trees()
    assert n.min() >= nodes
    states = bp.get_rng()
    bp.begin(0, B.MODE_PLANNING, 1 << 30)
    bp.run(window)
    v2, p2, n2 = bp.read_trees()
    assert bp.graph_stats()["fallbacks"] == 0
    for e in range(E):
        o = Oracle3D(problems[e], cap_iters, rng_state=states[e])
        o.load_tree(v[e, :n[e]], p[e, :n[e]])
        o.run(window, 0, 0)
        ov, op = o.tree()
        assert n2[e] == len(ov), e
        assert np.array_equal(p2[e, :n2[e]], op), e
        assert np.array_equal(v2[e, :n2[e]], ov), e
    # the eval variant (planning_random body: search_goal_parent + path length after every iteration,
    # rrt_star_3d.py:101-117,225-231) from the same snapshot: per-iteration path lengths and the trees
    b2 = B.BatchPlanner3D(problems, cap_iters, rng_states=states, record_capacity=window + 8)
    b2.load_trees(v, p, n)
    b2.begin(0, B.MODE_PLANNING_RANDOM, 1 << 30, 1 << 30)
    b2.run(window)
    recs = b2.records()
    v3, p3, n3 = b2.read_trees()
    gp, _ = b2.goal_parents()
    for e in range(E):
        o = Oracle3D(problems[e], cap_iters, rng_state=states[e])
        o.load_tree(v[e, :n[e]], p[e, :n[e]])
        want = o.run(window, 0, 1)["pathlen"]
        got = recs[e]
        assert len(got) == window and np.array_equal(np.isinf(got), np.isinf(want)), e
        f = np.isfinite(want)
        assert f.any() and np.allclose(got[f], want[f], rtol=1e-12, atol=0), e
        ov, op = o.tree()
        assert n3[e] == len(ov) and np.array_equal(p3[e, :n3[e]], op) and np.array_equal(v3[e, :n3[e]], ov), e
        assert gp[e] == o.search_goal_parent(), e
    b2.close()
    bp.close()


def test_eval_variant_goal_tracking_long_run(B):
    """RRT* planning_random keeps search_goal_parent's answer current incrementally (child lists + cached candidate
    costs, refreshed only below re-wired vertices).  Long runs on small worlds re-wire constantly, move whole
    subtrees (incl. through the duplicate guard) and improve the goal path many times: the per-iteration path
    lengths must equal the oracle's, which re-walks every candidate every iteration like the reference."""
    from oracle.planner_oracle import Oracle3D
    E, iter_max, iter_after = 6, 6000, 3000
    problems = [make_problem_3d(600 + i) for i in range(E)]
    seeds = [1234 + i for i in range(E)]
    bp = B.BatchPlanner3D(problems, iter_max, seeds=seeds, record_capacity=iter_max + iter_after + 8)
    bp.begin(0, B.MODE_PLANNING_RANDOM, iter_max, iter_after)
    bp.run_to_completion(chunk=512)
    lists = bp.path_len_lists()
    v, p, n = bp.read_trees()
    improved = 0
    for e in range(E):
        o = Oracle3D(problems[e], iter_max, seed=seeds[e])
        want = np.array(o.planning_random(iter_after, 0))
        got = np.array(lists[e])
        assert len(got) == len(want), (e, len(got), len(want))
        assert np.array_equal(np.isinf(got), np.isinf(want)), e
        f = np.isfinite(want)
        assert np.allclose(got[f], want[f], rtol=1e-12, atol=0), e
        improved += len(np.unique(want[f]))
        ov, op = o.tree()
        assert n[e] == len(ov) and np.array_equal(p[e, :n[e]], op) and np.array_equal(v[e, :n[e]], ov)
    assert improved > 10 * E          # the goal path really changed many times
    bp.close()


def test_graph_is_built_by_begin_and_reused(B):
    """nirrt_batch_run never captures a CUDA graph it can find: begin() builds one executable per (variant, mode),
    a second begin() with the same pair re-uses it, changing the run parameters (vertex limit, stop threshold,
    iteration counts) does not invalidate it, and graph replay == plain launches bit for bit."""
    E, iters = 64, 160
    problems = [make_problem_3d(500 + i) for i in range(E)]
    seeds = [9000 + i for i in range(E)]
    os.environ["NIRRT_PERSIST"] = "0"          # small trees would otherwise take the persistent path (no graph at all)
    try:
        bp = B.BatchPlanner3D(problems, 3 * iters, seeds=seeds)
    finally:
        os.environ.pop("NIRRT_PERSIST", None)
    bp.begin(0, B.MODE_PLANNING, 3 * iters)
    g = bp.graph_stats()
    assert g == {"builds": 1, "replays": 0, "fallbacks": 0}
    bp.run(iters)                                  # 160 = 10 replays of the 16-iteration graph
    g = bp.graph_stats()
    assert g["builds"] == 1 and g["replays"] == iters // 16 and g["fallbacks"] == 0
    bp.set_vertex_limit(1 << 20); bp.set_stop_threshold(1e9)
    bp.begin(0, B.MODE_PLANNING, 5 * iters, 7)
    bp.run(iters - 3)                              # 9 replays + 13 plain iterations
    g = bp.graph_stats()
    assert g["builds"] == 1 and g["replays"] == iters // 16 + (iters - 3) // 16
    bp.begin(1, B.MODE_PLANNING_RANDOM, iters, 10)
    assert bp.graph_stats()["builds"] == 2
    bp.begin(0, B.MODE_PLANNING, 3 * iters)
    assert bp.graph_stats()["builds"] == 2
    v, p, n = bp.read_trees()
    bp.close()
    ref = _run_batch(B, problems, seeds, 2 * iters - 3, 0, {"NIRRT_GRAPH": "0", "NIRRT_PERSIST": "0"})
    per = _run_batch(B, problems, seeds, 2 * iters - 3, 0)        # and the persistent path gives the same trees
    assert np.array_equal(per[2], ref[2]) and np.array_equal(per[1], ref[1]) and np.array_equal(per[0], ref[0])
    assert np.array_equal(n, ref[2])
    for e in range(E):
        assert np.array_equal(p[e, :n[e]], ref[1][e, :n[e]]) and np.array_equal(v[e, :n[e]], ref[0][e, :n[e]])


def test_dense_informed_tree_with_thousands_of_near_candidates(B):
    """Informed planners concentrate the tree in a thin ellipsoid, where one Near ball holds thousands of vertices (the
    reference has no limit on |Near|).  Beyond 1024 candidates k_expand stages its per-candidate arrays in HBM
    (near_capacity up to 8192): index sort, collision filter, parallel walks, ChooseParent and the sequential-equivalent
    Rewire must still reproduce the oracle bit for bit.  gamma is raised so that r = step_len throughout."""
    from oracle.planner_oracle import Oracle3D
    E, iters = 3, 3500
    problems = []
    for i in range(E):
        pr = dict(make_problem_3d(700 + i))
        pr["search_radius"] = 500.0
        problems.append(pr)
    seeds = [8100 + i for i in range(E)]
    bp = B.BatchPlanner3D(problems, iters, seeds=seeds, near_capacity=B.NEAR_CAPACITY_INFORMED)
    bp.begin(1, B.MODE_PLANNING, iters)
    biggest = 0
    for _ in range(iters // 250):
        bp.run(250)
        _, _, cnt, _, _ = bp.trace(near_stride=B.NEAR_CAPACITY_INFORMED)
        biggest = max(biggest, int(cnt.max()))
    running, need = bp.status()          # raises on any overflow
    v, p, n = bp.read_trees()
    assert biggest > 1024, biggest       # the HBM-staged path really ran
    for e in range(E):
        o = Oracle3D(problems[e], iters, seed=seeds[e])
        o.run(iters, 1, 0)
        ov, op = o.tree()
        assert n[e] == len(ov) and np.array_equal(p[e, :n[e]], op) and np.array_equal(v[e, :n[e]], ov), e
        assert np.array_equal(bp.solutions(e), o.solutions())
    bp.close()
    # without the large capacity the same run is a hard error, never a silent truncation
    from nirrt_star_b200._lib import NirrtError
    small = B.BatchPlanner3D(problems[:1], iters, seeds=seeds[:1])
    small.begin(1, B.MODE_PLANNING, iters)
    small.run(iters)
    with pytest.raises(NirrtError, match="near-candidate"):
        small.status()
    small.close()
