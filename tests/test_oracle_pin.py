"""Pins the CPU oracle (oracle/nirrt_oracle.c) against golden vectors produced by the REFERENCE's
own classes (tests/golden/make_golden_planner.py).  CPU-only."""
import glob
import os

import numpy as np
import pytest

from nirrt_star_b200.synthetic import make_problem_3d
from oracle.planner_oracle import Oracle3D

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "planner3d_*.npz")))


def _run_like_reference(g):
    """Re-runs the golden's driver on the oracle, returning the concatenated traces."""
    kind, mode = str(g["kind"]), str(g["mode"])
    variant = {"rrt": 0, "irrt": 1}[kind]
    problem = make_problem_3d(int(g["env_idx"]))
    o = Oracle3D(problem, int(g["iter_max"]), seed=int(g["seed"]))
    iter_max, iter_after = int(g["iter_max"]), int(g["iter_after"])
    traces = []
    if mode == "planning":
        traces.append(o.run(iter_max, variant, 0, trace=True))
        plist = np.zeros(0)
    else:
        r1 = o.run(iter_max, variant, 1, stop_on_first=True, trace=True)
        traces.append(r1)
        lst = list(r1["pathlen"])
        if variant == 0:
            if lst[-1] < np.inf:
                r2 = o.run(iter_after, variant, 1, trace=True)
                traces.append(r2); lst += list(r2["pathlen"])
        else:
            found = lst[-1] < np.inf
            lst = lst[1:]
            ok = True
            if not found:
                lst.append(o.best_cost()[0])
                ok = lst[-1] < np.inf
            if ok:
                lst = lst[:-1]
                r2 = o.run(iter_after, variant, 1, trace=True)
                traces.append(r2); lst += list(r2["pathlen"])
                lst.append(o.best_cost()[0])
        plist = np.array(lst)
    return o, traces, plist


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_oracle_matches_reference_trace(path):
    g = np.load(path)
    o, traces, plist = _run_like_reference(g)
    # the IRRT* driver's stop iteration does no expansion -> no trace entry for it
    nearest = np.concatenate([t["nearest"][: len(t["nearest"])] for t in traces])
    new = np.concatenate([t["new"] for t in traces])
    near_cnt = np.concatenate([t["near_cnt"] for t in traces])
    near = np.concatenate([t["near"] for t in traces])
    kind = str(g["kind"])
    if kind == "irrt" and str(g["mode"]) == "random":
        # oracle records the break iteration without expanding; drop entries with no expansion
        pass
    gn = g["nearest"]
    assert len(nearest) >= len(gn)
    keep = np.ones(len(nearest), dtype=bool)
    if len(nearest) > len(gn):
        # phase-1 break iteration of the IRRT* family: traced but not expanded by the oracle
        keep[len(traces[0]["nearest"]) - 1] = False
    assert np.array_equal(nearest[keep], gn)
    g_cnt = g["near_cnt"]
    o_cnt = np.where(new[keep] < 0, -1, near_cnt[keep])
    assert np.array_equal(o_cnt, g_cnt)
    assert np.array_equal(near, g["near"])
    v, p = o.tree()
    assert o.num_vertices == int(g["num_vertices"])
    assert np.array_equal(p, g["parents"])
    if kind == "rrt":
        assert np.array_equal(v, g["vertices"])          # RRT*: no transcendental in the loop -> bit exact
    else:
        assert np.array_equal(v, g["vertices"])          # np.sin / np.cos are the C library's: bit exact as well
        assert np.array_equal(o.solutions(), g["solutions"])
    gp = g["path_len_list"]
    assert len(plist) == len(gp)
    if len(gp):
        assert np.array_equal(np.isinf(plist), np.isinf(gp))
        f = np.isfinite(gp)
        assert np.allclose(plist[f], gp[f], rtol=1e-12, atol=0)


def test_nrrt_oracle_matches_reference_golden():
    """NRRT*-PNG 3D (RRT* driver + guidance cloud predicted once): the oracle regenerates the cloud
    from the numpy stream, takes the recorded prediction and must reproduce the reference's tree
    (tests/golden/make_golden_neural_planner.py)."""
    import glob
    import hashlib
    from oracle.planner_oracle import guidance_cloud_3d
    paths = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "neural3d_nrrt_*.npz")))
    assert paths
    for path in paths:
        g = np.load(path)
        pr = make_problem_3d(int(g["env_idx"]))
        seed, iter_max = int(g["seed"]), int(g["iter_max"])
        rs = np.random.RandomState(seed)
        o = Oracle3D(pr, iter_max, seed=seed)
        pc = guidance_cloud_3d(o, rs)
        assert hashlib.sha1(np.ascontiguousarray(pc.astype(np.float32)).tobytes()).hexdigest() == str(g["call_pc_sha1"][0])
        pred = np.unpackbits(g["call_pred"][0])[:len(pc)]
        st = rs.get_state()
        o2 = Oracle3D(pr, iter_max, rng_state=(st[1], st[2]))
        o2.set_cloud(pc[pred.nonzero()[0]], float(g["pc_sample_rate"]))
        if str(g["mode"]) == "planning":
            o2.run(iter_max, 3, 0)
        else:
            lst = np.array(o2.planning_random(int(g["iter_after"]), 3)); want = g["path_len_list"]
            assert len(lst) == len(want) and np.array_equal(np.isinf(lst), np.isinf(want))
            f = np.isfinite(want)
            assert np.allclose(lst[f], want[f], rtol=1e-12, atol=0)
        v, p = o2.tree()
        assert len(v) == int(g["num_vertices"]) and np.array_equal(p, g["parents"]) and np.array_equal(v, g["vertices"])
