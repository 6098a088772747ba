"""The numpy port timed by bench.py's CPU-baseline legs must grow exactly the same tree as the C
oracle (which is pinned against the reference goldens)."""
import time

import numpy as np

from nirrt_star_b200.synthetic import make_problem_3d
from oracle.numpy_port import RRTStar3DPort
from oracle.planner_oracle import Oracle3D


def test_numpy_port_equals_c_oracle():
    for env_idx, seed, iters in [(0, 7, 700), (6, 3, 500)]:
        pr = make_problem_3d(env_idx)
        o = Oracle3D(pr, iters, seed=seed)
        o.run(iters, 0, 0)
        ov, op = o.tree()
        port = RRTStar3DPort(pr, iters, rng=np.random.RandomState(seed))
        port.run(iters)
        n = port.num_vertices
        assert n == len(ov)
        assert np.array_equal(port.vertex_parents[:n], op)
        assert np.array_equal(port.vertices[:n], ov)


def test_numpy_port_continues_from_snapshot():
    pr = make_problem_3d(2)
    o = Oracle3D(pr, 4000, seed=5)
    o.run(3000, 0, 0)
    v, p = o.tree()
    key, pos = o.rng_state()
    rs = np.random.RandomState(0)
    rs.set_state(("MT19937", key, pos, 0, 0.0))
    port = RRTStar3DPort(pr, 4000, rng=rs)
    port.load_tree(v, p)
    port.run(150)
    o.run(150, 0, 0)
    ov, op = o.tree()
    n = port.num_vertices
    assert n == len(ov) and np.array_equal(port.vertex_parents[:n], op) and np.array_equal(port.vertices[:n], ov)
