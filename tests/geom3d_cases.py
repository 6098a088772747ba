"""Deterministic edge / point sets for the 3D geometry fixtures (shared by tests/golden/make_golden_geom3d.py, which
labels them with the REFERENCE's own predicates, and tests/test_gpu_geometry3d.py, which regenerates them on the GPU
box -- the fixture stores only the packed answers and a checksum of these inputs).
Covers what collision_check_utils_3d.py:3-84,151-216 branches on: zero-length edges, integer coordinates (touching
inflated faces / spheres exactly), axis-aligned edges, edges tangent to the inflated balls, edges whose end point lies
exactly on an inflated box face, far-away edges (AABB pre-filter), long edges through several obstacles."""
import hashlib

import numpy as np


def make_edges(env_dict, seed, m, clearance=2.0):
    rs = np.random.RandomState(seed)
    balls = np.asarray(env_dict["ball_obstacles"], dtype=np.float64).reshape(-1, 4)
    boxes = np.asarray(env_dict["box_obstacles"], dtype=np.float64).reshape(-1, 6)
    a = rs.uniform(-4, 54, (m, 3))
    d = rs.normal(size=(m, 3)); d /= np.linalg.norm(d, axis=1)[:, None]
    b = a + d * rs.uniform(0, 12, (m, 1))
    q = m // 20
    k = 0
    b[k:k + q] = a[k:k + q]; k += q                                            # zero length
    a[k:k + 2 * q] = np.round(a[k:k + 2 * q]); b[k:k + 2 * q] = np.round(b[k:k + 2 * q]); k += 2 * q   # integer coordinates
    for ax in range(3):                                                        # axis aligned
        oth = [i for i in range(3) if i != ax]
        b[k:k + q][:, oth] = a[k:k + q][:, oth]; k += q
    # tangent to an inflated ball: p on the sphere of radius r + clearance, edge in the tangent plane through p
    if len(balls):
        idx = rs.randint(0, len(balls), q)
        u = rs.normal(size=(q, 3)); u /= np.linalg.norm(u, axis=1)[:, None]
        p = balls[idx, :3] + u * (balls[idx, 3:4] + clearance)
        t = np.cross(u, rs.normal(size=(q, 3))); t /= np.linalg.norm(t, axis=1)[:, None]
        half = rs.uniform(0.5, 6, (q, 1))
        a[k:k + q] = p - t * half; b[k:k + q] = p + t * half; k += q
        # end point exactly on the inflated sphere along an integer axis direction (exact arithmetic)
        ax = rs.randint(0, 3, q); sgn = rs.choice([-1.0, 1.0], q)
        p2 = balls[idx, :3].copy(); p2[np.arange(q), ax] += sgn * (balls[idx, 3] + clearance)
        a[k:k + q] = p2; b[k:k + q] = p2 + np.round(rs.uniform(-6, 6, (q, 3))); k += q
    if len(boxes):
        # end point exactly on an inflated box face / edge / corner
        idx = rs.randint(0, len(boxes), 2 * q)
        lo = boxes[idx, :3] - clearance; hi = boxes[idx, :3] + boxes[idx, 3:] + clearance
        p = lo + (hi - lo) * rs.uniform(0, 1, (2 * q, 3))
        for j in range(2 * q):
            nfix = 1 + (j % 3)
            for ax in rs.permutation(3)[:nfix]:
                p[j, ax] = lo[j, ax] if rs.rand() < 0.5 else hi[j, ax]
        a[k:k + 2 * q] = p
        b[k:k + 2 * q] = p + np.round(rs.uniform(-8, 8, (2 * q, 3)) * 2) / 2
        k += 2 * q
        # edges sliding along an inflated face (coplanar)
        ax = rs.randint(0, 3, q)
        a2 = lo[:q] + (hi[:q] - lo[:q]) * rs.uniform(-0.3, 1.3, (q, 3)); b2 = lo[:q] + (hi[:q] - lo[:q]) * rs.uniform(-0.3, 1.3, (q, 3))
        face = np.where(rs.rand(q) < 0.5, lo[:q][np.arange(q), ax], hi[:q][np.arange(q), ax])
        a2[np.arange(q), ax] = face; b2[np.arange(q), ax] = face
        a[k:k + q] = a2; b[k:k + q] = b2; k += q
    # long edges across the world
    a[k:k + q] = rs.uniform(0, 50, (q, 3)); b[k:k + q] = rs.uniform(0, 50, (q, 3)); k += q
    return np.ascontiguousarray(np.stack([a, b], 1))


def make_points(env_dict, seed, m, clearance=2.0):
    rs = np.random.RandomState(seed)
    balls = np.asarray(env_dict["ball_obstacles"], dtype=np.float64).reshape(-1, 4)
    boxes = np.asarray(env_dict["box_obstacles"], dtype=np.float64).reshape(-1, 6)
    p = rs.uniform(-4, 54, (m, 3))
    q = m // 8
    p[:q] = np.round(p[:q])
    p[q:2 * q] = np.round(p[q:2 * q] * 2) / 2
    if len(balls):      # exactly on an inflated sphere (strict '<' in points_in_balls, collision_check_utils_3d.py:291)
        idx = rs.randint(0, len(balls), q); ax = rs.randint(0, 3, q); sgn = rs.choice([-1.0, 1.0], q)
        pp = balls[idx, :3].copy(); pp[np.arange(q), ax] += sgn * (balls[idx, 3] + clearance)
        p[2 * q:3 * q] = pp
    if len(boxes):      # exactly on an inflated box face
        idx = rs.randint(0, len(boxes), q)
        lo = boxes[idx, :3] - clearance; hi = boxes[idx, :3] + boxes[idx, 3:] + clearance
        pp = lo + (hi - lo) * rs.uniform(0, 1, (q, 3)); ax = rs.randint(0, 3, q)
        pp[np.arange(q), ax] = np.where(rs.rand(q) < 0.5, lo[np.arange(q), ax], hi[np.arange(q), ax])
        p[3 * q:4 * q] = pp
    # on / just beside the (clearance-shrunk) world boundary
    pp = rs.uniform(0, 50, (q, 3)); ax = rs.randint(0, 3, q)
    pp[np.arange(q), ax] = rs.choice([0.0, clearance, 50.0 - clearance, 50.0], q)
    p[4 * q:5 * q] = pp
    return np.ascontiguousarray(p)


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
