"""CPU checks of the PRODUCT's host/device math headers (compiled for the host with g++) against
the oracle and against CPython/numpy directly.  No GPU needed."""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest

from nirrt_star_b200.synthetic import make_problem_3d
from oracle.planner_oracle import Oracle3D, lib as olib

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hh(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("hh") / "libhost_harness.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-mfma", "-fPIC", "-shared",
                           "-x", "c++", os.path.join(HERE, "host_harness.cpp"), "-o", so])
    L = C.CDLL(so)
    for f in ("hh_hypot3", "hh_rownorm3", "hh_vecnorm3"):
        getattr(L, f).restype = C.c_double; getattr(L, f).argtypes = [C.c_double] * 3
    L.hh_hypot2.restype = C.c_double; L.hh_hypot2.argtypes = [C.c_double] * 2
    L.hh_sqrt_le_threshold.restype = C.c_double; L.hh_sqrt_le_threshold.argtypes = [C.c_double]
    L.hh_cr_sincos.argtypes = [C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.hh_pairwise_sum.restype = C.c_double
    return L


def test_hypot_matches_cpython(hh):
    rng = np.random.default_rng(0)
    for scale in (50.0, 1e-3, 1e-9, 1e5):
        w = rng.uniform(-scale, scale, (20000, 3))
        for a, b, c in w:
            assert hh.hh_hypot3(a, b, c) == math.hypot(a, b, c)
            assert hh.hh_hypot2(a, b) == math.hypot(a, b)
    for t in [(0, 0, 0), (3, 4, 0), (1, 1, 1), (0, 0, 5), (10, 0, 0), (1, 2, 2), (-7.5, 0.0, 7.5)]:
        assert hh.hh_hypot3(*map(float, t)) == math.hypot(*t)


def test_norms_match_numpy(hh):
    rng = np.random.default_rng(1)
    w = rng.uniform(-50, 50, (20000, 3))
    rows = np.linalg.norm(w, axis=-1)
    for i, (a, b, c) in enumerate(w):
        assert hh.hh_rownorm3(a, b, c) == rows[i]
        assert hh.hh_vecnorm3(a, b, c) == float(np.linalg.norm(w[i]))


def test_sqrt_threshold_is_exact(hh):
    rng = np.random.default_rng(2)
    for r in list(rng.uniform(0, 60, 3000)) + [0.0, 10.0, 1e-300, 1.0, 2.5]:
        t = hh.hh_sqrt_le_threshold(r)
        assert math.sqrt(t) <= r
        assert math.sqrt(np.nextafter(t, np.inf)) > r


def test_cr_sincos_same_as_oracle(hh):
    o = olib()
    o.orc_cr_sincos.argtypes = [C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    rng = np.random.default_rng(3)
    s1, c1, s2, c2 = C.c_double(), C.c_double(), C.c_double(), C.c_double()
    mism_libm = 0
    xs = np.concatenate([math.pi * rng.random(20000), 2 * math.pi * rng.random(20000), [0.0, math.pi, 2 * math.pi]])
    for x in xs:
        hh.hh_cr_sincos(x, C.byref(s1), C.byref(c1)); o.orc_cr_sincos(x, C.byref(s2), C.byref(c2))
        assert s1.value == s2.value and c1.value == c2.value
        assert abs(s1.value - math.sin(x)) <= 2.3e-16 and abs(c1.value - math.cos(x)) <= 2.3e-16
        mism_libm += (s1.value != math.sin(x)) + (c1.value != math.cos(x))
    assert mism_libm < 0.01 * len(xs)      # libm itself misrounds ~0.13 % of calls


def test_glibc_sin_cos_restatement_equals_libm(hh):
    """glibc_trig.cuh restates glibc 2.39's FMA-variant sin / cos kernels operation by operation: on this image's
    libm (what math.sin / np.sin of the reference call) the results must be bit-identical -- including the ~0.15 % of
    arguments where libm is NOT correctly rounded -- over every branch: |x| < 2^-26, Taylor (< 0.126), table
    (< 0.855469), pi/2 - x (< 2.426265), range reduction with n = 0..3 (< 105414350)."""
    rng = np.random.default_rng(7)
    xs = np.concatenate([rng.uniform(-2 * math.pi, 2 * math.pi, 1500000), rng.uniform(0, math.pi, 500000),
                         rng.uniform(-0.2, 0.2, 300000), rng.uniform(-1e-7, 1e-7, 100000), rng.uniform(-1e4, 1e4, 800000),
                         rng.uniform(-1e8, 1e8, 400000), rng.uniform(0.85, 0.86, 200000), rng.uniform(2.42, 2.43, 200000),
                         [0.0, -0.0, 0.126, 0.855469, 2.426265, math.pi, -math.pi, 2 * math.pi, math.pi / 2, 1e-300]])
    s = np.empty_like(xs); c = np.empty_like(xs)
    dp = C.POINTER(C.c_double)
    hh.hh_glibc_sincos.argtypes = [C.c_long, dp, dp, dp]
    hh.hh_glibc_sincos(len(xs), xs.ctypes.data_as(dp), s.ctypes.data_as(dp), c.ctypes.data_as(dp))
    assert np.array_equal(s, np.sin(xs)) and np.array_equal(c, np.cos(xs))
    assert all(s[i] == math.sin(xs[i]) and c[i] == math.cos(xs[i]) for i in range(0, len(xs), 997))
    # and the thing it is NOT: the correctly rounded value (that is why a restatement is needed)
    s1, c1 = C.c_double(), C.c_double()
    differ = 0
    for x in xs[:40000]:
        hh.hh_cr_sincos(x, C.byref(s1), C.byref(c1))
        differ += (s1.value != math.sin(x)) + (c1.value != math.cos(x))
    assert 0 < differ < 400


def test_glibc_atan2_restatement_equals_libm(hh):
    """math.atan2 of the 2D steer (rrt_star_2d.py:74): glibc 2.39's __ieee754_atan2 (FMA variant) restated -- all four
    quadrants, polynomial and table branches, axis-aligned and integer arguments, extreme ratios."""
    rng = np.random.default_rng(8)
    dp = C.POINTER(C.c_double)
    hh.hh_glibc_atan2.argtypes = [C.c_long, dp, dp, dp]
    for sy, sx in ((224, 224), (10, 10), (1, 1000), (1000, 1), (1e-3, 1), (1, 1e-3), (1e-9, 1e-9), (1e200, 1e-200)):
        y = rng.uniform(-sy, sy, 250000); x = rng.uniform(-sx, sx, 250000)
        y[::7] = np.rint(y[::7]); x[::11] = np.rint(x[::11])
        out = np.empty_like(y)
        hh.hh_glibc_atan2(len(y), y.ctypes.data_as(dp), x.ctypes.data_as(dp), out.ctypes.data_as(dp))
        # math.atan2 is the C library's; np.arctan2 on arrays is numpy's own SIMD kernel (differs in ~8 % of last bits)
        assert np.array_equal(out, np.array([math.atan2(a, b) for a, b in zip(y, x)]))
    for y, x in ((0.0, 1.0), (0.0, -1.0), (-0.0, 1.0), (-0.0, -1.0), (1.0, 0.0), (-1.0, 0.0), (1.0, -0.0), (0.0, 0.0), (0.0, -0.0),
                 (-0.0, -0.0), (3.0, 3.0), (-3.0, 3.0), (3.0, -3.0), (-3.0, -3.0), (1e-310, 1.0), (1.0, 1e-310)):
        a = np.array([y]); b = np.array([x]); out = np.empty(1)
        hh.hh_glibc_atan2(1, a.ctypes.data_as(dp), b.ctypes.data_as(dp), out.ctypes.data_as(dp))
        assert out[0] == math.atan2(y, x) and math.copysign(1, out[0]) == math.copysign(1, math.atan2(y, x)), (y, x)


def test_pairwise_sum_matches_numpy(hh):
    rng = np.random.default_rng(4)
    for n in (1, 2, 7, 8, 9, 15, 16, 17, 31, 100, 128, 129, 300, 1000):
        a = rng.uniform(0, 10, n)
        hh.hh_pairwise_sum.argtypes = [C.POINTER(C.c_double), C.c_long]
        got = hh.hh_pairwise_sum(a.ctypes.data_as(C.POINTER(C.c_double)), n)
        assert got == olib().orc_pairwise_sum(a.ctypes.data_as(C.POINTER(C.c_double)), n)


@pytest.mark.parametrize("env_idx", [0, 1, 2, 3])
def test_geometry_matches_oracle(hh, env_idx):
    pr = make_problem_3d(env_idx)
    o = Oracle3D(pr, 10)
    dpp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    hh.hh_set_geom(len(o.balls), dpp(o.balls), dpp(o.r2), len(o.boxes), dpp(o.boxes), C.c_double(2.0), dpp(o.range6))
    rng = np.random.default_rng(10 + env_idx)
    a = rng.uniform(-2, 52, (40000, 3)); d = rng.normal(size=(40000, 3)); d /= np.linalg.norm(d, axis=1)[:, None]
    b = a + d * rng.uniform(0, 12, (40000, 1))
    b[:200] = a[:200]                                   # degenerate zero-length edges
    edges = np.ascontiguousarray(np.stack([a, b], 1))
    out = np.zeros(len(edges), dtype=np.uint8)
    hh.hh_collide(len(edges), dpp(edges), out.ctypes.data_as(C.POINTER(C.c_uint8)))
    assert np.array_equal(out.astype(bool), o.collide_edges(edges))
    pts = np.ascontiguousarray(rng.uniform(-3, 53, (40000, 3)))
    pts[:1000] = np.round(pts[:1000])                   # integer points land exactly on faces
    hh.hh_inside(len(pts), dpp(pts), out.ctypes.data_as(C.POINTER(C.c_uint8)))
    assert np.array_equal(out.astype(bool), o.points_inside_obs(pts))
    hh.hh_valid(len(pts), dpp(pts), out.ctypes.data_as(C.POINTER(C.c_uint8)))
    assert np.array_equal(out.astype(bool), o.points_valid(pts))


# ---------------------------------------------------------------------------------------------- 2D
def test_np_hypot_is_glibc_hypot(hh):
    hh.hh_np_hypot.restype = C.c_double; hh.hh_np_hypot.argtypes = [C.c_double] * 2
    rng = np.random.default_rng(5)
    for scale in (224.0, 10.0, 1e-3, 1e6):
        w = rng.uniform(-scale, scale, (100000, 2))
        w[:2000, 1] *= 1e-9; w[2000:2500, 1] = 0.0; w[2500:3000] = np.round(w[2500:3000])
        want = np.hypot(w[:, 0], w[:, 1])
        got = np.array([hh.hh_np_hypot(a, b) for a, b in w])
        assert np.array_equal(got, want)


def test_cr_atan2_is_correctly_rounded(hh):
    hh.hh_cr_atan2.restype = C.c_double; hh.hh_cr_atan2.argtypes = [C.c_double] * 2
    rng = np.random.default_rng(6)
    w = rng.uniform(-224, 224, (40000, 2))
    w[:100, 0] = 0.0; w[100:200, 1] = 0.0
    mism = 0
    for y, x in w:
        got = hh.hh_cr_atan2(y, x)
        ref = math.atan2(y, x)
        assert abs(got - ref) <= 4.5e-16 * max(1.0, abs(ref))     # never more than one ulp from libm
        mism += got != ref
    assert mism < 0.01 * len(w)                                   # libm misrounds a fraction of a percent
    assert hh.hh_cr_atan2(0.0, 0.0) == 0.0 and hh.hh_cr_atan2(0.0, -1.0) == math.pi
    # exactness against an independent high-precision evaluation (decimal-free: long double on x86 has 64 bits,
    # enough to decide the rounding of almost every sample)
    ld = np.arctan2(w[:2000, 0].astype(np.longdouble), w[:2000, 1].astype(np.longdouble))
    got = np.array([hh.hh_cr_atan2(y, x) for y, x in w[:2000]])
    assert (got != ld.astype(np.float64)).sum() <= 2              # only double-rounding coincidences may differ


def test_norms2_match_numpy(hh):
    for f in ("hh_vecnorm2", "hh_rownorm2"):
        getattr(hh, f).restype = C.c_double; getattr(hh, f).argtypes = [C.c_double] * 2
    rng = np.random.default_rng(7)
    w = rng.uniform(-224, 224, (20000, 2))
    rows = np.linalg.norm(w, axis=1)
    for i, (a, b) in enumerate(w):
        assert hh.hh_rownorm2(a, b) == rows[i]
        assert hh.hh_vecnorm2(a, b) == float(np.linalg.norm(w[i]))


@pytest.mark.parametrize("env_idx", [0, 1, 2])
def test_geometry2d_matches_reference_golden(hh, env_idx):
    from nirrt_star_b200.synthetic import make_problem_2d
    g = np.load(os.path.join(HERE, "golden", f"geom2d_e{env_idx}.npz"))
    ed = make_problem_2d(env_idx)["env_dict"]
    circles = np.ascontiguousarray(ed["circle_obstacles"], dtype=np.float64)
    rects = np.ascontiguousarray(ed["rectangle_obstacles"], dtype=np.float64)
    dpp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    rng4 = np.array([0, ed["env_dims"][1], 0, ed["env_dims"][0]], dtype=np.float64)
    hh.hh_set_geom2(len(circles), dpp(circles), len(rects), dpp(rects), C.c_double(3.0), dpp(rng4))
    u8 = lambda a: a.ctypes.data_as(C.POINTER(C.c_ubyte))
    edges = np.ascontiguousarray(g["edges"]); pts = np.ascontiguousarray(g["pts"])
    out = np.zeros(len(edges), dtype=np.uint8)
    hh.hh_collide2(len(edges), dpp(edges), u8(out))
    assert np.array_equal(out.astype(bool), g["hit"])
    out = np.zeros(len(pts), dtype=np.uint8)
    hh.hh_inside2(len(pts), dpp(pts), u8(out)); assert np.array_equal(out.astype(bool), g["inside"])
    hh.hh_valid2(len(pts), dpp(pts), u8(out)); assert np.array_equal(out.astype(bool), g["valid"])


def test_ellipsoid_parameters_of_the_device_cloud_sampler():
    """batch.ellipsoid_params_{2d,3d}: the 12 host-side numbers of ellipsoid_point_cloud_sampling{,_3d}
    (datasets/point_cloud_mask_utils.py:104-133, datasets_3d/point_cloud_mask_utils_3d.py:132-160).  M = C @ L with C a proper
    rotation whose first column is the start -> goal direction and L = diag(c_max / 2, b, b), b = sqrt(c_max^2 - c_min^2) / 2;
    the centre is the midpoint (z = 0 in 2D)."""
    from nirrt_star_b200.batch import ellipsoid_params_2d, ellipsoid_params_3d
    rs = np.random.RandomState(4)
    for _ in range(50):
        a, b = rs.uniform(0, 50, 3), rs.uniform(0, 50, 3)
        ratio = rs.uniform(1.0005, 3.0)
        p = ellipsoid_params_3d(a, b, ratio)
        M, centre = p[:9].reshape(3, 3), p[9:]
        c_min = np.linalg.norm(b - a); c_max = c_min * ratio
        semi = np.array([c_max / 2, np.sqrt(c_max ** 2 - c_min ** 2) / 2, np.sqrt(c_max ** 2 - c_min ** 2) / 2])
        C = M / semi
        assert np.allclose(C.T @ C, np.eye(3), atol=1e-12) and abs(np.linalg.det(C) - 1) < 1e-12
        assert np.allclose(C[:, 0], (b - a) / c_min, atol=1e-12)
        assert np.array_equal(centre, (a + b) / 2.)
        # the foci are on the boundary's major axis: |x - start| + |x - goal| = c_max at the tip of the first semi-axis
        tip = centre + M[:, 0]
        assert abs(np.linalg.norm(tip - a) + np.linalg.norm(tip - b) - c_max) < 1e-9 * c_max
        a2, b2 = a[:2], b[:2]
        q = ellipsoid_params_2d(a2, b2, ratio)
        M2, centre2 = q[:9].reshape(3, 3), q[9:]
        c_min2 = math.hypot(*(b2 - a2)); c_max2 = c_min2 * ratio
        semi2 = np.array([c_max2 / 2, math.sqrt(c_max2 ** 2 - c_min2 ** 2) / 2, math.sqrt(c_max2 ** 2 - c_min2 ** 2) / 2])
        C2 = M2 / semi2
        assert np.allclose(C2.T @ C2, np.eye(3), atol=1e-12)
        assert np.allclose(C2[:2, 0], (b2 - a2) / c_min2, atol=1e-12) and abs(C2[2, 0]) < 1e-15
        assert np.array_equal(centre2, np.concatenate([(a2 + b2) / 2., [0.]]))
