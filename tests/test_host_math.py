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
