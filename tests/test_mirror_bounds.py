"""CPU checks of the arithmetic the mirror scans rely on (nirrt_star_b200/csrc/planner3d.cu, "Mirror scans"):
the fixed-point mirrors are conservative filters in front of the exact float64 tests, so what has to hold is
only the distance error bound they are used with, and the monotonic Near radius the speculative ball assumes."""
import numpy as np
import pytest

from nirrt_star_b200.batch import near_radius_table


def _cells(x, lo, scale, top):
    return np.clip(np.rint((x - lo) * scale), 0, top)


@pytest.mark.parametrize("ext,dim", [(50.0, 3), (224.0, 2), (1.0, 3), (1000.0, 3)])
def test_u16_mirror_distance_error_is_below_margin(ext, dim):
    """|sqrt(a) - d * S| < kMarginU16 = 2 cells: vertex and query rounded to a cell, the squared distance
    accumulated in float32 exactly like the kernel does (integers carried in floats, three roundings)."""
    rng = np.random.default_rng(1)
    lo = -3.0
    S = 65535.0 / ext
    v = rng.uniform(lo, lo + ext, (400000, dim))
    q = rng.uniform(lo, lo + ext, (400000, dim))
    # adversarial: coordinates exactly on cell boundaries (+- half a cell) and coincident points
    v[:1000] = lo + (np.floor(rng.uniform(0, 65535, (1000, dim))) + 0.5) / S
    q[:1000] = lo + (np.floor(rng.uniform(0, 65535, (1000, dim))) + 0.5) / S
    q[1000:2000] = v[1000:2000]
    cv = _cells(v, lo, S, 65535).astype(np.float32)
    cq = np.rint(((q - lo) * S).astype(np.float32)).astype(np.float32)     # mirror_query: float cast, then rintf
    d = cq - cv                                                            # exact: integers below 2^24
    a = (d[:, 0] * d[:, 0]).astype(np.float32)
    for k in range(1, dim):
        a = (d[:, k].astype(np.float64) * d[:, k].astype(np.float64) + a.astype(np.float64)).astype(np.float32)   # fmaf
    true = np.linalg.norm(q - v, axis=1) * S
    err = np.abs(np.sqrt(a.astype(np.float64)) - true)
    assert err.max() < 2.0, err.max()


@pytest.mark.parametrize("ext,dim", [(50.0, 3), (224.0, 2)])
def test_u8_mirror_distance_error_is_below_margin(ext, dim):
    """|sqrt(d8^2) - d * S8| <= sqrt(3) < kMarginU8 = 1.75 cells with exact integer arithmetic
    (|m|^2 - 2 m.q + |q|^2 from two DP4A)."""
    rng = np.random.default_rng(2)
    lo = 0.0
    S = 255.0 / ext
    v = rng.uniform(lo, lo + ext, (400000, dim))
    q = rng.uniform(lo, lo + ext, (400000, dim))
    v[:1000] = lo + (np.floor(rng.uniform(0, 255, (1000, dim))) + 0.5) / S
    q[:1000] = lo + (np.floor(rng.uniform(0, 255, (1000, dim))) + 0.5) / S
    m = _cells(v, lo, S, 255).astype(np.int64)
    c = _cells(q, lo, S, 255).astype(np.int64)
    d2 = (m * m).sum(1) - 2 * (m * c).sum(1) + (c * c).sum(1)
    assert d2.min() >= 0
    true = np.linalg.norm(q - v, axis=1) * S
    err = np.abs(np.sqrt(d2) - true)
    assert err.max() < 1.75, err.max()


@pytest.mark.parametrize("top,margin", [(65535, 2.0), (255, 1.75)])
def test_filters_keep_argmin_and_radius_members(top, margin):
    """What the scans conclude from the bound: the exact argmin lies within 2 * margin of the mirror minimum
    (in the distance domain), and every vertex within r passes `mirror distance <= r * S + margin`."""
    rng = np.random.default_rng(3)
    ext, lo = 50.0, 0.0
    S = top / ext
    for trial in range(40):
        n = int(rng.integers(50, 4000))
        v = rng.uniform(lo, lo + ext, (n, 3))
        q = rng.uniform(lo, lo + ext, 3)
        d = np.sqrt(((q - v) ** 2).sum(1))
        m = _cells(v, lo, S, top); c = _cells(q, lo, S, top)
        dm = np.sqrt(((c - m) ** 2).sum(1))
        assert dm[d.argmin()] <= dm.min() + 2 * margin
        r = float(rng.uniform(0.5, 10.0))
        assert np.all(dm[d <= r] <= r * S + margin)


@pytest.mark.parametrize("dim", [2, 3])
def test_near_radius_table_decreases(dim):
    """The speculative Near ball is drawn with gamma * max(f(n), f(n + 1)), f(n) = (ln n / n)^(1/D): it has to
    cover the radius of whichever vertex count the insertion produces (rrt_star_3d.py:134), for every n."""
    t = near_radius_table(200000, dim)
    assert t.shape[0] >= 200002 and np.all(np.isfinite(t[2:]))
    assert np.all(np.diff(t[3:]) <= 0)                 # non-increasing from n = 3 on
    n = np.arange(3, 200000)
    assert np.all(np.maximum(t[n], t[n + 1]) >= t[n]) and np.all(np.maximum(t[n], t[n + 1]) >= t[n + 1])
