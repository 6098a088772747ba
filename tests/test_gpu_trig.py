"""Device build of csrc/glibc_trig.cuh against the host C library (== math.sin / np.sin of the reference's runtime):
bit-identical on every branch, including the arguments where libm is not correctly rounded."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_device_sin_cos_equal_libm():
    from nirrt_star_b200.batch import sincos
    rng = np.random.default_rng(11)
    xs = np.concatenate([rng.uniform(-2 * math.pi, 2 * math.pi, 2000000), rng.uniform(0, math.pi, 500000),
                         rng.uniform(-0.2, 0.2, 300000), rng.uniform(-1e-7, 1e-7, 100000), rng.uniform(-1e4, 1e4, 600000),
                         rng.uniform(-1e8, 1e8, 300000), rng.uniform(0.85, 0.86, 100000), rng.uniform(2.42, 2.43, 100000),
                         [0.0, -0.0, 0.126, 0.855469, 2.426265, math.pi, -math.pi, 2 * math.pi, math.pi / 2]])
    s, c = sincos(xs)
    assert np.array_equal(s, np.sin(xs)) and np.array_equal(c, np.cos(xs))
    assert s[5] == math.sin(xs[5]) and c[5] == math.cos(xs[5])
    assert sincos(np.zeros(0))[0].shape == (0,)


def test_device_atan2_equals_libm():
    from nirrt_star_b200.batch import atan2
    rng = np.random.default_rng(12)
    for sy, sx in ((224, 224), (10, 10), (1, 1000), (1000, 1), (1e-3, 1), (1, 1e-3)):
        y = rng.uniform(-sy, sy, 300000); x = rng.uniform(-sx, sx, 300000)
        y[::7] = np.rint(y[::7]); x[::11] = np.rint(x[::11])
        assert np.array_equal(atan2(y, x), np.array([math.atan2(a, b) for a, b in zip(y, x)]))
