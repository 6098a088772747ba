"""Pins oracle/pointnet2_oracle.py against fixtures recorded from the reference's own PNGWrapper /
get_model (tests/golden/make_golden_pointnet2.py)."""
import glob
import os

import numpy as np
import pytest

from nirrt_star_b200.synthetic import make_pointnet2_state
from oracle import pointnet2_oracle as O

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "pointnet2_*.npz")))


def test_fixtures_present():
    assert len(GOLD) >= 3


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_oracle_matches_reference_golden(path):
    g = np.load(path)
    sd = make_pointnet2_state(int(g["ckpt_seed"]))
    tr = {}
    pred, score, logp = O.classify_path_points(sd, g["pc"], g["start_mask"], g["goal_mask"], g["fps_start"], trace=tr)
    for i in range(4):
        assert np.array_equal(tr["fps"][i][0], g[f"fps{i}"].astype(np.int64)), f"fps level {i}"
    for i in range(8):
        assert np.array_equal(tr["groups"][i][0], g[f"group{i}"].astype(np.int64)), f"ball-query group {i}"
    assert np.abs(logp - g["logp"]).max() <= 1e-4        # same torch kernels, BN written out by hand
    assert np.abs(score - g["score"]).max() <= 1e-4
    flip = pred != g["pred"]
    assert not np.any(flip & (np.abs(g["score"] - 0.5) > 1e-3))


def test_fps_start_draw_matches_recorded():
    for path in GOLD:
        g = np.load(path)
        assert np.array_equal(O.draw_fps_starts(int(g["seed"]), len(g["pc"]))[0], g["fps_start"])


def test_fp16_operand_emulation_is_within_the_stated_tolerance():
    """The CUDA path feeds fp16 operands (coordinates as hi+lo pairs) to the tensor cores with fp32
    accumulation; SURVEY.md 8c allows 2e-2 abs on the log-probabilities for reduced-precision
    operands.  (bf16 operands measure 3e-2 here and were rejected for that reason.)"""
    for path in GOLD:
        g = np.load(path)
        sd = make_pointnet2_state(int(g["ckpt_seed"]))
        _, score, logp = O.classify_path_points(sd, g["pc"], g["start_mask"], g["goal_mask"], g["fps_start"], emulate="fp16")
        err = np.abs(logp - g["logp"]).max()
        assert err <= 1e-2, err
