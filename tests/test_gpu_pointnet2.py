"""GPU parity tests of the PointNet++ path: CUDA (through the C ABI) vs the torch-fp32 oracle and
vs fixtures recorded from the reference's own PNGWrapper (tests/golden/pointnet2_*.npz).

Tolerances: FPS indices exact given identical start indices; ball-query groups exact
outside a |d^2 - r^2| < 4e-6 band; log-probabilities <= 2e-2 abs (fp16 tensor-core operands, fp32
accumulation); path_pred identical except where |score - 0.5| < 0.05."""
import glob
import os

import numpy as np
import pytest

from nirrt_star_b200.synthetic import make_cloud_3d, make_pointnet2_state

pytestmark = pytest.mark.gpu
GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "pointnet2_*.npz")))
LOGP_TOL = 2e-2
RADII = ((0.05, 0.1), (0.1, 0.2), (0.2, 0.4), (0.4, 0.8))
NP = (2048, 1024, 256, 64, 16)
CH = (6, 96, 256, 512, 1024)


@pytest.fixture(scope="module")
def engine():
    from nirrt_star_b200.pointnet2 import PointNet2Engine
    eng = PointNet2Engine(make_pointnet2_state(0), n_points=2048, max_batch=8)
    yield eng
    eng.close()


@pytest.mark.parametrize("m,n,k,mode,group", [
    (128, 16, 16, 0, 16), (128, 32, 64, 0, 16), (256, 64, 112, 0, 16), (384, 208, 272, 0, 16),
    (100, 128, 128, 0, 16), (512, 256, 528, 0, 16), (256, 384, 256, 0, 16), (256, 512, 384, 0, 16),
    (256, 32, 16, 1, 16), (256, 64, 32, 1, 32), (512, 512, 256, 1, 32), (2048, 128, 96, 1, 16),
])
def test_tensor_core_gemm_matches_fp32(m, n, k, mode, group):
    from nirrt_star_b200.pointnet2 import gemm_f16
    rng = np.random.default_rng(m * 7 + n * 3 + k)
    A = rng.standard_normal((m, k)).astype(np.float16)
    W = (rng.standard_normal((n, k)) / np.sqrt(k)).astype(np.float16)
    bias = rng.standard_normal(n).astype(np.float32) * 0.1
    got = gemm_f16(A, W, bias, mode, group).astype(np.float32)
    ref = A.astype(np.float32) @ W.astype(np.float32).T
    if mode == 0:
        want = np.maximum(ref + bias, 0)
    else:
        want = np.maximum(ref.reshape(m // group, group, n).max(axis=1) + bias, 0)
    assert got.shape == want.shape
    err = np.abs(got - want).max()
    assert err <= 4e-3 * max(1.0, np.abs(want).max()), err      # fp16 output rounding


def _ball_band_ok(got, want, sqd, r, K, N):
    """groups equal, or differ only at members whose expansion distance is within 4e-6 of r^2"""
    if np.array_equal(got, want):
        return True
    r2 = np.float32(r ** 2)
    for s in np.nonzero((got != want).any(axis=1))[0]:
        amb = set(np.nonzero(np.abs(sqd[s] - r2) < 4e-6)[0].tolist())
        if not amb:
            return False
        if (set(got[s].tolist()) ^ set(want[s].tolist())) - amb:
            return False
    return True


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_forward_matches_reference_golden(engine, path):
    from oracle import pointnet2_oracle as O
    g = np.load(path)
    pc, sm, gm, fs = g["pc"], g["start_mask"], g["goal_mask"], g["fps_start"]
    pred, score, logp = engine.classify(pc, sm, gm, fps_start=fs, return_logp=True)
    # normalisation is bit-exact
    pc3 = pc if pc.shape[1] == 3 else np.concatenate([pc, np.zeros((len(pc), 1), np.float32)], 1)
    assert np.array_equal(engine.read_buffer("xyz0", np.float32, (1, 2048, 3))[0], O.pc_normalize(pc3))
    # FPS: exact
    for l in range(4):
        got = engine.read_buffer(f"fps{l}", np.int32, (1, NP[l + 1]))[0]
        assert np.array_equal(got, g[f"fps{l}"].astype(np.int32)), f"fps level {l}"
    # ball query: exact outside the rounding band of the expansion distance
    tr = {}
    sd = make_pointnet2_state(int(g["ckpt_seed"]))
    O.classify_path_points(sd, pc, sm, gm, fs, trace=tr)
    for i in range(8):
        K = (16, 32)[i & 1]
        got = engine.read_buffer(f"group{i}", np.int32, (1, NP[i // 2 + 1], K))[0]
        assert _ball_band_ok(got, g[f"group{i}"].astype(np.int32), tr["sqd"][i][0], RADII[i // 2][i & 1], K, NP[i // 2]), f"group {i}"
    # network outputs
    assert np.abs(logp[0] - g["logp"]).max() <= LOGP_TOL, np.abs(logp[0] - g["logp"]).max()
    assert np.abs(score[0] - g["score"]).max() <= LOGP_TOL
    flip = pred[0] != g["pred"]
    assert not np.any(flip & (np.abs(g["score"] - 0.5) >= 0.05))
    assert pred.dtype == np.int64 and score.dtype == np.float32


def test_features_track_the_fp16_emulation(engine):
    """Layer by layer against the oracle's fp16-operand emulation (tight: same rounding points)."""
    from oracle import pointnet2_oracle as O
    g = np.load(GOLD[-1])
    sd = make_pointnet2_state(int(g["ckpt_seed"]))
    tr = {}
    _, _, want = O.classify_path_points(sd, g["pc"], g["start_mask"], g["goal_mask"], g["fps_start"], emulate="fp16", trace=tr)
    _, _, logp = engine.classify(g["pc"], g["start_mask"], g["goal_mask"], fps_start=g["fps_start"], return_logp=True)
    for l in range(1, 5):
        got = engine.read_buffer(f"feat{l}", np.float16, (1, NP[l], CH[l]))[0].astype(np.float32)
        ref = tr["feats"][l - 1][0]
        assert np.abs(got - ref).max() <= 2e-2 * max(1.0, np.abs(ref).max()), (l, np.abs(got - ref).max())
    assert np.abs(logp[0] - want).max() <= 5e-3


def test_batch_equals_singles_and_is_deterministic(engine):
    clouds = [make_cloud_3d(i) for i in range(5)]
    pc = np.stack([c[0] for c in clouds]); sm = np.stack([c[1] for c in clouds]); gm = np.stack([c[2] for c in clouds])
    fs = np.array([[7 * i % 2048, 5 * i % 1024, 3 * i % 256, i % 64] for i in range(5)], dtype=np.int32)
    pred, score, logp = engine.classify(pc, sm, gm, fps_start=fs, return_logp=True)
    pred2, score2, logp2 = engine.classify(pc, sm, gm, fps_start=fs, return_logp=True)
    assert np.array_equal(logp, logp2) and np.array_equal(pred, pred2)
    for i in (0, 3, 4):
        p1, s1, l1 = engine.classify(pc[i], sm[i], gm[i], fps_start=fs[i], return_logp=True)
        assert np.array_equal(l1[0], logp[i])
        assert np.array_equal(p1[0], pred[i]) and np.array_equal(s1[0], score[i])


def test_bad_arguments_raise(engine):
    from nirrt_star_b200._lib import NirrtError
    pc, sm, gm = make_cloud_3d(0)
    with pytest.raises(ValueError):
        engine.classify(pc[:1000], sm[:1000], gm[:1000])
    with pytest.raises(NirrtError):
        engine.classify(np.stack([pc] * 9), np.stack([sm] * 9), np.stack([gm] * 9), fps_start=np.zeros((9, 4), np.int32))


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_dropin_wrapper_matches_reference_golden(path, tmp_path):
    """wrapper{,_3d}.pointnet_pointnet2.pointnet2_wrapper.PNGWrapper: checkpoint file in, same
    outputs as the reference's wrapper, torch's global CPU generator advanced identically."""
    import importlib
    import torch
    from nirrt_star_b200 import dropin
    dropin.install()
    g = np.load(path)
    dim = int(g["dim"])
    d = tmp_path / f"results/model_training/pointnet2_{dim}d/checkpoints"
    d.mkdir(parents=True)
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in make_pointnet2_state(int(g["ckpt_seed"])).items()}
    torch.save({"model_state_dict": sd}, str(d / f"best_pointnet2_{dim}d.pth"))
    mod = importlib.import_module(("wrapper_3d" if dim == 3 else "wrapper") + ".pointnet_pointnet2.pointnet2_wrapper")
    w = mod.PNGWrapper(root_dir=str(tmp_path), device="cuda")
    torch.manual_seed(int(g["seed"]))
    pred, score = w.classify_path_points(g["pc"], g["start_mask"], g["goal_mask"])
    assert pred.shape == (2048,) and pred.dtype == np.int64 and score.dtype == np.float32
    assert np.abs(score - g["score"]).max() <= LOGP_TOL
    assert not np.any((pred != g["pred"]) & (np.abs(g["score"] - 0.5) >= 0.05))
    # the reference would have drawn exactly four randint()s from the global torch generator
    torch.manual_seed(int(g["seed"]))
    for n in (2048, 1024, 256, 64):
        torch.randint(0, n, (1,))
    expect_next = torch.rand(1)
    torch.manual_seed(int(g["seed"]))
    w.classify_path_points(g["pc"], g["start_mask"], g["goal_mask"])
    assert torch.equal(torch.rand(1), expect_next)


@pytest.mark.parametrize("n", [1500, 700, 3000])
def test_short_and_odd_sized_clouds(n):
    """The reference classifies clouds of ANY size: the samplers only down-sample `if len(point_cloud) > n_points`
    (point_cloud_mask_utils_3d.py:104-112), so a cluttered world yields fewer than 2048 points; with fewer than 1024
    points sa1's farthest point sampling re-selects index 0 once every point is taken (pointnet2_utils.py:79-85).
    FPS indices exact, log-probabilities within the fp16-operand tolerance of the fp32 oracle."""
    from nirrt_star_b200.pointnet2 import PointNet2Engine
    from oracle import pointnet2_oracle as O
    sd = make_pointnet2_state(0)
    pc, sm, gm = make_cloud_3d(3)
    rs = np.random.RandomState(n)
    if n <= len(pc):
        keep = np.sort(rs.choice(len(pc), n, replace=False))
        pc, sm, gm = pc[keep], sm[keep], gm[keep]
    else:
        extra = rs.randint(0, len(pc), n - len(pc))
        pc = np.concatenate([pc, pc[extra] + rs.uniform(-0.5, 0.5, (len(extra), 3)).astype(np.float32)])
        sm = np.concatenate([sm, sm[extra]]); gm = np.concatenate([gm, gm[extra]])
    eng = PointNet2Engine(sd, n_points=n, max_batch=1)
    fs = np.array([[n // 3, 11, 12, 13]], dtype=np.int32)
    pred, score, logp = eng.classify(pc, sm, gm, fps_start=fs, return_logp=True)
    tr = {}
    wpred, wscore, want = O.classify_path_points(sd, pc, sm, gm, fs[0], trace=tr)
    got_fps = eng.read_buffer("fps0", np.int32, (1, 1024))[0]
    assert np.array_equal(got_fps, np.asarray(tr["fps"][0][0]).astype(np.int32))
    if n < 1024:
        assert (got_fps[n:] == 0).all() and len(set(got_fps[:n].tolist())) == n
    assert np.abs(logp[0] - want).max() <= LOGP_TOL
    flip = pred[0] != wpred
    assert not np.any(flip & (np.abs(wscore - 0.5) >= 0.05))
    eng.close()


def test_retargeted_engine_equals_dedicated_engine():
    """nirrt_pn2_set_n_points: one engine sized for 2048 points, re-targeted to a 1300-point cloud, must give exactly
    what an engine created for 1300 points gives (the batch planner classifies short clouds this way)."""
    from nirrt_star_b200.pointnet2 import PointNet2Engine
    sd = make_pointnet2_state(0)
    pc, sm, gm = make_cloud_3d(2)
    keep = np.sort(np.random.RandomState(5).choice(len(pc), 1300, replace=False))
    fs = np.array([[100, 11, 12, 13]], dtype=np.int32)
    a = PointNet2Engine(sd, n_points=1300, max_batch=1)
    want = a.classify(pc[keep], sm[keep], gm[keep], fps_start=fs, return_logp=True)
    b = PointNet2Engine(sd, n_points=2048, max_batch=1)
    full = b.classify(pc, sm, gm, fps_start=fs, return_logp=True)
    b.set_n_points(1300)
    got = b.classify(pc[keep], sm[keep], gm[keep], fps_start=fs, return_logp=True)
    for x, y in zip(got, want):
        assert np.array_equal(x, y)
    b.set_n_points(2048)
    again = b.classify(pc, sm, gm, fps_start=fs, return_logp=True)
    for x, y in zip(again, full):
        assert np.array_equal(x, y)
    with pytest.raises(Exception):
        b.set_n_points(4096)
    a.close(); b.close()


def test_pruned_geometry_equals_the_exhaustive_scans(monkeypatch):
    """The bucket-pruned farthest point sampling (selected indices of every level), the slab-pruned ball query and 3-NN search (bitmaps of the points near each coordinate slab, candidates visited in index
    order, the reference's float32 test on each) must select exactly what the exhaustive scans select: group indices of every
    level and the network's outputs bit for bit -- also when the 3-NN candidate radius is so small that nearly every query
    falls back to the exhaustive scan, and on clouds squeezed into a corner of the normalised cube (end slabs)."""
    from nirrt_star_b200.pointnet2 import PointNet2Engine
    sd = make_pointnet2_state(0)
    clouds = [make_cloud_3d(i) for i in range(6)]
    pc = np.stack([c[0] for c in clouds]); sm = np.stack([c[1] for c in clouds]); gm = np.stack([c[2] for c in clouds])
    rs = np.random.RandomState(3)
    pc[4] = pc[4] * np.array([1.0, 0.02, 0.3], dtype=np.float32)             # a thin slab of a cloud
    pc[5, 1000:] = pc[5, 1000:] * 0.01 + 40.0                                 # two far-apart clusters
    fs = np.stack([rs.randint(0, n, 6) for n in (2048, 1024, 256, 64)], 1).astype(np.int32)

    def run(env):
        for k in ("NIRRT_PN2_BQ", "NIRRT_PN2_KNN", "NIRRT_PN2_KNN_REACH", "NIRRT_PN2_FPS_BUCKET"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        e = PointNet2Engine(sd, n_points=2048, max_batch=6)
        out = e.classify(pc, sm, gm, fps_start=fs, return_logp=True)
        groups = [e.read_buffer(f"group{g}", np.int32, (6, (1024, 256, 64, 16)[g // 2], (16, 32)[g & 1])).copy() for g in range(8)]
        groups += [e.read_buffer(f"fps{l}", np.int32, (6, (1024, 256, 64, 16)[l])).copy() for l in range(4)]
        e.close()
        return out, groups

    want, gw = run({"NIRRT_PN2_BQ": "0", "NIRRT_PN2_KNN": "0", "NIRRT_PN2_FPS_BUCKET": "0"})
    for env in ({}, {"NIRRT_PN2_KNN_REACH": "0.02"}, {"NIRRT_PN2_KNN_REACH": "0.9", "NIRRT_PN2_FPS_BUCKET": "1"}):
        got, gg = run(env)
        for a, b in zip(gg, gw):
            assert np.array_equal(a, b), env
        for a, b in zip(got, want):
            assert np.array_equal(a, b), env
