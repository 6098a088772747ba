"""Batched offline inference (eval_pointnet_pointnet2.py:108-161) on the sm_100a engine against the
same loop run with the torch-fp32 oracle forward: identical bookkeeping, metrics within the fp16-operand tolerance."""
import numpy as np
import pytest

from nirrt_star_b200.synthetic import make_cloud_3d, make_pointnet2_state

pytestmark = pytest.mark.gpu


def _dataset(n=37, N=2048):
    pcs, sms, gms, labels = [], [], [], []
    for i in range(n):
        pc, sm, gm = make_cloud_3d(i % 12)
        rs = np.random.RandomState(i)
        pc = pc + rs.uniform(-0.05, 0.05, pc.shape).astype(np.float32)
        a = pc[sm.argmax()]; b = pc[gm.argmax()]
        t = np.clip(((pc - a) @ (b - a)) / max(1e-9, float((b - a) @ (b - a))), 0, 1)
        labels.append((np.linalg.norm(pc - (a + t[:, None] * (b - a)), axis=1) < 6).astype(np.float32))
        pcs.append(pc); sms.append(sm); gms.append(gm)
    sm, gm = np.stack(sms), np.stack(gms)
    return {"pc": np.stack(pcs), "start": sm, "goal": gm, "free": 1 - ((sm + gm) > 0).astype(np.float32),
            "astar": np.stack(labels), "token": np.arange(n)}


def test_offline_eval_matches_oracle_loop():
    import torch
    from nirrt_star_b200.eval_pointnet2 import PathPlanArrays, evaluate
    from nirrt_star_b200.pointnet2 import draw_fps_starts
    from oracle import pointnet2_oracle as O
    sd = make_pointnet2_state(0)
    data = _dataset()
    ds = PathPlanArrays(data)
    torch.manual_seed(5)
    got = evaluate(sd, ds, batch_size=16, return_predictions=True)
    assert got["pred"].shape == (37, 2048) and got["scores"].shape == (37, 2048, 2)
    # the same loop with the oracle forward (same FPS start draws)
    torch.manual_seed(5)
    sdt = {k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}
    correct = seen = 0
    cc = np.zeros(2); deno = np.zeros(2); loss = 0.0; nb = 0
    w = ds.labelweights.astype(np.float64)
    flips = 0
    for b0 in range(0, 37, 16):
        sl = slice(b0, min(37, b0 + 16))
        fs = draw_fps_starts(sl.stop - sl.start, 2048)
        for k, i in enumerate(range(sl.start, sl.stop)):
            pred, score, logp = O.classify_path_points(sdt, ds.pc[i], ds.start_mask[i], ds.goal_mask[i], fs[k])
            lab = ds.astar_mask[i].astype(np.int64)
            flips += int(((got["pred"][i] != pred) & (np.abs(score - 0.5) >= 0.05)).sum())
            correct += int((pred == lab).sum()); seen += 2048
            for c in range(2):
                cc[c] += np.sum((pred == c) & (lab == c)); deno[c] += np.sum((pred == c) | (lab == c))
        # per-batch weighted NLL over the whole batch
        lab_b = ds.astar_mask[sl].astype(np.int64)
        lp = np.stack([O.classify_path_points(sdt, ds.pc[i], ds.start_mask[i], ds.goal_mask[i], fs[k])[2]
                       for k, i in enumerate(range(sl.start, sl.stop))]).astype(np.float64)
        picked = np.take_along_axis(lp, lab_b[..., None], axis=2)[..., 0]
        loss += float(-(w[lab_b] * picked).sum() / w[lab_b].sum()); nb += 1
    assert flips == 0
    assert abs(got["accuracy"] - correct / seen) < 5e-3
    assert abs(got["mIoU"] - float(np.mean(cc / (deno + 1e-6)))) < 1e-2
    assert abs(got["mean_loss"] - loss / nb) < 2e-2
    bad = dict(data); bad["free"] = np.ones_like(data["free"])
    with pytest.raises(ValueError):
        PathPlanArrays(bad)
