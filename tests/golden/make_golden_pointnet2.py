"""Generates tests/golden/pointnet2_*.npz by running the REFERENCE's own PNGWrapper / get_model
(/root/reference, via oracle/ref_shim.py) on synthetic clouds with the synthetic checkpoint of
nirrt_star_b200.synthetic.make_pointnet2_state.

Run in the build container only:
    python tests/golden/make_golden_pointnet2.py [--calibrate]
The GPU box never runs this (no /root/reference there); it only reads the committed .npz files.

Seeding convention: torch.manual_seed(s) right before the forward; the four FPS start
indices the reference then draws (pointnet2_utils.py:77) are recorded in the fixture as ``fps_start``.
"""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

ref_shim.install()
from nirrt_star_b200.synthetic import make_cloud_3d, make_pointnet2_state  # noqa: E402
from oracle import pointnet2_oracle as O  # noqa: E402
import pointnet_pointnet2.models.pointnet2_utils as ref_utils  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def write_checkpoint(root, sd, dim):
    d = os.path.join(root, f"results/model_training/pointnet2_{dim}d/checkpoints")
    os.makedirs(d, exist_ok=True)
    torch.save({"model_state_dict": {k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}},
               os.path.join(d, f"best_pointnet2_{dim}d.pth"))


def make_wrapper(sd, dim):
    root = tempfile.mkdtemp()
    write_checkpoint(root, sd, dim)
    if dim == 3:
        from wrapper_3d.pointnet_pointnet2.pointnet2_wrapper import PNGWrapper
    else:
        from wrapper.pointnet_pointnet2.pointnet2_wrapper import PNGWrapper
    return PNGWrapper(root_dir=root, device="cpu")


def traced_call(wrapper, pc, sm, gm, seed):
    """Runs the reference wrapper with hooks recording FPS indices and ball-query groups."""
    tr = {"fps": [], "groups": [], "fps_start": []}
    o_fps, o_qbp, o_randint = ref_utils.farthest_point_sample, ref_utils.query_ball_point, torch.randint

    def fps(xyz, npoint):
        out = o_fps(xyz, npoint)
        tr["fps"].append(out.numpy().copy()); tr["fps_start"].append(int(out[0, 0]))
        return out

    def qbp(radius, nsample, xyz, new_xyz):
        out = o_qbp(radius, nsample, xyz, new_xyz)
        tr["groups"].append(out.numpy().copy())
        return out

    ref_utils.farthest_point_sample, ref_utils.query_ball_point = fps, qbp
    captured = {}
    model = wrapper.model
    o_forward = model.forward

    def fwd(x):
        out = o_forward(x)
        captured["logp"] = out[0].detach().numpy().copy()
        return out

    model.forward = fwd
    try:
        torch.manual_seed(seed)
        pred, score = wrapper.classify_path_points(pc, sm, gm)
    finally:
        ref_utils.farthest_point_sample, ref_utils.query_ball_point = o_fps, o_qbp
        model.forward = o_forward
    return pred, score, captured["logp"][0], tr


def run_case(sd, ckpt_seed, env_idx, seed, dim=3, n_points=2048):
    pc, sm, gm = make_cloud_3d(env_idx, n_points)
    if dim == 2:
        pc = np.ascontiguousarray(pc[:, :2] * (224.0 / 50.0))
        c = pc.mean(0)
        sm = (np.linalg.norm(pc - pc[np.argmin(pc[:, 0])], axis=1) < 20).astype(np.float32)
        gm = (np.linalg.norm(pc - pc[np.argmax(pc[:, 0])], axis=1) < 20).astype(np.float32)
    wrapper = make_wrapper(sd, dim)
    pred, score, logp, tr = traced_call(wrapper, pc, sm, gm, seed)
    starts = O.draw_fps_starts(seed, n_points)[0]
    assert list(starts) == tr["fps_start"], (starts, tr["fps_start"])
    name = f"pointnet2_{dim}d_c{ckpt_seed}_e{env_idx}_s{seed}_n{n_points}.npz"
    kw = {f"fps{i}": tr["fps"][i][0].astype(np.int16) for i in range(4)}
    kw.update({f"group{i}": tr["groups"][i][0].astype(np.int16) for i in range(8)})
    np.savez_compressed(os.path.join(OUT, name), dim=dim, ckpt_seed=ckpt_seed, env_idx=env_idx, seed=seed,
                        pc=pc, start_mask=sm, goal_mask=gm, fps_start=starts.astype(np.int32),
                        logp=logp.astype(np.float32), pred=pred.astype(np.int64), score=score.astype(np.float32), **kw)
    print(name, "positives", int(pred.sum()), "/", n_points, "logit range", float(logp.min()), float(logp.max()))


def calibrate(ckpt_seed):
    """Median logit gap on a sample cloud: the conv2.bias[1] shift that makes ~half the points path."""
    sd = make_pointnet2_state(ckpt_seed)
    pc, sm, gm = make_cloud_3d(0)
    wrapper = make_wrapper(sd, 3)
    torch.manual_seed(0)
    x = O.model_inputs(pc, sm, gm)
    with torch.no_grad():
        m = wrapper.model
        # logits before log_softmax: recompute the head by hand from the reference's modules
        feats = {}
        h = m.conv1.register_forward_hook(lambda mod, i, o: feats.__setitem__("c1", o))
        m(x)
        h.remove()
        z = m.conv2(torch.relu(m.bn1(feats["c1"])))[0]
    gap = (z[0] - z[1]).numpy()
    print("ckpt", ckpt_seed, "median gap", float(np.median(gap)), "std", float(gap.std()))


if __name__ == "__main__":
    if "--calibrate" in sys.argv:
        calibrate(0)
        sys.exit(0)
    sd = make_pointnet2_state(0)
    run_case(sd, 0, 0, 3, dim=3)
    run_case(sd, 0, 4, 17, dim=3)
    run_case(sd, 0, 2, 5, dim=2)
