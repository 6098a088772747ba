"""CPU parity oracle for the PointNet++ guidance-state inference path (TEST INFRASTRUCTURE ONLY).

A plain torch-fp32 (CPU) restatement of the reference's forward, written functionally over a
``model_state_dict`` so that it needs none of the reference's modules at run time:

    classify_path_points   wrapper_3d/pointnet_pointnet2/pointnet2_wrapper.py:28-59
                           wrapper/pointnet_pointnet2/pointnet2_wrapper.py:28-64 (2D: z padded with 0)
    pc_normalize           pointnet_pointnet2/models/pointnet2_utils.py:13-18
    square_distance        pointnet2_utils.py:21-42
    farthest_point_sample  pointnet2_utils.py:65-86   (start index is an explicit input here;
                           the reference draws it from torch's CPU generator, :77)
    query_ball_point       pointnet2_utils.py:89-109
    SA-MSG forward         pointnet2_utils.py:226-264
    FP forward             pointnet2_utils.py:278-317
    get_model.forward      pointnet_pointnet2/models/pointnet2.py:24-42

Pinned against the reference's own get_model / PNGWrapper run in the build container:
tests/golden/make_golden_pointnet2.py -> tests/golden/pointnet2_*.npz, checked by
tests/test_oracle_pointnet2.py.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
reference arm may import this module; the product path (nirrt_star_b200/) never does.

``emulate='fp16'`` additionally rounds every GEMM operand (weights with folded BatchNorm, and
activations) to half precision at the points where the CUDA path does (coordinates travel as a
hi+lo pair), which gives a tight expected value for debugging the kernels; it is not the parity
reference.  ('bf16' and 'tf32' are there to document why fp16 operands were chosen: measured max
log-probability error vs the reference 3.0e-2 / 8e-3 / 6e-3 for bf16 / tf32 / fp16.)
"""
import numpy as np
import torch

SA_SPEC = [  # npoint, radii, nsamples   (pointnet2.py:11-14)
    (1024, (0.05, 0.1), (16, 32)),
    (256, (0.1, 0.2), (16, 32)),
    (64, (0.2, 0.4), (16, 32)),
    (16, (0.4, 0.8), (16, 32)),
]
FP_NAMES = ("fp4", "fp3", "fp2", "fp1")
BN_EPS = 1e-5


def pc_normalize(pc):
    centroid = np.mean(pc, axis=0)
    pc = pc - centroid
    m = np.max(np.sqrt(np.sum(pc ** 2, axis=1)))
    return pc / m


def square_distance(src, dst):
    dist = -2 * torch.matmul(src, dst.permute(0, 2, 1))
    dist += torch.sum(src ** 2, -1).unsqueeze(-1)
    dist += torch.sum(dst ** 2, -1).unsqueeze(1)
    return dist


def index_points(points, idx):
    B = points.shape[0]
    bi = torch.arange(B).view([B] + [1] * (idx.dim() - 1)).expand_as(idx)
    return points[bi, idx, :]


def farthest_point_sample(xyz, npoint, start):
    B, N, _ = xyz.shape
    centroids = torch.zeros(B, npoint, dtype=torch.long)
    distance = torch.ones(B, N) * 1e10
    farthest = start.clone().long()
    bi = torch.arange(B)
    for i in range(npoint):
        centroids[:, i] = farthest
        c = xyz[bi, farthest, :].view(B, 1, 3)
        dist = torch.sum((xyz - c) ** 2, -1)
        distance = torch.minimum(distance, dist)
        farthest = torch.max(distance, -1)[1]
    return centroids


def query_ball_point(radius, nsample, xyz, new_xyz):
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    group_idx = torch.arange(N).view(1, 1, N).repeat(B, S, 1)
    d = square_distance(new_xyz, xyz)
    group_idx[d > radius ** 2] = N
    group_idx = group_idx.sort(dim=-1)[0][:, :, :nsample]
    first = group_idx[:, :, 0:1].expand(-1, -1, nsample)
    mask = group_idx == N
    group_idx[mask] = first[mask]
    return group_idx, d


def _tf32(t):
    i = t.contiguous().view(torch.int32)
    return ((i + 0x0fff + ((i >> 13) & 1)) & ~0x1fff).view(torch.float32)


_ROUND = {
    "fp16": lambda t: t.clamp(-65504.0, 65504.0).to(torch.float16).to(torch.float32),
    "bf16": lambda t: t.to(torch.bfloat16).to(torch.float32),
    "tf32": _tf32,
}


def _folded(sd, conv, bn):
    """(W [Cout,Cin], b [Cout]) of conv followed by eval-mode BatchNorm."""
    w = torch.as_tensor(sd[conv + ".weight"]).float().flatten(1)
    b = torch.as_tensor(sd[conv + ".bias"]).float()
    g = torch.as_tensor(sd[bn + ".weight"]).float()
    be = torch.as_tensor(sd[bn + ".bias"]).float()
    mu = torch.as_tensor(sd[bn + ".running_mean"]).float()
    var = torch.as_tensor(sd[bn + ".running_var"]).float()
    s = g / torch.sqrt(var + BN_EPS)
    return w * s[:, None], (b - mu) * s + be


def _conv_bn_relu(sd, conv, bn, x, emulate, exact_cols=None):
    """x [..., Cin] -> relu(bn(conv(x))) [..., Cout].  Unfused when not emulating (mirrors
    F.relu(bn(conv(.))) of the reference); folded + rounded operands when emulating.
    exact_cols: input columns the CUDA path feeds as a hi+lo pair (coordinates)."""
    if not emulate:
        w = torch.as_tensor(sd[conv + ".weight"]).float().flatten(1)
        y = x @ w.t() + torch.as_tensor(sd[conv + ".bias"]).float()
        g = torch.as_tensor(sd[bn + ".weight"]).float()
        be = torch.as_tensor(sd[bn + ".bias"]).float()
        mu = torch.as_tensor(sd[bn + ".running_mean"]).float()
        var = torch.as_tensor(sd[bn + ".running_var"]).float()
        y = (y - mu) / torch.sqrt(var + BN_EPS) * g + be
        return torch.relu(y)
    rnd = _ROUND[emulate]
    w, b = _folded(sd, conv, bn)
    wq = rnd(w)
    xq = rnd(x)
    y = xq @ wq.t()
    if exact_cols is not None:
        lo = rnd(x[..., exact_cols] - xq[..., exact_cols])
        y = y + lo @ wq[:, exact_cols].t()
    return torch.relu(y + b)


def forward(sd, inputs, fps_start, emulate=None, trace=None):
    """inputs [B,6,N] f32 (normalised xyz + start/goal/free masks), fps_start [B,4] int.
    Returns log-probabilities [B,N,num_classes]; ``trace`` (dict) receives FPS indices, ball-query
    groups and per-level features."""
    inputs = torch.as_tensor(inputs).float()
    fps_start = torch.as_tensor(fps_start).long()
    em = emulate
    xyz = inputs[:, :3, :].permute(0, 2, 1).contiguous()       # [B,N,3]
    pts = inputs.permute(0, 2, 1).contiguous()                 # [B,N,6]
    levels = [(xyz, pts)]
    if trace is not None:
        trace.update(fps=[], groups=[], sqd=[], feats=[])
    for li, (npoint, radii, ks) in enumerate(SA_SPEC, start=1):
        xyz, pts = levels[-1]
        B, N, _ = xyz.shape
        fidx = farthest_point_sample(xyz, npoint, fps_start[:, li - 1])
        new_xyz = index_points(xyz, fidx)
        outs = []
        if trace is not None:
            trace["fps"].append(fidx.numpy().copy())
        for si, (radius, K) in enumerate(zip(radii, ks)):
            gidx, sqd = query_ball_point(radius, K, xyz, new_xyz)
            if trace is not None:
                trace["groups"].append(gidx.numpy().copy())
                trace["sqd"].append(sqd.numpy().copy())
            g_xyz = index_points(xyz, gidx) - new_xyz.view(B, npoint, 1, 3)
            g = torch.cat([index_points(pts, gidx), g_xyz], dim=-1)    # [B,S,K,D+3], features first
            cin = g.shape[-1]
            exact = None
            if em:   # coordinates (absolute at sa1, relative everywhere) travel as hi+lo pairs
                exact = list(range(cin - 3, cin)) + ([0, 1, 2] if li == 1 else [])
            for j in range(3):
                g = _conv_bn_relu(sd, f"sa{li}.conv_blocks.{si}.{j}", f"sa{li}.bn_blocks.{si}.{j}", g, em,
                                  exact if j == 0 else None)
            outs.append(g.max(dim=2)[0])
        new_pts = torch.cat(outs, dim=-1)                       # [B,S,C]
        if em:
            new_pts = _ROUND[em](new_pts)
        levels.append((new_xyz, new_pts))
        if trace is not None:
            trace["feats"].append(new_pts.numpy().copy())
    # feature propagation: fp4 (l3 <- l4), fp3 (l2 <- l3), fp2 (l1 <- l2), fp1 (l0 <- l1, no skip)
    feats = [lv[1] for lv in levels]
    up = feats[4]
    for k, name in enumerate(FP_NAMES):
        lo = 3 - k
        xyz1, xyz2 = levels[lo][0], levels[lo + 1][0]
        d = square_distance(xyz1, xyz2)
        d, idx = d.sort(dim=-1)
        d, idx = d[:, :, :3], idx[:, :, :3]
        rec = 1.0 / (d + 1e-8)
        w = rec / rec.sum(dim=2, keepdim=True)
        interp = (index_points(up, idx) * w.unsqueeze(-1)).sum(dim=2)
        x = interp if lo == 0 else torch.cat([feats[lo], interp], dim=-1)
        j = 0
        while f"{name}.mlp_convs.{j}.weight" in sd:
            x = _conv_bn_relu(sd, f"{name}.mlp_convs.{j}", f"{name}.mlp_bns.{j}", x, em)
            j += 1
        if em:
            x = _ROUND[em](x)
        up = x
        if trace is not None:
            trace["feats"].append(up.numpy().copy())
    x = _conv_bn_relu(sd, "conv1", "bn1", up, em)
    w2 = torch.as_tensor(sd["conv2.weight"]).float().flatten(1)
    b2 = torch.as_tensor(sd["conv2.bias"]).float()
    x = x @ w2.t() + b2
    return torch.log_softmax(x, dim=-1)


def model_inputs(pc, start_mask, goal_mask):
    """The [1,6,N] tensor classify_path_points builds (pointnet2_wrapper.py:46-59)."""
    pc = np.asarray(pc, dtype=np.float32)
    if pc.shape[1] == 2:
        pc = np.concatenate([pc, np.zeros((pc.shape[0], 1), dtype=np.float32)], axis=1)
    xyz = pc_normalize(pc)
    free = 1 - (start_mask + goal_mask).astype(bool)
    f = np.stack((start_mask, goal_mask, free.astype(np.float32)), axis=-1)
    x = np.concatenate([xyz, f], axis=1).astype(np.float32)
    return torch.from_numpy(x).permute(1, 0).unsqueeze(0).contiguous()


def classify_path_points(sd, pc, start_mask, goal_mask, fps_start, emulate=None, trace=None):
    with torch.no_grad():
        logp = forward(sd, model_inputs(pc, start_mask, goal_mask), np.asarray(fps_start).reshape(1, 4),
                       emulate, trace)
        pred = np.argmax(logp.numpy(), 2)[0]
        score = torch.softmax(logp, dim=-1)[0, :, 1].numpy()
    return pred, score, logp[0].numpy()


def draw_fps_starts(seed, n_points=2048, batch=1):
    """The four start indices the reference's forward draws after torch.manual_seed(seed)
    (torch.randint(0, N, (B,)) on the CPU generator, once per SA layer, pointnet2_utils.py:77)."""
    torch.manual_seed(seed)
    ns = [n_points, 1024, 256, 64]
    return np.stack([torch.randint(0, n, (batch,), dtype=torch.long).numpy() for n in ns], axis=1)
