"""Synthetic planning problems (no dataset files exist offline).

Restates the obstacle / start-goal distributions of the reference's dataset generators so that
benchmarks and parity tests run on inputs of the same shape as ``data/random_{2d,3d}``:

* 3D: ``env_configs/random_3d.yml`` + ``generate_random_world_env_3d_raw.py:15-87``
  (world 50^3, 6-9 boxes with sides 8-19, 6-9 balls with radius 8-11, integer coordinates,
  clearance 2, start/goal integer points with squared distance > 50^2 outside inflated obstacles).
* gamma_RRT* : ``datasets_3d/planning_problem_utils_3d.py:77-97`` (Monte-Carlo free volume).

The reference validates start/goal pairs with A* (offline labelling, out of scope); here a pair is
only required to be collision-free, so a problem may be infeasible -- planners then return all-inf.
"""
import math

import numpy as np

XYZ_MAX_3D = (50, 50, 50)


def make_env_3d(seed):
    """Returns an ``env_dict`` in the reference's json schema (rrt_env_3d.py:1-11)."""
    rs = np.random.RandomState(seed)
    xmax, ymax, zmax = XYZ_MAX_3D
    n_boxes = rs.randint(6, 10)
    n_balls = rs.randint(6, 10)
    boxes, balls = [], []
    while len(boxes) < n_boxes:
        x, y, z = rs.randint(0, xmax), rs.randint(0, ymax), rs.randint(0, zmax)
        w, h, d = rs.randint(8, 20), rs.randint(8, 20), rs.randint(8, 20)
        if x < xmax - w and y < ymax - h and z < zmax - d:
            boxes.append([int(x), int(y), int(z), int(w), int(h), int(d)])
    while len(balls) < n_balls:
        x, y, z = rs.randint(0, xmax), rs.randint(0, ymax), rs.randint(0, zmax)
        r = rs.randint(8, 12)
        if r < x < xmax - r and r < y < ymax - r and r < z < zmax - r:
            balls.append([int(x), int(y), int(z), int(r)])
    clearance = 2
    box_a = np.asarray(boxes, dtype=np.float64)
    ball_a = np.asarray(balls, dtype=np.float64)

    def blocked(p):
        in_box = np.any(np.all((box_a[:, :3] - clearance <= p) & (p <= box_a[:, :3] + box_a[:, 3:] + clearance), axis=1))
        in_ball = np.any(((ball_a[:, :3] - p) ** 2).sum(axis=1) <= (ball_a[:, 3] + clearance) ** 2)
        return bool(in_box or in_ball)

    min_d2 = 50 ** 2
    for attempt in range(100000):
        if attempt == 2000:
            min_d2 = 30 ** 2  # very cluttered worlds: relax, still a long query
        sg = rs.randint(low=clearance, high=np.array(XYZ_MAX_3D) - clearance, size=(2, 3))
        if ((sg[0] - sg[1]) ** 2).sum() > min_d2 and not blocked(sg[0]) and not blocked(sg[1]):
            break
    else:
        raise RuntimeError("could not place start/goal")
    return {
        "env_dims": list(XYZ_MAX_3D),
        "box_obstacles": boxes,
        "ball_obstacles": balls,
        "start": [sg[0].tolist()],
        "goal": [sg[1].tolist()],
    }


def gamma_rrt_star_3d(env_dict, seed, n_points=100000):
    """search_radius as compute_gamma_rrt_star_3d does it, with a private RandomState."""
    rs = np.random.RandomState(seed)
    dims = env_dict["env_dims"]  # (height, width, depth) -> y, x, z ranges (rrt_env_3d.py:6-9)
    pts = np.stack([rs.uniform(0, dims[1], n_points), rs.uniform(0, dims[0], n_points),
                    rs.uniform(0, dims[2], n_points)], axis=1)
    box = np.asarray(env_dict["box_obstacles"], dtype=np.float64).reshape(-1, 6)
    ball = np.asarray(env_dict["ball_obstacles"], dtype=np.float64).reshape(-1, 4)
    inside = np.zeros(n_points, dtype=bool)
    for b in box:
        inside |= np.all((b[:3] <= pts) & (pts <= b[:3] + b[3:]), axis=1)
    for b in ball:
        inside |= ((pts - b[:3]) ** 2).sum(axis=1) < b[3] ** 2
    free_vol = dims[0] * dims[1] * dims[2] * (1 - inside.mean())
    return math.ceil((2 * (1 + 1. / 3)) ** (1. / 3) * (free_vol / (4. / 3. * np.pi)) ** (1. / 3))


def make_problem_3d(env_idx, base_seed=100):
    """A ``problem`` dict with the keys get_random_3d_problem_input returns
    (planning_problem_utils_3d.py:62-75), minus the reference's own ``Env`` object (callers
    wrap ``env_dict`` with whichever Env class they use)."""
    env_dict = make_env_3d(base_seed + env_idx)
    return {
        "x_start": tuple(env_dict["start"][0]),
        "x_goal": tuple(env_dict["goal"][0]),
        "env_dict": env_dict,
        "search_radius": gamma_rrt_star_3d(env_dict, base_seed + env_idx),
    }


# ---------------------------------------------------------------------------------------------
# Synthetic PointNet++ checkpoint (the trained weights are a Google-Drive download, absent offline)

POINTNET2_SA = [  # npoint, radii, nsamples, in_channel, mlps            (pointnet2.py:11-14)
    (1024, (0.05, 0.1), (16, 32), 6, ((16, 16, 32), (32, 32, 64))),
    (256, (0.1, 0.2), (16, 32), 96, ((64, 64, 128), (64, 96, 128))),
    (64, (0.2, 0.4), (16, 32), 256, ((128, 196, 256), (128, 196, 256))),
    (16, (0.4, 0.8), (16, 32), 512, ((256, 256, 512), (256, 384, 512))),
]
POINTNET2_FP = [  # name, in_channel, mlp                                  (pointnet2.py:15-18)
    ("fp4", 1536, (256, 256)), ("fp3", 512, (256, 256)), ("fp2", 352, (256, 128)), ("fp1", 128, (128, 128, 128)),
]
# conv2.bias[1] shift that makes ~half of a uniform cloud "path" for the seed-0 checkpoint
# (SURVEY.md 8c item 4); measured once with tests/golden/make_golden_pointnet2.py --calibrate
POINTNET2_BIAS_SHIFT = {0: 0.84}


def make_pointnet2_state(seed=0, num_classes=2):
    """A ``model_state_dict`` with exactly the reference's 240 keys and shapes
    (pointnet_pointnet2/models/pointnet2.py:8-22), He-uniform weights, non-trivial BatchNorm
    statistics so that BN folding is exercised.  numpy arrays, generated from ``seed`` only."""
    rs = np.random.RandomState(1000 + seed)
    sd = {}

    def conv(prefix, cin, cout, kdims):
        b = math.sqrt(6.0 / cin)
        sd[prefix + ".weight"] = rs.uniform(-b, b, (cout, cin) + (1,) * kdims).astype(np.float32)
        sd[prefix + ".bias"] = rs.uniform(-0.1, 0.1, cout).astype(np.float32)

    def bn(prefix, c):
        sd[prefix + ".weight"] = rs.uniform(0.7, 1.3, c).astype(np.float32)
        sd[prefix + ".bias"] = rs.uniform(-0.2, 0.2, c).astype(np.float32)
        sd[prefix + ".running_mean"] = (0.2 * rs.standard_normal(c)).astype(np.float32)
        sd[prefix + ".running_var"] = rs.uniform(0.6, 1.6, c).astype(np.float32)
        sd[prefix + ".num_batches_tracked"] = np.array(0, dtype=np.int64)

    for li, (_, _, _, cin, mlps) in enumerate(POINTNET2_SA, start=1):
        for si, mlp in enumerate(mlps):
            last = cin + 3
            for j, cout in enumerate(mlp):
                conv(f"sa{li}.conv_blocks.{si}.{j}", last, cout, 2)
                bn(f"sa{li}.bn_blocks.{si}.{j}", cout)
                last = cout
    for name, cin, mlp in POINTNET2_FP:
        last = cin
        for j, cout in enumerate(mlp):
            conv(f"{name}.mlp_convs.{j}", last, cout, 1)
            bn(f"{name}.mlp_bns.{j}", cout)
            last = cout
    conv("conv1", 128, 128, 1)
    bn("bn1", 128)
    conv("conv2", 128, num_classes, 1)
    sd["conv2.bias"][1] += np.float32(POINTNET2_BIAS_SHIFT.get(seed, 0.0))
    return sd


def make_cloud_3d(env_idx, n_points=2048, base_seed=100):
    """(pc f32 (n,3), start_mask f32 (n,), goal_mask f32 (n,)) : uniform free-space samples of a
    synthetic random_3d world plus the start/goal neighbourhood masks the planners feed the
    network (datasets/point_cloud_mask_utils.py:20-31, nirrt_star_png_3d.py:157-166)."""
    pr = make_problem_3d(env_idx, base_seed)
    rs = np.random.RandomState(7000 + env_idx)
    ed = pr["env_dict"]
    box = np.asarray(ed["box_obstacles"], dtype=np.float64).reshape(-1, 6)
    ball = np.asarray(ed["ball_obstacles"], dtype=np.float64).reshape(-1, 4)
    pts = np.zeros((0, 3))
    while len(pts) < n_points:
        p = rs.uniform(0, 50, (4 * n_points, 3))
        inside = np.zeros(len(p), dtype=bool)
        for b in box:
            inside |= np.all((b[:3] - 2 <= p) & (p <= b[:3] + b[3:] + 2), axis=1)
        for b in ball:
            inside |= ((p - b[:3]) ** 2).sum(axis=1) <= (b[3] + 2) ** 2
        pts = np.concatenate([pts, p[~inside]])
    pc = pts[:n_points].astype(np.float32)
    r = 10.0
    sm = (np.linalg.norm(pc - np.array(pr["x_start"], dtype=np.float32), axis=1) < r).astype(np.float32)
    gm = (np.linalg.norm(pc - np.array(pr["x_goal"], dtype=np.float32), axis=1) < r).astype(np.float32)
    return pc, sm, gm


# ---------------------------------------------------------------------------------------------
# Synthetic random_2d worlds (env_configs/random_2d.yml + generate_random_world_env_2d.py:14-47)

IMG_2D = (224, 224)


def rasterize_2d(rects, circles, hw=IMG_2D):
    """binary_mask (1 = free) of the world image the reference draws with cv2.rectangle/circle
    (filled, inclusive pixel extents); the drawing itself is restated with numpy."""
    h, w = hw
    yy, xx = np.mgrid[0:h, 0:w]
    occ = np.zeros((h, w), dtype=bool)
    for x, y, rw, rh in rects:
        occ |= (xx >= x) & (xx <= x + rw) & (yy >= y) & (yy <= y + rh)
    for x, y, r in circles:
        occ |= (xx - x) ** 2 + (yy - y) ** 2 <= r * r
    return (~occ).astype(float)


def make_env_2d(seed):
    import random as _r
    rnd = _r.Random(seed)
    h, w = IMG_2D
    n_rect, n_circ = rnd.randint(8, 12), rnd.randint(8, 12)
    rects = [[rnd.randint(0, w), rnd.randint(0, h), rnd.randint(16, 24), rnd.randint(16, 24)] for _ in range(n_rect)]
    circles = [[rnd.randint(0, w), rnd.randint(0, h), rnd.randint(16, 24)] for _ in range(n_circ)]
    mask = rasterize_2d(rects, circles)
    c = 3

    def window_free(p):
        x, y = p
        if x - c < 0 or y - c < 0 or x + c >= w or y + c >= h:
            return False
        return bool(mask[y - c:y + c + 1, x - c:x + c + 1].all())

    for _ in range(100000):
        s = (rnd.randint(0, w - 1), rnd.randint(0, h - 1)); g = (rnd.randint(0, w - 1), rnd.randint(0, h - 1))
        if abs(s[0] - g[0]) >= 50 and abs(s[1] - g[1]) >= 50 and window_free(s) and window_free(g):
            break
    else:
        raise RuntimeError("could not place start/goal")
    return {"env_dims": [h, w], "rectangle_obstacles": rects, "circle_obstacles": circles,
            "start": [list(s)], "goal": [list(g)]}, mask


def make_problem_2d(env_idx, base_seed=100):
    """A 2D ``problem`` dict with the keys get_random_2d_problem_input returns
    (planning_problem_utils_2d.py:145-162), minus the reference's ``Env`` object."""
    env_dict, mask = make_env_2d(base_seed + env_idx)
    gamma = math.ceil((2 * (1 + 1. / 2)) ** (1. / 2) * (mask.sum() / np.pi) ** (1. / 2))   # compute_gamma_rrt_star
    return {"x_start": tuple(env_dict["start"][0]), "x_goal": tuple(env_dict["goal"][0]), "env_dict": env_dict,
            "binary_mask": mask, "search_radius": gamma}
