"""NIRRT*-PNG / NRRT*-PNG drop-in classes (3D and 2D) vs fixtures recorded from the reference's own classes
(tests/golden/make_golden_neural_planner.py).  The network is replaced on both sides by recorded
predictions, so this pins: the device loop body with the cloud / informed / free sampler switch, the
cloud-update trigger (c_best < ratio * c_update) and its pause/resume, the shared numpy stream
hand-over, and the guidance-cloud generation (uniform / ellipsoid draws, CUDA obstacle filters,
CUDA farthest-point down-sampling) -- every 3D cloud must hash identically to the reference's; 2D
clouds are compared to 1e-5 because the ellipse that is sampled depends on c_best, i.e. on vertex
coordinates that may differ from the reference in their last bit (libm steer, see
tests/test_gpu_planner2d.py)."""
import glob
import hashlib
import os
import random
import types

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "neural[23]d_*.npz")))


class ReplayWrapper:
    def __init__(self, g):
        self.g, self.k = g, 0

    def classify_path_points(self, pc, start_mask, goal_mask):
        g, k = self.g, self.k
        assert k < int(g["n_calls"]), "more cloud updates than the reference made"
        assert pc.dtype == np.float32 and len(pc) == int(g["call_n"][k])
        if int(g["dim"]) == 3:
            assert hashlib.sha1(np.ascontiguousarray(pc).tobytes()).hexdigest() == str(g["call_pc_sha1"][k]), f"cloud {k} differs"
            assert hashlib.sha1(start_mask.tobytes() + goal_mask.tobytes()).hexdigest() == str(g["call_mask_sha1"][k])
        else:
            assert np.array_equal(pc, g[f"call_pc{k}"]), f"cloud {k} differs"
        pred = np.unpackbits(g["call_pred"][k])[:len(pc)].astype(np.int64)
        self.k += 1
        return pred, pred.astype(np.float32)


@pytest.fixture(scope="module", autouse=True)
def dropin():
    from nirrt_star_b200 import dropin
    dropin.install()


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_neural_planner_matches_reference_golden(path):
    import importlib
    from nirrt_star_b200.synthetic import make_problem_2d, make_problem_3d
    g = np.load(path)
    kind, mode, dim = str(g["kind"]), str(g["mode"]), int(g["dim"])
    pkg = "path_planning_classes_3d." if dim == 3 else "path_planning_classes."
    mod = importlib.import_module(pkg + {"nirrt": "nirrt_star_png_", "nrrt": "nrrt_star_png_"}[kind] + f"{dim}d")
    problem = (make_problem_3d if dim == 3 else make_problem_2d)(int(g["env_idx"]))
    args = types.SimpleNamespace(step_len=10, iter_max=int(g["iter_max"]), clearance=2 if dim == 3 else 3, pc_n_points=2048,
                                 pc_over_sample_scale=5, pc_sample_rate=float(g["pc_sample_rate"]),
                                 pc_update_cost_ratio=float(g["ratio"]))
    seed = int(g["seed"])
    np.random.seed(seed); random.seed(seed)
    w = ReplayWrapper(g)
    planner = mod.get_path_planner(args, problem, w)
    if mode == "planning":
        planner.planning(False)
        if len(g["path"]):
            assert np.array_equal(planner.path, g["path"])
        else:
            assert len(planner.path) == 0
    else:
        lst = planner.planning_random(int(g["iter_after"]))
        want = g["path_len_list"]
        assert len(lst) == len(want)
        assert np.array_equal(np.isinf(lst), np.isinf(want))
        f = np.isfinite(want)
        assert np.allclose(np.array(lst)[f], want[f], rtol=1e-5, atol=0)
    assert w.k == int(g["n_calls"])
    n = planner.num_vertices
    assert n == int(g["num_vertices"])
    assert np.array_equal(planner.vertex_parents[:n], g["parents"])
    assert np.array_equal(planner.vertices[:n], g["vertices"])
    if kind == "nirrt":
        assert list(planner.path_solutions) == list(g["solutions"])
    assert np.random.random() == float(g["next_random"])
    assert random.random() == float(g["next_py_random"])


def test_fps_f64_matches_numpy():
    from nirrt_star_b200.batch import fps_f64
    rs = np.random.RandomState(3)
    for n, m in ((5000, 2048), (9000, 2048), (300, 300), (16384, 100)):
        pts = rs.uniform(0, 50, (n, 3))
        dist = np.full(n, np.inf); far = 0; want = []
        for _ in range(m):
            want.append(far)
            dist = np.minimum(dist, ((pts - pts[far]) ** 2).sum(axis=1))
            far = int(np.argmax(dist))
        assert np.array_equal(fps_f64(pts, m), np.array(want))


def test_nirrt_with_the_cuda_network_end_to_end(tmp_path):
    """NIRRTStarPNG3D + the sm_100a PNGWrapper (synthetic checkpoint): runs, returns a valid path."""
    import torch
    from nirrt_star_b200.synthetic import make_pointnet2_state, make_problem_3d
    from path_planning_classes_3d.nirrt_star_png_3d import get_path_planner
    from wrapper_3d.pointnet_pointnet2.pointnet2_wrapper import PNGWrapper
    d = tmp_path / "results/model_training/pointnet2_3d/checkpoints"
    d.mkdir(parents=True)
    torch.save({"model_state_dict": {k: torch.from_numpy(np.asarray(v)) for k, v in make_pointnet2_state(0).items()}},
               str(d / "best_pointnet2_3d.pth"))
    w = PNGWrapper(root_dir=str(tmp_path), device="cuda")
    problem = make_problem_3d(0)
    args = types.SimpleNamespace(step_len=10, iter_max=2000, clearance=2, pc_n_points=2048, pc_over_sample_scale=5,
                                 pc_sample_rate=0.5, pc_update_cost_ratio=0.9)
    np.random.seed(1); torch.manual_seed(1)
    planner = get_path_planner(args, problem, w)
    lst = planner.planning_random(200)
    assert np.isfinite(lst[-1]) and len(planner.path_point_cloud_pred) > 0
    assert all(b <= a + 1e-9 for a, b in zip(lst[-200:], lst[-199:]))      # best cost never increases


def test_device_cloud_sampler_matches_numpy_on_an_empty_world():
    """generate_rectangle_point_cloud_3d / ellipsoid_point_cloud_sampling_3d on the device vs the plain numpy recipe
    (datasets_3d/point_cloud_mask_utils_3d.py:83-113,132-200) in a world without obstacles: every raw sample survives,
    so the farthest point down-sampling runs on more candidates than its shared-memory staging holds (the L2 path),
    and the global numpy stream must end at the same position."""
    from datasets_3d.point_cloud_mask_utils_3d import ellipsoid_point_cloud_sampling_3d, generate_rectangle_point_cloud_3d
    from path_planning_utils_3d.rrt_env_3d import Env
    env = Env({"env_dims": [50, 50, 50], "box_obstacles": [], "ball_obstacles": []})

    def fps(pts, m):
        dist = np.full(len(pts), np.inf); far = 0; sel = []
        for _ in range(m):
            sel.append(far)
            dist = np.minimum(dist, ((pts - pts[far]) ** 2).sum(axis=1))
            far = int(np.argmax(dist))
        return pts[sel]

    np.random.seed(123)
    got = generate_rectangle_point_cloud_3d(env, 2048, over_sample_scale=5)
    after = np.random.random()
    np.random.seed(123)
    raw = np.random.uniform(low=(0, 0, 0), high=(50, 50, 50), size=(10240, 3))
    assert np.random.random() == after
    assert np.array_equal(got, fps(raw, 2048))
    # ellipsoid (all of it inside the world): (r, theta, phi) vectors, np.sin / np.cos, C @ L @ x + centre
    a, b = np.array([20., 25., 25.]), np.array([30., 25., 25.])
    np.random.seed(7)
    got = ellipsoid_point_cloud_sampling_3d(a, b, 1.5, env, n_points=2048, n_raw_samples=10240)
    after = np.random.random()
    np.random.seed(7)
    c_min = np.linalg.norm(b - a); c_max = c_min * 1.5
    r = np.array([c_max / 2, np.sqrt(c_max ** 2 - c_min ** 2) / 2, np.sqrt(c_max ** 2 - c_min ** 2) / 2])
    a1 = (b - a) / c_min
    U, _, V = np.linalg.svd(np.outer(a1, [1, 0, 0]))
    Crot = U @ np.diag([1, 1, np.linalg.det(U) * np.linalg.det(V)]) @ V.T
    rad = np.random.uniform(0.0, 1.0, 10240); th = np.random.uniform(0, np.pi, 10240); ph = np.random.uniform(0, 2 * np.pi, 10240)
    smp = np.array([rad * np.sin(th) * np.cos(ph), rad * np.sin(th) * np.sin(ph), rad * np.cos(th)]).T
    pc = np.dot(np.dot(Crot, np.diag(r)), smp.T).T + (a + b) / 2.
    assert np.random.random() == after
    assert np.array_equal(got, fps(pc, 2048))


def test_device_cloud_sampler_2d_matches_numpy_recipe():
    """generate_rectangle_point_cloud / ellipsoid_point_cloud_sampling on the device vs the plain numpy recipe
    (datasets/point_cloud_mask_utils.py:35-73,104-174) on an image with blocked rectangles: uniform draws, the 4-pixel
    free-space product, the unit-disc rejection, (C @ L) @ x + centre through a 3x3 dgemm, the range test, farthest point
    down-sampling, and the position of the global numpy stream afterwards."""
    import math
    from datasets.point_cloud_mask_utils import ellipsoid_point_cloud_sampling, generate_rectangle_point_cloud
    H, W = 160, 224
    mask = np.ones((H, W))
    mask[30:70, 50:90] = 0; mask[100:150, 120:200] = 0; mask[0:12, 180:224] = 0

    def free(pc):
        pix = pc.astype(int)
        nei = (pix + np.array([[0, 0], [0, 1], [1, 0], [1, 1]])[:, np.newaxis]).reshape(-1, 2)
        nei[:, 1] = np.clip(nei[:, 1], 0, H - 1); nei[:, 0] = np.clip(nei[:, 0], 0, W - 1)
        return np.prod(mask[nei[:, 1], nei[:, 0]].reshape(4, -1), axis=0).nonzero()[0]

    def fps(pts, m):
        dist = np.full(len(pts), np.inf); far = 0; sel = []
        for _ in range(m):
            sel.append(far)
            dist = np.minimum(dist, ((pts - pts[far]) ** 2).sum(axis=1))
            far = int(np.argmax(dist))
        return pts[sel]

    np.random.seed(11)
    got = generate_rectangle_point_cloud(mask, 2048, over_sample_scale=5)
    after = np.random.random()
    np.random.seed(11)
    raw = np.random.uniform(low=[0, 0], high=[W, H], size=(10240, 2))
    assert np.random.random() == after
    raw = raw[free(raw)]
    assert len(raw) > 2048 and np.array_equal(got, fps(raw, 2048))

    for seed, a, b, ratio, n_pts in ((5, [40., 90.], [180., 60.], 1.25, 2048), (6, [10., 10.], [200., 150.], 1.02, 2048),
                                     (7, [100., 80.], [110., 85.], 1.5, 2048), (8, [20., 20.], [215., 20.], 1.8, 512)):
        a, b = np.array(a), np.array(b)
        np.random.seed(seed)
        got = ellipsoid_point_cloud_sampling(a, b, ratio, mask, n_points=n_pts, n_raw_samples=10240)
        after = np.random.random()
        np.random.seed(seed)
        c_min = math.hypot(*(b - a)); c_max = c_min * ratio
        a1 = np.concatenate([(b - a) / c_min, [0.]])[:, np.newaxis]
        U, _, V_T = np.linalg.svd(a1 @ np.array([[1.0, 0.0, 0.0]]), True, True)
        C = U @ np.diag([1.0, 1.0, np.linalg.det(U) * np.linalg.det(V_T.T)]) @ V_T
        r = [c_max / 2.0, math.sqrt(c_max ** 2 - c_min ** 2) / 2.0, math.sqrt(c_max ** 2 - c_min ** 2) / 2.0]
        smp = np.random.uniform(-1, 1, size=(10240, 2))
        assert np.random.random() == after
        smp = smp[np.linalg.norm(smp, axis=1) <= 1]
        smp = np.concatenate([smp, np.zeros((len(smp), 1))], axis=1)
        pc = (np.dot(np.dot(C, np.diag(r)), smp.T).T + np.concatenate([(a + b) / 2., [0.]]))[:, :2]
        pc = pc[free(pc)]
        pc = pc[(pc[:, 0] >= 0) & (pc[:, 0] <= W) & (pc[:, 1] >= 0) & (pc[:, 1] <= H)]
        want = fps(pc, n_pts) if len(pc) > n_pts else pc
        assert np.array_equal(got, want), seed
