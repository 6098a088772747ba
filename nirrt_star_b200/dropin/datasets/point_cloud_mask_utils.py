"""datasets/point_cloud_mask_utils.py drop-in (the functions on the planning path)."""
import numpy as np


def get_point_cloud_mask_around_points(point_cloud, points, neighbor_radius=3):
    """mask_i = any_j |pc_i - p_j| < radius (strict), datasets/point_cloud_mask_utils.py:20-31.
    Host numpy (2048 x 1 distances per cloud update; the same expression as the reference)."""
    dist = point_cloud[:, np.newaxis] - points          # (n, m, C)
    dist = np.linalg.norm(dist, axis=2)
    neighbor_mask = dist < neighbor_radius
    return np.sum(neighbor_mask, axis=1) > 0


# ---- guidance point-cloud generation for 2D worlds ----------------------------
# The reference's function names and arguments, executed on the device (nirrt_batch_sample_clouds_sync): the draws come
# from the process-global numpy MT19937 stream, which is handed to the device and handed back at the reference's stream
# position; the free-space test is the reference's 4-neighbour pixel lookup in ``binary_mask`` (uploaded once per
# image); the down-sampling is farthest point sampling with open3d's semantics (parity unpinned).  Only
# the scalars that go through numpy's SVD / libm pow (batch.ellipsoid_params_2d) are evaluated on the host.
from nirrt_star_b200.batch import BatchPlanner2D, ellipsoid_params_2d

_cache = {}


def get_binary_mask(env_img):
    """datasets/point_cloud_mask_utils.py:8-17"""
    binary_mask = np.zeros(env_img.shape[:2]).astype(float)
    binary_mask[env_img[:, :, 0] != 0] = 1
    return binary_mask


def _context(binary_mask, n_points, n_raw):
    m = np.ascontiguousarray(np.asarray(binary_mask) != 0, dtype=np.uint8)
    key = (m.shape, m.tobytes(), int(n_points), int(n_raw))
    ctx = _cache.get(key)
    if ctx is None:
        if len(_cache) > 16:
            for c in _cache.values():
                c.close()
            _cache.clear()
        h, w = m.shape
        problem = {"x_start": (0., 0.), "x_goal": (1., 0.), "search_radius": 1.0,
                   "env_dict": {"env_dims": [h, w], "circle_obstacles": [], "rectangle_obstacles": []}}
        ctx = BatchPlanner2D([problem], 1, clearance=0)
        ctx.set_free_masks([m])
        _cache[key] = ctx
    return ctx


def _sample(binary_mask, kind, params, n_points, n_raw):
    ctx = _context(binary_mask, n_points, n_raw)
    st = np.random.get_state()
    ctx.set_rng([(st[1], st[2])])
    count = int(ctx.sample_clouds([0], [kind], params, n_points, n_raw, 1.0, None, None, None)[0])
    key, pos = ctx.get_rng()[0]
    np.random.set_state(("MT19937", key, pos, 0, 0.0))
    return ctx.read_sampled_clouds(0, 1, n_points)[0, :count].copy()


def generate_rectangle_point_cloud(binary_mask, n_points, over_sample_scale=5):
    """datasets/point_cloud_mask_utils.py:35-73"""
    return _sample(binary_mask, 0, np.zeros(12), n_points, n_points * over_sample_scale)


def ellipsoid_point_cloud_sampling(start_point, goal_point, max_min_ratio, binary_mask, n_points=1000, n_raw_samples=10000):
    """datasets/point_cloud_mask_utils.py:104-174"""
    return _sample(binary_mask, 1, ellipsoid_params_2d(start_point, goal_point, max_min_ratio), n_points, n_raw_samples)
