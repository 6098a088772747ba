"""datasets_3d/point_cloud_mask_utils_3d.py drop-in: guidance point-cloud generation with the
reference's function names and arguments, executed on the device (nirrt_batch_sample_clouds_sync): the uniform /
(r, theta, phi) draws come from the process-global numpy MT19937 stream, which is handed to the device and handed back
(the stream position afterwards is exactly the reference's), the obstacle / range filters are the CUDA predicates and
the down-sampling is farthest point sampling with open3d's semantics (start index 0, f64 squared distances, first
argmax; third-party arithmetic, parity unpinned -- SURVEY.md 8c).  Only the handful of scalars that go through
numpy's SVD / libm pow (batch.ellipsoid_params_3d) are evaluated on the host."""
import numpy as np

from nirrt_star_b200.batch import BatchPlanner3D, ellipsoid_params_3d, fps_f64

_cache = {}


def _context(env, n_points, n_raw):
    """A one-problem batch holding the world (obstacle table, ranges) and the sampler workspace; cached per world."""
    b = np.asarray(env.obs_ball, dtype=np.float64).reshape(-1, 4)
    x = np.asarray(env.obs_box, dtype=np.float64).reshape(-1, 6)
    key = (b.tobytes(), x.tobytes(), tuple(map(float, env.x_range + env.y_range + env.z_range)), int(n_points), int(n_raw))
    ctx = _cache.get(key)
    if ctx is None:
        if env.x_range[0] != 0 or env.y_range[0] != 0 or env.z_range[0] != 0:
            raise ValueError("ranges must start at 0 (Env.x_range = (0, width), rrt_env_3d.py:6-9)")
        if len(_cache) > 16:
            for c in _cache.values():
                c.close()
            _cache.clear()
        problem = {"x_start": (0., 0., 0.), "x_goal": (1., 0., 0.), "search_radius": 1.0,
                   "env_dict": {"env_dims": [env.y_range[1], env.x_range[1], env.z_range[1]],
                                "ball_obstacles": b.tolist(), "box_obstacles": x.tolist()}}
        ctx = BatchPlanner3D([problem], 1, clearance=0)
        _cache[key] = ctx
    return ctx


def _sample(env, kind, params, n_points, n_raw):
    ctx = _context(env, n_points, n_raw)
    st = np.random.get_state()
    ctx.set_rng([(st[1], st[2])])
    count = int(ctx.sample_clouds([0], [kind], params, n_points, n_raw, 1.0, None, None, None)[0])
    key, pos = ctx.get_rng()[0]
    np.random.set_state(("MT19937", key, pos, 0, 0.0))
    return ctx.read_sampled_clouds(0, 1, n_points)[0, :count].copy()


def farthest_point_sample_open3d(points, npoint):
    """point_cloud_mask_utils_3d.py:41-54"""
    points = np.asarray(points, dtype=np.float64)
    return points[fps_f64(points, npoint, start=0)]


def generate_rectangle_point_cloud_3d(env, n_points, over_sample_scale=5, use_open3d=True, clearance=0):
    """point_cloud_mask_utils_3d.py:83-113"""
    if clearance != 0:
        raise NotImplementedError("the device sampler implements the planners' call (clearance = 0)")
    return _sample(env, 0, np.zeros(12), n_points, n_points * over_sample_scale)


def ellipsoid_point_cloud_sampling_3d(start_point, goal_point, max_min_ratio, env, n_points=1000, n_raw_samples=10000,
                                      clearance=0):
    """point_cloud_mask_utils_3d.py:132-200"""
    if clearance != 0:
        raise NotImplementedError("the device sampler implements the planners' call (clearance = 0)")
    return _sample(env, 1, ellipsoid_params_3d(start_point, goal_point, max_min_ratio), n_points, n_raw_samples)
