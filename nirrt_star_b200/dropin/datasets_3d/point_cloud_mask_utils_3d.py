"""datasets_3d/point_cloud_mask_utils_3d.py drop-in: guidance point-cloud generation (SURVEY.md row
f1).  Random draws stay on the process-global numpy stream exactly as in the reference; the
obstacle / range filters run as CUDA predicates (nirrt_points_check_sync) and the farthest-point
down-sampling as a CUDA kernel (nirrt_fps_f64_sync) with open3d's semantics (start index 0, f64
squared distances, first argmax; third-party arithmetic, parity unpinned -- SURVEY.md 8c)."""
import numpy as np

from nirrt_star_b200.batch import fps_f64
from path_planning_classes_3d.collision_check_utils_3d import points_in_balls_boxes, points_validity_3d


def farthest_point_sample_open3d(points, npoint):
    """point_cloud_mask_utils_3d.py:41-54"""
    points = np.asarray(points, dtype=np.float64)
    return points[fps_f64(points, npoint, start=0)]


def generate_rectangle_point_cloud_3d(env, n_points, over_sample_scale=5, use_open3d=True, clearance=0):
    """point_cloud_mask_utils_3d.py:83-113"""
    point_cloud = np.random.uniform(
        low=(env.x_range[0] + clearance, env.y_range[0] + clearance, env.z_range[0] + clearance),
        high=(env.x_range[1] - clearance, env.y_range[1] - clearance, env.z_range[1] - clearance),
        size=(n_points * over_sample_scale, 3),
    )
    in_obs = points_in_balls_boxes(
        point_cloud,
        np.array(env.obs_ball).astype(np.float64),
        np.array(env.obs_box).astype(np.float64),
        clearance=clearance,
    )
    point_cloud = point_cloud[(1 - in_obs).astype(bool)]
    if len(point_cloud) > n_points:
        point_cloud = farthest_point_sample_open3d(point_cloud, n_points)
    return point_cloud


def RotationToWorldFrame(x_start, x_goal, L):
    """point_cloud_mask_utils_3d.py:117-129"""
    a1 = (x_goal - x_start) / L
    M = np.outer(a1, [1, 0, 0])
    U, S, V = np.linalg.svd(M)
    C = U @ np.diag([1, 1, np.linalg.det(U) * np.linalg.det(V)]) @ V.T
    return C


def ellipsoid_point_cloud_sampling_3d(start_point, goal_point, max_min_ratio, env, n_points=1000, n_raw_samples=10000,
                                      clearance=0):
    """point_cloud_mask_utils_3d.py:132-200"""
    c_min = np.linalg.norm(goal_point - start_point)
    C = RotationToWorldFrame(start_point, goal_point, c_min)
    x_center = (start_point + goal_point) / 2.
    c_max = c_min * max_min_ratio
    if c_max ** 2 - c_min ** 2 < 0:
        eps = 1e-6
    else:
        eps = 0
    r = np.zeros(3)
    r[0] = c_max / 2
    for i in [1, 2]:
        r[i] = np.sqrt(c_max ** 2 - c_min ** 2 + eps) / 2
    L = np.diag(r)

    radius = np.random.uniform(0.0, 1.0, n_raw_samples)
    theta = np.random.uniform(0, np.pi, n_raw_samples)
    phi = np.random.uniform(0, 2 * np.pi, n_raw_samples)
    samples_x = radius * np.sin(theta) * np.cos(phi)
    samples_y = radius * np.sin(theta) * np.sin(phi)
    samples_z = radius * np.cos(theta)
    samples = np.array([samples_x, samples_y, samples_z]).T
    point_cloud = np.dot(np.dot(C, L), samples.T).T + x_center

    obs_ball = np.array(env.obs_ball).astype(np.float64) if len(env.obs_ball) > 0 else None
    obs_box = np.array(env.obs_box).astype(np.float64) if len(env.obs_box) > 0 else None
    valid_flag = points_validity_3d(point_cloud, obs_ball, obs_box, env.x_range, env.y_range, env.z_range,
                                    obstacle_clearance=clearance, range_clearance=clearance)
    point_cloud = point_cloud[valid_flag]
    if len(point_cloud) > n_points:
        point_cloud = farthest_point_sample_open3d(point_cloud, n_points)
    return point_cloud
