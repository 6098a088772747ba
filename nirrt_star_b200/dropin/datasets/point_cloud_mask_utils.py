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
# Random draws stay on the process-global numpy stream exactly as in the reference; the free-space
# test is the reference's 4-neighbour pixel lookup in ``binary_mask`` (host numpy indexing, the
# image is an input of the planner); the farthest-point down-sampling runs as a CUDA kernel
# (nirrt_fps_f64_sync) with open3d's semantics (parity unpinned at that boundary).
import math

from nirrt_star_b200.batch import fps_f64
from path_planning_classes.collision_check_utils import points_in_range


def get_binary_mask(env_img):
    """datasets/point_cloud_mask_utils.py:8-17"""
    binary_mask = np.zeros(env_img.shape[:2]).astype(float)
    binary_mask[env_img[:, :, 0] != 0] = 1
    return binary_mask


def _free_pixels_mask(point_cloud, binary_mask):
    """product of the mask over the 4 pixels around each point (point_cloud_mask_utils.py:52-66)"""
    img_height, img_width = binary_mask.shape
    pix = point_cloud.astype(int)
    nei = (pix + np.array([[0, 0], [0, 1], [1, 0], [1, 1]])[:, np.newaxis]).reshape(-1, 2)
    nei[:, 1] = np.clip(nei[:, 1], 0, img_height - 1)
    nei[:, 0] = np.clip(nei[:, 0], 0, img_width - 1)
    return np.prod(binary_mask[nei[:, 1], nei[:, 0]].reshape(4, -1), axis=0)


def _fps_2d(point_cloud, n_points):
    fake_z = np.concatenate([point_cloud, np.zeros((point_cloud.shape[0], 1))], axis=1)
    return fake_z[fps_f64(fake_z, n_points, start=0)][:, :2]


def generate_rectangle_point_cloud(binary_mask, n_points, over_sample_scale=5):
    """datasets/point_cloud_mask_utils.py:35-73"""
    img_height, img_width = binary_mask.shape
    point_cloud = np.random.uniform(low=[0, 0], high=[img_width, img_height], size=(n_points * over_sample_scale, 2))
    point_cloud = point_cloud[_free_pixels_mask(point_cloud, binary_mask).nonzero()[0]]
    return _fps_2d(point_cloud, n_points)


def RotationToWorldFrame(start_point, goal_point, L):
    """datasets/point_cloud_mask_utils.py:77-92"""
    a1 = (goal_point - start_point) / L
    a1 = np.concatenate([a1, np.array([0.])], axis=0)[:, np.newaxis]
    e1 = np.array([[1.0], [0.0], [0.0]])
    M = a1 @ e1.T
    U, _, V_T = np.linalg.svd(M, True, True)
    return U @ np.diag([1.0, 1.0, np.linalg.det(U) * np.linalg.det(V_T.T)]) @ V_T


def ellipsoid_point_cloud_sampling(start_point, goal_point, max_min_ratio, binary_mask, n_points=1000, n_raw_samples=10000):
    """datasets/point_cloud_mask_utils.py:104-174"""
    dx, dy = goal_point - start_point
    c_min = math.hypot(dx, dy)
    C = RotationToWorldFrame(start_point, goal_point, c_min)
    x_center = np.concatenate([(start_point + goal_point) / 2., np.array([0.])], axis=0)
    c_max = c_min * max_min_ratio
    eps = 1e-6 if c_max ** 2 - c_min ** 2 < 0 else 0
    r = [c_max / 2.0, math.sqrt(c_max ** 2 - c_min ** 2 + eps) / 2.0, math.sqrt(c_max ** 2 - c_min ** 2 + eps) / 2.0]
    L = np.diag(r)
    samples = np.random.uniform(-1, 1, size=(n_raw_samples, 2))
    samples = samples[np.linalg.norm(samples, axis=1) <= 1]
    samples = np.concatenate([samples, np.zeros((len(samples), 1))], axis=1)
    point_cloud = (np.dot(np.dot(C, L), samples.T).T + x_center)[:, :2]
    point_cloud = point_cloud[_free_pixels_mask(point_cloud, binary_mask).nonzero()[0]]
    img_height, img_width = binary_mask.shape
    point_cloud = point_cloud[points_in_range(point_cloud, (0, img_width), (0, img_height), clearance=0)]
    if len(point_cloud) > n_points:
        point_cloud = _fps_2d(point_cloud, n_points)
    return point_cloud
