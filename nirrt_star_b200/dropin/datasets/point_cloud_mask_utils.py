"""datasets/point_cloud_mask_utils.py drop-in (the functions on the planning path)."""
import numpy as np


def get_point_cloud_mask_around_points(point_cloud, points, neighbor_radius=3):
    """mask_i = any_j |pc_i - p_j| < radius (strict), datasets/point_cloud_mask_utils.py:20-31.
    Host numpy (2048 x 1 distances per cloud update; the same expression as the reference)."""
    dist = point_cloud[:, np.newaxis] - points          # (n, m, C)
    dist = np.linalg.norm(dist, axis=2)
    neighbor_mask = dist < neighbor_radius
    return np.sum(neighbor_mask, axis=1) > 0
