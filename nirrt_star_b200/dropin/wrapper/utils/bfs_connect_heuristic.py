"""wrapper/utils/bfs_connect_heuristic.py drop-in (Neural Connect helpers).  The
O(n^2) graph work -- r-disc adjacency over the predicted path points, the component search and the
boundary test -- runs as one CUDA launch (nirrt_connect_analyse_sync); the rank heuristic over the
(few) boundary points stays numpy."""
import numpy as np

from nirrt_star_b200.pointnet2 import connect_analyse


def bfs_point_cloud(pc, path_mask, x_start, x_goal, step_len):
    """(has_path, visited_mask) -- bfs_connect_heuristic.py:32-78.  When a path exists the reference
    returns the vertices visited *so far*, which depends on its queue order and which its only caller
    ignores; here visited_mask is then the whole component of x_start."""
    has_path, visited, _ = connect_analyse(pc, path_mask, x_start, x_goal, step_len)
    return has_path, visited


def bfs_point_cloud_visualization(pc, path_mask, x_start, x_goal, step_len):
    """(has_path, path_line, visited_mask) -- :80-139; path_line is only drawn by the (out of scope)
    visualisers and is returned as None."""
    has_path, visited = bfs_point_cloud(pc, path_mask, x_start, x_goal, step_len)
    return has_path, None, visited


def get_boundary_mask(pc, path_mask, unvisited_mask, boundary_distance_threshold):
    """bfs_connect_heuristic.py:5-29 (general form, host numpy); generate_connected_path_points uses
    the fused CUDA analysis instead."""
    path_points = pc[path_mask.astype(bool)]
    unvisited_points = pc[unvisited_mask.astype(bool)]
    d = np.linalg.norm(path_points[:, np.newaxis] - unvisited_points, axis=2)
    on_path = (d < boundary_distance_threshold).astype(float).sum(axis=1) > 0
    out = np.zeros(len(pc))
    out[np.where(path_mask.astype(bool))[0][on_path]] = 1
    return out.astype(np.float32)


def select_heuristic_boundary_point(pc, boundary_mask, x_start, x_goal, cost_from_start_rank_weight=1):
    """Boundary point with the best (lowest) rank sum of total heuristic cost (ascending) and cost
    from start (descending) -- bfs_connect_heuristic.py:142-181."""
    boundary_points = pc[boundary_mask.astype(bool)]
    if len(boundary_points) == 0:
        return None, None, None
    from_start = np.linalg.norm(boundary_points - x_start, axis=1)
    to_goal = np.linalg.norm(boundary_points - x_goal, axis=1)
    total_rank = np.empty(len(boundary_points), dtype=np.int64)
    total_rank[np.argsort(from_start + to_goal)] = np.arange(len(boundary_points))
    start_rank = np.empty(len(boundary_points), dtype=np.int64)
    start_rank[np.flip(np.argsort(from_start))] = np.arange(len(boundary_points))
    heuristic = [int(-(total_rank[i] + cost_from_start_rank_weight * start_rank[i])) for i in range(len(boundary_points))]
    index = np.where(boundary_mask)[0][np.argmax(heuristic)]
    return index, pc[index], heuristic
