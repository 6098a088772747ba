"""collision_check_utils_3d with the reference's function names and argument conventions
(path_planning_classes_3d/collision_check_utils_3d.py), evaluated by the CUDA predicates
(nirrt_collide_edges_sync / nirrt_points_check_sync).  Points may be a tuple (3,) -> bool or an
(n,3) array -> bool[n] (collision_check_utils_3d.py:233-257)."""
import numpy as np

from nirrt_star_b200.batch import BatchPlanner3D

_BIG = 1.0e300
_cache = {}


def _context(balls, boxes, clearance, ranges=None):
    """A one-problem batch holding the obstacle table; cached per obstacle set."""
    b = np.zeros((0, 4)) if balls is None else np.asarray(balls, dtype=np.float64).reshape(-1, 4)
    x = np.zeros((0, 6)) if boxes is None else np.asarray(boxes, dtype=np.float64).reshape(-1, 6)
    key = (b.tobytes(), x.tobytes(), float(clearance), None if ranges is None else tuple(map(float, ranges)))
    ctx = _cache.get(key)
    if ctx is None:
        if len(_cache) > 64:
            _cache.clear()
        if ranges is None:
            dims = [2 * _BIG, 2 * _BIG, 2 * _BIG]
        else:
            dims = [ranges[3] - ranges[2], ranges[1] - ranges[0], ranges[5] - ranges[4]]
        problem = {"x_start": (0., 0., 0.), "x_goal": (1., 0., 0.), "search_radius": 1.0,
                   "env_dict": {"env_dims": dims, "ball_obstacles": b.tolist(), "box_obstacles": x.tolist()}}
        ctx = BatchPlanner3D([problem], 1, clearance=clearance)
        _cache[key] = ctx
    return ctx


def check_collision_line_balls_boxes(line, balls, boxes, clearance=0.):
    line = np.asarray(line, dtype=np.float64).reshape(1, 2, 3)
    return bool(_context(balls, boxes, clearance).collide_edges(0, line)[0])


def _as_points(points):
    if type(points) == tuple:
        return np.array(points, dtype=np.float64)[np.newaxis, :], True
    return np.asarray(points, dtype=np.float64).reshape(-1, 3), False


def points_in_balls_boxes(points, balls, boxes, clearance=0.):
    pts, single = _as_points(points)
    out = _context(balls, boxes, clearance).points_inside_obs(0, pts)
    return bool(out[0]) if single else out


def points_in_balls(points, balls, clearance=0.):
    return points_in_balls_boxes(points, balls, None, clearance)


def points_in_boxes(points, boxes, clearance=0.):
    return points_in_balls_boxes(points, None, boxes, clearance)


def points_in_range_3d(points, x_range, y_range, z_range, clearance=0.):
    pts, single = _as_points(points)
    ctx = _context(None, None, clearance, (x_range[0], x_range[1], y_range[0], y_range[1], z_range[0], z_range[1]))
    if x_range[0] != 0 or y_range[0] != 0 or z_range[0] != 0:
        raise ValueError("ranges must start at 0 (Env.x_range = (0, width), rrt_env_3d.py:6-9)")
    out = ctx.points_valid(0, pts)
    return bool(out[0]) if single else out


def points_validity_3d(points, ball_obstacles, box_obstacles, x_range, y_range, z_range,
                       obstacle_clearance=0., range_clearance=0.):
    if obstacle_clearance != range_clearance:
        in_range = points_in_range_3d(points, x_range, y_range, z_range, range_clearance)
        in_obs = points_in_balls_boxes(points, ball_obstacles, box_obstacles, obstacle_clearance)
        if type(points) == tuple:
            return bool(in_range and not in_obs)
        return in_range & ~in_obs
    pts, single = _as_points(points)
    if x_range[0] != 0 or y_range[0] != 0 or z_range[0] != 0:
        raise ValueError("ranges must start at 0 (Env.x_range = (0, width), rrt_env_3d.py:6-9)")
    ctx = _context(ball_obstacles, box_obstacles, obstacle_clearance,
                   (x_range[0], x_range[1], y_range[0], y_range[1], z_range[0], z_range[1]))
    out = ctx.points_valid(0, pts)
    return bool(out[0]) if single else out
