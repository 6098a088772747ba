"""collision_check_utils with the reference's function names and argument conventions
(path_planning_classes/collision_check_utils.py), evaluated by the CUDA predicates
(nirrt_collide_edges_sync / nirrt_points_check_sync on a 2D batch).  Points may be a tuple (2,) ->
bool or an (n,2) array -> bool[n] (collision_check_utils.py:233-258)."""
import numpy as np

from nirrt_star_b200.batch import BatchPlanner2D

_BIG = 1.0e300
_cache = {}


def _context(circles, rectangles, clearance, ranges=None):
    """A one-problem 2D batch holding the obstacle table; cached per obstacle set."""
    c = np.zeros((0, 3)) if circles is None else np.asarray(circles, dtype=np.float64).reshape(-1, 3)
    r = np.zeros((0, 4)) if rectangles is None else np.asarray(rectangles, dtype=np.float64).reshape(-1, 4)
    key = (c.tobytes(), r.tobytes(), float(clearance), None if ranges is None else tuple(map(float, ranges)))
    ctx = _cache.get(key)
    if ctx is None:
        if len(_cache) > 64:
            _cache.clear()
        dims = [2 * _BIG, 2 * _BIG] if ranges is None else [ranges[3] - ranges[2], ranges[1] - ranges[0]]
        problem = {"x_start": (0., 0.), "x_goal": (1., 0.), "search_radius": 1.0,
                   "env_dict": {"env_dims": dims, "circle_obstacles": c.tolist(), "rectangle_obstacles": r.tolist()}}
        ctx = BatchPlanner2D([problem], 1, clearance=clearance)
        _cache[key] = ctx
    return ctx


def check_collision_line_circles_rectangles(line, circles, rectangles, clearance=0):
    line = np.asarray(line, dtype=np.float64).reshape(1, 2, 2)
    return bool(_context(circles, rectangles, clearance).collide_edges(0, line)[0])


def _as_points(points):
    if type(points) == tuple:
        return np.array(points, dtype=np.float64)[np.newaxis, :], True
    return np.asarray(points, dtype=np.float64).reshape(-1, 2), False


def points_in_circles_rectangles(points, circles, rectangles, clearance=0):
    pts, single = _as_points(points)
    out = _context(circles, rectangles, clearance).points_inside_obs(0, pts)
    return bool(out[0]) if single else out


def points_in_circles(points, circles, clearance=0):
    return points_in_circles_rectangles(points, circles, None, clearance)


def points_in_rectangles(points, rectangles, clearance=0):
    return points_in_circles_rectangles(points, None, rectangles, clearance)


def _zero_based(x_range, y_range):
    if x_range[0] != 0 or y_range[0] != 0:
        raise ValueError("ranges must start at 0 (Env.x_range = (0, width), rrt_env.py:7-8)")


def points_in_range(points, x_range, y_range, clearance=0):
    pts, single = _as_points(points)
    _zero_based(x_range, y_range)
    out = _context(None, None, clearance, (x_range[0], x_range[1], y_range[0], y_range[1])).points_valid(0, pts)
    return bool(out[0]) if single else out


def points_validity(points, circle_obstacles, rectangle_obstacles, x_range, y_range, obstacle_clearance=0, range_clearance=0):
    if obstacle_clearance != range_clearance:
        in_range = points_in_range(points, x_range, y_range, range_clearance)
        in_obs = points_in_circles_rectangles(points, circle_obstacles, rectangle_obstacles, obstacle_clearance)
        if type(points) == tuple:
            return bool(in_range and not in_obs)
        return in_range & ~in_obs
    pts, single = _as_points(points)
    _zero_based(x_range, y_range)
    ctx = _context(circle_obstacles, rectangle_obstacles, obstacle_clearance, (x_range[0], x_range[1], y_range[0], y_range[1]))
    out = ctx.points_valid(0, pts)
    return bool(out[0]) if single else out
