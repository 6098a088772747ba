"""Visualisation is out of scope; the planners only need objects that
can be constructed cheaply (rrt_star_2d.py:30) and that refuse to draw."""


class _NoVisualizer:
    def __init__(self, x_start, x_goal, env, path_point_cloud_pred=None, img_path_score=None):
        self.x_start, self.x_goal, self.env = x_start, x_goal, env
        self.path_point_cloud_pred = path_point_cloud_pred
        self.path_point_cloud_other = None

    def set_path_point_cloud_pred(self, pc):
        self.path_point_cloud_pred = pc

    def set_path_point_cloud_other(self, pc):
        self.path_point_cloud_other = pc

    def animation(self, *args, **kwargs):
        raise NotImplementedError("nirrt_star_b200 does not ship the matplotlib visualisers; "
                                  "pass planner.vertices / vertex_parents / path to the reference's rrt_visualizer_2d")


RRTStarVisualizer = IRRTStarVisualizer = NRRTStarPNGVisualizer = NIRRTStarVisualizer = _NoVisualizer
