"""Visualisation is out of scope; the planners only need objects that
can be constructed cheaply (rrt_star_3d.py:30) and that refuse to draw."""


class _NoVisualizer:
    def __init__(self, x_start, x_goal, env, path_point_cloud_pred=None):
        self.x_start, self.x_goal, self.env = x_start, x_goal, env
        self.path_point_cloud_pred = path_point_cloud_pred

    def set_path_point_cloud_pred(self, pc):
        self.path_point_cloud_pred = pc

    def animation(self, *args, **kwargs):
        raise NotImplementedError("nirrt_star_b200 does not ship the matplotlib visualisers; "
                                  "pass planner.vertices / vertex_parents / path to the reference's rrt_visualizer_3d")


RRTStarVisualizer3D = IRRTStarVisualizer3D = NRRTStarPNGVisualizer3D = NIRRTStarVisualizer3D = _NoVisualizer
