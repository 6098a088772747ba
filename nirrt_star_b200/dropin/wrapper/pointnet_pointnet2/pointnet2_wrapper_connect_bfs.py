"""PNGWrapper with Neural Connect for 2D (reference:
wrapper/pointnet_pointnet2/pointnet2_wrapper_connect_bfs.py:76-240)."""
from wrapper_3d.pointnet_pointnet2.pointnet2_wrapper_connect_bfs import PNGWrapper as _PNGWrapperConnect3D


class PNGWrapper(_PNGWrapperConnect3D):
    _dim_tag = "2d"
    _banner = "PointNet++ wrapper with connect is initialized."
