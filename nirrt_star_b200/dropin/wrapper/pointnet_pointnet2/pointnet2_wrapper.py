"""PNGWrapper drop-in for 2D (reference: wrapper/pointnet_pointnet2/pointnet2_wrapper.py): (n,2)
clouds are z-padded with 0 on the device (:46-50); 2D checkpoint path."""
from wrapper_3d.pointnet_pointnet2.pointnet2_wrapper import PNGWrapper as _PNGWrapper3D


class PNGWrapper(_PNGWrapper3D):
    _dim_tag = "2d"
    _banner = "PointNet++ wrapper is initialized."
