"""PNGWrapper drop-in (reference: wrapper_3d/pointnet_pointnet2/pointnet2_wrapper.py).  Same
constructor, checkpoint path and ``classify_path_points`` contract; the forward runs on the sm_100a
engine (nirrt_pn2_classify_sync) instead of torch modules."""
from os.path import join

import numpy as np
import torch

from nirrt_star_b200.pointnet2 import PointNet2Engine


class PNGWrapper:
    _dim_tag = "3d"
    _banner = "PointNet++ wrapper 3d is initialized."

    def __init__(self, num_classes=2, root_dir='.', device='cuda'):
        if num_classes != 2:
            raise ValueError("the sm_100a PointNet++ engine is built for num_classes=2 (path / not path)")
        if not str(device).startswith('cuda'):
            raise RuntimeError("nirrt_star_b200 has no CPU path: PNGWrapper needs device='cuda' on a B200")
        model_filepath = join(root_dir, f'results/model_training/pointnet2_{self._dim_tag}/checkpoints/'
                                        f'best_pointnet2_{self._dim_tag}.pth')
        checkpoint = torch.load(model_filepath, map_location=torch.device('cpu'))
        self.state_dict = checkpoint['model_state_dict']
        self.device = device
        dev = torch.device(device)
        self._device_index = dev.index if dev.index is not None else 0
        self._engines = {}
        print(self._banner)

    def _engine(self, n_points):
        """One engine sized for the largest cloud seen so far, re-targeted to each call's size (the samplers hand over
        clouds of every size up to pc_n_points)."""
        eng = self._engines.get("e")
        if eng is None or eng.capacity < n_points:
            if eng is not None:
                eng.close()
            eng = PointNet2Engine(self.state_dict, n_points=max(n_points, 2048), max_batch=1, device=self._device_index)
            self._engines["e"] = eng
        eng.set_n_points(n_points)
        return eng

    def classify_path_points(self, pc, start_mask, goal_mask):
        """
        - inputs:
            - pc: np float32 (n_points, 2) for XY or (n_points, 3) for XYZ
            - start_mask: np float32 (n_points,) 1-0 mask
            - goal_mask: np float32 (n_points,) 1-0 mask
        - outputs:
            - path_pred: np int64 (n_points, ), 1 is path point, 0 is not.
            - path_score: np float32 (n_points, ), probability of being a path point.
        """
        pc = np.asarray(pc)
        pred, score = self._engine(pc.shape[0]).classify(pc, start_mask, goal_mask)
        return pred[0], score[0]
