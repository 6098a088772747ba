"""Deterministic stand-in for the network used by the Neural Connect fixtures (generator and test
share it): "path" = points within `reach` of the centroid of the start-mask points or of the
goal-mask points.  As generate_connected_path_points re-centres the masks on boundary points, the
two blobs hop towards each other, so several trials are exercised."""
import hashlib

import numpy as np


class FakeClassifier:
    def __init__(self, reach):
        self.reach = reach
        self.calls = []

    def classify_path_points(self, pc, start_mask, goal_mask):
        p = pc.astype(np.float64)
        pred = np.zeros(len(pc), dtype=np.int64)
        for m in (start_mask, goal_mask):
            sel = np.asarray(m) > 0
            if sel.any():
                c = p[sel].mean(axis=0)
                pred |= (np.linalg.norm(p - c, axis=1) < self.reach).astype(np.int64)
        self.calls.append(hashlib.sha1(np.asarray(start_mask, dtype=np.float32).tobytes() +
                                       np.asarray(goal_mask, dtype=np.float32).tobytes()).hexdigest())
        return pred, pred.astype(np.float32)
