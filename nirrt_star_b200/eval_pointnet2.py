"""Batched offline inference of the guidance network: the test loop of the reference's
eval_pointnet_pointnet2.py:108-161 -- B = 16 clouds per forward, argmax, per-class IoU / accuracy, weighted NLL --
on the sm_100a engine instead of torch modules.  Inference only (training needs backward kernels: out of scope).

Dataset format == PathPlanDataset (pointnet_pointnet2/PathPlanDataLoader.py:7-46): an .npz with ``pc`` (n, N, 2|3),
``start``, ``goal``, ``free``, ``astar`` (n, N) and ``token``; 2D clouds are z-padded.  The network's free-space channel
is 1 - (start + goal > 0) as in the planners' wrapper (pointnet2_wrapper.py:52-56); a dataset whose ``free`` mask says
otherwise is rejected instead of silently evaluated on different inputs."""
import numpy as np

from .pointnet2 import PointNet2Engine, draw_fps_starts


class PathPlanArrays:
    """PathPlanDataset without torch: arrays + the class weights of PathPlanDataLoader.py:30-34."""

    def __init__(self, source):
        data = np.load(source) if isinstance(source, (str, bytes)) or hasattr(source, "read") else source
        self.pc = np.asarray(data["pc"], dtype=np.float32)
        self.start_mask = np.asarray(data["start"], dtype=np.float32)
        self.goal_mask = np.asarray(data["goal"], dtype=np.float32)
        self.free_mask = np.asarray(data["free"], dtype=np.float32)
        self.astar_mask = np.asarray(data["astar"], dtype=np.float32)
        self.token = data["token"] if "token" in data else np.arange(len(self.pc))
        if self.pc.shape[2] == 2:
            self.pc = np.concatenate((self.pc, np.zeros((self.pc.shape[0], self.pc.shape[1], 1), np.float32)), axis=2)
        if self.pc.shape[2] != 3:
            raise RuntimeError("Point cloud is not 3D.")
        if not np.array_equal(self.free_mask, 1 - ((self.start_mask + self.goal_mask) > 0).astype(np.float32)):
            raise ValueError("free mask differs from 1 - (start + goal > 0): the engine derives that channel itself")
        lw, _ = np.histogram(self.astar_mask, range(3))
        lw = lw.astype(np.float32)
        lw = lw / np.sum(lw)
        self.labelweights = np.power(np.amax(lw) / lw, 1 / 3.0)

    def __len__(self):
        return len(self.pc)


def evaluate(state_dict, dataset, batch_size=16, num_classes=2, device=0, return_predictions=False, engine=None):
    """eval_pointnet_pointnet2.py:108-161.  Returns {'mean_loss', 'mIoU', 'accuracy', 'class_acc', 'labelweights'}
    (+ 'pred', 'scores' when asked).  Like the reference's DataLoader(drop_last=False, shuffle=False) loop, the last
    batch may be smaller; FPS start indices are drawn per batch from torch's global CPU generator exactly as the
    reference's forward draws them (pointnet2_utils.py:77)."""
    if num_classes != 2:
        raise ValueError("the sm_100a PointNet++ engine is built for num_classes=2")
    ds = dataset if isinstance(dataset, PathPlanArrays) else PathPlanArrays(dataset)
    n, N = ds.pc.shape[0], ds.pc.shape[1]
    own = engine is None
    eng = PointNet2Engine(state_dict, n_points=N, max_batch=batch_size, device=device) if own else engine
    w = ds.labelweights.astype(np.float64)
    total_correct = total_seen = 0
    loss_sum, num_batches = 0.0, 0
    labelweights = np.zeros(num_classes)
    seen = np.zeros(num_classes); correct_c = np.zeros(num_classes); iou_deno = np.zeros(num_classes)
    preds, scores_all = [], []
    for b0 in range(0, n, batch_size):
        sl = slice(b0, min(n, b0 + batch_size))
        B = sl.stop - sl.start
        fs = draw_fps_starts(B, N)
        pred, score, logp = eng.classify(ds.pc[sl], ds.start_mask[sl], ds.goal_mask[sl], fps_start=fs, return_logp=True)
        label = ds.astar_mask[sl].astype(np.int64)
        # F.nll_loss(seg_pred, target, weight=weights): weighted mean of -log p[target] (pointnet2.py get_loss)
        picked = np.take_along_axis(logp.astype(np.float64), label[..., None], axis=2)[..., 0]
        wt = w[label]
        loss_sum += float(-(wt * picked).sum() / wt.sum())
        num_batches += 1
        total_correct += int((pred == label).sum())
        total_seen += B * N
        tmp, _ = np.histogram(label, range(num_classes + 1))
        labelweights += tmp
        for c in range(num_classes):
            seen[c] += np.sum(label == c)
            correct_c[c] += np.sum((pred == c) & (label == c))
            iou_deno[c] += np.sum((pred == c) | (label == c))
        if return_predictions:
            preds.append(pred.copy())
            scores_all.append(np.stack([1 - score, score], axis=-1))
    if own:
        eng.close()
    out = {"mean_loss": loss_sum / max(1, num_batches),
           "mIoU": float(np.mean(correct_c / (iou_deno.astype(np.float32) + 1e-6))),
           "accuracy": total_correct / float(max(1, total_seen)),
           "class_acc": float(np.mean(correct_c / (seen.astype(np.float32) + 1e-6))),
           "labelweights": (labelweights.astype(np.float32) / max(1.0, float(labelweights.sum())))}
    if return_predictions:
        out["pred"] = np.concatenate(preds) if preds else np.zeros((0, N), np.int64)
        out["scores"] = np.concatenate(scores_all) if scores_all else np.zeros((0, N, 2), np.float32)
    return out
