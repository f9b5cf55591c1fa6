"""3-D keypoint error measures with the reference's interface
(/root/reference/mano_train/evaluation/zimeval.py:22-129, ``EvalUtil.feed`` / ``get_measures``).

The reference appends one Python float per (sample, joint) to 21 lists after a per-step ``.cpu()`` of the predicted
joints (epochpass3d.py:138-150).  Here the per-joint distances of a whole step arrive as one array (computed on the
device, copied once per epoch by ``epoch_pass``) and the measures are evaluated with array operations; the numbers
are the same: per-joint mean / median end-point error, PCK curve over ``steps`` thresholds in [val_min, val_max],
AUC = trapezoid(pck) / trapezoid(1), all averaged over the joints that have at least one visible measurement.
"""
import warnings

import numpy as np


def _trapezoid(y, x):
    return float(np.sum((y[1:] + y[:-1]) * np.diff(x)) / 2.0)


class EvalUtil(object):
    def __init__(self, num_kp=21):
        self.num_kp = num_kp
        self._dists = []   # list of (n, num_kp) float arrays
        self._vis = []     # list of (n, num_kp) bool arrays

    def feed(self, keypoint_gt, keypoint_pred, keypoint_vis=None):
        """One sample: (num_kp, D) ground truth and prediction (tensors or arrays), optional visibility (num_kp,)."""
        gt = np.squeeze(np.asarray(keypoint_gt, dtype=np.float64))
        pred = np.squeeze(np.asarray(keypoint_pred, dtype=np.float64))
        assert gt.ndim == 2 and pred.ndim == 2
        dist = np.sqrt(((gt - pred) ** 2).sum(1))[None]
        vis = np.ones_like(dist, dtype=bool) if keypoint_vis is None else np.squeeze(np.asarray(keypoint_vis)).astype(bool)[None]
        self.feed_distances(dist, vis)

    def feed_distances(self, dists, vis=None):
        """A batch of Euclidean distances (n, num_kp) [+ visibility (n, num_kp)] - the epoch_pass fast path."""
        dists = np.asarray(dists, dtype=np.float64).reshape(-1, self.num_kp)
        self._dists.append(dists)
        self._vis.append(np.ones_like(dists, dtype=bool) if vis is None else np.asarray(vis, dtype=bool).reshape(dists.shape))

    @property
    def data(self):
        """The reference's storage: one list of distances per keypoint."""
        if not self._dists:
            return [[] for _ in range(self.num_kp)]
        d, v = np.concatenate(self._dists), np.concatenate(self._vis)
        return [d[v[:, k], k].tolist() for k in range(self.num_kp)]

    def get_measures(self, val_min, val_max, steps):
        thresholds = np.linspace(val_min, val_max, steps)
        norm_factor = _trapezoid(np.ones_like(thresholds), thresholds)
        epe_mean, epe_median, aucs, curves = [], [], [], []
        if self._dists:
            d, v = np.concatenate(self._dists), np.concatenate(self._vis)
            for k in range(self.num_kp):
                col = d[v[:, k], k]
                if col.size == 0:
                    continue  # no valid measurement for this keypoint
                epe_mean.append(col.mean())
                epe_median.append(np.median(col))
                curve = (col[None, :] <= thresholds[:, None]).mean(1)
                curves.append(curve)
                aucs.append(_trapezoid(curve, thresholds) / norm_factor)
        epe_mean_joint = epe_mean
        with np.errstate(invalid="ignore"), warnings.catch_warnings():
            warnings.simplefilter("ignore")  # empty evaluator -> nan, like the reference
            epe_mean_all = np.mean(np.array(epe_mean))
            epe_median_all = np.mean(np.array(epe_median))
            auc_all = np.mean(np.array(aucs))
            pck_curve_all = np.mean(np.array(curves), 0)
        return epe_mean_all, epe_mean_joint, epe_median_all, auc_all, pck_curve_all, thresholds
