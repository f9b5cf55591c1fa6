"""One pass over a data loader - drop-in for /root/reference/mano_train/netscripts/epochpass3d.py:17-214
(``epoch_pass(loader, model, epoch, optimizer=..., train=...) -> (avg_meters, pck_info)``), re-scheduled for one
process per GPU.

What changed relative to the reference, and why:
* ``optimizer`` is a ``FlatAdamTrainer``: zero-grad, forward, backward, the single NCCL all-reduce of the flat gradient
  buffer and the fused Adam are one call (``trainer.step`` or, with ``use_graph``, one CUDA-graph replay);
* no ``.item()`` per loss per step (the reference synchronises ~12 times per step, epochpass3d.py:111-121): the loss
  scalars of a step are stacked into one device vector and accumulated on the device; the host reads ONE vector per
  ``log_every`` steps (progress line) and once at the end;
* joint errors for PCK / AUC are computed on the device per step (one (B,21) distance tensor) and copied once at the
  end of the pass instead of a per-step ``.cpu()`` of the predictions plus a Python loop over samples (:138-150);
* with ``world_size > 1`` loss sums / counts are all-reduced and the distance tables all-gathered at the end, so every
  rank returns the global averages (rank 0 is the one that logs and checkpoints);
* visualisation (``displaymano``), result pickles and the MANO face tables read from ``misc/mano`` are out of scope.
"""
import time

import torch
import torch.distributed as dist

from ..evaluation.evalutils import AverageMeters
from ..evaluation.zimeval import EvalUtil
from ..queries import TransQueries


class _DeviceLossLog(object):
    """Per-key running sums kept on the device; keys are fixed by the first step that reports them."""

    def __init__(self):
        self.keys = None
        self.sums = None
        self.steps = 0

    def add(self, losses):
        items = [(k, v) for k, v in losses.items() if v is not None]
        if self.keys is None:
            self.keys = [k for k, _ in items]
        elif [k for k, _ in items] != self.keys:
            raise RuntimeError("epoch_pass: the set of reported losses changed inside an epoch: {} vs {}".format(
                [k for k, _ in items], self.keys))
        vec = torch.stack([torch.as_tensor(v).detach().reshape(-1)[0].float() for _, v in items])
        self.sums = vec.clone() if self.sums is None else self.sums.add_(vec)
        self.last = vec
        self.steps += 1

    def read(self, world_size=1):
        """-> ({key: mean over steps (and ranks)}, {key: last value on this rank})."""
        if self.keys is None:
            return {}, {}
        sums = torch.cat([self.sums, self.sums.new_tensor([float(self.steps)])])
        if world_size > 1:
            dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        host = sums.cpu().tolist()
        last = self.last.cpu().tolist()
        return ({k: s / host[-1] for k, s in zip(self.keys, host)}, dict(zip(self.keys, last)))


def epoch_pass(loader, model, epoch, optimizer=None, debug=False, freeze_batchnorm=True, display=False,
               display_freq=10, save_path=None, idxs=None, train=True, inspect_weights=False, fig=None,
               save_results=False, world_size=1, rank=0, log_every=50, use_graph=False, out=print):
    """Returns ``(avg_meters, pck_info)`` like the reference.  ``optimizer``: a ``FlatAdamTrainer`` when ``train``."""
    if display or save_results or inspect_weights:
        raise NotImplementedError("epoch_pass: visualisation / result dumps / weight inspection are outside the "
                                  "B200 hot path (SURVEY.md §2 rows 14-16)")
    if train and optimizer is None:
        raise ValueError("epoch_pass(train=True) needs the FlatAdamTrainer as `optimizer`")
    if rank == 0:
        out("epoch: {}".format(epoch))
    idxs = list(range(21)) if idxs is None else list(idxs)
    if train and not freeze_batchnorm:
        model.train()   # BatchNorm on batch statistics (csrc/bn_train.cu)
    else:
        model.eval()    # the README recipe (--freeze_batchnorm): eval-mode BN with trainable gamma / beta
    log = _DeviceLossLog()
    time_meters = AverageMeters()
    dists, vis_rows = [], []
    end = time.time()
    n_steps = len(loader) if hasattr(loader, "__len__") else None
    for batch_idx, sample in enumerate(loader):
        time_meters.add_loss_value("data_time", time.time() - end)
        if train:
            if use_graph and getattr(optimizer, "_graph", None) is None:
                optimizer.capture(dict(sample))       # the first batch's tensors become the static input buffers
            if use_graph and optimizer.matches_captured(sample):
                optimizer.replay(sample)              # copies this batch (and its left/right mask) into them
                _, results, losses = optimizer.static_outputs()
                if "joints" in results:               # static output: the next replay overwrites it
                    results = dict(results, joints=results["joints"].clone())
            else:
                # eager launches: no graph, or a batch the captured step was not built for (a partial last batch, a
                # key that appears / disappears)
                _, results, losses = optimizer.step(sample, return_all=True)
        else:
            with torch.no_grad():
                _, results, losses = model.forward(sample)
        log.add(losses)
        if "joints" in results and TransQueries.joints3d in sample:
            gt = sample[TransQueries.joints3d]
            pred = results["joints"].detach()
            gt = gt.to(pred.device, non_blocking=True)
            dists.append(torch.norm(pred[:, idxs] - gt[:, idxs], dim=2))
            if "vis" in sample:
                vis_rows.append(torch.as_tensor(sample["vis"]).to(pred.device)[:, idxs] != 0)
        time_meters.add_loss_value("batch_time", time.time() - end)
        if rank == 0 and log_every and (batch_idx + 1) % log_every == 0:
            means, _ = _DeviceLossLog.read(log)  # local averages only: no collective inside the loop
            out("({}/{}) Data: {:.6f}s | Batch: {:.3f}s | Loss: {:.4f}".format(
                batch_idx + 1, n_steps if n_steps is not None else "?",
                time_meters.average_meters["data_time"].val, time_meters.average_meters["batch_time"].avg,
                means.get("total_loss", float("nan"))))
        end = time.time()
    means, _ = log.read(world_size)
    avg_meters = AverageMeters()
    for key, val in means.items():
        avg_meters.add_loss_value(key, val, n=max(1, log.steps))
    pck_info = {}
    if dists:
        d = torch.cat(dists)
        v = torch.cat(vis_rows) if vis_rows else None
        if world_size > 1:
            parts = [torch.zeros_like(d) for _ in range(world_size)]
            dist.all_gather(parts, d)
            d = torch.cat(parts)
            if v is not None:
                vparts = [torch.zeros_like(v) for _ in range(world_size)]
                dist.all_gather(vparts, v)
                v = torch.cat(vparts)
        evaluator = EvalUtil(num_kp=len(idxs))
        evaluator.feed_distances(d.cpu().numpy(), None if v is None else v.cpu().numpy())
        epe_mean_all, _, epe_median_all, auc_all, pck_curve_all, thresholds = evaluator.get_measures(0, 50, 20)
        pck_info = {"auc": auc_all, "thres": thresholds, "pck_curve": pck_curve_all, "epe_mean": epe_mean_all,
                    "epe_median": epe_median_all, "evaluator": evaluator}
    return avg_meters, pck_info
