"""One-process-per-GPU training driver - the B200 counterpart of /root/reference/traineval.py:25-415
(SURVEY.md §8f rank 2).  Launch with torchrun (or plain ``python -m`` for one GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29500 \\
        -m obman_train_b200.netscripts.train --exp_id checkpoints/run0 --epochs 3 --batch_size 64

What it keeps from the reference: seeding (:28-33), the README model recipe as defaults, Adam lr 1e-4 (:113-116),
``--resume`` through ``modelio.load_checkpoint`` with the optimizer state (:148-166), the lr override after a resume
(:168-171), StepLR (:173-176, ``--lr_decay_step/--lr_decay_gamma``), per-epoch train + validation passes, best-score
tracking on ``total_loss`` and ``modelio.save_checkpoint`` (:375-398), ``decay_regul`` every ``--regul_decay_step``
epochs (:401-402).  What changes: ``torch.nn.DataParallel`` (:130) becomes one rank per GPU with identical replicas and
one NCCL all-reduce per step (``FlatAdamTrainer``); every rank draws its own shard; only rank 0 logs and writes
checkpoints (keys carry the ``module.`` prefix the reference's loader expects); the step runs as a CUDA graph.

Datasets are outside the hot-path scope, so the loader here is the synthetic stream of
bench.py (SURVEY.md §8d config 2/3 statistics); any iterable of sample dicts keyed by the query enums (Appendix B)
can be passed to ``run()`` instead.
"""
import argparse
import os
import random

import numpy as np
import torch
import torch.distributed as dist

from ..modelutils import modelio
from ..queries import BaseQueries, TransQueries
from .epochpass3d import epoch_pass

README_RECIPE = dict(resnet_version=18, mano_root="synthetic", mano_comps=30, mano_use_shape=True, mano_use_pca=True,
                     mano_neurons=[1024, 256], mano_center_idx=0, mano_lambda_verts=0.167, mano_lambda_joints3d=0.167,
                     mano_lambda_shape=0.167, mano_lambda_pose_reg=0.167, atlas_lambda=0.167, atlas_final_lambda=0.167,
                     atlas_mesh=True, atlas_predict_trans=True, atlas_predict_scale=True, atlas_trans_weight=0.167,
                     atlas_scale_weight=0.167, atlas_separate_encoder=True, atlas_ico_divisions=3, atlas_points_nb=600)


class SyntheticLoader(object):
    """``steps`` batches of ``batch`` samples, generated on the device from a per-rank seed (fixed per epoch index so
    that validation sees the same data every epoch)."""

    def __init__(self, steps, batch, img=256, n_gt=600, seed=0, device="cuda"):
        self.steps, self.batch, self.img, self.n_gt, self.seed, self.device = steps, batch, img, n_gt, seed, device

    def __len__(self):
        return self.steps

    def __iter__(self):
        g = torch.Generator(device=self.device).manual_seed(self.seed)
        B = self.batch
        for _ in range(self.steps):
            yield {TransQueries.images: torch.rand(B, 3, self.img, self.img, generator=g, device=self.device) - 0.5,
                   BaseQueries.sides: ["right" if i % 2 == 0 else "left" for i in range(B)], "root": "wrist",
                   TransQueries.joints3d: torch.randn(B, 21, 3, generator=g, device=self.device) * 40,
                   TransQueries.verts3d: torch.randn(B, 778, 3, generator=g, device=self.device) * 40,
                   TransQueries.objpoints3d: torch.randn(B, self.n_gt, 3, generator=g, device=self.device) * 40 + 30}


def run(args, model_kwargs=None, train_loader=None, val_loader=None, out=print):
    from ..networks.handnet import HandNet
    from ..trainer import FlatAdamTrainer
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # identical replicas: the same seed builds the same weights on every rank (no broadcast needed)
    torch.cuda.manual_seed_all(args.manual_seed)
    torch.manual_seed(args.manual_seed)
    np.random.seed(args.manual_seed)
    random.seed(args.manual_seed)
    if rank == 0:
        os.makedirs(args.exp_id, exist_ok=True)
    cfg = dict(README_RECIPE)
    cfg.update(model_kwargs or {})
    model = HandNet(**cfg).eval().cuda()
    trainer = FlatAdamTrainer(model, lr=args.lr, weight_decay=args.weight_decay, world_size=world)
    start_epoch, best_score = 0, None
    if args.resume:
        # the stored best score is dropped on resume, as the reference does (`start_epoch, _ = load_checkpoint`,
        # traineval.py:160-166): it may be an AUC (higher is better) or a loss (lower is better) depending on who wrote it
        start_epoch, _ = modelio.load_checkpoint(model, args.resume, optimizer=trainer, strict=False)
        trainer.lr = args.lr                      # "Override loaded learning rate" (traineval.py:168-171)
        trainer.set_lr_scale(1.0)
    n_gt = cfg.get("atlas_points_nb", 600)
    if train_loader is None:
        train_loader = SyntheticLoader(args.steps_per_epoch, args.batch_size, args.img_size, n_gt,
                                       seed=args.manual_seed * 1000 + rank)
    if val_loader is None:
        val_loader = SyntheticLoader(max(1, args.steps_per_epoch // 4), args.batch_size, args.img_size, n_gt,
                                     seed=args.manual_seed * 1000 + 500 + rank)
    history = []
    for epoch in range(start_epoch, args.epochs):
        if args.lr_decay_gamma:
            trainer.step_lr(epoch - start_epoch, args.lr_decay_step, args.lr_decay_gamma)
        train_meters, _ = epoch_pass(train_loader, model, epoch, optimizer=trainer, train=True,
                                     freeze_batchnorm=True, world_size=world, rank=rank, log_every=args.log_every,
                                     use_graph=not args.no_graph, out=out)
        val_meters, val_pck = epoch_pass(val_loader, model, epoch, train=False, world_size=world, rank=rank,
                                         log_every=0, out=out)
        val_total = val_meters.average_meters["total_loss"].avg
        # best-score bookkeeping of traineval.py:375-388: the validation AUC (higher is better) when joint errors were
        # evaluated, else the validation loss (lower is better); "best_score" in the checkpoint is that number.  (The
        # first evaluation counts as best here; the reference's strict comparison against itself never marks it.)
        if "auc" in val_pck:
            score = float(val_pck["auc"])
            is_best = best_score is None or score > best_score
            best_score = score if best_score is None else max(best_score, score)
        else:
            is_best = best_score is None or val_total < best_score
            best_score = val_total if best_score is None else min(best_score, val_total)
        history.append({"epoch": epoch + 1, "train_total": train_meters.average_meters["total_loss"].avg,
                        "val_total": val_total, "val_auc": float(val_pck.get("auc", float("nan")))})
        if rank == 0:
            out("epoch {}: train {:.4f} val {:.4f} auc {:.4f}".format(epoch + 1, history[-1]["train_total"], val_total,
                                                                     history[-1]["val_auc"]))
            modelio.save_checkpoint(
                {"epoch": epoch + 1, "network": "handnet",
                 "state_dict": {"module." + k: v for k, v in model.state_dict().items()},
                 "best_score": best_score, "optimizer": trainer.state_dict()},
                is_best=is_best, checkpoint=args.exp_id, snapshot=args.snapshot)
        if args.regul_decay_step and epoch % args.regul_decay_step == 0:
            model.decay_regul(args.regul_decay_gamma)
        if world > 1:
            dist.barrier()
    return history


def build_parser():
    ap = argparse.ArgumentParser(description="B200 data-parallel training of the obman hot path")
    ap.add_argument("--exp_id", default="checkpoints/b200_debug")
    ap.add_argument("--epochs", type=int, default=2)
    ap.add_argument("--batch_size", type=int, default=64, help="per GPU")
    ap.add_argument("--steps_per_epoch", type=int, default=20)
    ap.add_argument("--img_size", type=int, default=256)
    ap.add_argument("--lr", type=float, default=1e-4)
    ap.add_argument("--weight_decay", type=float, default=0.0)
    ap.add_argument("--lr_decay_step", type=int, default=300)
    ap.add_argument("--lr_decay_gamma", type=float, default=0.5)
    ap.add_argument("--regul_decay_step", type=int, default=300)
    ap.add_argument("--regul_decay_gamma", type=float, default=0.5)
    ap.add_argument("--manual_seed", type=int, default=0)
    ap.add_argument("--resume", default=None)
    ap.add_argument("--snapshot", type=int, default=None)
    ap.add_argument("--log_every", type=int, default=10)
    ap.add_argument("--no_graph", action="store_true")
    return ap


if __name__ == "__main__":
    run(build_parser().parse_args())
    # no destroy_process_group(): tearing NCCL down while captured graphs that contain its kernels are alive can hang
