"""Host logic of the epoch loop and the evaluation mirrors on CPU: device-side loss log,
PCK / AUC measures against the reference's own EvalUtil, world-size-2 reduction over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from obman_train_b200.evaluation.evalutils import AverageMeters
from obman_train_b200.evaluation.zimeval import EvalUtil
from obman_train_b200.netscripts.epochpass3d import epoch_pass
from obman_train_b200.queries import TransQueries


class _StubModel(object):
    """Stands in for HandNet on CPU: (total, results, losses) with the keys epoch_pass consumes."""

    def __init__(self):
        self.mode = None

    def eval(self):
        self.mode = "eval"

    def train(self):
        self.mode = "train"

    def forward(self, sample):
        joints = sample[TransQueries.joints3d] + sample["offset"].view(-1, 1, 1)
        loss = (sample["offset"] ** 2).mean().reshape(1)
        return loss, {"joints": joints}, {"mano_joints3d": loss[0] * 2, "atlas_objpoints3d": None, "total_loss": loss}


class _StubTrainer(object):
    def __init__(self, model):
        self.model, self.steps = model, 0

    def step(self, sample, return_all=False):
        self.steps += 1
        out = self.model.forward(sample)
        return out if return_all else out[0]


def _loader(n_steps, B, seed, rank=0):
    g = torch.Generator().manual_seed(seed + 100 * rank)
    return [{TransQueries.joints3d: torch.randn(B, 21, 3, generator=g) * 40,
             "offset": torch.rand(B, generator=g) * 30} for _ in range(n_steps)]


def test_epoch_pass_losses_and_pck_single_process():
    model = _StubModel()
    trainer = _StubTrainer(model)
    loader = _loader(5, 4, 0)
    lines = []
    meters, pck = epoch_pass(loader, model, epoch=3, optimizer=trainer, train=True, freeze_batchnorm=True,
                             log_every=2, out=lines.append)
    assert model.mode == "eval" and trainer.steps == 5   # --freeze_batchnorm: eval-mode BN while training
    assert isinstance(meters, AverageMeters)
    exp_total = np.mean([float((s["offset"] ** 2).mean()) for s in loader])
    assert meters.average_meters["total_loss"].avg == pytest.approx(exp_total, rel=1e-6)
    assert meters.average_meters["mano_joints3d"].avg == pytest.approx(2 * exp_total, rel=1e-6)
    assert "atlas_objpoints3d" not in meters.average_meters   # None losses are skipped like the reference does
    # every joint of sample b is displaced by offset*(1,1,1): distance = sqrt(3)*offset
    d = np.concatenate([np.repeat((np.sqrt(3) * s["offset"].numpy())[:, None], 21, 1) for s in loader])
    ref = EvalUtil()
    ref.feed_distances(d)
    exp = ref.get_measures(0, 50, 20)
    assert pck["epe_mean"] == pytest.approx(exp[0], rel=1e-5) and pck["auc"] == pytest.approx(exp[3], rel=1e-5)
    assert np.allclose(pck["pck_curve"], exp[4], atol=1e-6) and len(pck["thres"]) == 20
    assert lines[0] == "epoch: 3" and sum("Loss:" in l for l in lines) == 2
    # evaluation pass: no optimizer, no_grad
    meters_val, _ = epoch_pass(loader, model, epoch=3, train=False, out=lines.append)
    assert meters_val.average_meters["total_loss"].avg == pytest.approx(exp_total, rel=1e-6)
    with pytest.raises(ValueError):
        epoch_pass(loader, model, epoch=0, train=True, optimizer=None, out=lines.append)
    with pytest.raises(NotImplementedError):
        epoch_pass(loader, model, epoch=0, train=False, display=True, out=lines.append)


def test_evalutil_matches_bruteforce_definition():
    rng = np.random.RandomState(1)
    ev = EvalUtil(num_kp=5)
    gts, prs, viss = [], [], []
    for i in range(30):
        gt, pr = rng.randn(5, 3) * 30, rng.randn(5, 3) * 30
        vis = rng.rand(5) > 0.3
        vis[4] = False  # a keypoint that is never visible is left out of every average
        ev.feed(torch.tensor(gt), torch.tensor(pr), keypoint_vis=vis)
        gts.append(gt); prs.append(pr); viss.append(vis)
    epe_mean, per_joint, epe_median, auc, curve, thr = ev.get_measures(0, 50, 20)
    d = np.sqrt(((np.array(gts) - np.array(prs)) ** 2).sum(2))
    v = np.array(viss)
    means = [d[v[:, k], k].mean() for k in range(4)]
    assert len(per_joint) == 4 and epe_mean == pytest.approx(np.mean(means))
    assert epe_median == pytest.approx(np.mean([np.median(d[v[:, k], k]) for k in range(4)]))
    pck_k = np.array([[np.mean(d[v[:, k], k] <= t) for t in thr] for k in range(4)])
    assert np.allclose(curve, pck_k.mean(0))
    trap = lambda y: np.sum((y[1:] + y[:-1]) * np.diff(thr)) / 2  # noqa: E731
    assert auc == pytest.approx(np.mean([trap(pck_k[k]) / trap(np.ones_like(thr)) for k in range(4)]))
    assert [len(x) for x in ev.data] == [int(v[:, k].sum()) for k in range(5)]


@pytest.mark.reference
def test_evalutil_matches_reference_class(mano_tables_np):
    from oracle import refhook
    refhook.set_mano_tables(mano_tables_np["right"], mano_tables_np["left"])
    refhook.install()
    if not hasattr(np, "trapz"):
        np.trapz = np.trapezoid  # numpy >= 2.4 dropped the alias the reference calls (zimeval.py:87,113)
    from mano_train.evaluation.zimeval import EvalUtil as RefEvalUtil
    rng = np.random.RandomState(0)
    a, b = RefEvalUtil(), EvalUtil()
    for i in range(40):
        gt, pr = rng.randn(21, 3) * 40, rng.randn(21, 3) * 40
        vis = (rng.rand(21) > 0.2) if i % 3 else None
        a.feed(torch.tensor(gt), torch.tensor(pr), keypoint_vis=vis)
        b.feed(gt, pr, keypoint_vis=vis)
    for x, y in zip(a.get_measures(0, 50, 20), b.get_measures(0, 50, 20)):
        assert np.allclose(np.asarray(x), np.asarray(y))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    model = _StubModel()
    meters, pck = epoch_pass(_loader(3, 4, 7, rank), model, epoch=0, optimizer=_StubTrainer(model), train=True,
                             world_size=world, rank=rank, log_every=0, out=lambda *_: None)
    if rank == 0:
        torch.save({"total": meters.average_meters["total_loss"].avg, "auc": float(pck["auc"]),
                    "n": len(pck["evaluator"].data[0])}, out)
    dist.destroy_process_group()


def test_epoch_pass_world2_gloo_reduces_losses_and_gathers_distances(tmp_path):
    out = str(tmp_path / "r.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = torch.load(out)
    shards = [_loader(3, 4, 7, r) for r in range(2)]
    exp_total = np.mean([float((s["offset"] ** 2).mean()) for sh in shards for s in sh])
    assert res["total"] == pytest.approx(exp_total, rel=1e-6)
    assert res["n"] == 2 * 3 * 4   # distances of both ranks reach the evaluator
    d = np.concatenate([np.repeat((np.sqrt(3) * s["offset"].numpy())[:, None], 21, 1) for sh in shards for s in sh])
    ref = EvalUtil()
    ref.feed_distances(d)
    assert res["auc"] == pytest.approx(float(ref.get_measures(0, 50, 20)[3]), rel=1e-6)


def test_freeze_helpers_and_trainer_parameter_selection():
    """netutils.rec_freeze / freeze_batchnorm_stats (netutils.py:4-19) and their effect on FlatAdamTrainer: frozen
    parameters stay out of the flat buffers, and the optimizer indices follow the reference's
    ``filter(requires_grad, model.parameters())`` numbering (traineval.py:105-116)."""
    from obman_train_b200.networks import netutils
    from obman_train_b200.trainer import FlatAdamTrainer
    torch.manual_seed(0)
    enc = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3), torch.nn.BatchNorm2d(4))
    head = torch.nn.Sequential(torch.nn.Linear(4, 5), torch.nn.BatchNorm1d(5), torch.nn.Linear(5, 2))
    model = torch.nn.ModuleDict({"base_net": enc, "head": head})
    netutils.freeze_batchnorm_stats(model)
    assert all(m.momentum == 0 for m in model.modules() if isinstance(m, torch.nn.modules.batchnorm._BatchNorm))
    assert all(p.requires_grad for p in model.parameters())
    netutils.rec_freeze(model["base_net"])
    assert not any(p.requires_grad for p in enc.parameters()) and all(p.requires_grad for p in head.parameters())
    trainer = FlatAdamTrainer(model)
    assert trainer.names == ["head." + n for n, _ in head.named_parameters()]
    assert trainer.optim_len == len(list(head.parameters())) and trainer.optim_index == list(range(trainer.optim_len))
    ref_opt = torch.optim.Adam(filter(lambda p: p.requires_grad, model.parameters()), lr=1e-4)
    assert len(ref_opt.state_dict()["param_groups"][0]["params"]) == trainer.optim_len
    # the frozen encoder's storage was left alone, the trainable parameters became views of the flat buffer
    base = trainer.flat_p.untyped_storage().data_ptr()
    assert all(p.untyped_storage().data_ptr() == base for p in head.parameters())
    assert all(p.untyped_storage().data_ptr() != base for p in enc.parameters())
    m = AverageMeters()
    m.add_loss_value("a", 2.0, n=3)
    m.add_loss_value("a", 4.0)
    a = m.average_meters["a"]
    assert (a.val, a.sum, a.count, a.avg) == (4.0, 10.0, 4, 2.5)
