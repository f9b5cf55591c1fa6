"""Data-parallel equivalence on hardware: one training step of a batch of 8 on ONE GPU equals
the same batch as 2 shards of 4 on TWO GPUs after the NCCL all-reduce of the flat gradient buffer (with the overlapped
layer3+4 range all-reduce on) and the fused Adam - eagerly and as a captured CUDA graph, followed by an orderly
teardown (graphs released before the process group is destroyed).  Needs two GPUs: `gpurun --gpus 2`."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu

B = 8


def _cfg():
    from tests.util import FULL_CFG
    cfg = dict(FULL_CFG)
    # masked_mean_loss of the contact terms is a ratio of batch-global sums (per-shard semantics under data parallelism,
    # DESIGN.md §6), every other term is a mean over samples: equal shards average exactly
    cfg.update(contact_lambda=0, collision_lambda=0, atlas_lambda_regul_edges=0.1)
    return cfg


def _build(world):
    from obman_train_b200.networks.handnet import HandNet
    from obman_train_b200.trainer import FlatAdamTrainer
    from tests.util import randomise_bn
    torch.manual_seed(21)
    model = HandNet(**_cfg())
    randomise_bn(model, 22)
    model = model.eval().cuda()
    return FlatAdamTrainer(model, lr=1e-4, world_size=world)


def _shard(rank, world):
    from tests.util import enum_sample, make_sample
    host = make_sample(B, 64, 23)
    per = B // world
    sl = {k: (v[rank * per:(rank + 1) * per] if torch.is_tensor(v) or isinstance(v, list) else v) for k, v in host.items()}
    return enum_sample(sl)


def _worker(rank, world, port, out_dir, use_graph):
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    trainer = _build(world)
    assert trainer.overlap_allreduce
    sample = _shard(rank, world)
    if use_graph:
        trainer.capture(dict(sample))
        loss = trainer.replay().clone()
    else:
        loss = trainer.step(dict(sample))
    torch.cuda.synchronize()
    if not use_graph:
        assert trainer._sink is not None
    torch.save({"g": trainer.flat_g.cpu(), "p": trainer.flat_p.cpu(), "loss": loss.cpu()},
               os.path.join(out_dir, "rank%d.pt" % rank))
    trainer.release_graph()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    dist.destroy_process_group()


@pytest.mark.parametrize("use_graph", [False, True])
def test_two_gpu_shards_equal_one_gpu_batch(tmp_path, use_graph):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path), use_graph), nprocs=2, join=True)
    r0 = torch.load(str(tmp_path / "rank0.pt"))
    r1 = torch.load(str(tmp_path / "rank1.pt"))
    assert torch.equal(r0["p"], r1["p"]), "replicas diverged"
    assert torch.equal(r0["g"], r1["g"])
    single = _build(1)
    p0 = single.flat_p.clone().cpu()
    loss = single.step(_shard(0, 1))
    torch.cuda.synchronize()
    g1 = single.flat_g.cpu()
    g2 = r0["g"] / 2.0                      # the all-reduce sums, Adam applies grad_scale = 1 / world
    rel = ((g2 - g1).norm() / g1.norm()).item()
    worst = ((g2 - g1).abs().max() / g1.abs().max()).item()
    print("gradient: L2 rel %.2e, max rel %.2e" % (rel, worst))
    assert rel < 1e-5 and worst < 1e-5
    assert abs(0.5 * (r0["loss"] + r1["loss"]).item() - loss.item()) <= 1e-5 * abs(loss.item())
    # parameters: the first Adam step moves every entry by ~lr * g / (|g| + eps), i.e. by ~lr * sign(g): an entry whose
    # gradient is within rounding of zero may legitimately move the other way (|difference| <= 2 lr), everything else
    # moves identically.  Hence: an absolute bound of 2 lr everywhere, almost all entries equal to 1e-6 relative, and the
    # update equal norm-wise.
    p1 = single.flat_p.cpu()
    diff = (r0["p"] - p1).abs()
    upd_rel = (diff.norm() / (p1 - p0).norm()).item()
    moved = (p1 - p0).abs() > 0
    off = (diff > 1e-6 * p1.abs() + 1e-9) & moved
    frac_off = off.float().sum().item() / max(1.0, moved.float().sum().item())
    print("update rel %.2e, max |dp| %.2e (lr 1e-4), entries off %.2e" % (upd_rel, diff.max().item(), frac_off))
    assert diff.max().item() <= 2.1e-4
    assert frac_off < 2e-2 and upd_rel < 5e-2
