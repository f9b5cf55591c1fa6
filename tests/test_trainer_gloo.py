"""Host-side logic of the data-parallel step on CPU with the gloo backend, world size 2: flat parameter /
gradient buffers, autograd accumulating in place into the flat gradient views, ONE all-reduce per step whose
result equals the sum of the per-rank gradients (so that grad_scale = 1/world gives the global-batch mean)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


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
    from obman_train_b200.trainer import FlatAdamTrainer
    torch.manual_seed(0)  # identical replicas
    model = torch.nn.Sequential(torch.nn.Linear(7, 33), torch.nn.ReLU(), torch.nn.Linear(33, 3))
    trainer = FlatAdamTrainer(model, world_size=world)
    assert trainer.grads_are_views()
    base = trainer.flat_p.data_ptr()
    assert all((p.data_ptr() - base) % 256 == 0 for p in trainer.params)  # CUDA allocations are 512-B aligned
    g = torch.Generator().manual_seed(100 + rank)  # different shard per rank
    x, y = torch.randn(5, 7, generator=g), torch.randn(5, 3, generator=g)
    for p in trainer.params:
        p.grad = None
    loss = ((model(x) - y) ** 2).mean()
    loss.backward()
    trainer.gather_grads()
    assert trainer.grads_are_views()  # gradients live in the flat buffer
    local = trainer.flat_g.clone()
    trainer.all_reduce_grads()
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    assert torch.allclose(trainer.flat_g, sum(gathered), atol=1e-6)
    # replicas agree after the exchange
    ref = trainer.flat_g.clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(ref, trainer.flat_g)
    try:
        trainer.adam_update()
        raised = False
    except RuntimeError:
        raised = True
    assert raised  # there is no CPU Adam: the product path fails loudly without CUDA
    if rank == 0:
        torch.save({"ok": True, "numel": trainer.numel, "param_numel": trainer.param_numel}, out)
    dist.destroy_process_group()


def test_flat_gradient_allreduce_world2_gloo(tmp_path):
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = torch.load(out)
    assert res["ok"] and res["numel"] >= res["param_numel"]


def test_bench_reference_arm_runs_on_cpu():
    """`bench.py --impl reference` must print one JSON line with the contract keys (tiny batch here)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, OBMAN_BENCH_CPU_BATCH="1")
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, env=env)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0


def test_bench_product_arm_fails_loudly_without_cuda():
    """The product path has no CPU fallback: without a CUDA device bench.py must stop with an error, not fall back."""
    import subprocess
    import sys
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("needs a machine without CUDA")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode != 0
    assert "no CPU fallback" in r.stderr and not any(l.startswith("{") for l in r.stdout.splitlines())


def test_complement_ranges_and_side_mask_helpers():
    from obman_train_b200.trainer import FlatAdamTrainer, complement_ranges
    assert complement_ranges([], 10) == [(0, 10)]
    assert complement_ranges([(4, 6), (0, 2)], 10) == [(2, 4), (6, 10)]
    assert complement_ranges([(0, 10)], 10) == []
    assert complement_ranges([(3, 10)], 10) == [(0, 3)]
    from obman_train_b200.queries import BaseQueries
    assert FlatAdamTrainer._sides_of({BaseQueries.sides: ["left", "right"], "root": "wrist"}) == ["left", "right"]
    assert FlatAdamTrainer._sides_of({"sides": ("right",)}) == ["right"]
    assert FlatAdamTrainer._sides_of({"images": 0}) is None
