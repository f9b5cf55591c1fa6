"""Kernel-level time breakdown of one eager training step with torch.profiler (CUPTI): which kernels the step spends
its GPU time in.  PROF_CONFIG=2|3 selects the bench workload.  Not a benchmark: the profiler serialises streams."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    wl = bench.Workload(int(os.environ.get("PROF_CONFIG", "3")))
    from obman_train_b200 import dense
    from obman_train_b200.networks.handnet import HandNet
    from obman_train_b200.trainer import FlatAdamTrainer
    from obman_train_b200.queries import TransQueries, BaseQueries
    dense.set_precision("bf16x3", "bf16x3")
    torch.manual_seed(0)
    model = HandNet(**wl.cfg).eval().cuda()
    trainer = FlatAdamTrainer(model, lr=1e-4)
    host = wl.sample(wl.batch, 1000)
    sample = {TransQueries.images: host["images"].cuda(), BaseQueries.sides: host["sides"], "root": host["root"],
              TransQueries.joints3d: host["joints3d"].cuda(), TransQueries.verts3d: host["verts3d"].cuda(),
              TransQueries.objpoints3d: host["objpoints3d"].cuda()}
    for _ in range(3):
        trainer.step(sample)
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    n = 3
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(n):
            trainer.step(sample)
        torch.cuda.synchronize()
    rows = []
    for e in prof.key_averages():
        t = getattr(e, "device_time_total", None)
        if t is None:
            t = getattr(e, "cuda_time_total", 0)
        if t > 0:
            rows.append((t / n, e.count / n, e.key))
    rows.sort(reverse=True)
    total = sum(r[0] for r in rows)
    print("# one step, config %s: %.1f us of kernel time in %d launches" % (
        os.environ.get("PROF_CONFIG", "2"), total, sum(r[1] for r in rows)))
    for t, c, k in rows[:45]:
        print("%9.1f us %5.1f%%  n=%5.1f  %s" % (t, 100 * t / total, c, k[:110]))


if __name__ == "__main__":
    main()
