"""Time the nearest-neighbour search kernel at the workload sizes (Chamfer 2562 x 2500, contact 778 x 2562), B = 256."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from obman_train_b200._lib import call, ptr, stream_ptr  # noqa: E402

B = int(os.environ.get("PROF_B", "256"))
for N, M in ((2562, 2500), (778, 2562), (10000, 10000)):
    b = B if N < 10000 else 64
    x = torch.randn(b, N, 3, device="cuda") * 50
    y = torch.randn(b, M, 3, device="cuda") * 50
    minx, miny = torch.empty(b, N, device="cuda"), torch.empty(b, M, device="cuda")
    idxx = torch.empty(b, N, dtype=torch.int32, device="cuda")
    idxy = torch.empty(b, M, dtype=torch.int32, device="cuda")
    fn = lambda: call("obman_nn_fwd", ptr(x), ptr(y), b, N, M, ptr(minx), ptr(idxx), ptr(miny), ptr(idxy), 3, stream_ptr())  # noqa: E731
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    pairs = 2.0 * b * N * M
    # exact check against a direct evaluation on a few samples
    d = ((x[:2, :, None, :] - y[:2, None, :, :]) ** 2).sum(-1)
    ok = bool((d.min(2).values - minx[:2]).abs().max() < 1e-3) and bool((d.argmin(2).int() == idxx[:2]).float().mean() > 0.999)
    print("nn %5d x %5d  B=%3d  %.4f ms  %.2f Tpairs/s  packed=%s  ok=%s" % (N, M, b, ms, pairs / ms / 1e9, os.environ.get("OBMAN_NN_PACKED", "1"), ok))
