"""A/B micro-benchmark of one convolution shape: which epilogue / tap-order factor costs what (CUDA events)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from obman_train_b200 import dense  # noqa: E402

B = int(os.environ.get("AB_B", "64"))
REPS = int(os.environ.get("AB_REPS", "30"))


def run(h, cin, cout, variant):
    x = torch.randn(B, h, h, cin, device="cuda")
    w = dense.pack_bf16(torch.randn(cout, 9 * cin, device="cuda") / (9 * cin) ** 0.5)
    dh, dw, phase, slot, step = dense.fprop_taps(3, 1, 1)
    if "rev" in variant:
        dh, dw, slot = dense.dgrad_taps(3, 1, 1, (0, 0))
        phase = None
    out = torch.empty(B, h, h, cout, device="cuda")
    kw = {}
    if "bias" in variant:
        kw["bias"] = torch.randn(cout, device="cuda")
    if "relu" in variant:
        kw["relu"] = True
    if "add" in variant:
        kw["addend"] = torch.randn(B, h, h, cout, device="cuda")
    if "mask" in variant:
        kw["mask_src"] = torch.randn(B, h, h, cout, device="cuda")
    fn = lambda: dense.conv_nhwc(x, w, cout, (dh, dw, phase, slot), 1, out, h, h, passes=2, **kw)  # noqa: E731
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(REPS):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / REPS
    fl = 2.0 * B * h * h * cout * 9 * cin
    print("%3dx%-3d c%d->%d %-18s %7.4f ms %7.1f TFLOP/s" % (h, h, cin, cout, variant, ms, fl / ms / 1e9))


for shape in ((64, 64, 64), (32, 128, 128), (16, 256, 256), (8, 512, 512)):
    for variant in ("plain", "bias+relu", "bias+relu+add", "rev", "rev+mask", "rev+mask+add", "mask"):
        run(*shape, variant)
