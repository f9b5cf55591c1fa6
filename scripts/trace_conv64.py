"""Wait-time accounting of the persistent 64-channel convolution kernel (obman_debug_trace): which warp role waits
for which barrier, in SM clocks per CTA (mean over CTAs).  `python scripts/trace_conv64.py [> profiles/...]`"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from obman_train_b200 import dense  # noqa: E402
from obman_train_b200._lib import call  # noqa: E402

B = int(os.environ.get("PROF_B", "256"))
NAMES_V2 = {0: "kernel body (thread 0)", 1: "producer: wait halo buffer free", 2: "converter: wait halo landed",
            5: "converter: loop total", 13: "MMA: wait box converted", 7: "MMA: wait accumulator drained", 8: "MMA: loop total",
            9: "epilogue group 0: wait accumulator", 10: "epilogue group 0: loop total",
            3: "epilogue group 0: operand fetch (next round)", 4: "epilogue group 0: tcgen05.ld + staging stores",
            6: "epilogue group 0: barrier after staging", 11: "epilogue group 0: read + global stores + barrier",
            14: "epilogue group 1: wait accumulator", 15: "epilogue group 1: loop total", 12: "tiles per CTA"}
NAMES = {0: "kernel body (thread 0)", 1: "producer: wait halo buffer free", 2: "splitter: wait halo landed",
         3: "splitter: wait A stage free", 4: "splitter: in-place split + barrier", 5: "splitter: loop total",
         13: "MMA: wait A stage written", 7: "MMA: wait accumulator drained", 8: "MMA: loop total",
         9: "epilogue: wait accumulator complete", 10: "epilogue: loop total", 11: "MMA: wait weights", 12: "tiles per CTA"}


def run(h, cin, cout, variant):
    x = torch.randn(B, h, h, cin, device="cuda")
    if "stem" in variant:   # the stem's tap pattern: four vertical taps (-2 .. 1), no horizontal ones
        w = dense.pack_bf16(torch.randn(cout, 4 * cin, device="cuda") / (4 * cin) ** 0.5)
        dh, dw, phase, slot, step = [-2, -1, 0, 1], [0, 0, 0, 0], [0] * 4, list(range(4)), 1
    else:
        w = dense.pack_bf16(torch.randn(cout, 9 * cin, device="cuda") / (9 * cin) ** 0.5)
        dh, dw, phase, slot, step = dense.fprop_taps(3, 1, 1)
    out = torch.empty(B, h, h, cout, device="cuda")
    kw = {}
    if "add" in variant:
        kw["addend"] = torch.randn(B, h, h, cout, device="cuda")
    if "mask" in variant:
        kw["mask_src"] = torch.randn(B, h, h, cout, device="cuda")
    fn = lambda: dense.conv_nhwc(x, w, cout, (dh, dw, phase, slot), 1, out, h, h, passes=2, **kw)  # noqa: E731
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    buf = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
    call("obman_debug_trace", buf.data_ptr(), buf.numel())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    call("obman_debug_trace", None, 0)
    t = buf.view(148, 16).double()
    print("conv %dx%d c%d->%d %s, B=%d: %.4f ms (traced launch)" % (h, h, cin, cout, variant, B, e0.elapsed_time(e1)))
    names = NAMES if os.environ.get("OBMAN_CONV64_GEN") == "1" else NAMES_V2
    for k in sorted(names):
        print("   %-40s %12.0f" % (names[k], t[:, k].mean().item()))


for variant in ("plain", "mask+add"):
    run(64, 64, 64, variant)
run(128, 64, 64, "stem-like plain")
