"""Launch a few representative tensor-core kernels of the B=64 train step once each (for `ncu --set full`)
or repeatedly with CUDA-event timing (default)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from obman_train_b200 import dense  # noqa: E402

B = int(os.environ.get("PROF_B", "64"))
PASSES = int(os.environ.get("PROF_PASSES", "3"))
WG_PASSES = int(os.environ.get("PROF_WG_PASSES", "3"))
REPS = int(os.environ.get("PROF_REPS", "1"))


def conv_case(h, cin, cout, k=3, stride=1):
    x = torch.randn(B, h, h, cin, device="cuda")
    w = torch.randn(cout, k * k * cin, device="cuda") / (k * k * cin) ** 0.5
    dh, dw, phase, slot, step = dense.fprop_taps(k, stride, k // 2)
    ho = h // stride
    out = torch.empty(B, ho, ho, cout, device="cuda")
    bias = torch.randn(cout, device="cuda")
    flops = 2.0 * B * ho * ho * cout * k * k * cin
    w_lo = None
    if PASSES == 3 and os.environ.get("PROF_PRESPLIT", "1") == "1":
        w, w_lo = dense.split_tf32(w)
    if PASSES == 2:
        w = dense.pack_bf16(w)
    return (lambda: dense.conv_nhwc(x, w, cout, (dh, dw, phase, slot), step, out, ho, ho, bias=bias, relu=True, passes=PASSES, w_lo=w_lo)), flops


def wgrad_case(h, cin, cout, k=3, stride=1):
    ho = h // stride
    x = torch.randn(B, h, h, cin, device="cuda")
    dy = torch.randn(B, ho, ho, cout, device="cuda")
    dh, dw, phase, slot, step = dense.fprop_taps(k, stride, k // 2)
    dwt = torch.empty(cout, k * k * cin, device="cuda")
    flops = 2.0 * B * ho * ho * cout * k * k * cin
    return (lambda: dense.wgrad_nhwc(dy, x, (dh, dw, phase, slot), step, dwt, passes=WG_PASSES)), flops


def gemm_case(M, N, K):
    a = torch.randn(M, (K + 31) // 32 * 32, device="cuda")
    w = torch.randn(N, (K + 3) // 4 * 4, device="cuda")
    w_lo = None
    if PASSES == 3 and os.environ.get("PROF_PRESPLIT", "1") == "1":
        w, w_lo = dense.split_tf32(w)
    if PASSES == 2:
        w = dense.pack_bf16(w, K)
    return (lambda: dense.gemm(a, w, relu=True, passes=PASSES, n=N, k=K, w_lo=w_lo, packed=PASSES == 2)), 2.0 * M * N * K


CASES = [
    ("conv3x3 64x64 64->64", conv_case(64, 64, 64)),
    ("conv3x3 32x32 128->128", conv_case(32, 128, 128)),
    ("conv3x3 16x16 256->256", conv_case(16, 256, 256)),
    ("conv3x3 8x8 512->512", conv_case(8, 512, 512)),
    ("wgrad3x3 64x64 64->64", wgrad_case(64, 64, 64)),
    ("wgrad3x3 32x32 128->128", wgrad_case(32, 128, 128)),
    ("wgrad3x3 16x16 256->256", wgrad_case(16, 256, 256)),
    ("wgrad3x3 8x8 512->512", wgrad_case(8, 512, 512)),
    ("gemm 41088x257x515", gemm_case(B * 642, 257, 515)),
    ("stem-like wgrad 128x128 32->64 (16 taps)", None),
]


def main():
    for name, case in CASES:
        if case is None:
            continue
        fn, flops = case
        fn()
        torch.cuda.synchronize()
        if REPS > 1:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(REPS):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / REPS
            print("%-28s %8.3f ms  %7.1f TFLOP/s (algorithmic, passes=%d)" % (name, ms, flops / ms / 1e9, PASSES))


if __name__ == "__main__":
    main()
