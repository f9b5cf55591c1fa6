"""Wait-time accounting of the persistent decoder GEMM (gemm_persist.cu) on the point decoder's shapes: which warp
role is blocked on which barrier, SM clocks per CTA (mean over CTAs).  Run with OBMAN_GEMM_PERSIST=2.

    OBMAN_GEMM_PERSIST=2 python scripts/trace_gemm_persist.py [> profiles/...]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from obman_train_b200 import dense  # noqa: E402
from obman_train_b200._lib import call  # noqa: E402

M = int(os.environ.get("PROF_M", str(256 * 2562)))
NAMES = {0: "kernel body (thread 0)", 1: "producer: wait stage free", 2: "MMA: wait TMA landed", 3: "MMA: wait A converted",
         4: "MMA: wait accumulator drained", 5: "MMA: loop total", 6: "splitter: wait TMA landed", 7: "splitter: loop total",
         8: "splitter: tcgen05.st + wait::st", 9: "epilogue: wait accumulator complete", 10: "epilogue: loop total",
         12: "tiles per CTA"}


def run(N, K, variant):
    r32 = lambda v: (v + 31) // 32 * 32  # noqa: E731
    a = torch.randn(M, r32(K), device="cuda")
    w = dense.pack_bf16(torch.randn(N, K, device="cuda") / K ** 0.5, K)
    out = torch.empty(M, r32(N), device="cuda")
    kw = {}
    if "bias" in variant:
        kw["bias"] = torch.randn(N, device="cuda")
        kw["relu"] = True
    if "mask" in variant:
        kw["mask_src"] = torch.randn(M, r32(N), device="cuda")
    fn = lambda: dense.gemm(a, w, out=out, passes=2, n=N, k=K, packed=True, **kw)  # noqa: E731
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    buf = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
    call("obman_debug_trace", buf.data_ptr(), buf.numel())
    fn()
    torch.cuda.synchronize()
    call("obman_debug_trace", None, 0)
    t = buf.view(148, 16).double()
    kb = r32(K) // 32
    tiles = t[:, 12].mean().item()
    print("gemm M%d N%d K%d %s: %.4f ms untraced; %d K blocks per tile" % (M, N, K, variant, ms, kb))
    for k in sorted(NAMES):
        v = t[:, k].mean().item()
        per = "" if k == 12 or tiles == 0 else "   %8.0f per tile %7.0f per K block" % (v / tiles, v / tiles / kb)
        print("   %-36s %12.0f%s" % (NAMES[k], v, per))


run(257, 515, "bias+relu")
run(515, 257, "mask")
run(257, 128, "mask")
run(128, 257, "bias+relu")
