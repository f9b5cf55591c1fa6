"""Per-CTA phase timing of the tensor-core kernels (obman_debug_trace): where does a CTA's lifetime go?

For a few layer shapes of the B=64 step, every CTA stamps clock64 at: entry, setup done, first / last TMA issued,
first operands ready at the MMA warp, last MMA issued, accumulator complete, exit.  Prints the mean phase lengths
in SM clocks and the CTA lifetime against the pure tensor-core time of its tile.

    python scripts/trace_kernels.py [> profiles/cta_phases_*.txt]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from obman_train_b200 import dense  # noqa: E402
from obman_train_b200._lib import call  # noqa: E402

B = int(os.environ.get("PROF_B", "64"))
P = int(os.environ.get("PROF_PASSES", "2"))


def conv_case(h, cin, cout, k=3, stride=1, addend=False, mask=False):
    x = torch.randn(B, h, h, cin, device="cuda")
    w = torch.randn(cout, k * k * cin, device="cuda") / (k * k * cin) ** 0.5
    dh, dw, phase, slot, step = dense.fprop_taps(k, stride, k // 2)
    ho = h // stride
    out = torch.empty(B, ho, ho, cout, device="cuda")
    bias = torch.randn(cout, device="cuda")
    w_lo = None
    if P == 3:
        w, w_lo = dense.split_tf32(w)
    if P == 2:
        w = dense.pack_bf16(w)
    add = torch.randn_like(out) if addend else None
    msk = torch.randn_like(out) if mask else None
    return lambda: dense.conv_nhwc(x, w, cout, (dh, dw, phase, slot), step, out, ho, ho, bias=bias, relu=True,
                                   addend=add, mask_src=msk, passes=P, w_lo=w_lo)


def wgrad_case(h, cin, cout, k=3, stride=1):
    ho = h // stride
    x = torch.randn(B, h, h, cin, device="cuda")
    dy = torch.randn(B, ho, ho, cout, device="cuda")
    dh, dw, phase, slot, step = dense.fprop_taps(k, stride, k // 2)
    dwt = torch.empty(cout, k * k * cin, device="cuda")
    return lambda: dense.wgrad_nhwc(dy, x, (dh, dw, phase, slot), step, dwt, passes=P)


def gemm_case(M, N, K):
    a = torch.randn(M, (K + 31) // 32 * 32, device="cuda")
    w = torch.randn(N, (K + 3) // 4 * 4, device="cuda")
    kw = {}
    if P == 2:
        w, kw = dense.pack_bf16(w, K), {"packed": True}
    return lambda: dense.gemm(a, w, relu=True, passes=P, n=N, k=K, **kw)


CASES = [
    ("conv3x3 64x64 64->64", conv_case(64, 64, 64)),
    ("conv3x3 64x64 64->64 +addend", conv_case(64, 64, 64, addend=True)),
    ("conv3x3 64x64 64->64 +addend+mask", conv_case(64, 64, 64, addend=True, mask=True)),
    ("conv3x3 32x32 128->128", conv_case(32, 128, 128)),
    ("conv3x3 16x16 256->256", conv_case(16, 256, 256)),
    ("conv3x3 8x8 512->512", conv_case(8, 512, 512)),
    ("wgrad3x3 64x64 64->64", wgrad_case(64, 64, 64)),
    ("wgrad3x3 32x32 128->128", wgrad_case(32, 128, 128)),
    ("wgrad3x3 8x8 512->512", wgrad_case(8, 512, 512)),
    ("gemm 41088x257x515", gemm_case(B * 642, 257, 515)),
]
NAMES = ["setup", "first TMA issued", "all TMA issued", "first operands ready", "last MMA issued",
         "accumulator complete", "epilogue+exit"]


def main():
    cap = 1 << 22
    buf = torch.zeros(cap, dtype=torch.int64, device="cuda")
    only = os.environ.get("TRACE_ONLY")
    for name, fn in CASES:
        if only and only not in name:
            continue
        fn()  # warm (tensor maps, attributes, L2)
        torch.cuda.synchronize()
        buf.zero_()
        call("obman_debug_trace", buf.data_ptr(), cap)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        call("obman_debug_trace", None, 0)
        t = buf.view(-1, 16).cpu()
        fine = (t[:, 8:16] & 0x0000ffffffffffff).double()
        keep = t[:, 0] != 0
        t = t[keep]
        fine = fine[keep]
        smid = (t[:, 7] >> 48) & 0xffff
        t7 = t[:, 7] & 0x0000ffffffffffff
        t = torch.cat([t[:, :7] & 0x0000ffffffffffff, t7[:, None]], 1).double()
        life = (t[:, 7] - t[:, 0])
        print("%-36s %4d CTAs on %3d SMs, kernel %.1f us; CTA lifetime mean %.0f clk (min %.0f max %.0f)" % (
            name, t.shape[0], smid.unique().numel(), e0.elapsed_time(e1) * 1e3, life.mean(), life.min(), life.max()))
        # stamps are taken by different warps; report each relative to entry
        order = [1, 2, 4, 3, 5, 6, 7]
        labels = {1: "setup done", 2: "first TMA issued", 4: "first operands ready (MMA warp)", 3: "last TMA issued",
                  5: "last MMA issued", 6: "accumulator complete", 7: "exit"}
        for k in order:
            d = t[:, k] - t[:, 0]
            print("      %-34s +%8.0f clk (median %8.0f)" % (labels[k], d.mean(), d.median()))
        if (fine[:, 0] > 0).any():
            ok = fine[:, 0] > 0
            f = fine[ok]
            wg = name.startswith("wgrad")
            names = ["producer: raw slot free (it=12)", "producer: loads issued", "splitter: raw tile landed",
                     "splitter: converted slot free", "splitter: conversion done", "MMA warp: operands ready",
                     "MMA warp: 6 MMAs + commit issued", "splitter: raw tile landed (it=13)"] if wg else [
                     "splitter: slot free (it=12)", "splitter: row read + split done", "splitter: TMEM store complete",
                     "MMA warp: weight tile landed", "MMA warp: A operand ready", "MMA warp: 6 MMAs + commit issued",
                     "splitter: slot free (it=13)", "MMA warp: A operand ready (it=13)"]
            for k in range(8):
                d = f[:, k] - f[:, 0]
                print("      [it 12] %-38s %+8.0f clk (median %+8.0f)" % (names[k], d.mean(), d.median()))
        per_sm = {}
        for s, a, b in zip(smid.tolist(), t[:, 0].tolist(), t[:, 7].tolist()):
            lo, hi, n = per_sm.get(s, (a, b, 0))
            per_sm[s] = (min(lo, a), max(hi, b), n + 1)
        spans = torch.tensor([hi - lo for lo, hi, _ in per_sm.values()])
        print("      per-SM busy span mean %.0f clk, CTAs/SM mean %.1f (clock64 is per SM: spans are not comparable across SMs)" % (
            spans.mean(), sum(n for _, _, n in per_sm.values()) / len(per_sm)))


if __name__ == "__main__":
    main()
