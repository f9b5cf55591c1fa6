"""Bring-up probe for the tcgen05 dense kernels: each case runs in its own subprocess (a device fault in
one case must not poison the others) and reports its error against an fp64 torch reference.

    python scripts/probe_dense.py            # run all cases, write gpurun_out/probe_dense.json
    python scripts/probe_dense.py <case>     # run one case in-process
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _stats(got, ref):
    import torch
    got = got.double()
    err = (got - ref).abs()
    scale = ref.abs().max().item() + 1e-30
    return {"max_err": err.max().item(), "rel": err.max().item() / scale, "ref_max": scale,
            "nan": bool(torch.isnan(got).any().item()),
            "frac_bad": (err > 1e-2 * scale).double().mean().item()}


def case_gemm(M, N, K, passes, epilogue=False):
    import torch
    from obman_train_b200 import dense
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    kp = (K + 3) // 4 * 4
    a = torch.zeros(M, kp, device="cuda"); a[:, :K] = torch.randn(M, K, device="cuda", generator=g)
    w = torch.zeros(N, kp, device="cuda"); w[:, :K] = torch.randn(N, K, device="cuda", generator=g)
    ref = a.double() @ w.double().t()
    kw = {}
    if epilogue:
        bias = torch.randn(N, device="cuda", generator=g)
        addend = torch.randn(M, N, device="cuda", generator=g)
        mask = torch.randn(M, N, device="cuda", generator=g)
        ref = torch.relu(0.5 * ref + bias.double() + addend.double()) * (mask > 0).double()
        kw = dict(bias=bias, addend=addend, mask_src=mask, alpha=0.5, relu=True)
    if os.environ.get("PROBE_PRESPLIT", "1") == "1" and passes == 3:
        w, kw["w_lo"] = dense.split_tf32(w)
    if passes == 2:
        w, kw["packed"] = dense.pack_bf16(w, K), True
    out = dense.gemm(a, w, passes=passes, n=N, k=K, **kw)
    torch.cuda.synchronize()
    return _stats(out, ref)


def _nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def case_conv(n, h, cin, cout, k, stride, passes, epilogue=False):
    import torch
    import torch.nn.functional as F
    from obman_train_b200 import dense
    g = torch.Generator(device="cuda").manual_seed(h + cin + cout + k + stride)
    pad = k // 2
    x = torch.randn(n, cin, h, h, device="cuda", generator=g)
    w = torch.randn(cout, cin, k, k, device="cuda", generator=g) / (cin * k * k) ** 0.5
    ref = F.conv2d(x.double(), w.double(), None, stride=stride, padding=pad)
    ho = ref.shape[2]
    wk = w.permute(0, 2, 3, 1).reshape(cout, k * k * cin).contiguous()
    dh, dw, phase, slot, step = dense.fprop_taps(k, stride, pad)
    out = torch.empty(n, ho, ho, cout, device="cuda")
    kw = {}
    if epilogue:
        bias = torch.randn(cout, device="cuda", generator=g)
        res = torch.randn(n, ho, ho, cout, device="cuda", generator=g)
        ref = torch.relu(ref + bias.double().view(1, -1, 1, 1) + res.double().permute(0, 3, 1, 2))
        kw = dict(bias=bias, addend=res, relu=True)
    if os.environ.get("PROBE_PRESPLIT", "1") == "1" and passes == 3:
        wk, kw["w_lo"] = dense.split_tf32(wk)
    if passes == 2:
        wk = dense.pack_bf16(wk)
    dense.conv_nhwc(_nhwc(x), wk, cout, (dh, dw, phase, slot), step, out, ho, ho, passes=passes, **kw)
    torch.cuda.synchronize()
    return _stats(out.permute(0, 3, 1, 2), ref)


def case_dgrad(n, h, cin, cout, k, stride, passes):
    import torch
    import torch.nn.functional as F
    from obman_train_b200 import dense
    g = torch.Generator(device="cuda").manual_seed(h + cin + cout + k + stride + 7)
    pad = k // 2
    x = torch.randn(n, cin, h, h, device="cuda", generator=g).double().requires_grad_(True)
    w = (torch.randn(cout, cin, k, k, device="cuda", generator=g) / (cout * k * k) ** 0.5)
    y = F.conv2d(x, w.double(), None, stride=stride, padding=pad)
    gy = torch.randn(y.shape, device="cuda", generator=g)
    y.backward(gy.double())
    ref = x.grad
    ho = y.shape[2]
    wt = w.permute(1, 2, 3, 0).reshape(cin, k * k * cout).contiguous()  # (C_in, KH*KW*C_out)
    gy_nhwc = _nhwc(gy)
    wt_lo = None
    if os.environ.get("PROBE_PRESPLIT", "1") == "1" and passes == 3:
        wt, wt_lo = dense.split_tf32(wt)
    if passes == 2:
        wt = dense.pack_bf16(wt)
    dx = torch.zeros(n, h, h, cin, device="cuda")
    for ph in range(stride):
        for pw in range(stride):
            dh, dw, slot = dense.dgrad_taps(k, stride, pad, (ph, pw))
            if not dh:
                continue
            dense.conv_nhwc(gy_nhwc, wt, cin, (dh, dw, None, slot), 1, dx, h // stride, h // stride,
                            out_strides=(h * h * cin, stride * h * cin, stride * cin),
                            out_offset=(ph * h + pw) * cin, passes=passes, w_slots=k * k, w_lo=wt_lo)
    torch.cuda.synchronize()
    return _stats(dx.permute(0, 3, 1, 2), ref)


def case_wgrad(n, h, cin, cout, k, stride, passes):
    import torch
    import torch.nn.functional as F
    from obman_train_b200 import dense
    g = torch.Generator(device="cuda").manual_seed(h + cin + cout + k + stride + 13)
    pad = k // 2
    x = torch.randn(n, cin, h, h, device="cuda", generator=g)
    w = torch.randn(cout, cin, k, k, device="cuda", generator=g).double().requires_grad_(True)
    y = F.conv2d(x.double(), w, None, stride=stride, padding=pad)
    gy = torch.randn(y.shape, device="cuda", generator=g)
    y.backward(gy.double())
    ref = w.grad.permute(0, 2, 3, 1).reshape(cout, k * k * cin)
    dh, dw, phase, slot, step = dense.fprop_taps(k, stride, pad)
    dwt = torch.empty(cout, k * k * cin, device="cuda")
    dense.wgrad_nhwc(_nhwc(gy), _nhwc(x), (dh, dw, phase, slot), step, dwt, passes=passes)
    torch.cuda.synchronize()
    return _stats(dwt, ref)


def case_wgrad_matrix(M, N, K, passes):
    import torch
    from obman_train_b200 import dense
    g = torch.Generator(device="cuda").manual_seed(M + N + K + 3)
    dy = torch.randn(M, N, device="cuda", generator=g)
    x = torch.randn(M, K, device="cuda", generator=g)
    ref = dy.double().t() @ x.double()
    out = dense.wgrad_matrix(dy, x, passes=passes)
    torch.cuda.synchronize()
    return _stats(out, ref)


CASES = {
    "gemm_128x128x32_p1": lambda: case_gemm(128, 128, 32, 1),
    "gemm_128x128x64_p1": lambda: case_gemm(128, 128, 64, 1),
    "gemm_256x256x512_p1": lambda: case_gemm(256, 256, 512, 1),
    "gemm_256x256x512_p3": lambda: case_gemm(256, 256, 512, 3),
    "gemm_64x33x256_p3": lambda: case_gemm(64, 33, 256, 3),
    "gemm_1000x515x515_p3_epi": lambda: case_gemm(1000, 515, 515, 3, True),
    "gemm_4096x512x2304_p3": lambda: case_gemm(4096, 512, 2304, 3),
    "gemm_20000x257x515_p1_epi": lambda: case_gemm(20000, 257, 515, 1, True),
    "conv3x3_s1_16_64_64_p1": lambda: case_conv(2, 16, 64, 64, 3, 1, 1),
    "conv3x3_s1_16_64_64_p3_epi": lambda: case_conv(2, 16, 64, 64, 3, 1, 3, True),
    "conv3x3_s1_64_64_128_p3": lambda: case_conv(3, 64, 64, 128, 3, 1, 3),
    "conv3x3_s1_8_512_512_p3": lambda: case_conv(5, 8, 512, 512, 3, 1, 3),
    "conv3x3_s2_32_64_128_p3": lambda: case_conv(2, 32, 64, 128, 3, 2, 3),
    "conv1x1_s2_32_64_128_p3": lambda: case_conv(2, 32, 64, 128, 1, 2, 3),
    "conv3x3_s1_20_36_40_p3": lambda: case_conv(2, 20, 36, 40, 3, 1, 3),
    "small_conv3x3_s1_4_512_512_p3": lambda: case_conv(4, 4, 512, 512, 3, 1, 3),
    "small_conv3x3_s1_2_512_512_p3": lambda: case_conv(2, 2, 512, 512, 3, 1, 3, True),
    "small_dgrad3x3_s1_4_512_512_p3": lambda: case_dgrad(4, 4, 512, 512, 3, 1, 3),
    "small_dgrad3x3_s1_2_512_512_p3": lambda: case_dgrad(2, 2, 512, 512, 3, 1, 3),
    "small_dgrad3x3_s2_4_256_512_p3": lambda: case_dgrad(2, 4, 256, 512, 3, 2, 3),
    "small_wgrad3x3_s1_4_512_512_p3": lambda: case_wgrad(4, 4, 512, 512, 3, 1, 3),
    "small_wgrad3x3_s1_2_512_512_p3": lambda: case_wgrad(2, 2, 512, 512, 3, 1, 3),
    "small_wgrad3x3_s2_4_256_512_p3": lambda: case_wgrad(2, 4, 256, 512, 3, 2, 3),
    "dgrad3x3_s1_16_64_64_p3": lambda: case_dgrad(2, 16, 64, 64, 3, 1, 3),
    "dgrad3x3_s2_32_64_128_p3": lambda: case_dgrad(2, 32, 64, 128, 3, 2, 3),
    "dgrad1x1_s2_32_64_128_p3": lambda: case_dgrad(2, 32, 64, 128, 1, 2, 3),
    "wgrad_matrix_4096x128x256_p1": lambda: case_wgrad_matrix(4096, 128, 256, 1),
    "wgrad_matrix_4096x128x256_p3": lambda: case_wgrad_matrix(4096, 128, 256, 3),
    "wgrad_matrix_1000x64x544_p3": lambda: case_wgrad_matrix(1000, 64, 544, 3),
    "wgrad3x3_s1_16_64_64_p3": lambda: case_wgrad(4, 16, 64, 64, 3, 1, 3),
    "wgrad3x3_s1_8_512_512_p3": lambda: case_wgrad(4, 8, 512, 512, 3, 1, 3),
    "wgrad3x3_s2_32_64_128_p3": lambda: case_wgrad(2, 32, 64, 128, 3, 2, 3),
    "wgrad1x1_s2_32_64_128_p3": lambda: case_wgrad(2, 32, 64, 128, 1, 2, 3),
}


def case_rounding_mode():
    """Does the tensor core truncate or round fp32 operands to tf32?  1-pass GEMM vs fp64 references built
    from truncated / nearest-rounded inputs."""
    import torch
    from obman_train_b200 import dense
    g = torch.Generator(device="cuda").manual_seed(1)
    a = torch.randn(128, 64, device="cuda", generator=g)
    w = torch.randn(128, 64, device="cuda", generator=g)
    out = dense.gemm(a, w, passes=1).double()

    def trunc(t):
        return (t.view(torch.int32) & ~0x1FFF).view(torch.float32)

    def rna(t):
        return ((t.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
    res = {}
    for name, f in (("trunc", trunc), ("rna", rna)):
        ref = f(a).double() @ f(w).double().t()
        res["err_vs_" + name] = (out - ref).abs().max().item()
    res.update({"max_err": min(res.values()), "rel": 0.0, "ref_max": 1.0, "nan": False, "frac_bad": 0.0})
    return res


# 3xBF16 (packed bf16 hi|lo weights, A split into tensor memory)
CASES.update({
    "gemm_128x128x32_b3": lambda: case_gemm(128, 128, 32, 2),
    "gemm_256x256x512_b3": lambda: case_gemm(256, 256, 512, 2),
    "gemm_64x33x256_b3": lambda: case_gemm(64, 33, 256, 2),
    "gemm_1000x515x515_b3_epi": lambda: case_gemm(1000, 515, 515, 2, True),
    "gemm_4096x512x2304_b3": lambda: case_gemm(4096, 512, 2304, 2),
    "gemm_20000x3x128_b3": lambda: case_gemm(20000, 3, 128, 2),
    "conv3x3_s1_16_64_64_b3_epi": lambda: case_conv(2, 16, 64, 64, 3, 1, 2, True),
    "conv3x3_s1_64_64_128_b3": lambda: case_conv(3, 64, 64, 128, 3, 1, 2),
    "conv3x3_s1_8_512_512_b3": lambda: case_conv(5, 8, 512, 512, 3, 1, 2),
    "conv3x3_s2_32_64_128_b3": lambda: case_conv(2, 32, 64, 128, 3, 2, 2),
    "conv1x1_s2_32_64_128_b3": lambda: case_conv(2, 32, 64, 128, 1, 2, 2),
    "small_conv3x3_s1_2_512_512_b3": lambda: case_conv(2, 2, 512, 512, 3, 1, 2, True),
    "conv3x3_s1_20_96_160_b3_epi": lambda: case_conv(2, 20, 96, 160, 3, 1, 2, True),
    "conv3x3_s1_16_256_256_b3": lambda: case_conv(3, 16, 256, 256, 3, 1, 2),
    "conv3x3_s1_12_64_64_b3": lambda: case_conv(5, 12, 64, 64, 3, 1, 2),
    "dgrad3x3_s1_20_96_160_b3": lambda: case_dgrad(2, 20, 96, 160, 3, 1, 2),
    "dgrad3x3_s1_8_512_512_b3": lambda: case_dgrad(3, 8, 512, 512, 3, 1, 2),
    "small_dgrad3x3_s2_4_256_512_b3": lambda: case_dgrad(2, 4, 256, 512, 3, 2, 2),
    "dgrad3x3_s1_16_64_64_b3": lambda: case_dgrad(2, 16, 64, 64, 3, 1, 2),
    "dgrad3x3_s2_32_64_128_b3": lambda: case_dgrad(2, 32, 64, 128, 3, 2, 2),
    "dgrad1x1_s2_32_64_128_b3": lambda: case_dgrad(2, 32, 64, 128, 1, 2, 2),
})
CASES.update({
    "small_wgrad3x3_s1_4_512_512_b3": lambda: case_wgrad(4, 4, 512, 512, 3, 1, 2),
    "small_wgrad3x3_s1_2_512_512_b3": lambda: case_wgrad(2, 2, 512, 512, 3, 1, 2),
    "small_wgrad3x3_s2_4_256_512_b3": lambda: case_wgrad(2, 4, 256, 512, 3, 2, 2),
    "wgrad_matrix_4096x128x256_b3": lambda: case_wgrad_matrix(4096, 128, 256, 2),
    "wgrad_matrix_1000x64x544_b3": lambda: case_wgrad_matrix(1000, 64, 544, 2),
    "wgrad_matrix_1000x32x128_b3": lambda: case_wgrad_matrix(1000, 32, 128, 2),
    "wgrad_matrix_777x544x64_b3": lambda: case_wgrad_matrix(777, 544, 64, 2),
    "wgrad3x3_s1_16_64_64_b3": lambda: case_wgrad(4, 16, 64, 64, 3, 1, 2),
    "wgrad3x3_s1_8_512_512_b3": lambda: case_wgrad(4, 8, 512, 512, 3, 1, 2),
    "wgrad3x3_s2_32_64_128_b3": lambda: case_wgrad(2, 32, 64, 128, 3, 2, 2),
    "wgrad1x1_s2_32_64_128_b3": lambda: case_wgrad(2, 32, 64, 128, 1, 2, 2),
    "stemlike_wgrad_4x4_128_32_64_b3": lambda: case_wgrad(2, 64, 32, 64, 3, 1, 2),
    "wgrad3x3_s1_32_128_128_b3": lambda: case_wgrad(3, 32, 128, 128, 3, 1, 2),
    "wgrad3x3_s1_20_96_160_b3": lambda: case_wgrad(2, 20, 96, 160, 3, 1, 2),
})
# persistent 64-output-channel kernel (conv64.cu): many tiles per CTA, ragged tiles, one K block, strided outputs
CASES.update({
    "conv3x3_s1_64_64_64_b3_epi_persistent": lambda: case_conv(40, 64, 64, 64, 3, 1, 2, True),
    "conv3x3_s1_50_64_64_b3_epi_ragged": lambda: case_conv(7, 50, 64, 64, 3, 1, 2, True),
    "conv3x3_s1_24_32_32_b3": lambda: case_conv(3, 24, 32, 32, 3, 1, 2),
    "conv3x3_s1_9_64_48_b3_epi": lambda: case_conv(5, 9, 64, 48, 3, 1, 2, True),
    "dgrad3x3_s1_64_64_64_b3_persistent": lambda: case_dgrad(20, 64, 64, 64, 3, 1, 2),
    "dgrad3x3_s2_64_64_128_b3": lambda: case_dgrad(6, 64, 64, 128, 3, 2, 2),
})
# TAIL variant of gemm_tc_kernel: N = 256 j + (1..4) on >= 148 row tiles (the decoder's 515 -> 257 -> 128 layers and their
# data gradients); the last case stays below the row-tile threshold and takes the ordinary column tiles
CASES.update({
    "gemm_tail_19000x257x515_b3_epi": lambda: case_gemm(19000, 257, 515, 2, True),
    "gemm_tail_19000x515x257_b3_epi": lambda: case_gemm(19000, 515, 257, 2, True),
    "gemm_tail_19000x257x128_b3": lambda: case_gemm(19000, 257, 128, 2),
    "gemm_tail_19001x260x96_b3_epi": lambda: case_gemm(19001, 260, 96, 2, True),
    "gemm_tail_19000x514x40_b3": lambda: case_gemm(19000, 514, 40, 2),
    "gemm_notail_3000x257x515_b3_epi": lambda: case_gemm(3000, 257, 515, 2, True),
})
# persistent kernel (gemm_persist.cu): >= 4 x 148 tiles of 128 columns, K <= 768; with and without tail columns, ragged rows
CASES.update({
    "gemm_persist_40001x257x515_b3_epi": lambda: case_gemm(40001, 257, 515, 2, True),
    "gemm_persist_40000x515x257_b3_epi": lambda: case_gemm(40000, 515, 257, 2, True),
    "gemm_persist_40000x257x128_b3": lambda: case_gemm(40000, 257, 128, 2),
    "gemm_persist_80000x128x257_b3_epi": lambda: case_gemm(80000, 128, 257, 2, True),
    "gemm_persist_77777x100x40_b3_epi": lambda: case_gemm(77777, 100, 40, 2, True),
    "gemm_persist_40000x384x768_b3": lambda: case_gemm(40000, 384, 768, 2),
})
CASES["rounding_mode_p1"] = case_rounding_mode
CASES["stemlike_wgrad_4x4_128_32_64_p3"] = lambda: case_wgrad(2, 64, 32, 64, 3, 1, 3)
CASES["wgrad3x3_s1_32_128_128_p3"] = lambda: case_wgrad(3, 32, 128, 128, 3, 1, 3)
CASES["wgrad3x3_s1_20_96_160_p3"] = lambda: case_wgrad(2, 20, 96, 160, 3, 1, 3)


def main():
    if len(sys.argv) > 1:
        name = sys.argv[1]
        try:
            res = CASES[name]()
        except Exception as e:  # noqa: BLE001
            res = {"error": repr(e)[:400]}
        print("PROBE_RESULT " + json.dumps({name: res}))
        return
    results = {}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    names = [n for n in CASES if (len(sys.argv) <= 1)]
    only = os.environ.get("PROBE_ONLY")
    if only:
        names = [n for n in CASES if only in n]
    for name in names:
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), name], capture_output=True,
                               text=True, timeout=180)
            line = [l for l in r.stdout.splitlines() if l.startswith("PROBE_RESULT ")]
            if line:
                results.update(json.loads(line[-1][len("PROBE_RESULT "):]))
            else:
                results[name] = {"error": "rc=%d " % r.returncode + (r.stdout + r.stderr)[-400:]}
        except subprocess.TimeoutExpired:
            results[name] = {"error": "timeout"}
        print(name, json.dumps(results[name]), flush=True)
        with open(os.path.join(ROOT, "gpurun_out", "probe_dense.json"), "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
