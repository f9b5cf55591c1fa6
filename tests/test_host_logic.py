"""CPU checks of the host-side contracts the CUDA convolution path is driven by (no GPU, no library calls):
the tap tables of ``dense.fprop_taps`` / ``dense.dgrad_taps`` and the weight / stem layouts documented in
include/obman_b200.h, executed by a small torch emulator of ``obman_conv_nhwc``'s semantics (shifted-box taps over
NHWC phase views, zero fill outside) and compared with ``torch.nn.functional.conv2d`` and its autograd."""
import pytest
import torch
import torch.nn.functional as F

from obman_train_b200 import dense
from obman_train_b200.encoder import resnet18_units, stem_view


def emu_conv_nhwc(x, w, c_out, taps, in_step, h_out, w_out):
    """out[n,h,w,:] = sum_t xview_t[n, h+dh[t], w+dw[t], :] @ w[:, slot[t]*C:(slot[t]+1)*C]^T (header contract)."""
    dh, dw, phase, slot = taps
    n, _, _, c = x.shape
    out = torch.zeros(n, h_out, w_out, c_out, dtype=x.dtype)
    for t in range(len(dh)):
        ph = phase[t] if (phase is not None and in_step == 2) else 0
        view = x[:, (ph >> 1)::in_step, (ph & 1)::in_step]
        hv, wv = view.shape[1], view.shape[2]
        shifted = torch.zeros(n, h_out, w_out, c, dtype=x.dtype)
        h_lo, h_hi = max(0, -dh[t]), min(h_out, hv - dh[t])
        w_lo, w_hi = max(0, -dw[t]), min(w_out, wv - dw[t])
        if h_lo < h_hi and w_lo < w_hi:
            shifted[:, h_lo:h_hi, w_lo:w_hi] = view[:, h_lo + dh[t]:h_hi + dh[t], w_lo + dw[t]:w_hi + dw[t]]
        out += shifted @ w[:, slot[t] * c:(slot[t] + 1) * c].t()
    return out


def fold_layouts(w):
    """(O,I,KH,KW) -> fprop operand (O, KH*KW*I) and dgrad operand (I, KH*KW*O) of obman_fold_conv (scale 1)."""
    O, I, KH, KW = w.shape
    wf = w.permute(0, 2, 3, 1).reshape(O, KH * KW * I)
    wft = w.permute(1, 2, 3, 0).reshape(I, KH * KW * O)
    return wf, wft


@pytest.mark.parametrize("k,stride", [(3, 1), (3, 2), (1, 2), (1, 1)])
def test_fprop_and_dgrad_tap_tables_reproduce_conv2d_and_its_gradient(k, stride):
    g = torch.Generator().manual_seed(k * 10 + stride)
    N, H, C, O = 2, 8, 5, 7
    x = torch.randn(N, H, H, C, generator=g, dtype=torch.float64)
    w = torch.randn(O, C, k, k, generator=g, dtype=torch.float64)
    wf, wft = fold_layouts(w)
    ho = H // stride
    dh, dw, phase, slot, step = dense.fprop_taps(k, stride, k // 2)
    assert step == stride and len(dh) == k * k and slot == list(range(k * k))
    y = emu_conv_nhwc(x, wf, O, (dh, dw, phase, slot), step, ho, ho)
    xr = x.permute(0, 3, 1, 2).clone().requires_grad_(True)
    ref = F.conv2d(xr, w, stride=stride, padding=k // 2)
    assert torch.allclose(y.permute(0, 3, 1, 2), ref, atol=1e-10)
    # data gradient: one emulated launch per output phase, written interleaved (encoder._Unit.dgrad)
    gy = torch.randn(N, ho, ho, O, generator=g, dtype=torch.float64)
    ref.backward(gy.permute(0, 3, 1, 2))
    gx = torch.zeros(N, H, H, C, dtype=torch.float64)
    covered = 0
    for ph in range(stride):
        for pw in range(stride):
            tdh, tdw, tslot = dense.dgrad_taps(k, stride, k // 2, (ph, pw))
            covered += len(tdh)
            if not tdh:
                continue
            gx[:, ph::stride, pw::stride] = emu_conv_nhwc(gy, wft, C, (tdh, tdw, None, tslot), 1, H // stride, H // stride)
    assert covered == k * k          # every filter tap belongs to exactly one output phase
    assert torch.allclose(gx.permute(0, 3, 1, 2), xr.grad, atol=1e-10)


def emu_stem_pack(img):
    """obman_stem_pack: (B,3,H,W) -> (B, H/2, W/2 + 4, 16), channel (ph*2+pw)*3 + c, two zero pixels either side."""
    B, _, H, W = img.shape
    out = torch.zeros(B, H // 2, W // 2 + 4, 16, dtype=img.dtype)
    for ph in range(2):
        for pw in range(2):
            for c in range(3):
                out[:, :, 2:-2, (ph * 2 + pw) * 3 + c] = img[:, c, ph::2, pw::2]
    return out


def stem_weight_layout(w):
    """obman_fold_conv(stem=1): (O,3,7,7) -> (O, 4*64); slot = a + 2 for the vertical tap a in [-2, 1],
    channel q*16 + (ph*2+pw)*3 + c  <->  kh = 2a + ph + 3, kw = 2(q-2) + pw + 3 (zero outside the 7x7 window)."""
    O = w.shape[0]
    out = torch.zeros(O, 256, dtype=w.dtype)
    for slot in range(4):
        for q in range(4):
            for ph in range(2):
                for pw in range(2):
                    kh, kw = 2 * (slot - 2) + ph + 3, 2 * (q - 2) + pw + 3
                    if 0 <= kh < 7 and 0 <= kw < 7:
                        for c in range(3):
                            out[:, slot * 64 + q * 16 + (ph * 2 + pw) * 3 + c] = w[:, c, kh, kw]
    return out


def test_stem_as_four_tap_convolution_over_the_overlapping_view():
    g = torch.Generator().manual_seed(3)
    B, H, O = 2, 16, 6
    img = torch.randn(B, 3, H, H, generator=g, dtype=torch.float64)
    w = torch.randn(O, 3, 7, 7, generator=g, dtype=torch.float64)
    xs = emu_stem_pack(img)
    n, ho, wo, c, sN, sH, sW = stem_view(xs)
    assert (n, ho, wo, c, sN, sH, sW) == (B, H // 2, H // 2, 64, (H // 2) * (H // 2 + 4) * 16, (H // 2 + 4) * 16, 16)
    # materialise what the overlapping TMA view reads: pixel j = 4 neighbouring 16-channel pixels j .. j+3 (padded)
    view = torch.as_strided(xs, (n, ho, wo, c), (sN, sH, sW, 1))
    assert torch.equal(view[:, :, 5, 16:32], xs[:, :, 6, :])
    taps = ([-2, -1, 0, 1], [0, 0, 0, 0], [0] * 4, [0, 1, 2, 3])     # encoder._Unit(stem=True).taps
    y = emu_conv_nhwc(view.contiguous(), stem_weight_layout(w), O, taps, 1, ho, wo)
    ref = F.conv2d(img, w, stride=2, padding=3)
    assert torch.allclose(y.permute(0, 3, 1, 2), ref, atol=1e-10)


def test_resnet18_unit_table_matches_the_reference_state_dict_order():
    """20 conv+BN units in the order their parameters appear in ResNet-18's state dict (resnet.py:99-152)."""
    from obman_train_b200.networks.bases import resnet
    units = resnet18_units()
    assert len(units) == 20 and units[0] == ("conv1", "bn1", 64, 3, 7, 2)
    model = resnet.resnet18(pretrained=False)
    names = [n for n, _ in model.named_parameters() if not n.startswith("fc.")]
    expected = []
    for conv, bn, O, I, k, s in units:
        expected += [conv + ".weight", bn + ".weight", bn + ".bias"]
        wshape = tuple(dict(model.named_parameters())[conv + ".weight"].shape)
        assert wshape == (O, I, k, k), (conv, wshape)
    assert names == expected
    # 2.369 GMAC forward per 256x256 image (fc excluded)
    macs = 128 * 128 * 64 * 3 * 7 * 7          # stem: 256 -> 128
    size = {"layer1": 64, "layer2": 32, "layer3": 16, "layer4": 8}   # output size of every unit of a stage
    for conv, _, O, I, k, s in units[1:]:
        h_out = size[conv.split(".")[0]]
        macs += h_out * h_out * O * I * k * k
    assert abs(macs / 1e9 - 2.369) < 0.005, macs / 1e9


def test_streams_module_is_a_no_op_when_disabled():
    """OBMAN_OVERLAP=0 / streams.set_enabled(False): fork / join / on_aux must not touch CUDA at all (CPU-safe)."""
    from obman_train_b200 import streams
    prev = streams.set_enabled(False)
    try:
        assert streams.enabled() is False
        streams.fork()
        streams.fork(streams.CHAIN)
        with streams.on_aux(streams.BRANCH):
            x = torch.ones(3) * 2
        streams.join(streams.BRANCH)
        streams.join()
        assert x.sum().item() == 6.0
        assert (streams.WGRAD, streams.CHAIN, streams.BRANCH) == (0, 1, 2)
    finally:
        assert streams.set_enabled(prev) is False


def test_loss_weights_host_mirror_and_slots():
    """losshead.LossWeights: named slots, None / 0 stored as 0, host values readable without a device."""
    from obman_train_b200.losshead import LossWeights
    w = LossWeights(["a", "b", "one"])
    w["a"], w["b"], w["one"] = 0.167, None, 1
    assert w.slot == {"a": 0, "b": 1, "one": 2}
    assert w["a"] == pytest.approx(0.167) and w["b"] == 0.0 and w["one"] == 1.0
    w["a"] = 0.167 * 0.5
    assert w.host[0] == pytest.approx(0.0835)
    with pytest.raises(KeyError):
        w["missing"] = 1.0


def test_bench_workloads_follow_baseline_configs():
    """bench.Workload: --config k = BASELINE.json configs[k-1]; default = configs[2] on one GPU, configs[3] under torchrun."""
    import json
    import os
    import bench
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    configs = json.load(open(os.path.join(root, "BASELINE.json")))["configs"]
    assert bench.default_config(1) == 3 and bench.default_config(8) == 4
    w1, w2, w3, w4 = (bench.Workload(k) for k in (1, 2, 3, 4))
    assert w1.hand_only and w1.batch == 1 and "single 256" in configs[0]
    assert (w2.batch, w2.n_obj, w2.n_gt) == (64, 642, 600) and "batch 64" in configs[1]
    assert (w3.batch, w3.n_obj, w3.n_gt) == (256, 2562, 2500) and "batch 256" in configs[2]
    assert w3.cfg["atlas_separate_encoder"] and w3.cfg["contact_zones"] == "zones" and w3.cfg["contact_lambda"]
    assert w4.batch == 128 and w4.cfg["atlas_lambda_regul_edges"] > 0 and "batch 1024" in configs[3]
    for k, name in ((2, "configs[1]"), (3, "configs[2]"), (4, "configs[3]")):
        assert name in bench.Workload(k).name
    s = w3.sample(2, 0)
    assert s["images"].shape == (2, 3, 256, 256) and s["objpoints3d"].shape == (2, 2500, 3) and s["verts3d"].shape == (2, 778, 3)
    assert sorted(w1.sample(1, 0).keys()) == ["images", "joints3d", "root", "sides"]


def test_bench_traffic_record_goes_stale_with_the_kernel_sources(tmp_path, monkeypatch):
    """roofline.traffic is only reported while csrc/ hashes to the profiled build (it used to be a constant)."""
    import json
    import os
    import bench
    prof = tmp_path / "profiles"
    prof.mkdir()
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    monkeypatch.setattr(bench, "kernel_source_hash", lambda: "abc")
    rec = {"source": "x", "precision": "bf16x3", "kernel_source_sha256": "abc", "dram_bytes_per_launch": 1.0}
    (prof / "gemm_traffic_current.json").write_text(json.dumps(rec))
    tr, src = bench.traffic_record("bf16x3")
    assert tr["dram_bytes_per_launch"] == 1.0 and src == "x"
    rec["kernel_source_sha256"] = "other"
    (prof / "gemm_traffic_current.json").write_text(json.dumps(rec))
    tr, src = bench.traffic_record("bf16x3")
    assert tr is None and "stale" in src
    assert bench.traffic_record("tf32") == (None, None) or bench.traffic_record("tf32")[0] is None


def test_ray_triangle_predicate_folding_is_exact_in_fp32():
    """contact.cu's packed ray / triangle kernels test  min(u, v, 1 - (u + v)) > 0  instead of the reference's
    (u > 0) & (u < 1) & (v > 0) & (u + v < 1)  (contactutils.py:117-127).  In IEEE fp32 the two are the same predicate for
    all finite u, v: 1 - s > 0 <=> s < 1 exactly (a difference of two floats is zero only if they are equal and has the
    right sign), and u < 1 follows from v > 0 and fl(u + v) < 1 because rounding is monotonic.  Checked here on random
    values, on values within a few ulps of the boundaries and on the boundaries themselves."""
    import numpy as np
    rng = np.random.default_rng(3)
    one = np.float32(1.0)
    parts = [rng.uniform(-0.5, 1.5, 400000).astype(np.float32)]
    edge = np.array([0.0, 1.0, 0.5, 1e-30, -1e-30, 1e-7, 1.0 - 2.0 ** -24, 1.0 + 2.0 ** -23], dtype=np.float32)
    for base in edge:   # the value itself and its 8 neighbours on either side
        parts.append(np.array([base], np.float32))
        lo, hi = parts[-1].copy(), parts[-1].copy()
        for _ in range(8):
            lo = np.nextafter(lo, np.float32(-np.inf))
            hi = np.nextafter(hi, np.float32(np.inf))
            parts.append(lo.copy())
            parts.append(hi.copy())
    vals = np.concatenate(parts).astype(np.float32)
    special = vals[-(len(vals) - 400000):]
    u = np.concatenate([vals[:200000], np.repeat(special, len(special)), rng.choice(special, 50000)])
    v = np.concatenate([vals[200000:400000], np.tile(special, len(special)), rng.uniform(-0.5, 1.5, 50000).astype(np.float32)])
    # also pairs that sum to (almost) exactly one
    u2 = rng.uniform(0, 1, 100000).astype(np.float32)
    v2 = (one - u2).astype(np.float32)
    for k in range(3):
        u = np.concatenate([u, u2])
        v = np.concatenate([v, np.nextafter(v2, np.float32(np.inf)) if k == 1 else (np.nextafter(v2, np.float32(-np.inf)) if k == 2 else v2)])
    s = (u + v).astype(np.float32)
    reference = (u > 0) & (u < 1) & (v > 0) & (s < 1)
    w = (one - s).astype(np.float32)
    folded = np.minimum(np.minimum(u, v), w) > 0
    assert u.dtype == np.float32 and w.dtype == np.float32
    assert np.array_equal(reference, folded), int((reference != folded).sum())
    assert reference.any() and (~reference).any()
