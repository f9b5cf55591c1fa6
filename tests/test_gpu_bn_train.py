"""BatchNorm with batch statistics (model.train(): training without --freeze_batchnorm, epochpass3d.py:48-52) on the
ResNet-18 encoder: features, every parameter gradient and the running-statistics update against the fp64 oracle
(oracle.nets.resnet18_features(bn_training=True) = torch.nn.functional.batch_norm in training mode)."""
import pytest
import torch

from oracle import nets

pytestmark = pytest.mark.gpu


def _stats_reference(state, images, momentum=0.1):
    """Running statistics after one training-mode forward, from torchvision-equivalent modules in fp64."""
    from obman_train_b200.networks.bases.resnet import resnet18
    ref = resnet18().double()
    ref.load_state_dict(state)
    ref.train()
    # the product module's forward needs CUDA; run the same layers through torch ops instead
    x = images.double()
    x = ref.relu(ref.bn1(ref.conv1(x)))
    x = ref.maxpool(x)
    for layer in (ref.layer1, ref.layer2, ref.layer3, ref.layer4):
        for blk in layer:
            idt = x
            out = blk.relu(blk.bn1(blk.conv1(x)))
            out = blk.bn2(blk.conv2(out))
            if blk.downsample is not None:
                idt = blk.downsample(x)
            x = blk.relu(out + idt)
    return {k: v for k, v in ref.state_dict().items() if "running_" in k}


@pytest.mark.parametrize("B,H", [(4, 64), (3, 128)])
def test_resnet18_train_mode_batchnorm_vs_oracle(B, H):
    from obman_train_b200.networks.bases.resnet import resnet18
    torch.manual_seed(0)
    model = resnet18()
    g = torch.Generator().manual_seed(3)
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.weight.data = 0.5 + torch.rand(m.weight.shape, generator=g)
            m.bias.data = torch.randn(m.bias.shape, generator=g) * 0.1
    model.train()
    state0 = {k: v.detach().double().clone() for k, v in model.state_dict().items()}
    state64 = {"base_net." + k: v.clone() for k, v in state0.items()}
    for k, v in state64.items():
        if v.is_floating_point() and "running_" not in k:
            v.requires_grad_(True)
    images = torch.rand(B, 3, H, H, generator=g) - 0.5
    wts = torch.randn(B, 512, generator=g)
    ref = nets.resnet18_features(state64, images.double(), "base_net", True)
    (ref * wts.double()).sum().backward()
    ref_stats = _stats_reference(state0, images)

    model = model.cuda()
    feats, extra = model(images.cuda())
    assert extra == {}
    (feats * wts.cuda()).sum().backward()
    scale = ref.abs().max().item()
    err = (feats.detach().cpu().double() - ref.detach()).abs().max().item()
    assert err < 2e-4 * scale, (err, scale)
    l2s = []
    for name, p in model.named_parameters():
        if name.startswith("fc."):
            assert p.grad is None
            continue
        gref = state64["base_net." + name].grad
        d = p.grad.cpu().double() - gref
        l2s.append((d.norm().item() / (gref.norm().item() + 1e-30), name))
    l2s.sort(reverse=True)
    print("train-mode BN: features rel err %.2e; worst gradients (L2 rel): %s" % (
        err / scale, ["%s %.2e" % (n, r) for r, n in l2s[:4]]))
    # norm-wise bound as in tests/test_gpu_encoder.py (a ReLU pre-activation within rounding of zero may take the other
    # branch than in fp64; batch statistics couple all pixels of a channel, so a flip is diluted, not amplified)
    assert l2s[0][0] < 2e-2, l2s[:5]
    sd = model.state_dict()
    for k, v in ref_stats.items():
        if "num_batches" in k:
            continue
        got = sd[k].cpu().double()
        assert (got - v).abs().max().item() <= 1e-4 * v.abs().max().item() + 1e-6, k


def test_eval_mode_still_uses_running_statistics():
    from obman_train_b200.networks.bases.resnet import resnet18
    torch.manual_seed(1)
    model = resnet18().cuda()
    x = torch.rand(2, 3, 64, 64, device="cuda") - 0.5
    model.eval()
    before = {k: v.clone() for k, v in model.state_dict().items() if "running_" in k}
    f_eval, _ = model(x)
    assert all(torch.equal(v, model.state_dict()[k]) for k, v in before.items())
    model.train()
    f_train, _ = model(x)
    assert not torch.allclose(f_eval, f_train)
    assert any(not torch.equal(v, model.state_dict()[k]) for k, v in before.items())


def test_handnet_in_training_mode_matches_oracle():
    """The whole graph with every BatchNorm (both encoders, the three BatchNorm1d of the AtlasNet decoder) on batch
    statistics: losses, vertices and parameter gradients against oracle.nets.handnet_forward(bn_training=True)."""
    import numpy as np
    from obman_train_b200.assets import load_contacts
    from obman_train_b200.networks.handnet import HandNet
    from tests.util import FULL_CFG, enum_sample, make_sample
    cfg = dict(FULL_CFG)
    torch.manual_seed(2)
    model = HandNet(**cfg)
    model.train()
    B, H = 6, 128
    sample = make_sample(B, H, 5)
    state = {k: v.detach().double().clone() for k, v in model.state_dict().items()}
    for k, v in state.items():
        if v.is_floating_point() and "running_" not in k and "th_" not in k:
            v.requires_grad_(True)
    tables = {s: {k: v.detach().cpu().double() for k, v in getattr(model.mano_branch, "mano_layer_" + s).named_buffers()
                  if k != "th_faces"} for s in ("right", "left")}
    grid = model.atlas_branch.test_verts.double()
    _, zones = load_contacts()
    s64 = {k: (v.double() if torch.is_tensor(v) else v) for k, v in sample.items()}
    ototal, oresults, olosses = nets.handnet_forward(state, cfg, s64, tables, grid, model.atlas_branch.test_faces, zones,
                                                     bn_training=True)
    ototal.backward()
    model = model.cuda()
    total, results, losses = model.forward(enum_sample(sample))
    total.backward()
    # Tolerance 5e-4 (eval mode: 1e-4): the deepest BatchNorm layers normalise with statistics of only B * 4 * 4 = 96
    # values per channel, which amplifies the 1e-5-level differences of the convolution outputs.
    assert abs(total.item() - ototal.item()) < 5e-4 * abs(ototal.item()), (total.item(), ototal.item())
    for key in ("verts", "joints", "objpointscentered3d", "objtrans", "objscale", "objpoints3d"):
        o = oresults[key].detach().numpy()
        got = results[key].detach().cpu().numpy()
        rel = np.abs(got - o).max() / np.abs(o).max()
        print("  %-22s rel err %.2e" % (key, rel))
        assert rel < 5e-4, (key, rel)
    rels = []
    for name, p in model.named_parameters():
        og = state[name].grad
        if p.grad is None:
            assert og is None or og.abs().max() == 0, name
            continue
        d = p.grad.cpu().double() - og
        if og.norm().item() < 1e-9:   # biases in front of a batch-statistics BatchNorm: exact gradient zero
            wgrad = dict(model.named_parameters())[name.replace("bias", "weight")].grad
            assert p.grad.abs().max().item() <= 1e-4 * wgrad.abs().max().item() + 1e-3, (name, p.grad.abs().max().item())
            continue
        rels.append((d.norm().item() / (og.norm().item() + 1e-12), name))
    rels.sort(reverse=True)
    print("train-mode HandNet: total %.6f (oracle %.6f); worst gradients: %s" % (
        total.item(), ototal.item(), ["%s %.2e" % (n, r) for r, n in rels[:5]]))
    assert rels[0][0] < 5e-2, rels[:5]
    # the decoder's running statistics moved (BatchNorm1d, momentum 0.1)
    assert not torch.equal(model.atlas_branch.decoder.bn1.running_mean.cpu(), torch.zeros(515))


@pytest.mark.parametrize("B,N", [(4, 162), (2, 642)])
def test_point_decoder_train_mode_vs_oracle(B, N):
    """PointGenCon with BatchNorm1d on batch statistics (atlasutils.py:65-75 under model.train()) against the fp64 oracle:
    vertices, every parameter gradient, the feature gradient and the running-statistics update."""
    from oracle import icosphere
    from obman_train_b200.networks.branches.atlasutils import PointGenCon
    torch.manual_seed(4)
    dec = PointGenCon(bottleneck_size=515, out_factor=200)
    g = torch.Generator().manual_seed(6)
    for bn in (dec.bn1, dec.bn2, dec.bn3):
        bn.weight.data = 0.5 + torch.rand(bn.weight.shape, generator=g)
        bn.bias.data = torch.randn(bn.bias.shape, generator=g) * 0.1
    dec.train()
    grid = torch.tensor(icosphere.icosphere(2 if N == 162 else 3)[0], dtype=torch.float32)
    feat = torch.randn(B, 512, generator=g)
    wts = torch.randn(B, N, 3, generator=g)
    state = {"d." + k: v.detach().double().clone() for k, v in dec.state_dict().items()}
    for k, v in state.items():
        if v.is_floating_point() and "running_" not in k:
            v.requires_grad_(True)
    f64 = feat.double().requires_grad_(True)
    x = torch.cat([grid.double().t().unsqueeze(0).expand(B, -1, -1), f64.unsqueeze(2).expand(-1, -1, N)], 1)
    ref = nets.point_decoder(state, x, "d", True, 200).transpose(2, 1)
    (ref * wts.double()).sum().backward()
    dec = dec.cuda()
    fc = feat.cuda().requires_grad_(True)
    out = dec.decode(fc, grid.cuda())
    (out * wts.cuda()).sum().backward()
    rel = ((out.detach().cpu().double() - ref.detach()).abs().max() / ref.abs().max()).item()
    print("decoder (train-mode BN) output rel err %.2e" % rel)
    assert rel < 1e-4
    worst = []
    for name, p in dec.named_parameters():
        gref = state["d." + name].grad.reshape(p.shape)
        if name in ("conv1.bias", "conv2.bias", "conv3.bias"):
            # a bias in front of a batch-statistics BatchNorm has NO effect on the output: its exact gradient is zero
            # (sum of dz over the batch), the oracle's is ~1e-20; ours must be rounding noise of that sum
            scale = dec.get_parameter(name.replace("bias", "weight")).grad.abs().max().item()
            assert p.grad.abs().max().item() <= 1e-4 * scale + 1e-6, (name, p.grad.abs().max().item(), scale)
            assert gref.abs().max().item() < 1e-9
            continue
        worst.append((((p.grad.cpu().double() - gref).norm() / (gref.norm() + 1e-30)).item(), name))
    worst.append((((fc.grad.cpu().double() - f64.grad).norm() / f64.grad.norm()).item(), "features"))
    worst.sort(reverse=True)
    print("worst gradients:", ["%s %.2e" % (n, r) for r, n in worst[:4]])
    assert worst[0][0] < 1e-2, worst[:4]
    # running statistics: momentum 0.1, unbiased variance
    z = torch.nn.functional.conv1d(x.detach(), state["d.conv1.weight"].detach(), state["d.conv1.bias"].detach())
    m_ref = 0.1 * z.mean((0, 2))
    v_ref = 0.9 + 0.1 * z.transpose(0, 1).reshape(515, -1).var(1, unbiased=True)
    assert ((dec.bn1.running_mean.cpu().double() - m_ref).abs().max() <= 1e-4 * m_ref.abs().max() + 1e-6)
    assert ((dec.bn1.running_var.cpu().double() - v_ref).abs().max() <= 1e-4 * v_ref.abs().max())
