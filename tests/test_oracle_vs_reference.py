"""Pins oracle.nets / oracle.geometry against the reference's OWN code executed through oracle.refhook.
Runs only where /root/reference is mounted (the build container)."""
import numpy as np
import pytest
import torch

from oracle import geometry, icosphere, nets, refhook
from obman_train_b200.assets import load_contacts

pytestmark = pytest.mark.reference


@pytest.fixture(scope="module")
def ref(mano_tables_np):
    refhook.set_mano_tables(mano_tables_np["right"], mano_tables_np["left"])
    refhook.install()
    return True


def _tables(layer):
    return {k: v.detach().double() for k, v in layer.named_buffers() if k != "th_faces"}


def _randomise_bn(model, seed):
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
            m.weight.data = 0.5 + torch.rand(m.weight.shape, generator=g)
            m.bias.data = torch.randn(m.bias.shape, generator=g) * 0.1
            m.running_mean.data = torch.randn(m.running_mean.shape, generator=g) * 0.1
            m.running_var.data = 0.5 + torch.rand(m.running_var.shape, generator=g)


CFG = dict(resnet_version=18, mano_root="synthetic", mano_comps=30, mano_use_shape=True, mano_neurons=[1024, 256],
           mano_center_idx=0, mano_lambda_verts=0.167, mano_lambda_joints3d=0.167, mano_lambda_shape=0.167,
           mano_lambda_pose_reg=0.167, atlas_lambda=0.167, atlas_final_lambda=0.167, atlas_predict_trans=True,
           atlas_predict_scale=True, atlas_trans_weight=0.167, atlas_scale_weight=0.167,
           atlas_separate_encoder=True, atlas_ico_divisions=2, atlas_lambda_regul_edges=0.1, contact_lambda=1,
           collision_lambda=1, contact_zones="zones", contact_mode="dist_tanh", collision_mode="dist_tanh",
           contact_thresh=10, collision_thresh=20)


def make_sample(B, H, seed, device="cpu"):
    # the reference model is the consumer here: key the dict with ITS enums (served by the import hook), whatever
    # obman_train_b200.queries resolved to at import time
    from handobjectdatasets.queries import TransQueries, BaseQueries
    g = torch.Generator().manual_seed(seed)
    verts, _ = load_contacts()
    hand = torch.tensor(verts * 1000, dtype=torch.float32).unsqueeze(0).repeat(B, 1, 1)
    sample = {
        TransQueries.images: torch.rand(B, 3, H, H, generator=g) - 0.5,
        BaseQueries.sides: ["right" if i % 2 == 0 else "left" for i in range(B)],
        "root": "wrist",
        TransQueries.joints3d: torch.randn(B, 21, 3, generator=g) * 40,
        TransQueries.verts3d: hand + torch.randn(B, 778, 3, generator=g) * 5,
        TransQueries.objpoints3d: torch.randn(B, 600, 3, generator=g) * 40 + 30,
    }
    return {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in sample.items()}


def plain_sample(sample):
    return {getattr(k, "value", k).strip(): v for k, v in sample.items()}


def test_handnet_oracle_matches_reference_end_to_end(ref, mano_tables_np):
    from mano_train.networks.handnet import HandNet
    torch.manual_seed(0)
    with refhook.cwd():
        model = HandNet(**{k: v for k, v in CFG.items()})
    _randomise_bn(model, 3)
    model.eval()
    sample = make_sample(3, 64, 5)
    with refhook.cwd():
        total, results, losses = model.forward(dict(sample))
    total.backward()
    state = {k: v.detach().double().clone() for k, v in model.state_dict().items()}
    for k, v in state.items():
        if v.is_floating_point() and "running_" not in k and "th_" not in k:
            v.requires_grad_(True)
    tables = {"right": _tables(model.mano_branch.mano_layer_right), "left": _tables(model.mano_branch.mano_layer_left)}
    grid, faces = icosphere.icosphere(2)
    _, zones = load_contacts()
    s64 = {k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in plain_sample(sample).items()}
    ototal, oresults, olosses = nets.handnet_forward(state, CFG, s64, tables, torch.tensor(grid), faces, zones)
    ototal.backward()
    assert abs(total.item() - ototal.item()) < 1e-4 * abs(ototal.item())
    for key in ("mano_verts3d", "mano_joints3d", "mano_shape", "pose_reg", "atlas_trans3d", "atlas_scale3d",
                "final_chamfer_loss", "atlas_objpoints3d", "atlas_edge_regul", "penetration_loss",
                "attraction_loss", "max_penetr", "mean_penetr"):
        assert abs(float(losses[key]) - float(olosses[key])) <= 1e-4 * abs(float(olosses[key])) + 1e-7, key
    np.testing.assert_allclose(results["verts"].detach().numpy(), oresults["verts"].detach().numpy(), rtol=1e-3, atol=1e-2)
    np.testing.assert_allclose(results["objpoints3d"].detach().numpy(), oresults["objpoints3d"].detach().numpy(), rtol=1e-3, atol=1e-2)
    # the in-place "+=" aliasing quirk of the reference
    assert float(losses["mano_total_loss"]) == float(total)
    worst = 0.0
    for name, p in model.named_parameters():
        if p.grad is None:
            assert state[name].grad is None or state[name].grad.abs().max() == 0, name
            continue
        g = state[name].grad
        rel = (p.grad.double() - g).abs().max().item() / (g.abs().max().item() + 1e-12)
        worst = max(worst, rel)
    assert worst < 5e-3, worst


def test_chamfer_and_exterior_random_vs_reference(ref):
    from mano_train.networks.branches import atlasutils, contactutils
    g = torch.Generator().manual_seed(9)
    p = torch.randn(4, 200, 3, generator=g) * 40
    t = torch.randn(4, 150, 3, generator=g) * 40 + 10
    l1, l2 = atlasutils.ChamferLoss()(p, t)
    o1, o2 = geometry.chamfer(p.double(), t.double())
    np.testing.assert_allclose(l1.numpy(), o1.numpy(), rtol=2e-5)
    np.testing.assert_allclose(l2.numpy(), o2.numpy(), rtol=2e-5)
    v, f = icosphere.icosphere(2)
    obj = torch.tensor(v, dtype=torch.float32).unsqueeze(0) * 30 + torch.randn(2, v.shape[0], 3, generator=g)
    pts = torch.randn(2, 300, 3, generator=g) * 25
    ext = contactutils.batch_mesh_contains_points(pts, obj[:, torch.tensor(f)])
    oext, margin = geometry.mesh_exterior(pts.double(), obj.double()[:, torch.tensor(f)], return_margin=True)
    assert ((ext == oext) | (margin < 1e-5)).all()
