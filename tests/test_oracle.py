"""CPU checks of the oracle itself: golden vectors made by the reference, brute-force restatements,
and self-consistency of the (unpinned) MANO restatement."""
import numpy as np
import pytest
import torch

from oracle import geometry, icosphere, mano
from obman_train_b200.assets import load_contacts


def test_icosphere_counts():
    for sub, (nv, nf) in {0: (12, 20), 1: (42, 80), 2: (162, 320), 3: (642, 1280), 4: (2562, 5120)}.items():
        v, f = icosphere.icosphere(sub)
        assert v.shape == (nv, 3) and f.shape == (nf, 3)
        assert np.allclose(np.linalg.norm(v, axis=1), 1.0)
        assert f.min() == 0 and f.max() == nv - 1
        # closed manifold: every edge shared by exactly two faces
        e = np.sort(np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]]), axis=1)
        _, counts = np.unique(e, axis=0, return_counts=True)
        assert (counts == 2).all()


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_chamfer_matches_reference_golden(golden, tag):
    preds = torch.tensor(golden["chamfer_%s_preds" % tag], dtype=torch.float64, requires_grad=True)
    gts = torch.tensor(golden["chamfer_%s_gts" % tag], dtype=torch.float64)
    l1, l2 = geometry.chamfer(preds, gts)
    (l1 + l2).mean().backward()
    # the reference computes in fp32 with the expansion formula: loss scalars agree to ~1e-6 relative
    np.testing.assert_allclose(l1.detach().numpy(), golden["chamfer_%s_loss1" % tag], rtol=2e-5)
    np.testing.assert_allclose(l2.detach().numpy(), golden["chamfer_%s_loss2" % tag], rtol=2e-5)
    np.testing.assert_allclose(preds.grad.numpy(), golden["chamfer_%s_gpreds" % tag], rtol=1e-3, atol=1e-3)


def test_chamfer_bruteforce_loops():
    rng = np.random.RandomState(0)
    p, g = rng.randn(1, 7, 3), rng.randn(1, 5, 3)
    l1, l2 = geometry.chamfer(torch.tensor(p), torch.tensor(g))
    b1 = np.mean([min(((p[0, j] - g[0, i]) ** 2).sum() for i in range(5)) for j in range(7)])
    b2 = np.mean([min(((p[0, j] - g[0, i]) ** 2).sum() for j in range(7)) for i in range(5)])
    assert abs(l1.item() - b1) < 1e-12 and abs(l2.item() - b2) < 1e-12


def test_exterior_matches_reference_golden_and_sphere_truth(golden):
    hand = torch.tensor(golden["contact_hand"], dtype=torch.float64)
    obj = torch.tensor(golden["contact_obj"], dtype=torch.float64)
    faces = torch.tensor(golden["contact_faces"])
    ext = geometry.mesh_exterior(hand, obj[:, faces])
    assert (ext.numpy() == golden["contact_exterior"]).all()
    # analytic truth on an exact sphere mesh
    v, f = icosphere.icosphere(3)
    rng = np.random.RandomState(1)
    pts = rng.randn(1, 500, 3)
    pts = pts / np.linalg.norm(pts, axis=2, keepdims=True) * rng.uniform(0.2, 1.8, (1, 500, 1))
    r = np.linalg.norm(pts, axis=2)
    keep = (r < 0.97) | (r > 1.0)  # inside the inscribed radius of the faceted sphere, or outside
    ext = geometry.mesh_exterior(torch.tensor(pts), torch.tensor(v)[torch.tensor(f)].unsqueeze(0))
    assert ((ext.numpy() == (r > 1.0)) | ~keep).all()


@pytest.mark.parametrize("zones_mode", ["all", "tips", "zones"])
@pytest.mark.parametrize("mode", ["dist_sq", "dist", "dist_tanh"])
def test_contact_loss_matches_reference_golden(golden, zones_mode, mode):
    _, zones = load_contacts()
    targets = ("all", "obj", "hand") if (zones_mode == "zones" and mode == "dist_tanh") else ("all",)
    for target in targets:
        k = "contact_%s_%s_%s_" % (zones_mode, mode, target)
        hand = torch.tensor(golden["contact_hand"], dtype=torch.float64, requires_grad=True)
        obj = torch.tensor(golden["contact_obj"], dtype=torch.float64, requires_grad=True)
        missed, penetr, info, metrics = geometry.contact_loss(
            hand, obj, golden["contact_faces"], zones, contact_thresh=10, contact_mode=mode,
            collision_thresh=20, collision_mode=mode, contact_target=target, contact_zones=zones_mode)
        assert (info["attraction_masks"].numpy() == golden[k + "attr"].astype(bool)).all()
        assert (info["repulsion_masks"].numpy() == golden[k + "rep"].astype(bool)).all()
        np.testing.assert_allclose(missed.detach().numpy().reshape(-1), golden[k + "missed"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(penetr.detach().numpy().reshape(-1), golden[k + "penetr"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(metrics["max_penetr"].item(), golden[k + "max_penetr"][0], rtol=1e-5)
        np.testing.assert_allclose(metrics["mean_penetr"].item(), golden[k + "mean_penetr"][0], rtol=1e-5)
        total = (missed + 0.5 * penetr).sum()
        if total.requires_grad:
            total.backward()
        gh = hand.grad.numpy() if hand.grad is not None else np.zeros(hand.shape)
        go = obj.grad.numpy() if obj.grad is not None else np.zeros(obj.shape)
        np.testing.assert_allclose(gh, golden[k + "ghand"], rtol=1e-3, atol=1e-6)
        np.testing.assert_allclose(go, golden[k + "gobj"], rtol=1e-3, atol=1e-6)


def test_edge_loss_golden(golden):
    obj = torch.tensor(golden["contact_obj"], dtype=torch.float64)
    np.testing.assert_allclose(geometry.edge_loss(obj, golden["contact_faces"]).item(),
                               golden["edge_loss"][0], rtol=1e-5)


# ---- MANO restatement: self-consistency only (parity unpinned, see oracle/mano.py) ----------------------
def _tables64(t, ncomps=30, flat=True):
    f = lambda a: torch.tensor(np.asarray(a), dtype=torch.float64)  # noqa: E731
    return {
        "th_shapedirs": f(t["shapedirs"]), "th_posedirs": f(t["posedirs"]),
        "th_v_template": f(t["v_template"]).unsqueeze(0), "th_J_regressor": f(t["J_regressor"]),
        "th_weights": f(t["weights"]), "th_betas": f(t["betas"]).view(1, 10),
        "th_hands_mean": (torch.zeros(1, 45, dtype=torch.float64) if flat else f(t["hands_mean"]).view(1, 45)),
        "th_selected_comps": f(t["hands_components"][:ncomps]),
    }


def test_mano_zero_pose_is_template(mano_tables_np):
    T = _tables64(mano_tables_np["right"])
    pose = torch.zeros(2, 33, dtype=torch.float64)
    verts, joints = mano.mano_forward(T, pose, torch.zeros(2, 10, dtype=torch.float64), center_idx=0)
    j0 = (T["th_J_regressor"] @ T["th_v_template"][0])[0]
    expect = (T["th_v_template"][0] - j0) * 1000
    assert (verts[0] - expect).abs().max() < 1e-4  # the 1e-8 Rodrigues epsilon leaves ~1e-5 mm
    assert joints[0, 0].abs().max() < 1e-9


def test_mano_global_rotation_equivariance(mano_tables_np):
    T = _tables64(mano_tables_np["right"])
    g = torch.Generator().manual_seed(0)
    pose = torch.randn(1, 33, generator=g, dtype=torch.float64) * 0.3
    betas = torch.randn(1, 10, generator=g, dtype=torch.float64)
    pose0 = pose.clone()
    pose0[:, :3] = 0
    v_rot, _ = mano.mano_forward(T, pose, betas, center_idx=0)
    v_0, _ = mano.mano_forward(T, pose0, betas, center_idx=0)
    R = mano.rodrigues(pose[:, :3])[0]
    R0 = mano.rodrigues(pose0[:, :3])[0]
    assert (v_rot[0] - v_0[0] @ (R @ R0.t()).t()).abs().max() < 1e-6


def test_mano_rodrigues_is_rotation_and_grad_is_finite_difference(mano_tables_np):
    a = torch.tensor([[0.3, -0.2, 0.9], [0.0, 0.0, 0.0]], dtype=torch.float64)
    R = mano.rodrigues(a)
    assert (R @ R.transpose(1, 2) - torch.eye(3, dtype=torch.float64)).abs().max() < 1e-12
    T = _tables64(mano_tables_np["left"], ncomps=6)
    pose = (torch.randn(1, 9, dtype=torch.float64, generator=torch.Generator().manual_seed(1)) * 0.5).requires_grad_(True)
    betas = torch.randn(1, 10, dtype=torch.float64, generator=torch.Generator().manual_seed(2)).requires_grad_(True)
    assert torch.autograd.gradcheck(
        lambda p, b: mano.mano_forward(T, p, b, side="left", center_idx=9, root_palm=True)[1][:, :, :2],
        (pose, betas), eps=1e-6, atol=1e-5)
