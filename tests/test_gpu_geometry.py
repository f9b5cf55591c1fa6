"""Parity of the CUDA geometry kernels (through the C ABI) against the fp64 oracle and the
reference-generated golden vectors.  Tolerance: 1e-4 relative on loss scalars / vertex coordinates
(BASELINE.json north_star); masks and indices exact except on numerically degenerate samples."""
import numpy as np
import pytest
import torch

from oracle import geometry, icosphere, mano
from obman_train_b200 import functional as Fb
from obman_train_b200.assets import load_contacts

pytestmark = pytest.mark.gpu
RTOL = 1e-4


def _cuda(a, grad=False):
    t = torch.tensor(np.asarray(a), dtype=torch.float32, device="cuda")
    return t.requires_grad_(grad)


def _rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


# ---- Chamfer ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_chamfer_golden(golden, tag):
    preds = _cuda(golden["chamfer_%s_preds" % tag], True)
    gts = _cuda(golden["chamfer_%s_gts" % tag])
    l1, l2 = Fb.chamfer(preds, gts)
    (l1 + l2).mean().backward()
    np.testing.assert_allclose(l1.detach().cpu().numpy(), golden["chamfer_%s_loss1" % tag], rtol=RTOL)
    np.testing.assert_allclose(l2.detach().cpu().numpy(), golden["chamfer_%s_loss2" % tag], rtol=RTOL)
    assert _rel(preds.grad.cpu().numpy(), golden["chamfer_%s_gpreds" % tag]) < 1e-3


@pytest.mark.parametrize("B,N,M", [(4, 642, 600), (2, 2500, 2100), (3, 1, 1), (2, 1025, 1023), (1, 130, 4100)])
def test_chamfer_vs_fp64_oracle(B, N, M):
    g = torch.Generator().manual_seed(B * 1000 + N)
    p = torch.randn(B, N, 3, generator=g) * 40
    t = torch.randn(B, M, 3, generator=g) * 40 + 30
    pc = p.cuda().requires_grad_(True)
    tc = t.cuda().requires_grad_(True)
    l1, l2 = Fb.chamfer(pc, tc)
    w = torch.linspace(0.5, 1.5, B, device="cuda")
    ((l1 * w).sum() + (l2 / w).sum()).backward()
    pd = p.double().requires_grad_(True)
    td = t.double().requires_grad_(True)
    o1, o2 = geometry.chamfer(pd, td)
    wd = w.cpu().double()
    ((o1 * wd).sum() + (o2 / wd).sum()).backward()
    np.testing.assert_allclose(l1.detach().cpu().numpy(), o1.detach().numpy(), rtol=RTOL)
    np.testing.assert_allclose(l2.detach().cpu().numpy(), o2.detach().numpy(), rtol=RTOL)
    assert _rel(pc.grad.cpu().numpy(), pd.grad.numpy()) < RTOL
    assert _rel(tc.grad.cpu().numpy(), td.grad.numpy()) < RTOL


def test_nn_indices_and_properties_full_size():
    # BASELINE config 5 scale; properties only (the oracle would need 100M-entry matrices per sample)
    B, N, M = 32, 10000, 9000
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(B, N, 3, device="cuda", generator=g) * 60
    y = torch.randn(B, M, 3, device="cuda", generator=g) * 60
    minx, idxx, miny, idxy = Fb.nearest_neighbours(x, y)
    # (1) the reported index reproduces the reported minimum exactly
    d = ((x - torch.gather(y, 1, idxx.long().unsqueeze(2).expand(-1, -1, 3))) ** 2)
    assert (idxx >= 0).all() and (idxx < M).all() and (idxy >= 0).all() and (idxy < N).all()
    np.testing.assert_allclose((d[..., 0] + d[..., 1] + d[..., 2]).cpu().numpy(), minx.cpu().numpy(), rtol=1e-5, atol=1e-4)
    # (2) swapping the clouds swaps the outputs bit-for-bit
    miny2, idxy2, minx2, idxx2 = Fb.nearest_neighbours(y, x)
    assert torch.equal(minx, minx2) and torch.equal(miny, miny2)
    assert torch.equal(idxx, idxx2) and torch.equal(idxy, idxy2)
    # (3) permuting the candidates leaves every minimum unchanged bit-for-bit
    perm = torch.randperm(M, device="cuda")
    minx3, idxx3, _, _ = Fb.nearest_neighbours(x, y[:, perm], dirs=1)
    assert torch.equal(minx, minx3)
    # (4) no candidate beats the reported minimum on a sampled slab
    sl = slice(0, 257)
    full = ((x[:2, sl].unsqueeze(2) - y[:2].unsqueeze(1)) ** 2).sum(-1).min(2)[0]
    np.testing.assert_allclose(full.cpu().numpy(), minx[:2, sl].cpu().numpy(), rtol=1e-5, atol=1e-4)


def test_chamfer_is_deterministic_in_forward():
    g = torch.Generator().manual_seed(5)
    p = (torch.randn(8, 642, 3, generator=g) * 40).cuda()
    t = (torch.randn(8, 600, 3, generator=g) * 40).cuda()
    a = Fb.chamfer(p, t)
    b = Fb.chamfer(p, t)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


# ---- ray casting / contact loss ----------------------------------------------------------------------------
def test_exterior_golden(golden):
    hand = _cuda(golden["contact_hand"])
    obj = _cuda(golden["contact_obj"])
    faces = torch.tensor(golden["contact_faces"], dtype=torch.int32, device="cuda")
    ext, _ = Fb.mesh_exterior(hand, obj, faces)
    ref, margin = geometry.mesh_exterior(hand.cpu().double(), obj.cpu().double()[:, faces.cpu().long()], return_margin=True)
    ok = (ext.cpu() == ref) | (margin < 1e-5)
    assert ok.all()
    assert (ext.cpu().numpy() == golden["contact_exterior"]).mean() > 0.999


@pytest.mark.parametrize("sub,B", [(3, 5), (4, 2)])
def test_exterior_vs_oracle_icosphere(sub, B):
    v, f = icosphere.icosphere(sub)
    g = torch.Generator().manual_seed(sub)
    obj = torch.tensor(v, dtype=torch.float32).unsqueeze(0) * 45 + torch.randn(B, v.shape[0], 3, generator=g) * 1.5
    pts = torch.randn(B, 778, 3, generator=g) * 35
    faces = torch.tensor(f, dtype=torch.int32, device="cuda")
    ext, hits = Fb.mesh_exterior(pts.cuda(), obj.cuda(), faces)
    ref, margin = geometry.mesh_exterior(pts.double(), obj.double()[:, torch.tensor(f)], return_margin=True)
    assert ((ext.cpu() == ref) | (margin < 1e-5)).all()
    assert 0.05 < (~ext).float().mean().item() < 0.95  # both classes are exercised


@pytest.mark.parametrize("zones_mode", ["all", "tips", "zones"])
@pytest.mark.parametrize("mode", ["dist_sq", "dist", "dist_tanh"])
def test_contact_loss_golden_and_oracle(golden, zones_mode, mode):
    from obman_train_b200.networks.branches.contactloss import compute_contact_loss
    _, zones = load_contacts()
    targets = ("all", "obj", "hand") if (zones_mode == "zones" and mode == "dist_tanh") else ("all",)
    for target in targets:
        k = "contact_%s_%s_%s_" % (zones_mode, mode, target)
        hand = _cuda(golden["contact_hand"], True)
        obj = _cuda(golden["contact_obj"], True)
        missed, penetr, info, metrics = compute_contact_loss(
            hand, None, obj, golden["contact_faces"], contact_thresh=10, contact_mode=mode,
            collision_thresh=20, collision_mode=mode, contact_target=target, contact_zones=zones_mode)
        hd = torch.tensor(golden["contact_hand"], dtype=torch.float64, requires_grad=True)
        od = torch.tensor(golden["contact_obj"], dtype=torch.float64, requires_grad=True)
        om, op, oinfo, ometrics = geometry.contact_loss(
            hd, od, golden["contact_faces"], zones, contact_thresh=10, contact_mode=mode,
            collision_thresh=20, collision_mode=mode, contact_target=target, contact_zones=zones_mode)
        assert (info["attraction_masks"].cpu().numpy().astype(bool) == golden[k + "attr"].astype(bool)).all()
        assert (info["repulsion_masks"].cpu().numpy().astype(bool) == golden[k + "rep"].astype(bool)).all()
        np.testing.assert_allclose(missed.item(), om.item(), rtol=RTOL, atol=1e-7)
        np.testing.assert_allclose(penetr.item(), op.item(), rtol=RTOL, atol=1e-7)
        np.testing.assert_allclose(missed.item(), golden[k + "missed"][0], rtol=RTOL, atol=1e-7)
        np.testing.assert_allclose(penetr.item(), golden[k + "penetr"][0], rtol=RTOL, atol=1e-7)
        np.testing.assert_allclose(metrics["max_penetr"].item(), ometrics["max_penetr"].item(), rtol=RTOL)
        np.testing.assert_allclose(metrics["mean_penetr"].item(), ometrics["mean_penetr"].item(), rtol=RTOL)
        np.testing.assert_allclose(info["min_dists"].cpu().numpy(), oinfo["min_dists"].detach().numpy(), rtol=1e-4, atol=1e-3)
        np.testing.assert_allclose(info["contact_points"].cpu().numpy(), oinfo["contact_points"].detach().numpy(), rtol=1e-6)
        total = (missed + 0.5 * penetr).sum()
        ototal = (om + 0.5 * op).sum()
        if ototal.requires_grad:
            total.backward()
            ototal.backward()
            gh = hand.grad.cpu().numpy() if hand.grad is not None else np.zeros(hd.shape)
            go = obj.grad.cpu().numpy() if obj.grad is not None else np.zeros(od.shape)
            ogh = hd.grad.numpy() if hd.grad is not None else np.zeros(hd.shape)
            ogo = od.grad.numpy() if od.grad is not None else np.zeros(od.shape)
            assert _rel(gh, ogh) < RTOL or np.abs(ogh).max() == 0
            assert _rel(go, ogo) < RTOL or np.abs(ogo).max() == 0


def test_contact_loss_bad_mode_raises(golden):
    from obman_train_b200.networks.branches.contactloss import compute_contact_loss
    hand = _cuda(golden["contact_hand"])
    obj = _cuda(golden["contact_obj"])
    with pytest.raises(ValueError):
        compute_contact_loss(hand, None, obj, golden["contact_faces"], contact_mode="nope")
    with pytest.raises(ValueError):
        compute_contact_loss(hand, None, obj, golden["contact_faces"], contact_zones="nope")
    with pytest.raises(ValueError):
        compute_contact_loss(hand, None, obj, golden["contact_faces"], contact_target="nope")


# ---- MANO -----------------------------------------------------------------------------------------------------
def _oracle_tables(layer):
    return {k: v.detach().cpu().double() for k, v in layer.named_buffers() if k != "th_faces"}


@pytest.mark.parametrize("side,center_idx,root_palm,ncomps,with_betas,flat",
                         [("right", 0, False, 30, True, True), ("left", 9, True, 6, True, False),
                          ("right", None, False, 45, False, True), ("left", 4, False, 30, True, True)])
def test_mano_layer_fwd_bwd_vs_fp64_oracle(mano_tables_np, side, center_idx, root_palm, ncomps, with_betas, flat):
    from obman_train_b200.manopth.manolayer import ManoLayer
    layer = ManoLayer(center_idx=center_idx, flat_hand_mean=flat, ncomps=ncomps, side=side,
                      tables=mano_tables_np[side]).cuda()
    B = 7
    g = torch.Generator().manual_seed(11)
    pose = torch.randn(B, 3 + ncomps, generator=g) * 0.4
    pose[0] = 0  # rest pose: exercises the Rodrigues epsilon
    betas = torch.randn(B, 10, generator=g)
    pc = pose.cuda().requires_grad_(True)
    bc = betas.cuda().requires_grad_(True) if with_betas else None
    verts, joints = layer(pc, th_betas=bc, th_trans=torch.Tensor([0]), root_palm=root_palm)
    pd = pose.double().requires_grad_(True)
    bd = betas.double().requires_grad_(True) if with_betas else None
    overts, ojoints = mano.mano_forward(_oracle_tables(layer), pd, bd, torch.zeros(1), root_palm, side,
                                        center_idx, True, ncomps)
    scale = overts.abs().max().item()
    assert (verts.cpu().double() - overts).abs().max().item() < RTOL * scale
    assert (joints.cpu().double() - ojoints).abs().max().item() < RTOL * scale
    wv = torch.randn(B, 778, 3, generator=g)
    wj = torch.randn(B, 21, 3, generator=g)
    ((verts * wv.cuda()).sum() + (joints * wj.cuda()).sum()).backward()
    ((overts * wv.double()).sum() + (ojoints * wj.double()).sum()).backward()
    assert _rel(pc.grad.cpu().numpy(), pd.grad.numpy()) < RTOL
    if with_betas:
        assert _rel(bc.grad.cpu().numpy(), bd.grad.numpy()) < RTOL


def test_mano_layer_large_batch_matches_small_batches(mano_tables_np):
    from obman_train_b200.manopth.manolayer import ManoLayer
    layer = ManoLayer(center_idx=0, ncomps=30, side="right", tables=mano_tables_np["right"]).cuda()
    g = torch.Generator().manual_seed(3)
    pose = (torch.randn(1024, 33, generator=g) * 0.4).cuda()
    betas = torch.randn(1024, 10, generator=g).cuda()
    v_all, j_all = layer(pose, betas)
    v_part, j_part = layer(pose[300:307].contiguous(), betas[300:307].contiguous())
    assert torch.equal(v_all[300:307], v_part) and torch.equal(j_all[300:307], j_part)


def test_nearest_neighbours_propagate_nan_like_torch_min():
    """A NaN coordinate must give NaN distances (the reference's torch.min over the distance matrix does), not a huge
    finite value with an arbitrary neighbour: the NaN query's own row, and - for a NaN candidate - every row of that
    sample.  Other samples of the batch are unaffected."""
    import torch
    from obman_train_b200 import functional as Fb
    g = torch.Generator().manual_seed(0)
    x = torch.randn(3, 700, 3, generator=g).cuda() * 40
    y = torch.randn(3, 650, 3, generator=g).cuda() * 40
    x[1, 5, 2] = float("nan")
    minx, _, miny, _ = Fb.nearest_neighbours(x, y)
    assert torch.isnan(minx[1, 5]) and torch.isfinite(minx[1, :5]).all() and torch.isfinite(minx[1, 6:]).all()
    assert torch.isnan(miny[1]).all()                      # x[1] is the candidate cloud of the y -> x direction
    assert torch.isfinite(minx[0]).all() and torch.isfinite(miny[0]).all() and torch.isfinite(minx[2]).all()
    l1, l2 = Fb.chamfer(x, y)
    assert torch.isnan(l1[1]) and torch.isnan(l2[1]) and torch.isfinite(l1[0]) and torch.isfinite(l2[2])


def test_packed_and_scalar_search_kernels_agree_bit_for_bit(tmp_path):
    """The packed fp32x2 nearest-neighbour / ray-triangle kernels (default), their scalar forms and the staged / streamed
    ray variants evaluate the same operations per pair: distances, arg-mins (ties -> lowest index), NaN propagation and
    hit counts must be IDENTICAL.  The switches are read once per process, so each variant runs in its own."""
    import os
    import subprocess
    import sys
    import numpy as np
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = os.path.join(root, "scripts", "dump_search_kernels.py")
    variants = {"default": {},
                "scalar": {"OBMAN_NN_PACKED": "0", "OBMAN_RAYCAST_PACKED": "0"},
                "staged": {"OBMAN_RAYCAST_STREAM": "0"},
                "staged4": {"OBMAN_RAYCAST_STREAM": "0", "OBMAN_RAYCAST_PT": "4"}}
    dumps = {}
    for name, env in variants.items():
        full_env = dict(os.environ)
        full_env.update(env)
        path = str(tmp_path / (name + ".npz"))
        out = subprocess.run([sys.executable, script, path], env=full_env, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0 and "dumped" in out.stdout, out.stdout[-1500:] + out.stderr[-1500:]
        dumps[name] = dict(np.load(path))
    ref = dumps["scalar"]
    assert np.isnan(ref["small_minx"][1, 5]) and np.isnan(ref["small_minx"][2]).all() and not np.isnan(ref["small_minx"][0]).any()
    assert (ref["chamfer_minx"][0, :50] == 0).all()
    for name, d in dumps.items():
        for key, val in ref.items():
            assert np.array_equal(val, d[key], equal_nan=True), (name, key)
