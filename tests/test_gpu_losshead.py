"""Device-side scalar-loss stage (csrc/loss_head.cu, obman_train_b200/losshead.py) against plain torch in fp64:
fused mse terms (manobranch.py:251-324), GT object statistics (atlasbranch.py:211-227), the weighted total with
device-side lambdas, and the scatter-free Chamfer backward (atlasutils.py:11-39)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import geometry
from obman_train_b200 import functional as Fb
from obman_train_b200 import losshead

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


def test_sq_terms_match_mse_loss_and_its_gradients():
    g = torch.Generator().manual_seed(0)
    B = 5
    verts = (torch.randn(B, 778, 3, generator=g) * 40).cuda().requires_grad_(True)
    verts_t = (torch.randn(B, 778, 3, generator=g) * 40).cuda()
    joints = (torch.randn(B, 21, 3, generator=g) * 40).cuda().requires_grad_(True)
    joints_t = (torch.randn(B, 21, 3, generator=g) * 40).cuda()
    shape = torch.randn(B, 10, generator=g).cuda().requires_grad_(True)
    pose = torch.randn(B, 33, generator=g).cuda().requires_grad_(True)
    w = losshead.LossWeights(["a", "b", "c", "d"])
    for name, val in zip("abcd", (0.167, 0.3, 1.5, 0.01)):
        w[name] = val
    ws = losshead._Workspace()
    terms = [(verts, verts_t, None, 0), (joints, joints_t, None, 1), (shape, None, None, 2), (pose, None, (3, 33), 3)]
    for rep in range(2):   # twice: the ticket counter must have been left at zero
        wsum, vals = losshead.sq_terms(terms, w.device("cuda"), ws)
        ref_vals = [F.mse_loss(verts.double(), verts_t.double()), F.mse_loss(joints.double(), joints_t.double()),
                    (shape.double() ** 2).mean(), (pose.double()[:, 3:] ** 2).mean()]
        ref = 0.167 * ref_vals[0] + 0.3 * ref_vals[1] + 1.5 * ref_vals[2] + 0.01 * ref_vals[3]
        for k in range(4):
            assert abs(vals[k].item() - ref_vals[k].item()) <= 2e-6 * abs(ref_vals[k].item()), (k, rep)
        assert abs(wsum.item() - ref.item()) <= 2e-6 * abs(ref.item())
    got = torch.autograd.grad(wsum * 2.0, [verts, joints, shape, pose])
    want = torch.autograd.grad(ref * 2.0, [verts, joints, shape, pose])
    for a, b in zip(got, want):
        assert _rel(a, b) < 2e-6
    assert (got[3][:, :3] == 0).all()


def test_object_targets_match_torch():
    g = torch.Generator().manual_seed(1)
    gt = (torch.randn(7, 601, 3, generator=g) * 40 + 30).cuda()
    centroid, scale, centred = losshead.object_targets(gt)
    c = gt.double().mean(1)
    cen = gt.double() - c.unsqueeze(1)
    s = torch.norm(cen, 2, 2).max(1)[0]
    assert _rel(centroid, c) < 1e-6 and _rel(centred, cen) < 1e-6 and _rel(scale[:, 0], s) < 1e-6


def test_combine_total_groups_and_device_side_weights():
    dev = "cuda"
    a = torch.tensor([2.0], device=dev, requires_grad=True)
    l1 = torch.rand(9, device=dev, requires_grad=True)
    l2 = torch.rand(9, device=dev, requires_grad=True)
    c = torch.tensor(3.0, device=dev, requires_grad=True)      # 0-dim, like edge_loss / laplacian_loss
    w = losshead.LossWeights(["one", "lam", "edge"])
    w["one"], w["lam"], w["edge"] = 1.0, 0.167, 0.1
    terms = [(a, 1.0, 0, 0), ((l1, l2), 1.0 / 9, 1, 0), (c, 1.0, 2, 1)]
    total, groups, vals = losshead.combine(terms, w.device(dev))
    sym = (l1 + l2).mean()
    assert total.item() == pytest.approx((a + 0.167 * sym + 0.1 * c).item(), rel=1e-6)
    assert vals[1].item() == pytest.approx(sym.item(), rel=1e-6)
    assert groups[1].item() == pytest.approx(0.3, rel=1e-6) and groups[0].item() == pytest.approx((a + 0.167 * sym).item(), rel=1e-6)
    ga, g1, g2, gc = torch.autograd.grad(total, [a, l1, l2, c])
    assert ga.item() == pytest.approx(1.0) and gc.item() == pytest.approx(0.1)
    assert torch.allclose(g1, torch.full_like(g1, 0.167 / 9)) and torch.allclose(g2, g1)
    # a weight changed on the host reaches the device vector: the same launch parameters now give another total
    w["edge"] = 0.05
    total2, _, _ = losshead.combine(terms, w.device(dev))
    assert total2.item() == pytest.approx((a + 0.167 * sym + 0.05 * c).item(), rel=1e-6)


@pytest.mark.parametrize("shape", [(3, 642, 600), (2, 50, 777), (2, 2562, 2500), (1, 1, 1), (2, 5, 4000)])
@pytest.mark.parametrize("scalar_grad", [False, True])
def test_chamfer_backward_gather_kernel_matches_fp64_autograd(shape, scalar_grad):
    B, N, M = shape
    g = torch.Generator().manual_seed(N + M)
    preds = (torch.randn(B, N, 3, generator=g) * 40)
    gts = (torch.randn(B, M, 3, generator=g) * 40 + 10)
    p64, t64 = preds.double().requires_grad_(True), gts.double().requires_grad_(True)
    o1, o2 = geometry.chamfer(p64, t64)
    pc, tc = preds.cuda().requires_grad_(True), gts.cuda().requires_grad_(True)
    l1, l2 = Fb.chamfer(pc, tc)
    if scalar_grad:
        (o1 + o2).mean().backward()
        (l1 + l2).mean().backward()      # arrives as one expanded scalar: the g_stride = 0 path when B > 1
    else:
        w1 = torch.rand(B, generator=g).double()
        w2 = torch.rand(B, generator=g).double()
        ((o1 * w1).sum() + (o2 * w2).sum()).backward()
        ((l1 * w1.float().cuda()).sum() + (l2 * w2.float().cuda()).sum()).backward()
    assert _rel(pc.grad, p64.grad) < 1e-5
    assert _rel(tc.grad, t64.grad) < 1e-5
    # bit-reproducible: no float atomics
    pc2 = preds.cuda().requires_grad_(True)
    a1, a2 = Fb.chamfer(pc2, gts.cuda())
    (a1 + a2).mean().backward()
    pc3 = preds.cuda().requires_grad_(True)
    b1, b2 = Fb.chamfer(pc3, gts.cuda())
    (b1 + b2).mean().backward()
    assert torch.equal(pc2.grad, pc3.grad)


def test_decay_regul_is_followed_by_a_captured_step():
    """lambdas baked into a captured graph made HandNet.decay_regul a silent no-op under use_graph.  The
    lambdas now live in device memory: a replay after decay_regul must equal an eager step with the decayed value."""
    from obman_train_b200.networks.handnet import HandNet
    from obman_train_b200.trainer import FlatAdamTrainer
    from tests.util import FULL_CFG, enum_sample, make_sample
    cfg = dict(FULL_CFG)
    cfg.update(atlas_lambda_regul_edges=5000.0, contact_lambda=0, collision_lambda=0)   # large: the term must matter
    sample = enum_sample(make_sample(2, 64, 3))

    def build():
        torch.manual_seed(11)
        model = HandNet(**cfg).eval().cuda()
        return model, FlatAdamTrainer(model, lr=1e-3)

    model_g, tr_g = build()
    tr_g.capture(dict(sample))
    model_g.decay_regul(0.25)
    loss_g = tr_g.replay().clone()
    model_e, tr_e = build()
    model_e.decay_regul(0.25)
    loss_e = tr_e.step(dict(sample))
    model_u, tr_u = build()     # undecayed, for contrast
    loss_u = tr_u.step(dict(sample))
    assert model_g.atlas_loss.edge_regul_lambda == pytest.approx(1250.0)
    assert abs(loss_g.item() - loss_e.item()) <= 1e-5 * abs(loss_e.item()), (loss_g.item(), loss_e.item())
    assert abs(loss_u.item() - loss_e.item()) > 1e-3 * abs(loss_e.item())
    assert ((tr_g.flat_p - tr_e.flat_p).norm() / (tr_e.flat_p - tr_u.flat_p).norm()).item() < 0.05


def test_contact_iou_kernel_matches_reference_formula():
    """meshiou / thresh_ious (contactloss.py:20-47) restated with torch ops, including samples with an empty union."""
    from obman_train_b200.networks.branches.contactloss import meshiou
    g = torch.Generator().manual_seed(5)
    B, P = 9, 778
    gt = (torch.rand(B, P, generator=g) * 30).cuda()
    pred = (torch.rand(B, P, generator=g) * 30).cuda()
    gt[3] += 100.0      # no contact at any threshold in either map: union empty -> IoU 0
    pred[3] += 100.0
    threshs = [1, 2, 3, 4, 5, 6, 7, 8, 9, 10]
    ious, auc = meshiou(gt, pred, threshs)
    rows = []
    for t in threshs:
        a, b = gt <= t, pred <= t
        inter, union = (a & b).sum(1).double(), (a | b).sum(1).double()
        rows.append(torch.where(union != 0, inter / union.clamp(min=1), torch.zeros_like(union)))
    table = torch.stack(rows)                                   # (T, B)
    ref_auc = np.mean(np.trapezoid(table.cpu().numpy(), axis=0, x=threshs))
    assert torch.allclose(ious.double(), table.mean(1), rtol=1e-6, atol=1e-7)
    assert abs(float(auc) - ref_auc) <= 1e-6 * abs(ref_auc)
    assert table[:, 3].abs().max() == 0
