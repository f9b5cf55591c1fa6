"""Geometry losses of the obman_train hot path, restated on CPU (oracle; test infrastructure only).

Each function cites the reference code it follows.  Written with direct
``(x - y)^2`` distances (the reference uses the ``|x|^2 + |y|^2 - 2 x.y`` expansion,
/root/reference/mano_train/networks/branches/atlasutils.py:20-39, which loses ~2e-2 mm^2
in fp32); run in float64 this is the arbiter for both the reference and the CUDA path.
Pinned against the reference's own files through ``oracle.refhook``
(tests/test_oracle_vs_reference.py) and the golden vectors in tests/golden/.
"""
import numpy as np
import torch

RAY_DIRECTION = (0.4395064455, 0.617598629942, 0.652231566745)  # contactutils.py:65
RAY_TOL = 0.0000001  # contactutils.py:78
TIP_IDXS = (745, 317, 444, 556, 673)  # contactloss.py:258


def pairwise_sqdist(x, y):
    """P[b,i,j] = |x_i - y_j|^2 ; follows batch_pairwise_dist (contactloss.py:60-79)."""
    return ((x.unsqueeze(2) - y.unsqueeze(1)) ** 2).sum(-1)


def chamfer(preds, gts):
    """ChamferLoss.forward (atlasutils.py:11-18): P = pairwise(gts, preds);
    loss_1[b] = mean_j min_i P (pred -> nearest gt), loss_2[b] = mean_i min_j P."""
    P = pairwise_sqdist(gts, preds)
    loss_1 = P.min(1)[0].mean(1)
    loss_2 = P.min(2)[0].mean(1)
    return loss_1, loss_2


def chamfer_with_idx(preds, gts):
    P = pairwise_sqdist(gts, preds)
    m1, i1 = P.min(1)  # per pred j: nearest gt
    m2, i2 = P.min(2)  # per gt i: nearest pred
    return m1, i1, m2, i2


def mesh_exterior(points, triangles, return_margin=False):
    """batch_mesh_contains_points (contactutils.py:62-159): Moller-Trumbore along a fixed
    direction against every triangle; exterior <=> even number of hits.

    points (B,P,3), triangles (B,F,3,3) -> bool (B,P).  With ``return_margin`` also returns,
    per point, the smallest distance of any decision quantity (det, u, v, u+v, t) to its
    threshold, so that tests can skip numerically degenerate rays.
    """
    dt = points.dtype
    d = torch.tensor(RAY_DIRECTION, dtype=dt, device=points.device)
    v0 = triangles[:, :, 0]
    e1 = triangles[:, :, 1] - v0
    e2 = triangles[:, :, 2] - v0
    pvec = torch.cross(d.expand_as(e2), e2, dim=2)  # (B,F,3)
    det = (e1 * pvec).sum(2)  # (B,F)
    parallel = det.abs() < RAY_TOL
    invdet = 1 / (det + 0.1 * RAY_TOL)
    tvec = points.unsqueeze(2) - v0.unsqueeze(1)  # (B,P,F,3)
    u = (tvec * pvec.unsqueeze(1)).sum(3) * invdet.unsqueeze(1)
    qvec = torch.cross(tvec, e1.unsqueeze(1).expand_as(tvec), dim=3)
    v = (qvec * d).sum(3) * invdet.unsqueeze(1)
    t = (qvec * e2.unsqueeze(1)).sum(3) * invdet.unsqueeze(1)
    hit = (u > 0) & (u < 1) & (v > 0) & (u + v < 1) & (t >= RAY_TOL) & (~parallel).unsqueeze(1)
    exterior = hit.sum(2) % 2 == 0
    if not return_margin:
        return exterior
    # A decision can only flip when the quantity closest to ITS threshold is tiny and all other
    # conditions of that triangle hold; take a conservative per-point margin.
    big = torch.full_like(u, 1e30)
    others_u = (v > 0) & (u + v < 1) & (t >= RAY_TOL)
    others_v = (u > 0) & (u < 1) & (t >= RAY_TOL)
    others_t = (u > 0) & (u < 1) & (v > 0) & (u + v < 1)
    m = torch.minimum(torch.where(others_u, torch.minimum(u.abs(), (1 - u).abs()), big),
                      torch.where(others_v, torch.minimum(v.abs(), (1 - u - v).abs()), big))
    m = torch.minimum(m, torch.where(others_t, (t - RAY_TOL).abs(), big))
    return exterior, m.min(2)[0]


def masked_mean(vals, mask):
    """masked_mean_loss (contactloss.py:50-57): batch-global sum(mask*vals)/sum(mask); 0 if empty."""
    maskf = mask.to(vals.dtype)
    n = maskf.sum()
    if n > 0:
        return (maskf * vals).sum() / n
    return torch.zeros(1, dtype=vals.dtype, device=vals.device)


def contact_loss(hand, obj, obj_faces, zones=None, contact_thresh=5, contact_mode="dist_sq",
                 collision_thresh=10, collision_mode="dist_sq", contact_target="all",
                 contact_sym=False, contact_zones="all"):
    """compute_contact_loss (contactloss.py:149-308).

    hand (B,778,3), obj (B,N,3), obj_faces (F,3) int.  ``zones``: dict zone-id -> list of hand
    vertex ids (assets/contact_zones.pkl), required for contact_zones == "zones".
    Returns (missed_loss, penetr_loss, contact_info, metrics).
    """
    faces = torch.as_tensor(np.asarray(obj_faces).astype(np.int64)).to(obj.device)
    P = pairwise_sqdist(hand, obj)  # (B,778,N)
    mins12, idx12 = P.min(1)
    mins21, idx21 = P.min(2)
    tri = obj[:, faces]  # (B,F,3,3)
    exterior = mesh_exterior(hand.detach(), tri.detach())
    penetr_mask = ~exterior
    close = torch.gather(obj, 1, idx21.unsqueeze(2).expand(-1, -1, 3))

    def diff():
        if contact_target == "all":
            return close - hand
        if contact_target == "obj":
            return close - hand.detach()
        if contact_target == "hand":
            return close.detach() - hand
        raise ValueError("contact_target {} not in [all|obj|hand]".format(contact_target))

    anchor = torch.sqrt((diff() ** 2).sum(2))

    def values(mode, thresh):
        if mode == "dist_sq":
            return (diff() ** 2).sum(2)
        if mode == "dist":
            return anchor
        if mode == "dist_tanh":
            return thresh * torch.tanh(anchor / thresh)
        raise ValueError("mode {} not in [dist_sq|dist|dist_tanh]".format(mode))

    contact_vals = values(contact_mode, contact_thresh)
    if contact_mode == "dist_sq":
        below = mins21 < contact_thresh ** 2
    elif contact_mode == "dist":
        below = mins21 < contact_thresh  # squared distance vs unsquared threshold: reference quirk
    else:
        below = torch.ones_like(mins21, dtype=torch.bool)
    collision_vals = values(collision_mode, collision_thresh)

    missed_mask = below & exterior
    if contact_zones == "tips":
        tips = torch.zeros_like(missed_mask)
        tips[:, list(TIP_IDXS)] = True
        missed_mask = missed_mask & tips
    elif contact_zones == "zones":
        matching = torch.zeros_like(missed_mask)
        for _, zone_idxs in zones.items():
            zone_idxs = torch.as_tensor(np.asarray(zone_idxs).astype(np.int64)).to(hand.device)
            _, arg = mins21[:, zone_idxs].min(1)
            matching[torch.arange(hand.shape[0], device=hand.device), zone_idxs[arg]] = True
        missed_mask = missed_mask & matching
    elif contact_zones != "all":
        raise ValueError("contact_zones {} not in [tips|zones|all]".format(contact_zones))

    missed_loss = masked_mean(contact_vals, missed_mask)
    penetr_loss = masked_mean(collision_vals, penetr_mask)
    if contact_sym:
        missed_loss = missed_loss + masked_mean(torch.sqrt(mins12), mins12 < contact_thresh)
    pen = anchor.detach() * penetr_mask.to(anchor.dtype)
    metrics = {"max_penetr": pen.max(1)[0].mean(), "mean_penetr": pen.mean(1).mean()}
    info = {"attraction_masks": missed_mask, "repulsion_masks": penetr_mask,
            "contact_points": close, "min_dists": mins21}
    return missed_loss, penetr_loss, info, metrics


def edge_loss(verts, faces):
    """edge_loss (atlasbranch.py:153-167): mean |e - mean_b(e)| over squared edge lengths."""
    faces = torch.as_tensor(np.asarray(faces).astype(np.int64)).to(verts.device)
    a, b, c = verts[:, faces[:, 0]], verts[:, faces[:, 1]], verts[:, faces[:, 2]]
    la = ((b - a) ** 2).sum(2)
    lb = ((c - b) ** 2).sum(2)
    lc = ((a - c) ** 2).sum(2)
    e = torch.cat([lc, lb, la], dim=1)
    return (e - e.mean(1, keepdim=True)).abs().mean()


def laplacian_matrix(sphere_verts, faces):
    """Dense (N,N) float64 cotangent Laplacian of the unit sphere mesh, as laplacianloss.py builds it:
    cotangent (:153-185) -> C (F,3) = [l2^2+l3^2-l1^2, l1^2+l3^2-l2^2, l1^2+l2^2-l3^2] / (2*Heron area) / 4;
    entries at (rows, cols) = (f[:, [1,2,0]], f[:, [2,0,1]]) (:117-123); L = L + L^T; L = L - diag(rowsum) (:124-127)."""
    v = np.asarray(sphere_verts, dtype=np.float64)
    f = np.asarray(faces).astype(np.int64)
    v1, v2, v3 = v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]
    l1 = np.sqrt(((v2 - v3) ** 2).sum(1))
    l2 = np.sqrt(((v3 - v1) ** 2).sum(1))
    l3 = np.sqrt(((v1 - v2) ** 2).sum(1))
    sp = (l1 + l2 + l3) * 0.5
    A = 2 * np.sqrt(sp * (sp - l1) * (sp - l2) * (sp - l3))
    C = np.stack([l2 ** 2 + l3 ** 2 - l1 ** 2, l1 ** 2 + l3 ** 2 - l2 ** 2, l1 ** 2 + l2 ** 2 - l3 ** 2], 1)
    C = C / A[:, None] / 4
    n = v.shape[0]
    L = np.zeros((n, n))
    np.add.at(L, (f[:, [1, 2, 0]].reshape(-1), f[:, [2, 0, 1]].reshape(-1)), C.reshape(-1))
    L = L + L.T
    return L - np.diag(L.sum(1))


def laplacian_loss(verts, L):
    """LaplacianLoss.__call__ (laplacianloss.py:36-41): mean over all B*N rows of ||(L V_b)_i||_2.
    ``L`` from laplacian_matrix (numpy or tensor); differentiable in ``verts`` (the reference's backward is
    L^T g = L g, :137-150, which is what autograd gives here)."""
    Lt = torch.as_tensor(L, dtype=verts.dtype).to(verts.device)
    lx = torch.einsum("ij,bjc->bic", Lt, verts)
    return torch.sqrt((lx ** 2).sum(2)).mean(), lx


def mse(a, b):
    return ((a - b) ** 2).mean()


def atlas_loss(preds, gt_points, lambda_atlas, final_lambda_atlas, trans_weight, scale_weight,
               edge_regul_lambda=None, lambda_laplacian=0, laplacian=None):
    """AtlasLoss.compute_loss (atlasbranch.py:199-287), translation-predicted branch
    (:206-253) and the plain branch (:255-264); ``laplacian`` = laplacian_matrix(sphere verts, faces) when
    ``lambda_laplacian`` is set (:275-280)."""
    losses = {}
    if "objtrans" in preds and "objpointscentered3d" in preds:
        centroid = gt_points.mean(1)
        trans_l = mse(preds["objtrans"], centroid)
        losses["atlas_trans3d"] = trans_l
        centered = gt_points - centroid.unsqueeze(1)
        if "objscale" in preds:
            scales = torch.sqrt((centered ** 2).sum(2)).max(1)[0]
            scale_l = mse(preds["objscale"], scales.unsqueeze(1))
            losses["atlas_scale3d"] = scale_l
        else:
            scale_l = 0
        l1, l2 = chamfer(preds["objpointscentered3d"], centered)
        sym = (l1 + l2).mean()
        mesh = preds["objpointscentered3d"]
        f1, f2 = chamfer(preds["objpoints3d"], gt_points)
        sym_final = (f1 + f2).mean()
        losses["final_chamfer_loss"] = sym_final
        total = (lambda_atlas * sym + final_lambda_atlas * sym_final
                 + trans_weight * trans_l + scale_weight * scale_l)
    else:
        l1, l2 = chamfer(preds["objpoints3d"], gt_points)
        sym = (l1 + l2).mean()
        total = lambda_atlas * sym
        mesh = preds["objpoints3d"]
    if edge_regul_lambda is not None and edge_regul_lambda > 0:
        el = edge_loss(mesh, preds["objfaces"])
        losses["atlas_edge_regul"] = el
        total = total + edge_regul_lambda * el
    if lambda_laplacian:
        ll, _ = laplacian_loss(mesh, laplacian)
        losses["atlas_laplac"] = ll
        total = total + lambda_laplacian * ll
    losses["atlas_objpoints3d"] = sym
    return total, losses


def mano_loss(preds, target_verts=None, target_joints=None, lambda_verts=None,
              lambda_joints3d=None, lambda_shape=None, lambda_pose_reg=None):
    """ManoLoss.compute_loss (manobranch.py:251-324); PCA-supervision term excluded
    (BaseQueries.hand_pcas is never in the training sample)."""
    total = 0
    losses = {}
    if target_verts is not None and lambda_verts:
        losses["mano_verts3d"] = mse(preds["verts"], target_verts)
        total = total + lambda_verts * losses["mano_verts3d"]
    if target_joints is not None and lambda_joints3d:
        losses["mano_joints3d"] = mse(preds["joints"], target_joints)
        total = total + lambda_joints3d * losses["mano_joints3d"]
    if lambda_shape:
        losses["mano_shape"] = (preds["shape"] ** 2).mean()
        total = total + lambda_shape * losses["mano_shape"]
    if lambda_pose_reg:
        losses["pose_reg"] = (preds["pose"][:, 3:] ** 2).mean()
        total = total + lambda_pose_reg * losses["pose_reg"]
    losses["mano_total_loss"] = total
    return total, losses
