"""torch.autograd wrappers over the C ABI (device pointers in, device pointers out).

PyTorch is plumbing here: it owns the device memory, the current stream and the autograd tape; every
arithmetic step of the ops below runs in libobman_b200.so.  All functions require CUDA fp32 tensors and
raise if the library is missing - there is no CPU path.
"""
import torch

from . import _lib
from ._lib import call, ptr, stream_ptr

CONTACT_MODES = {"dist_sq": 0, "dist": 1, "dist_tanh": 2}
CONTACT_ZONES = {"all": 0, "tips": 1, "zones": 2}
CONTACT_TARGETS = {"all": 0, "obj": 1, "hand": 2}


def _prep(t, name):
    if not t.is_cuda:
        raise RuntimeError("{}: expected a CUDA tensor (obman_train_b200 has no CPU path)".format(name))
    if t.dtype != torch.float32:
        raise RuntimeError("{}: expected float32, got {}".format(name, t.dtype))
    return t.contiguous()


# ---------------------------------------------------------------------------------------------------
# nearest neighbours / Chamfer
# ---------------------------------------------------------------------------------------------------
def nearest_neighbours(x, y, dirs=3):
    """x (B,N,3), y (B,M,3) -> (minx (B,N), idxx (B,N) int32, miny (B,M), idxy (B,M) int32);
    squared distances; entries of a direction that was not requested are None."""
    x = _prep(x.detach(), "x")
    y = _prep(y.detach(), "y")
    B, N, _ = x.shape
    M = y.shape[1]
    minx = idxx = miny = idxy = None
    if dirs & 1:
        minx = torch.empty((B, N), device=x.device, dtype=torch.float32)
        idxx = torch.empty((B, N), device=x.device, dtype=torch.int32)
    if dirs & 2:
        miny = torch.empty((B, M), device=x.device, dtype=torch.float32)
        idxy = torch.empty((B, M), device=x.device, dtype=torch.int32)
    call("obman_nn_fwd", ptr(x), ptr(y), B, N, M, ptr(minx), ptr(idxx), ptr(miny), ptr(idxy), dirs,
         stream_ptr())
    return minx, idxx, miny, idxy


class _ChamferFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, preds, gts):
        preds = _prep(preds, "preds")
        gts = _prep(gts, "gts")
        B, N, _ = preds.shape
        M = gts.shape[1]
        if gts.shape[0] != B or preds.shape[2] != 3 or gts.shape[2] != 3:
            raise RuntimeError("chamfer: expected preds (B,N,3) and gts (B,M,3)")
        dev = preds.device
        loss1 = torch.empty(B, device=dev)
        loss2 = torch.empty(B, device=dev)
        min1 = torch.empty((B, N), device=dev)
        idx1 = torch.empty((B, N), device=dev, dtype=torch.int32)
        min2 = torch.empty((B, M), device=dev)
        idx2 = torch.empty((B, M), device=dev, dtype=torch.int32)
        call("obman_chamfer_fwd", ptr(preds), ptr(gts), B, N, M, ptr(loss1), ptr(loss2), ptr(min1),
             ptr(idx1), ptr(min2), ptr(idx2), stream_ptr())
        ctx.save_for_backward(preds, gts, idx1, idx2)
        return loss1, loss2

    @staticmethod
    def backward(ctx, g1, g2):
        preds, gts, idx1, idx2 = ctx.saved_tensors
        B, N, _ = preds.shape
        M = gts.shape[1]
        # the gradient of torch.mean(loss_1 + loss_2) (atlasbranch.py:235,243) and of the device-side loss total
        # (losshead.combine) arrives as ONE scalar expanded over the batch: pass it as such (g_stride 0) instead of
        # materialising the (B,) vector
        if g1.dim() == 1 and g2.dim() == 1 and B > 1 and g1.stride(0) == 0 and g2.stride(0) == 0:
            g_stride = 0
        else:
            g1 = _prep(g1, "g1")
            g2 = _prep(g2, "g2")
            g_stride = 1
        if not (g1.is_cuda and g2.is_cuda and g1.dtype == torch.float32 and g2.dtype == torch.float32):
            raise RuntimeError("chamfer backward: expected CUDA float32 gradients")
        gpreds = torch.empty_like(preds)
        ggts = torch.empty_like(gts) if ctx.needs_input_grad[1] else None
        call("obman_chamfer_bwd", ptr(preds), ptr(gts), ptr(idx1), ptr(idx2), ptr(g1), ptr(g2), g_stride, B, N,
             M, ptr(gpreds), ptr(ggts), stream_ptr())
        return gpreds, ggts


def chamfer(preds, gts):
    """ChamferLoss.forward: (loss_1 (B,), loss_2 (B,)); atlasutils.py:11-18."""
    return _ChamferFn.apply(preds, gts)


# ---------------------------------------------------------------------------------------------------
# contact loss
# ---------------------------------------------------------------------------------------------------
def mesh_exterior(points, obj_verts, faces_i32):
    """bool (B,P): True where the fixed-direction ray from the point crosses the mesh an even number of
    times (batch_mesh_contains_points, contactutils.py:62-159).  faces_i32 (F,3) int32 CUDA."""
    points = _prep(points.detach(), "points")
    obj_verts = _prep(obj_verts.detach(), "obj_verts")
    B, P, _ = points.shape
    N = obj_verts.shape[1]
    F = faces_i32.shape[0]
    hits = torch.empty((B, P), device=points.device, dtype=torch.int32)
    scratch = torch.empty((B, (F + 1) // 2, 32), device=points.device, dtype=torch.float32)
    call("obman_raycast_hits", ptr(points), ptr(obj_verts), ptr(faces_i32), B, P, N, F, ptr(hits), ptr(scratch),
         stream_ptr())
    return (hits & 1) == 0, hits


class _ContactFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, hand, obj, faces_i32, zone_ids, zone_ptr, cfg):
        hand = _prep(hand, "hand_verts")
        obj = _prep(obj, "obj_verts")
        B, P, _ = hand.shape
        N = obj.shape[1]
        dev = hand.device
        st = stream_ptr()
        mins21 = torch.empty((B, P), device=dev)
        idx21 = torch.empty((B, P), device=dev, dtype=torch.int32)
        call("obman_nn_fwd", ptr(hand), ptr(obj), B, P, N, ptr(mins21), ptr(idx21), None, None, 1, st)
        hits = torch.empty((B, P), device=dev, dtype=torch.int32)
        n_faces = faces_i32.shape[0]
        scratch = torch.empty((B, (n_faces + 1) // 2, 32), device=dev, dtype=torch.float32)
        call("obman_raycast_hits", ptr(hand), ptr(obj), ptr(faces_i32), B, P, N, n_faces, ptr(hits), ptr(scratch), st)
        attr = torch.empty((B, P), device=dev, dtype=torch.uint8)
        rep = torch.empty((B, P), device=dev, dtype=torch.uint8)
        close = torch.empty((B, P, 3), device=dev)
        anchor = torch.empty((B, P), device=dev)
        partial = torch.empty((B, 6), device=dev)
        out = torch.empty(6, device=dev)
        n_zones = 0 if zone_ptr is None else zone_ptr.numel() - 1
        call("obman_contact_fwd", ptr(hand), ptr(obj), ptr(mins21), ptr(idx21), ptr(hits),
             ptr(zone_ids), ptr(zone_ptr), n_zones, cfg["zones_mode"], B, P, N,
             float(cfg["contact_thresh"]), cfg["contact_mode"], float(cfg["collision_thresh"]),
             cfg["collision_mode"], ptr(attr), ptr(rep), ptr(close), ptr(anchor), ptr(partial),
             ptr(out), st)
        ctx.cfg = cfg
        ctx.save_for_backward(hand, close, anchor, idx21, attr, rep, out)
        ctx.n_obj = N
        missed = out[0:1].clone()
        penetr = out[1:2].clone()
        ctx.mark_non_differentiable(attr, rep, close, mins21, out)
        return missed, penetr, attr, rep, close, mins21, out

    @staticmethod
    def backward(ctx, g_missed, g_penetr, *unused):
        hand, close, anchor, idx21, attr, rep, out = ctx.saved_tensors
        cfg = ctx.cfg
        B, P, _ = hand.shape
        N = ctx.n_obj
        g_missed = _prep(g_missed, "g_missed")
        g_penetr = _prep(g_penetr, "g_penetr")
        ghand = torch.empty_like(hand) if ctx.needs_input_grad[0] else None
        gobj = torch.empty((B, N, 3), device=hand.device) if ctx.needs_input_grad[1] else None
        call("obman_contact_bwd", ptr(hand), ptr(close), ptr(anchor), ptr(idx21), ptr(attr), ptr(rep),
             ptr(out), ptr(g_missed), ptr(g_penetr), B, P, N, float(cfg["contact_thresh"]),
             cfg["contact_mode"], float(cfg["collision_thresh"]), cfg["collision_mode"],
             cfg["target"], ptr(ghand), ptr(gobj), stream_ptr())
        return ghand, gobj, None, None, None, None


def contact_loss(hand, obj, faces_i32, zone_ids, zone_ptr, contact_thresh, contact_mode,
                 collision_thresh, collision_mode, contact_target, contact_zones):
    """Fused compute_contact_loss core; returns
    (missed_loss (1,), penetr_loss (1,), attr_mask u8 (B,P), rep_mask u8 (B,P), close (B,P,3),
     mins21 (B,P), stats (6,) = [missed, penetr, max_penetr, mean_penetr, n_attr, n_rep])."""
    cfg = {
        "contact_thresh": contact_thresh, "collision_thresh": collision_thresh,
        "contact_mode": CONTACT_MODES[contact_mode], "collision_mode": CONTACT_MODES[collision_mode],
        "zones_mode": CONTACT_ZONES[contact_zones], "target": CONTACT_TARGETS[contact_target],
    }
    return _ContactFn.apply(hand, obj, faces_i32, zone_ids, zone_ptr, cfg)


def contact_iou(gt_dists, pred_dists, threshs):
    """(batch_ious (T,), auc 0-dim): meshiou of contactloss.py:35-47 on (B,P) squared-distance maps."""
    import ctypes
    gt = _prep(gt_dists.detach(), "gt_dists")
    pred = _prep(pred_dists.detach(), "pred_dists")
    if gt.shape != pred.shape or gt.dim() != 2:
        raise RuntimeError("contact_iou: expected two (B,P) tensors")
    B, P = gt.shape
    T = len(threshs)
    th = (ctypes.c_float * T)(*[float(t) for t in threshs])
    ws = torch.empty((B, T), device=gt.device)
    ious = torch.empty(T, device=gt.device)
    auc = torch.empty((), device=gt.device)
    call("obman_contact_iou", ptr(gt), ptr(pred), B, P, th, T, ptr(ws), ptr(ious), ptr(auc), stream_ptr())
    return ious, auc


# ---------------------------------------------------------------------------------------------------
# MANO layer
# ---------------------------------------------------------------------------------------------------
class _ManoFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pose, betas, trans, tables, side_left, root_palm, center_idx):
        pose = _prep(pose, "th_pose_coeffs")
        B = pose.shape[0]
        ncomps = tables["comps"].shape[0]
        if pose.shape[1] != 3 + ncomps:
            raise RuntimeError("mano: pose has {} columns, expected 3+ncomps={}".format(pose.shape[1], 3 + ncomps))
        betas_c = None if betas is None else _prep(betas, "th_betas")
        trans_c = None if trans is None else _prep(trans, "th_trans")
        V = tables["v_template"].shape[0]
        dev = pose.device
        ws_pm = torch.empty((B, 135), device=dev)
        ws_gp = torch.empty((B, 192), device=dev)
        ws_tw = torch.empty((B, 48), device=dev)
        ws_be = torch.empty((B, 10), device=dev)
        verts = torch.empty((B, V, 3), device=dev)
        joints = torch.empty((B, 21, 3), device=dev)
        t = tables
        call("obman_mano_fwd", ptr(t["v_template"]), ptr(t["shapedirs"]), ptr(t["posedirs"]),
             ptr(t["weights"]), ptr(t["j_template"]), ptr(t["j_shapedirs"]), ptr(t["hands_mean"]),
             ptr(t["comps"]), ptr(t["betas"]), V, ncomps, ptr(pose), ptr(betas_c), ptr(trans_c), B,
             int(side_left), int(root_palm), int(center_idx), ptr(ws_pm), ptr(ws_gp), ptr(ws_tw),
             ptr(ws_be), ptr(verts), ptr(joints), stream_ptr())
        ctx.tables = tables
        ctx.flags = (int(side_left), int(root_palm), int(center_idx), trans is not None)
        ctx.has_betas = betas is not None
        ctx.save_for_backward(pose, betas_c if betas_c is not None else pose.new_empty(0), ws_pm, ws_gp, ws_be)
        return verts, joints

    @staticmethod
    def backward(ctx, gverts, gjoints):
        pose, betas, ws_pm, ws_gp, ws_be = ctx.saved_tensors
        t = ctx.tables
        side_left, root_palm, center_idx, has_trans = ctx.flags
        B = pose.shape[0]
        V = t["v_template"].shape[0]
        ncomps = t["comps"].shape[0]
        dev = pose.device
        gverts = None if gverts is None else _prep(gverts, "gverts")
        gjoints = None if gjoints is None else _prep(gjoints, "gjoints")
        ws_gv = torch.empty((B, V, 3), device=dev)
        ws_gtw = torch.empty((B, 48), device=dev)
        ws_gacc = torch.empty((B, 337), device=dev)
        gpose = torch.empty_like(pose)
        gbetas = torch.empty((B, 10), device=dev) if (ctx.has_betas and ctx.needs_input_grad[1]) else None
        call("obman_mano_bwd", ptr(t["v_template"]), ptr(t["shapedirs"]), ptr(t["posedirs"]),
             ptr(t["weights"]), ptr(t["j_template"]), ptr(t["j_shapedirs"]), ptr(t["hands_mean"]),
             ptr(t["comps"]), ptr(t["betas"]), V, ncomps, ptr(pose),
             ptr(betas) if ctx.has_betas else None, int(has_trans), B, side_left, root_palm,
             center_idx, ptr(ws_pm), ptr(ws_gp), ptr(ws_be), ptr(gverts), ptr(gjoints), ptr(ws_gv),
             ptr(ws_gtw), ptr(ws_gacc), ptr(gpose), ptr(gbetas), stream_ptr())
        return gpose, gbetas, None, None, None, None, None


def mano_layer(pose, betas, trans, tables, side_left, root_palm, center_idx):
    """(verts (B,778,3) mm, joints (B,21,3) mm); tables: dict of contiguous CUDA fp32 tensors
    v_template, shapedirs, posedirs, weights, j_template, j_shapedirs, hands_mean, comps, betas."""
    return _ManoFn.apply(pose, betas, trans, tables, side_left, root_palm, center_idx)


def launches():
    return _lib.launch_count


# ---------------------------------------------------------------------------------------------------
# mesh regularisers (fixed icosphere topology)
# ---------------------------------------------------------------------------------------------------
class _LaplacianFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, verts, nbr, w):
        verts = _prep(verts, "verts")
        B, N, _ = verts.shape
        K = nbr.shape[1]
        if nbr.shape[0] != N or tuple(w.shape) != tuple(nbr.shape):
            raise RuntimeError("laplacian_loss: the Laplacian was built for {} vertices, got {}".format(
                nbr.shape[0], N))
        lx = torch.empty_like(verts)
        partial = torch.empty((B * N + 255) // 256, device=verts.device)
        loss = torch.empty(1, device=verts.device)
        call("obman_laplacian_fwd", ptr(verts), ptr(nbr), ptr(w), B, N, K, ptr(lx), ptr(partial), ptr(loss),
             stream_ptr())
        ctx.save_for_backward(lx, nbr, w)
        ctx.mark_non_differentiable(lx)
        return loss, lx

    @staticmethod
    def backward(ctx, gloss, _glx):
        lx, nbr, w = ctx.saved_tensors
        B, N, _ = lx.shape
        gv = torch.empty_like(lx)
        call("obman_laplacian_bwd", ptr(lx), ptr(_prep(gloss, "gloss")), ptr(nbr), ptr(w), B, N, nbr.shape[1],
             ptr(gv), stream_ptr())
        return gv, None, None


def laplacian_loss(verts, nbr, w):
    """verts (B,N,3); nbr (N,K) int32 / w (N,K) fp32: ELL form of the constant cotangent Laplacian.
    Returns (loss (1,), Lx (B,N,3)); laplacianloss.py:36-41."""
    return _LaplacianFn.apply(verts, nbr, w)


class _EdgeLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, verts, faces_i32, vert_faces):
        verts = _prep(verts, "verts")
        B, N, _ = verts.shape
        F = faces_i32.shape[0]
        stats = torch.empty((B, 3), device=verts.device)
        loss = torch.empty(1, device=verts.device)
        call("obman_edge_loss_fwd", ptr(verts), ptr(faces_i32), B, N, F, ptr(stats), ptr(loss), stream_ptr())
        ctx.save_for_backward(verts, faces_i32, vert_faces, stats)
        return loss

    @staticmethod
    def backward(ctx, gloss):
        verts, faces_i32, vert_faces, stats = ctx.saved_tensors
        B, N, _ = verts.shape
        gv = torch.empty_like(verts)
        call("obman_edge_loss_bwd", ptr(verts), ptr(faces_i32), ptr(vert_faces), ptr(stats),
             ptr(_prep(gloss, "gloss")), B, N, faces_i32.shape[0], vert_faces.shape[1], ptr(gv), stream_ptr())
        return gv, None, None


def edge_loss(verts, faces_i32, vert_faces):
    """verts (B,N,3); faces_i32 (F,3) int32; vert_faces (N,Kf) int32 (-1 padded incident-face table).
    Returns the (1,) edge regulariser of atlasbranch.py:153-167."""
    return _EdgeLossFn.apply(verts, faces_i32, vert_faces)
