"""Functional CPU restatement of the network part of the hot path (oracle; test infrastructure only).

Every function takes a flat ``state`` dict keyed by the REFERENCE's state-dict names
(SURVEY.md §5 checkpoint row) so that the product modules' ``state_dict()`` can be fed in directly.
Plain torch CPU ops (``F.conv2d`` ...), dtype-generic: cast ``state`` and inputs to float64 to arbitrate.

Follows:
* ResNet-18           /root/reference/mano_train/networks/bases/resnet.py:25-54,99-188
* ManoBranch.forward  /root/reference/mano_train/networks/branches/manobranch.py:115-218
* PointGenCon.forward /root/reference/mano_train/networks/branches/atlasutils.py:65-75
* AtlasBranch         /root/reference/mano_train/networks/branches/atlasbranch.py:78-150
* HandNet.forward     /root/reference/mano_train/networks/handnet.py:198-392
"""
import torch
import torch.nn.functional as F

from . import geometry, mano

BN_EPS = 1e-5


def _bn(x, state, prefix, training):
    w, b = state[prefix + ".weight"], state[prefix + ".bias"]
    rm, rv = state[prefix + ".running_mean"], state[prefix + ".running_var"]
    if training:
        return F.batch_norm(x, None, None, w, b, True, 0.0, BN_EPS)
    return F.batch_norm(x, rm, rv, w, b, False, 0.0, BN_EPS)


class _ReluWithMask(torch.autograd.Function):
    """relu(z) whose backward uses an externally supplied branch mask.  Test-only: lets a gradient check
    follow the ReLU branches the implementation under test actually took (a pre-activation within rounding
    of zero may legitimately fall on the other side), so that the comparison measures arithmetic, not flips."""

    @staticmethod
    def forward(ctx, z, mask):
        ctx.save_for_backward(mask)
        return z.clamp(min=0)

    @staticmethod
    def backward(ctx, g):
        (mask,) = ctx.saved_tensors
        return g * mask.to(g.dtype), None


def _relu(z, masks, name):
    if masks is None:
        return F.relu(z)
    return _ReluWithMask.apply(z, masks[name])


def _maxpool_with_idx(x, idx):
    """3x3/2 pad-1 max pooling that takes the window slot (dy*3+dx, NCHW int64) chosen elsewhere (tests only)."""
    xp = F.pad(x, (1, 1, 1, 1), value=float("-inf"))
    win = xp.unfold(2, 3, 2).unfold(3, 3, 2)  # (B,C,Ho,Wo,3,3)
    win = win.reshape(win.shape[:4] + (9,))
    return torch.gather(win, 4, idx.unsqueeze(4)).squeeze(4)


def resnet18_features(state, images, prefix="base_net", bn_training=False, relu_masks=None, pool_idx=None):
    """(B,3,H,W) -> (B,512); conv7x7/2, BN, ReLU, maxpool3/2, 4x2 BasicBlocks, spatial mean.
    ``relu_masks`` (tests only): {"c1", "a0".."a7", "out0".."out7"} -> bool NCHW branch masks;
    ``pool_idx`` (tests only): arg-max window slots of the max pooling (near-ties resolved as the caller did)."""
    p = prefix + "."
    x = F.conv2d(images, state[p + "conv1.weight"], None, stride=2, padding=3)
    x = _relu(_bn(x, state, p + "bn1", bn_training), relu_masks, "c1")
    if pool_idx is None:
        x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    else:
        x = _maxpool_with_idx(x, pool_idx)
    b = 0
    for li, stride in ((1, 1), (2, 2), (3, 2), (4, 2)):
        for bi in range(2):
            q = "{}layer{}.{}.".format(p, li, bi)
            s = stride if bi == 0 else 1
            out = F.conv2d(x, state[q + "conv1.weight"], None, stride=s, padding=1)
            out = _relu(_bn(out, state, q + "bn1", bn_training), relu_masks, "a%d" % b)
            out = F.conv2d(out, state[q + "conv2.weight"], None, stride=1, padding=1)
            out = _bn(out, state, q + "bn2", bn_training)
            if (q + "downsample.0.weight") in state:
                res = F.conv2d(x, state[q + "downsample.0.weight"], None, stride=s)
                res = _bn(res, state, q + "downsample.1", bn_training)
            else:
                res = x
            x = _relu(out + res, relu_masks, "out%d" % b)
            b += 1
    return x.mean(3).mean(2)


def mano_branch(state, tables_right, tables_left, feats, sides, root_palm=False, ncomps=30,
                center_idx=0, use_shape=True, prefix="mano_branch"):
    """ManoBranch.forward with use_pca=True, use_trans=False, adapt_skeleton=False, dropout=0."""
    p = prefix + "."
    x = feats
    i = 0
    while (p + "base_layer.{}.weight".format(i)) in state:
        x = F.relu(F.linear(x, state[p + "base_layer.{}.weight".format(i)],
                            state[p + "base_layer.{}.bias".format(i)]))
        i += 2
    pose = F.linear(x, state[p + "pose_reg.weight"], state[p + "pose_reg.bias"])
    shape = None
    if use_shape:
        shape = F.linear(x, state[p + "shape_reg.0.weight"], state[p + "shape_reg.0.bias"])
    B = feats.shape[0]
    verts = feats.new_zeros((B, 778, 3))
    joints = feats.new_zeros((B, 21, 3))
    is_right = torch.tensor([s == "right" for s in sides][:B])
    for flag, tables, side in ((is_right, tables_right, "right"), (~is_right, tables_left, "left")):
        if flag.sum() == 0:
            continue
        v, j = mano.mano_forward(tables, pose[flag], None if shape is None else shape[flag],
                                 trans=torch.zeros(1), root_palm=root_palm, side=side,
                                 center_idx=center_idx, use_pca=True, ncomps=ncomps)
        verts = verts.clone()
        joints = joints.clone()
        verts[flag] = v
        joints[flag] = j
    return {"verts": verts, "joints": joints, "shape": shape, "pose": pose}


def point_decoder(state, x, prefix, bn_training=False, out_factor=200):
    """PointGenCon: (B,515,N) -> (B,3,N)."""
    p = prefix + "."
    for i in (1, 2, 3):
        x = F.conv1d(x, state[p + "conv{}.weight".format(i)], state[p + "conv{}.bias".format(i)])
        x = F.relu(_bn(x, state, p + "bn{}".format(i), bn_training))
    x = F.conv1d(x, state[p + "conv4.weight"], state[p + "conv4.bias"])
    return out_factor * x


def _mlp2(state, prefix, x):
    h = F.relu(F.linear(x, state[prefix + ".0.weight"], state[prefix + ".0.bias"]))
    return F.linear(h, state[prefix + ".2.weight"], state[prefix + ".2.bias"])


def atlas_branch(state, feats, grid, faces=None, sep_feats=None, predict_trans=True,
                 predict_scale=True, mesh_mode=True, bn_training=False, out_factor=200,
                 prefix="atlas_branch"):
    """AtlasBranch.forward_inference (mesh_mode, grid = icosphere verts (N,3)) or
    AtlasBranch.forward (point mode; ``grid`` (B,N,3) is the injected unit-sphere sample, the
    reference draws it from the global RNG, atlasbranch.py:83-89; scale head and separate
    encoder ignored there)."""
    p = prefix + "."
    B = feats.shape[0]
    trans = _mlp2(state, p + "decode_trans", feats) if predict_trans else None
    if mesh_mode:
        scale = _mlp2(state, p + "decode_scale", feats) if predict_scale else None
        g = grid.to(feats.dtype).t().unsqueeze(0).expand(B, -1, -1)  # (B,3,N)
        dec_feat = sep_feats if sep_feats is not None else feats
    else:
        scale = None
        g = grid.to(feats.dtype).transpose(2, 1)
        dec_feat = feats
    x = torch.cat([g, dec_feat.unsqueeze(2).expand(-1, -1, g.shape[2])], 1)
    verts = point_decoder(state, x, p + "decoder", bn_training, out_factor).transpose(2, 1)
    res = {}
    pts = verts
    if scale is not None:
        pts = scale.unsqueeze(1) * verts
    if trans is not None:
        pts = pts + trans.unsqueeze(1)
        res.update({"objpoints3d": pts, "objtrans": trans, "objpointscentered3d": verts})
    else:
        res["objpoints3d"] = verts
    if mesh_mode:
        res["objfaces"] = faces
    if scale is not None:
        res["objscale"] = scale
    return res


def handnet_forward(state, cfg, sample, mano_tables, grid, faces, zones=None, bn_training=False):
    """HandNet.forward -> (total_loss, results, losses).

    cfg keys (HandNet kwargs, handnet.py:20-63): atlas_lambda, atlas_final_lambda, atlas_mesh,
    atlas_lambda_regul_edges, atlas_predict_trans, atlas_trans_weight, atlas_predict_scale,
    atlas_scale_weight, atlas_separate_encoder, atlas_out_factor, contact_target, contact_zones,
    contact_lambda, contact_thresh, contact_mode, collision_thresh, collision_mode,
    collision_lambda, mano_comps, mano_use_shape, mano_lambda_pose_reg, mano_center_idx,
    mano_lambda_joints3d, mano_lambda_verts, mano_lambda_shape.
    sample keys: images, sides, root, [joints3d], [verts3d], [objpoints3d].
    """
    g = cfg.get
    feats = resnet18_features(state, sample["images"], "base_net", bn_training)
    sep = None
    if g("atlas_separate_encoder", False):
        sep = resnet18_features(state, sample["images"], "atlas_base_net", bn_training)
    results, losses = {}, {}
    total = None
    mano_on = bool(g("mano_lambda_verts") or g("mano_lambda_joints3d"))
    if ("joints3d" in sample or "verts3d" in sample) and "sides" in sample and mano_on:
        mres = mano_branch(state, mano_tables["right"], mano_tables["left"], feats, sample["sides"],
                           root_palm=(sample["root"] == "palm"), ncomps=g("mano_comps", 6),
                           center_idx=g("mano_center_idx", 9), use_shape=g("mano_use_shape", False))
        if not g("no_loss", False):
            mtotal, ml = geometry.mano_loss(
                mres, sample.get("verts3d"), sample.get("joints3d"), g("mano_lambda_verts"),
                g("mano_lambda_joints3d"), g("mano_lambda_shape"), g("mano_lambda_pose_reg", 0))
            total = mtotal
            losses.update(ml)
        results.update(mres)
    if "objpoints3d" in sample and (g("atlas_lambda") or g("atlas_final_lambda")):
        mesh_mode = g("atlas_mesh", True)
        ares = atlas_branch(state, feats, grid, faces, sep if mesh_mode else None,
                            g("atlas_predict_trans", False),
                            g("atlas_predict_scale", False), mesh_mode, bn_training,
                            g("atlas_out_factor", 200))
        if g("contact_lambda", 0) or g("collision_lambda", 0):
            attr, pen, info, metrics = geometry.contact_loss(
                results["verts"], ares["objpoints3d"], faces, zones,
                contact_thresh=g("contact_thresh", 25), contact_mode=g("contact_mode", "dist_sq"),
                collision_thresh=g("collision_thresh", 25),
                collision_mode=g("collision_mode", "dist_sq"),
                contact_target=g("contact_target", "all"), contact_zones=g("contact_zones", "all"))
            if not g("no_loss", False):
                closs = g("contact_lambda", 0) * attr + g("collision_lambda", 0) * pen
                total = total + closs
                losses.update({"penetration_loss": pen, "attraction_loss": attr,
                               "contact_loss": closs})
                losses.update(metrics)
            results["contact_info"] = info
        results.update(ares)
        if not g("no_loss", False):
            atotal, al = geometry.atlas_loss(
                ares, sample["objpoints3d"], g("atlas_lambda") or 0, g("atlas_final_lambda") or 0,
                g("atlas_trans_weight", 1), g("atlas_scale_weight", 1),
                g("atlas_lambda_regul_edges", 0), g("atlas_lambda_laplacian", 0),
                geometry.laplacian_matrix(grid.detach().cpu().numpy(), faces) if g("atlas_lambda_laplacian", 0)
                else None)
            total = atotal if total is None else total + atotal
            losses.update(al)
    losses["total_loss"] = total
    return total, results, losses
