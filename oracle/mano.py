"""MANO layer restatement (oracle side; test infrastructure only).  PARITY UNPINNED.

``manopth`` (hassony2/manopth, un-pinned; install instruction at
/root/reference/README.md:48-50) is absent from /root/reference and from this
image.  This file restates its published algorithm (``manopth/manolayer.py``
``ManoLayer.forward``, ``rodrigues_layer.batch_rodrigues`` / ``quat2mat``,
``tensutils.th_posemap_axisang`` / ``subtract_flat_id`` / ``th_with_zeros``),
anchored on the reference's call sites:

* ctor   /root/reference/mano_train/networks/branches/manobranch.py:92-105
* call   /root/reference/mano_train/networks/branches/manobranch.py:170-182
  (``th_pose_coeffs, th_betas=, th_trans=Tensor([0]), root_palm=``)
* faces  /root/reference/mano_train/networks/branches/manobranch.py:113
* tips   /root/reference/mano_train/networks/branches/contactloss.py:258

Differentiable torch code, dtype-generic (run it in float64 to arbitrate).
"""
import torch

PARENTS = (-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 0, 10, 11, 0, 13, 14)
TIPS_RIGHT = (745, 317, 444, 556, 673)
TIPS_LEFT = (745, 317, 445, 556, 673)
JOINT_REORDER = (0, 13, 14, 15, 16, 1, 2, 3, 17, 4, 5, 6, 18, 10, 11, 12, 19, 7, 8, 9, 20)
PALM_VERTS = (95, 22)


def rodrigues(axisang):
    """(R,3) axis-angle -> (R,3,3); epsilon added per component before the norm."""
    angle = torch.sqrt(((axisang + 1e-8) ** 2).sum(1, keepdim=True))
    axis = axisang / angle
    half = angle * 0.5
    quat = torch.cat([torch.cos(half), torch.sin(half) * axis], dim=1)
    quat = quat / torch.sqrt((quat ** 2).sum(1, keepdim=True))
    w, x, y, z = quat[:, 0], quat[:, 1], quat[:, 2], quat[:, 3]
    rows = [
        w * w + x * x - y * y - z * z, 2 * x * y - 2 * w * z, 2 * w * y + 2 * x * z,
        2 * w * z + 2 * x * y, w * w - x * x + y * y - z * z, 2 * y * z - 2 * w * x,
        2 * x * z - 2 * w * y, 2 * w * x + 2 * y * z, w * w - x * x - y * y + z * z,
    ]
    return torch.stack(rows, dim=1).view(-1, 3, 3)


def mano_forward(tables, pose_coeffs, betas=None, trans=None, root_palm=False,
                 side="right", center_idx=None, use_pca=True, ncomps=None):
    """Return (verts (B,778,3) mm, joints (B,21,3) mm).

    ``tables``: dict of tensors named like the upstream registered buffers:
    th_shapedirs (778,3,10), th_posedirs (778,3,135), th_v_template (1,778,3),
    th_J_regressor (16,778), th_weights (778,16), th_hands_mean (1,45),
    th_selected_comps (C,45), th_betas (1,10).
    """
    dt = pose_coeffs.dtype
    T = {k: (v.to(dt) if torch.is_floating_point(v) else v) for k, v in tables.items()}
    B = pose_coeffs.shape[0]
    sel = T["th_selected_comps"]
    C = sel.shape[0] if ncomps is None else ncomps
    hand = pose_coeffs[:, 3:3 + C]
    if use_pca:
        hand = hand @ sel
    full_pose = torch.cat([pose_coeffs[:, :3], T["th_hands_mean"] + hand], dim=1)  # (B,48)
    rots = rodrigues(full_pose.reshape(-1, 3)).view(B, 16, 3, 3)
    eye = torch.eye(3, dtype=dt, device=pose_coeffs.device)
    pose_map = (rots[:, 1:] - eye).reshape(B, 135)

    if betas is None or betas.numel() == 1:
        betas = T["th_betas"].expand(B, 10)
    v_shaped = torch.einsum("vck,bk->bvc", T["th_shapedirs"], betas) + T["th_v_template"]
    joints = torch.einsum("jv,bvc->bjc", T["th_J_regressor"], v_shaped)  # (B,16,3)
    v_posed = v_shaped + torch.einsum("vck,bk->bvc", T["th_posedirs"], pose_map)

    # kinematic chain: world rotation / translation per joint
    Rw = [None] * 16
    tw = [None] * 16
    for j in range(16):
        p = PARENTS[j]
        if p < 0:
            Rw[j] = rots[:, 0]
            tw[j] = joints[:, 0]
        else:
            Rw[j] = Rw[p] @ rots[:, j]
            tw[j] = tw[p] + (Rw[p] @ (joints[:, j] - joints[:, p]).unsqueeze(2)).squeeze(2)
    Rw = torch.stack(Rw, 1)  # (B,16,3,3)
    tw = torch.stack(tw, 1)  # (B,16,3)
    # remove the rest-pose joint location from the translation column
    t_rel = tw - (Rw @ joints.unsqueeze(3)).squeeze(3)

    W = T["th_weights"]  # (778,16)
    R_v = torch.einsum("vj,bjrc->bvrc", W, Rw)
    t_v = torch.einsum("vj,bjr->bvr", W, t_rel)
    verts = (R_v @ v_posed.unsqueeze(3)).squeeze(3) + t_v

    tips = TIPS_RIGHT if side == "right" else TIPS_LEFT
    jtr = tw
    if bool(root_palm):
        palm = (verts[:, PALM_VERTS[0]] + verts[:, PALM_VERTS[1]]).unsqueeze(1) / 2
        jtr = torch.cat([palm, jtr[:, 1:]], 1)
    jtr = torch.cat([jtr, verts[:, list(tips)]], 1)[:, list(JOINT_REORDER)]

    if trans is None or bool(torch.norm(trans) == 0):
        if center_idx is not None:
            c = jtr[:, center_idx].unsqueeze(1)
            jtr = jtr - c
            verts = verts - c
    else:
        jtr = jtr + trans.unsqueeze(1)
        verts = verts + trans.unsqueeze(1)
    return verts * 1000, jtr * 1000
