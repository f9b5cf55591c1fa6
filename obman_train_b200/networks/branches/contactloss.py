"""Mirror of mano_train/networks/branches/contactloss.py: contact attraction / repulsion loss.

Same function names, argument meaning, return structure and error behaviour as the reference
(/root/reference/mano_train/networks/branches/contactloss.py:50-79,149-308); the arithmetic is the
fused nearest-neighbour + ray-parity + value/mask kernels of libobman_b200.so.
"""
import numpy as np
import torch

from ... import functional as F_b200
from ...assets import load_contacts
from .contactutils import batch_mesh_contains_points  # noqa: F401

TIP_IDXS = [745, 317, 444, 556, 673]
_cache = {}


def _device_faces(obj_faces, device):
    arr = np.ascontiguousarray(np.asarray(obj_faces).astype(np.int32))
    key = ("faces", arr.shape, hash(arr.tobytes()), str(device))
    if key not in _cache:
        _cache[key] = torch.from_numpy(arr).to(device)
    return _cache[key]


def _device_zone_table(mode, device):
    key = ("zones", mode, str(device))
    if key not in _cache:
        if mode == "tips":
            ids, ptr = TIP_IDXS, [0, len(TIP_IDXS)]
        else:
            _, zones = load_contacts("assets/contact_zones.pkl")
            ids, ptr = [], [0]
            for z in sorted(zones.keys()):
                ids.extend(int(i) for i in zones[z])
                ptr.append(len(ids))
        _cache[key] = (torch.tensor(ids, dtype=torch.int32, device=device),
                       torch.tensor(ptr, dtype=torch.int32, device=device))
    return _cache[key]


def batch_pairwise_dist(x, y, use_cuda=True):
    """(B,Nx,3), (B,Ny,3) -> (B,Nx,Ny) squared distances (contactloss.py:60-79).  Kept for callers
    that want the full matrix (HandNet's GT contact metric); the losses never materialise it."""
    return ((x.unsqueeze(2) - y.unsqueeze(1)) ** 2).sum(-1)


def masked_mean_loss(dists, mask):
    """contactloss.py:50-57 (batch-global; grad-less zero when the mask is empty)."""
    mask = mask.float()
    valid_vals = mask.sum()
    if valid_vals > 0:
        return (mask * dists).sum() / valid_vals
    return torch.zeros(1, device=dists.device)


def thresh_ious(gt_dists, pred_dists, thresh):
    gt_contacts = gt_dists <= thresh
    pred_contacts = pred_dists <= thresh
    inter = (gt_contacts & pred_contacts).sum(1).float()
    union = (gt_contacts | pred_contacts).sum(1).float()
    return torch.where(union != 0, inter / union.clamp(min=1), torch.zeros_like(union))


def meshiou(gt_dists, pred_dists, threshs=(1, 2, 3, 4, 5, 6, 7, 8, 9, 10)):
    """contactloss.py:35-47: IoU of thresholded contact maps, averaged over the batch, and its AUC - two kernel
    launches (csrc/contact.cu) instead of ~12 ATen launches per threshold and a host-side ``np.trapz``.  The AUC is
    returned as a 0-dim device tensor (``float()`` / ``.item()`` give the number; the reference returns a numpy float
    after a device synchronisation)."""
    return F_b200.contact_iou(gt_dists, pred_dists, threshs)


def compute_contact_loss(hand_verts_pt, hand_faces, obj_verts_pt, obj_faces, contact_thresh=5,
                         contact_mode="dist_sq", collision_thresh=10, collision_mode="dist_sq",
                         contact_target="all", contact_sym=False, contact_zones="all"):
    """Returns (missed_loss, penetr_loss, contact_info, metrics) like contactloss.py:149-308."""
    if contact_target not in F_b200.CONTACT_TARGETS:
        raise ValueError("contact_target {} not in [all|obj|hand]".format(contact_target))
    if contact_mode not in F_b200.CONTACT_MODES:
        raise ValueError("contact_mode {} not in [dist_sq|dist|dist_tanh]".format(contact_mode))
    if collision_mode not in F_b200.CONTACT_MODES:
        raise ValueError("collision_mode {} not in [dist_sq|dist|dist_tanh]".format(collision_mode))
    if contact_zones not in F_b200.CONTACT_ZONES:
        raise ValueError("contact_zones {} not in [tips|zones|all]".format(contact_zones))
    dev = hand_verts_pt.device
    faces = _device_faces(obj_faces, dev)
    zone_ids = zone_ptr = None
    if contact_zones != "all":
        zone_ids, zone_ptr = _device_zone_table(contact_zones, dev)
    missed_loss, penetr_loss, attr, rep, close, mins21, stats = F_b200.contact_loss(
        hand_verts_pt, obj_verts_pt, faces, zone_ids, zone_ptr, contact_thresh, contact_mode,
        collision_thresh, collision_mode, contact_target, contact_zones)
    if contact_sym:
        _, _, mins12, _ = F_b200.nearest_neighbours(hand_verts_pt, obj_verts_pt, dirs=2)
        missed_loss = missed_loss + masked_mean_loss(torch.sqrt(mins12), mins12 < contact_thresh)
    contact_info = {
        "attraction_masks": attr.bool(),
        "repulsion_masks": rep.bool(),
        "contact_points": close,
        "min_dists": mins21,
    }
    metrics = {"max_penetr": stats[2], "mean_penetr": stats[3]}
    return missed_loss, penetr_loss, contact_info, metrics
