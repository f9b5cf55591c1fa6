"""Generate tests/golden/*.npz by RUNNING THE REFERENCE's own code (container only).

The reference ships no tests and no golden vectors, so the fixtures are produced by
executing its unmodified files from /root/reference on CPU through ``oracle.refhook`` on seeded
inputs.  They pin ``oracle.geometry`` on machines where /root/reference is absent (the GPU box) and are
used directly by the CUDA parity tests.

    python scripts/make_golden.py
"""
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from oracle import icosphere, refhook  # noqa: E402
from obman_train_b200.assets import load_contacts  # noqa: E402
from obman_train_b200.manopth.synthetic import synthetic_mano_tables  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    refhook.set_mano_tables(synthetic_mano_tables("right"), synthetic_mano_tables("left"))
    refhook.install()
    from mano_train.networks.branches import atlasutils, contactloss, contactutils
    from mano_train.networks.branches.atlasbranch import edge_loss

    os.makedirs(OUT, exist_ok=True)
    g = torch.Generator().manual_seed(1234)
    out = {}

    # ---- Chamfer (atlasutils.py:11-18), incl. ragged N != M and a tiny cloud --------------------------
    for tag, (B, N, M) in {"a": (2, 50, 40), "b": (3, 33, 129), "c": (1, 1, 7)}.items():
        preds = torch.randn(B, N, 3, generator=g) * 40
        gts = torch.randn(B, M, 3, generator=g) * 40 + 30
        preds.requires_grad_(True)
        l1, l2 = atlasutils.ChamferLoss()(preds, gts)
        (l1 + l2).mean().backward()
        out["chamfer_%s_preds" % tag] = preds.detach().numpy()
        out["chamfer_%s_gts" % tag] = gts.numpy()
        out["chamfer_%s_loss1" % tag] = l1.detach().numpy()
        out["chamfer_%s_loss2" % tag] = l2.detach().numpy()
        out["chamfer_%s_gpreds" % tag] = preds.grad.numpy()

    # ---- contact loss (contactloss.py:149-308) on a hand-shaped cloud vs a noisy sphere ---------------
    verts, _ = load_contacts()
    sv, sf = icosphere.icosphere(2)  # 162 verts / 320 faces
    B = 2
    hand = torch.tensor(verts * 1000, dtype=torch.float32).unsqueeze(0).repeat(B, 1, 1)
    hand = hand + torch.randn(B, 1, 3, generator=g) * 5
    obj = (torch.tensor(sv, dtype=torch.float32).unsqueeze(0) * 40 + torch.tensor([30.0, 0.0, 20.0])
           + torch.randn(B, sv.shape[0], 3, generator=g))
    out["contact_hand"] = hand.numpy()
    out["contact_obj"] = obj.numpy()
    out["contact_faces"] = sf.astype(np.int64)
    with refhook.cwd():
        tri = obj[:, torch.as_tensor(sf)]
        out["contact_exterior"] = contactutils.batch_mesh_contains_points(hand, tri).numpy()
        for zones_mode in ("all", "tips", "zones"):
            for mode in ("dist_sq", "dist", "dist_tanh"):
                for target in ("all", "obj", "hand"):
                    if target != "all" and (zones_mode != "zones" or mode != "dist_tanh"):
                        continue
                    h = hand.clone().requires_grad_(True)
                    o = obj.clone().requires_grad_(True)
                    missed, penetr, info, metrics = contactloss.compute_contact_loss(
                        h, None, o, sf, contact_thresh=10, contact_mode=mode, collision_thresh=20,
                        collision_mode=mode, contact_target=target, contact_zones=zones_mode)
                    (missed + 0.5 * penetr).sum().backward()
                    k = "contact_%s_%s_%s_" % (zones_mode, mode, target)
                    out[k + "missed"] = missed.detach().numpy().reshape(-1)
                    out[k + "penetr"] = penetr.detach().numpy().reshape(-1)
                    out[k + "attr"] = info["attraction_masks"].numpy()
                    out[k + "rep"] = info["repulsion_masks"].numpy()
                    out[k + "max_penetr"] = metrics["max_penetr"].numpy().reshape(-1)
                    out[k + "mean_penetr"] = metrics["mean_penetr"].numpy().reshape(-1)
                    out[k + "ghand"] = h.grad.numpy() if h.grad is not None else np.zeros_like(hand.numpy())
                    out[k + "gobj"] = o.grad.numpy() if o.grad is not None else np.zeros_like(obj.numpy())
    out["edge_loss"] = edge_loss(obj, sf).numpy().reshape(-1)

    np.savez_compressed(os.path.join(OUT, "geometry_golden.npz"), **out)
    print("wrote", os.path.join(OUT, "geometry_golden.npz"), len(out), "arrays")


if __name__ == "__main__":
    main()
