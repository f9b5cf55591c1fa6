"""Constant data of the hot path: contact zones / MANO template fixture.

The reference reads ``assets/contact_zones.pkl`` relative to the working directory
(/root/reference/handobjectdatasets/contactutils.py:8-45, used at
/root/reference/mano_train/networks/branches/contactloss.py:262-274).  The same lookup is kept
(so the path drops in under ``traineval.py``); when that file is absent the packaged,
pickle-free copy ``obman_train_b200/assets/contact_zones.npz`` (made by scripts/make_assets.py)
is used.
"""
import os
import pickle
from functools import lru_cache

import numpy as np

_PKG_NPZ = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets", "contact_zones.npz")


@lru_cache(maxsize=8)
def load_contacts(save_contact_paths="assets/contact_zones.pkl"):
    """Return (hand_verts (778,3) float64 metres, {zone_id: [vertex ids]})."""
    if os.path.exists(save_contact_paths):
        with open(save_contact_paths, "rb") as p_f:
            data = pickle.load(p_f, encoding="latin1")
        return data["verts"], {k: [int(i) for i in v] for k, v in data["contact_zones"].items()}
    data = np.load(_PKG_NPZ)
    ptr = data["zone_ptr"]
    ids = data["zone_ids"]
    zones = {z: [int(i) for i in ids[ptr[z]:ptr[z + 1]]] for z in range(len(ptr) - 1)}
    return data["verts"], zones


@lru_cache(maxsize=2)
def template_mesh():
    """MANO right-hand template (verts (778,3) float64 metres, faces (1538,3) int64)."""
    data = np.load(_PKG_NPZ)
    return data["verts"], data["faces"]
