"""Drop-in for ``manopth.manolayer.ManoLayer`` backed by the fused sm_100a MANO kernels.

Keeps the constructor / forward signature, the ``th_*`` buffer names (they are state-dict keys in the
released checkpoints) and the return convention ``(verts_mm (B,778,3), joints_mm (B,21,3))``
of the external dependency the reference calls at
/root/reference/mano_train/networks/branches/manobranch.py:92-105 (ctor) and :170-182 (forward).
"""
import os
import pickle

import numpy as np
import torch
from torch import nn

from .. import functional as F_b200
from .synthetic import synthetic_mano_tables


class _ChStub(object):
    """Stand-in for chumpy.Ch objects inside the licensed MANO pickles (chumpy is not required)."""

    def __setstate__(self, state):
        self.__dict__.update(state if isinstance(state, dict) else {"x": state})

    @property
    def r(self):
        return np.asarray(self.__dict__.get("x"))


class _ManoUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.startswith("chumpy"):
            return _ChStub
        return super().find_class(module, name)


def _to_np(v):
    if isinstance(v, _ChStub):
        return v.r
    if hasattr(v, "toarray"):
        return np.asarray(v.toarray())
    return np.asarray(v)


def load_mano_pickle(path):
    with open(path, "rb") as f:
        data = _ManoUnpickler(f, encoding="latin1").load()
    keys = ("v_template", "shapedirs", "posedirs", "J_regressor", "weights", "hands_mean",
            "hands_components", "f")
    out = {k: _to_np(data[k]) for k in keys}
    out["betas"] = np.zeros(10)
    return out


class ManoLayer(nn.Module):
    __constants__ = ["use_pca", "rot", "ncomps", "ncomps", "kintree_parents", "check", "side",
                     "center_idx", "joint_rot_mode"]

    def __init__(self, center_idx=None, flat_hand_mean=True, ncomps=6, side="right",
                 mano_root="mano/models", use_pca=True, root_rot_mode="axisang",
                 joint_rot_mode="axisang", robust_rot=False, tables=None):
        """``tables`` (dict of numpy arrays, see synthetic_mano_tables) overrides the pickle lookup;
        ``mano_root="synthetic"`` (or ``"synthetic:<seed>"``) builds synthetic tables explicitly."""
        super().__init__()
        if root_rot_mode != "axisang" or joint_rot_mode != "axisang":
            raise NotImplementedError("only axis-angle rotation modes are on the hot path "
                                      "(README training recipe uses --mano_use_pca)")
        self.center_idx = center_idx
        self.robust_rot = robust_rot
        self.rot = 3
        self.flat_hand_mean = flat_hand_mean
        self.side = side
        self.use_pca = use_pca
        self.joint_rot_mode = joint_rot_mode
        self.root_rot_mode = root_rot_mode
        self.ncomps = ncomps if use_pca else 45
        if tables is None:
            if str(mano_root).startswith("synthetic"):
                seed = int(mano_root.split(":")[1]) if ":" in mano_root else 0
                tables = synthetic_mano_tables(side, seed)
            else:
                fname = "MANO_RIGHT.pkl" if side == "right" else "MANO_LEFT.pkl"
                self.mano_path = os.path.join(mano_root, fname)
                if not os.path.exists(self.mano_path):
                    raise FileNotFoundError(
                        "{} not found (licensed MANO model); pass mano_root='synthetic' for "
                        "synthetic tables".format(self.mano_path))
                tables = load_mano_pickle(self.mano_path)
        comps = np.asarray(tables["hands_components"], dtype=np.float64)
        hands_mean = np.zeros(comps.shape[1]) if flat_hand_mean else np.asarray(tables["hands_mean"]).copy()
        f32 = lambda a: torch.tensor(np.asarray(a), dtype=torch.float32)  # noqa: E731
        self.register_buffer("th_betas", f32(tables["betas"]).view(1, 10))
        self.register_buffer("th_shapedirs", f32(tables["shapedirs"]))
        self.register_buffer("th_posedirs", f32(tables["posedirs"]))
        self.register_buffer("th_v_template", f32(tables["v_template"]).unsqueeze(0))
        self.register_buffer("th_J_regressor", f32(tables["J_regressor"]))
        self.register_buffer("th_weights", f32(tables["weights"]))
        self.register_buffer("th_faces", torch.tensor(np.asarray(tables["f"]).astype(np.int32)).long())
        self.register_buffer("th_hands_mean", f32(hands_mean).unsqueeze(0))
        self.register_buffer("th_comps", f32(comps))
        sel = comps[:ncomps] if use_pca else np.eye(45)
        self.register_buffer("th_selected_comps", f32(sel))
        self.kintree_parents = [-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 0, 10, 11, 0, 13, 14]
        self._cache_key = None
        self._cache = None

    def _device_tables(self):
        bufs = (self.th_v_template, self.th_shapedirs, self.th_posedirs, self.th_weights,
                self.th_J_regressor, self.th_hands_mean, self.th_selected_comps, self.th_betas)
        key = tuple((b.data_ptr(), b._version) for b in bufs)
        if key != self._cache_key:
            with torch.no_grad():
                vt = self.th_v_template[0].contiguous()
                jreg = self.th_J_regressor
                self._cache = {
                    "v_template": vt,
                    "shapedirs": self.th_shapedirs.contiguous(),
                    "posedirs": self.th_posedirs.contiguous(),
                    "weights": self.th_weights.contiguous(),
                    "j_template": (jreg.double() @ vt.double()).float().contiguous(),
                    "j_shapedirs": torch.einsum("jv,vck->jck", jreg.double(),
                                                self.th_shapedirs.double()).float().contiguous(),
                    "hands_mean": self.th_hands_mean.reshape(-1).contiguous(),
                    "comps": self.th_selected_comps.contiguous(),
                    "betas": self.th_betas.reshape(-1).contiguous(),
                }
            self._cache_key = key
        return self._cache

    def forward(self, th_pose_coeffs, th_betas=torch.zeros(1), th_trans=torch.zeros(1),
                root_palm=torch.Tensor([0]), share_betas=torch.Tensor([0])):
        """th_pose_coeffs (B, 3+ncomps); th_betas (B,10) or None / 1 element => layer betas;
        th_trans (B,3) or None / zero => centre on ``center_idx``.  Returns (verts, joints) in mm."""
        tables = self._device_tables()
        betas = None
        if th_betas is not None and th_betas.numel() != 1:
            betas = th_betas
            if bool(share_betas):
                betas = betas.mean(0, keepdim=True).expand(betas.shape[0], 10)
        trans = None
        if th_trans is not None and bool(torch.norm(th_trans) != 0):
            trans = th_trans.to(th_pose_coeffs.device)
            if trans.requires_grad:
                raise NotImplementedError("gradient w.r.t. th_trans is not on the hot path "
                                          "(ManoBranch is built with use_trans=False, handnet.py:137)")
        cidx = -1 if self.center_idx is None else int(self.center_idx)
        return F_b200.mano_layer(th_pose_coeffs, betas, trans, tables, self.side != "right",
                                 bool(root_palm), cidx)
