"""Mirror of mano_train/networks/branches/manobranch.py: MANO regression branch and its loss.

Same constructor / forward signature and result keys as the reference
(/root/reference/mano_train/networks/branches/manobranch.py:12-218) and the same ManoLoss.compute_loss
contract (:229-324).  The MLP runs on the tcgen05 GEMM kernel, the left/right ManoLayers on the fused MANO
kernels.  Out of scope (never enabled by the README recipe): use_pca=False (rotation-matrix output with SVD
projection), use_trans, adapt_skeleton, dropout, normalize_hand, PCA supervision.
"""
import torch
from torch import nn
from ... import losshead, mlp
from ...manopth.manolayer import ManoLayer
from ...queries import TransQueries, BaseQueries


class _ReluMLP(nn.Sequential):
    """nn.Sequential([Linear, ReLU] * k) evaluated with fused Linear+ReLU GEMMs."""

    def forward(self, x):
        mods = list(self)
        i = 0
        while i < len(mods):
            if isinstance(mods[i], nn.Linear):
                relu = i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU)
                x = mlp.linear(x, mods[i].weight, mods[i].bias, relu=relu)
                i += 2 if relu else 1
            else:
                x = mods[i](x)
                i += 1
        return x


class ManoBranch(nn.Module):
    def __init__(self, ncomps=6, base_neurons=[1024, 512], center_idx=9, use_shape=False, use_trans=False,
                 use_pca=True, mano_root="misc/mano", adapt_skeleton=True, dropout=0):
        super(ManoBranch, self).__init__()
        if not use_pca:
            raise NotImplementedError("use_pca=False (rotation-matrix regression) is not on the hot path")
        if use_trans or dropout:
            raise NotImplementedError("use_trans / dropout are not on the hot path (HandNet builds "
                                      "ManoBranch with use_trans=False, handnet.py:137; fc_dropout defaults to 0)")
        self.adapt_skeleton = adapt_skeleton
        self.use_trans = use_trans
        self.use_shape = use_shape
        self.use_pca = use_pca
        self.register_buffer("stereo_shape", torch.Tensor(
            [-0.00298099, -0.0013994, -0.00840144, 0.00362311, 0.00248761, 0.00044125, 0.00381337,
             -0.00183374, -0.00149655, 0.00137479]), persistent=False)
        mano_pose_size = ncomps + 3
        base_layers = []
        for inp_neurons, out_neurons in zip(base_neurons[:-1], base_neurons[1:]):
            base_layers.append(nn.Linear(inp_neurons, out_neurons))
            base_layers.append(nn.ReLU())
        self.base_layer = _ReluMLP(*base_layers)
        self.pose_reg = nn.Linear(base_neurons[-1], mano_pose_size)
        if self.use_shape:
            self.shape_reg = torch.nn.Sequential(nn.Linear(base_neurons[-1], 10))
        self.mano_layer_right = ManoLayer(ncomps=ncomps, center_idx=center_idx, side="right",
                                          mano_root=mano_root, use_pca=use_pca)
        self.mano_layer_left = ManoLayer(ncomps=ncomps, center_idx=center_idx, side="left",
                                         mano_root=mano_root, use_pca=use_pca)
        if self.adapt_skeleton:
            joint_nb = 21
            self.left_skeleton_reg = nn.Linear(joint_nb, joint_nb, bias=False)
            self.left_skeleton_reg.weight.data = torch.eye(joint_nb)
            self.right_skeleton_reg = nn.Linear(joint_nb, joint_nb, bias=False)
            self.right_skeleton_reg.weight.data = torch.eye(joint_nb)
        self.faces = self.mano_layer_right.th_faces

    def _side_indices(self, flags, device):
        """Row indices of the right / left hands, cached per side pattern (built from the Python list: no
        device synchronisation, and safe to call while a CUDA graph is being captured once warmed up)."""
        cache = self.__dict__.setdefault("_side_cache", {})
        key = (flags, str(device))
        if key not in cache:
            cache[key] = (torch.tensor([i for i, f in enumerate(flags) if f], dtype=torch.long, device=device),
                          torch.tensor([i for i, f in enumerate(flags) if not f], dtype=torch.long, device=device))
        return cache[key]

    def forward(self, inp, sides, root_palm=False, shape=None, pose=None, use_stereoshape=False, side_mask=None):
        """``side_mask`` (extension, optional): device bool tensor (B,), True = right hand.  When given, BOTH MANO
        layers run on the whole batch and the result is selected per sample on the device, so the launch sequence does
        not depend on the left/right pattern of the batch (a captured CUDA graph stays valid for any ``sides``); the
        MANO layer is 0.55 MMAC per sample, the duplicate is cheaper than a re-capture.  Without it the reference's
        gather / scatter by side (manobranch.py:133-207) is used."""
        base_features = self.base_layer(inp)
        pose = mlp.linear(base_features, self.pose_reg.weight, self.pose_reg.bias)
        mano_pose = pose
        B = pose.shape[0]
        if side_mask is not None and not use_stereoshape:
            if self.use_shape:
                shape = mlp.linear(base_features, self.shape_reg[0].weight, self.shape_reg[0].bias)
            else:
                shape = None
            trans = torch.Tensor([0])
            verts_r, joints_r = self.mano_layer_right(mano_pose, th_betas=shape, th_trans=trans, root_palm=root_palm)
            verts_l, joints_l = self.mano_layer_left(mano_pose, th_betas=shape, th_trans=trans, root_palm=root_palm)
            if self.adapt_skeleton:
                joints_r = torch.einsum("ij,bjc->bic", self.right_skeleton_reg.weight, joints_r)
                joints_l = torch.einsum("ij,bjc->bic", self.left_skeleton_reg.weight, joints_l)
            m = side_mask[:B].to(torch.bool).view(B, 1, 1)
            return {"verts": torch.where(m, verts_r, verts_l), "joints": torch.where(m, joints_r, joints_l),
                    "shape": shape, "pose": pose}
        flags = [side == "right" for side in sides][:B]
        n_right = int(sum(flags))
        n_left = B - n_right
        if use_stereoshape:
            shape = self.stereo_shape.unsqueeze(0).repeat(B, 1)
            assert n_right == 0, "When stereoshape is used only left hands expected"
        elif self.use_shape:
            shape = mlp.linear(base_features, self.shape_reg[0].weight, self.shape_reg[0].bias)
        else:
            shape = None
        trans = torch.Tensor([0])
        def adapt(joints_, reg):
            # 21x21 joint-mixing map of the reference (manobranch.py:183-191), identity at init
            return torch.einsum("ij,bjc->bic", reg.weight, joints_) if self.adapt_skeleton else joints_

        if n_right == B or n_left == B:
            layer = self.mano_layer_right if n_right == B else self.mano_layer_left
            verts, joints = layer(mano_pose, th_betas=shape, th_trans=trans, root_palm=root_palm)
            if self.adapt_skeleton:
                joints = adapt(joints, self.right_skeleton_reg if n_right == B else self.left_skeleton_reg)
        else:
            idx_r, idx_l = self._side_indices(tuple(flags), inp.device)
            verts_r, joints_r = self.mano_layer_right(
                mano_pose[idx_r], th_betas=None if shape is None else shape[idx_r], th_trans=trans,
                root_palm=root_palm)
            verts_l, joints_l = self.mano_layer_left(
                mano_pose[idx_l], th_betas=None if shape is None else shape[idx_l], th_trans=trans,
                root_palm=root_palm)
            joints_r = adapt(joints_r, self.right_skeleton_reg if self.adapt_skeleton else None)
            joints_l = adapt(joints_l, self.left_skeleton_reg if self.adapt_skeleton else None)
            verts = inp.new_zeros((B, 778, 3)).index_copy(0, idx_r, verts_r).index_copy(0, idx_l, verts_l)
            joints = inp.new_zeros((B, 21, 3)).index_copy(0, idx_r, joints_r).index_copy(0, idx_l, joints_l)
        results = {"verts": verts, "joints": joints, "shape": shape, "pose": pose}
        return results


class ManoLoss:
    """Weighted sum of the MANO supervision terms (manobranch.py:229-324): vertices, 3-D joints, shape and pose
    regularisers, optional PCA supervision - every active term in ONE fused kernel launch (losshead.sq_terms) whose
    weights are read from device memory.  ``compute_loss(preds, target) -> (final_loss (1,), mano_losses)`` with the
    reference's keys: "mano_verts3d", "mano_shape", "mano_pca" are always reported (None when off), "mano_joints3d"
    and "pose_reg" only when on, plus "mano_total_loss"."""

    # (attribute holding the weight, logged key, always reported?)
    _TERMS = (("lambda_verts", "mano_verts3d", True), ("lambda_joints3d", "mano_joints3d", False),
              ("lambda_shape", "mano_shape", True), ("lambda_pose_reg", "pose_reg", False),
              ("lambda_pca", "mano_pca", True))

    def __init__(self, lambda_verts=None, lambda_joints3d=None, lambda_shape=None, lambda_pose_reg=None,
                 lambda_pca=None, center_idx=9, normalize_hand=False):
        self.lambda_verts = lambda_verts
        self.lambda_joints3d = lambda_joints3d
        self.lambda_shape = lambda_shape
        self.lambda_pose_reg = lambda_pose_reg
        self.lambda_pca = lambda_pca
        self.center_idx = center_idx
        self.normalize_hand = normalize_hand
        self._weights = losshead.LossWeights([attr for attr, _, _ in self._TERMS])
        self._scratch = losshead._Workspace()

    def _active(self, preds, target):
        """(weight attribute, prediction, target or None, column range or None) of every term that is switched on."""
        out = []
        if TransQueries.verts3d in target and self.lambda_verts:
            out.append(("lambda_verts", preds["verts"], target[TransQueries.verts3d], None))
        if TransQueries.joints3d in target and self.lambda_joints3d:
            out.append(("lambda_joints3d", preds["joints"], target[TransQueries.joints3d], None))
        if self.lambda_shape:
            out.append(("lambda_shape", preds["shape"], None, None))
        if self.lambda_pose_reg:
            # the three global-rotation coefficients are not regularised (manobranch.py:307-312)
            out.append(("lambda_pose_reg", preds["pose"], None, (3, preds["pose"].shape[1])))
        if BaseQueries.hand_pcas in target and self.lambda_pca:
            out.append(("lambda_pca", preds["pcas"], target[BaseQueries.hand_pcas], None))
        return out

    def compute_loss(self, preds, target):
        dev = preds["pose"].device
        for attr, _, _ in self._TERMS:
            self._weights[attr] = getattr(self, attr)
        mano_losses = {key: None for _, key, always in self._TERMS if always}
        active = self._active(preds, target)
        if active:
            final_loss, values = losshead.sq_terms(
                [(pred, tgt, cols, self._weights.slot[attr]) for attr, pred, tgt, cols in active],
                self._weights.device(dev), self._scratch)
            key_of = {attr: key for attr, key, _ in self._TERMS}
            for k, (attr, _, _, _) in enumerate(active):
                mano_losses[key_of[attr]] = values[k:k + 1]
        else:
            final_loss = torch.zeros(1, device=dev)
        mano_losses["mano_total_loss"] = final_loss
        return final_loss, mano_losses
