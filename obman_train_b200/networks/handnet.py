"""Drop-in for mano_train/networks/handnet.py: the per-image training graph of obman_train.

Boundary: ``HandNet(**kwargs)`` takes the reference's keyword arguments (handnet.py:20-63),
``forward(sample, no_loss=False, return_features=False, force_objects=False) -> (total_loss, results, losses)`` returns
the reference's keys (:198-392), the attributes the training scripts read (``base_net``, ``atlas_base_net``,
``atlas_branch.decoder``, ``decay_regul``, ``mano_branch.faces``, ``atlas_branch.test_faces/test_verts``) and the
state-dict keys are the reference's, so it drops in under traineval.py / epochpass3d.py.

Schedule of one forward (every arithmetic stage is a kernel of libobman_b200.so):

    encoder(s)            ResNet-18 on the tcgen05 convolution kernels                      encoder.py
    hand lane (BRANCH)    ManoBranch MLP -> MANO layers -> ManoLoss (one fused kernel)      own stream, next to:
    object lane           AtlasNet decoder (+ translation / scale heads)
    contact               nearest neighbours + ray parity + value / mask kernels            needs both lanes
    object loss           GT statistics, Chamfer x2, regression terms, regularisers
    total                 one weighted-sum kernel over the lanes' results (device-side lambdas)

Not on the hot path and rejected at construction: ResNet-50, the absolute / 2-D joint branches
(``absolute_lambda``, ``mano_lambda_joints2d``; dead or default-off in the reference),
the residual decoder, ``fc_dropout``.
"""
from copy import deepcopy

import torch
from torch import nn

from .. import functional as F_b200
from .. import losshead, mlp, streams
from ..queries import TransQueries, BaseQueries
from .bases import resnet
from .branches.manobranch import ManoBranch, ManoLoss
from .branches.atlasbranch import AtlasBranch, AtlasLoss
from .branches.contactloss import compute_contact_loss, meshiou

FEATURE_SIZE = {18: 512}


class HandNet(nn.Module):
    def __init__(self, absolute_lambda=None, atlas_lambda=None, atlas_loss="chamfer", atlas_final_lambda=None,
                 atlas_mesh=True, atlas_residual=False, atlas_lambda_regul_edges=0, atlas_lambda_laplacian=0,
                 atlas_points_nb=600, atlas_predict_trans=False, atlas_trans_weight=1,
                 atlas_predict_scale=False, atlas_scale_weight=1, atlas_use_tanh=False, atlas_ico_divisions=3,
                 atlas_separate_encoder=False, atlas_out_factor=200, contact_target="all",
                 contact_zones="all", contact_lambda=0, contact_thresh=25, contact_mode="dist_sq",
                 collision_thresh=25, collision_mode="dist_sq", collision_lambda=0, fc_dropout=0,
                 resnet_version=50, mano_adapt_skeleton=False, mano_neurons=[512], mano_comps=6,
                 mano_use_shape=False, mano_lambda_pose_reg=0, mano_use_pca=True, mano_center_idx=9,
                 mano_root="misc/mano", mano_lambda_joints3d=None, mano_lambda_joints2d=None,
                 mano_lambda_verts=None, mano_lambda_shape=None, mano_lambda_pca=None,
                 adapt_atlas_decoder=False):
        super(HandNet, self).__init__()
        if int(resnet_version) not in FEATURE_SIZE:
            raise NotImplementedError("Resnet {} not supported on the B200 hot path (ResNet-18 only)".format(resnet_version))
        if absolute_lambda or mano_lambda_joints2d:
            raise NotImplementedError("absolute / 2-D joint branches are not on the hot path")
        width = FEATURE_SIZE[int(resnet_version)]
        # hyper-parameters under the reference's attribute names (traineval.py / reload.py read some of them)
        for name, value in (("absolute_lambda", absolute_lambda), ("adapt_atlas_decoder", adapt_atlas_decoder),
                            ("atlas_separate_encoder", atlas_separate_encoder), ("atlas_mesh", atlas_mesh),
                            ("atlas_lambda", atlas_lambda), ("atlas_final_lambda", atlas_final_lambda),
                            ("atlas_trans_weight", atlas_trans_weight), ("atlas_scale_weight", atlas_scale_weight),
                            ("contact_target", contact_target), ("contact_zones", contact_zones),
                            ("contact_lambda", contact_lambda), ("contact_thresh", contact_thresh),
                            ("contact_mode", contact_mode), ("collision_lambda", collision_lambda),
                            ("collision_thresh", collision_thresh), ("collision_mode", collision_mode),
                            ("mano_adapt_skeleton", mano_adapt_skeleton), ("lambda_joints2d", mano_lambda_joints2d)):
            setattr(self, name, value)
        self.need_collisions = bool(contact_lambda or collision_lambda)
        self.mano_lambdas = bool(mano_lambda_verts or mano_lambda_joints3d or mano_lambda_joints2d or mano_lambda_pca)
        # sub-modules in the reference's construction order (= RNG order of a seeded init) and registration order
        # (= state-dict order)
        encoder = resnet.resnet18(pretrained=True)
        if adapt_atlas_decoder:
            self.atlas_adapter = torch.nn.Linear(width, width)
        self.base_net = encoder
        if atlas_separate_encoder:
            self.atlas_base_net = deepcopy(self.base_net)
        self.mano_branch = ManoBranch(ncomps=mano_comps, base_neurons=[width] + mano_neurons,
                                      adapt_skeleton=mano_adapt_skeleton, dropout=fc_dropout, use_trans=False,
                                      mano_root=mano_root, center_idx=mano_center_idx,
                                      use_shape=mano_use_shape, use_pca=mano_use_pca)
        self.mano_loss = ManoLoss(lambda_verts=mano_lambda_verts, lambda_joints3d=mano_lambda_joints3d,
                                  lambda_shape=mano_lambda_shape, lambda_pose_reg=mano_lambda_pose_reg,
                                  lambda_pca=mano_lambda_pca)
        self.atlas_branch = AtlasBranch(mode="sphere", use_residual=atlas_residual, points_nb=atlas_points_nb,
                                        predict_trans=atlas_predict_trans, predict_scale=atlas_predict_scale,
                                        inference_ico_divisions=atlas_ico_divisions, bottleneck_size=width,
                                        use_tanh=atlas_use_tanh, out_factor=atlas_out_factor,
                                        separate_encoder=atlas_separate_encoder)
        self.atlas_loss = AtlasLoss(atlas_loss=atlas_loss, lambda_atlas=atlas_lambda,
                                    final_lambda_atlas=atlas_final_lambda, trans_weight=atlas_trans_weight,
                                    scale_weight=atlas_scale_weight, edge_regul_lambda=atlas_lambda_regul_edges,
                                    lambda_laplacian=atlas_lambda_laplacian,
                                    laplacian_faces=self.atlas_branch.test_faces,
                                    laplacian_verts=self.atlas_branch.test_verts)
        # weights of the final total: the branch losses enter with 1, the contact terms with their lambdas
        self._total_weights = losshead.LossWeights(("one", "contact_lambda", "collision_lambda"))
        self._total_weights["one"] = 1.0

    def decay_regul(self, gamma):
        """handnet.py:188-196: rescale the two mesh-regulariser weights.  The new values reach the device weight vector
        at once, so a captured training step follows the decay from its next replay on."""
        for name in ("edge_regul_lambda", "lambda_laplacian"):
            current = getattr(self.atlas_loss, name)
            if current is not None:
                setattr(self.atlas_loss, name, gamma * current)
        self.atlas_loss.sync_weights()

    # ---- stages of forward -----------------------------------------------------------------------------------------
    def _wants_hand(self, sample):
        supervised = (TransQueries.joints3d in sample or TransQueries.verts3d in sample
                      or (TransQueries.joints2d in sample and TransQueries.camintrs in sample))
        return supervised and BaseQueries.sides in sample and self.mano_lambdas

    def _object_lane(self, features, atlas_features):
        if not self.atlas_mesh:
            # random-points mode ignores the separate encoder and the scale head
            return self.atlas_branch(features)
        if self.adapt_atlas_decoder:
            features = mlp.linear(features, self.atlas_adapter.weight, self.atlas_adapter.bias)
        if self.atlas_separate_encoder:
            return self.atlas_branch.forward_inference(features, separate_encoder_features=atlas_features)
        return self.atlas_branch.forward_inference(features)

    def _contact_stage(self, hand, obj, sample, no_loss, results, losses, parts):
        if hand is None:
            raise RuntimeError("HandNet: the contact loss needs the MANO branch in the same forward (hand supervision "
                               "keys + 'sides' in the sample and a non-zero MANO lambda; handnet.py:337)")
        attraction, repulsion, info, metrics = compute_contact_loss(
            hand["verts"], self.mano_branch.faces, obj["objpoints3d"], self.atlas_branch.test_faces,
            contact_thresh=self.contact_thresh, contact_mode=self.contact_mode,
            collision_thresh=self.collision_thresh, collision_mode=self.collision_mode,
            contact_target=self.contact_target, contact_zones=self.contact_zones)
        results["contact_info"] = info
        if no_loss:
            return
        if TransQueries.verts3d in sample and TransQueries.objpoints3d in sample:
            # contact-IoU metric against the GT hand -> GT object distances (handnet.py:349-362); nearest-neighbour
            # kernel instead of the reference's (B,778,M) matrix
            gt_dists, _, _, _ = F_b200.nearest_neighbours(sample[TransQueries.verts3d],
                                                          sample[TransQueries.objpoints3d], dirs=1)
            info["batch_ious"], losses["contact_auc"] = meshiou(gt_dists, info["min_dists"])
        slot = self._total_weights.slot
        parts.append((attraction, 1.0, slot["contact_lambda"], 1))
        parts.append((repulsion, 1.0, slot["collision_lambda"], 1))
        losses["penetration_loss"] = repulsion
        losses["attraction_loss"] = attraction
        losses.update(metrics)

    def _total(self, parts, losses, device):
        """One weighted-sum kernel over the branch totals and the contact terms; group 1 of its by-products is the
        reference's ``contact_loss`` (handnet.py:363-367)."""
        if not parts:
            return None
        if len(parts) == 1 and parts[0][2] == self._total_weights.slot["one"]:
            return parts[0][0]
        self._total_weights["contact_lambda"] = self.contact_lambda
        self._total_weights["collision_lambda"] = self.collision_lambda
        total, groups, _ = losshead.combine(parts, self._total_weights.device(device))
        if "attraction_loss" in losses:
            losses["contact_loss"] = groups[1:2]
        return total

    def forward(self, sample, no_loss=False, return_features=False, force_objects=False):
        if force_objects and TransQueries.objpoints3d not in sample:
            sample[TransQueries.objpoints3d] = None
        # the reference receives every tensor on-device from DataParallel.scatter
        for key in (TransQueries.joints3d, TransQueries.verts3d, TransQueries.objpoints3d):
            if torch.is_tensor(sample.get(key)) and not sample[key].is_cuda:
                sample[key] = sample[key].cuda()
        image = sample[TransQueries.images].cuda()
        results, losses, parts = {}, {}, []
        one = self._total_weights.slot["one"]

        features, _ = self.base_net(image)
        atlas_features = None
        if self.atlas_separate_encoder:
            atlas_features, _ = self.atlas_base_net(image)
        if return_features:
            results["img_features"] = features
            if atlas_features is not None:
                results["atlas_features"] = atlas_features

        # hand lane: 3 small GEMMs, the MANO layers and the fused loss kernel are launch-latency bound and only share
        # the image features with the object lane -> issued on the BRANCH stream, joined before the first consumer of
        # the hand vertices; autograd replays the same two lanes in the backward pass
        hand = None
        if self._wants_hand(sample):
            streams.fork(streams.BRANCH)
            with streams.on_aux(streams.BRANCH):
                hand = self.mano_branch(features, sides=sample[BaseQueries.sides], root_palm=sample["root"] == "palm",
                                        use_stereoshape=False, side_mask=sample.get("sides_mask"))
                if not no_loss:
                    hand_total, hand_report = self.mano_loss.compute_loss(hand, sample)
                    parts.append((hand_total, 1.0, one, 0))
                    losses.update(hand_report)
            results.update(hand)

        wants_object = TransQueries.objpoints3d in sample and (self.atlas_lambda or self.atlas_final_lambda)
        obj = self._object_lane(features, atlas_features) if wants_object else None
        if hand is not None:
            streams.join(streams.BRANCH)
        if obj is not None:
            if self.need_collisions:
                self._contact_stage(hand, obj, sample, no_loss, results, losses, parts)
            results.update(obj)
            if not no_loss:
                obj_total, obj_report = self.atlas_loss.compute_loss(obj, sample)
                parts.append((obj_total, 1.0, one, 0))
                losses.update(obj_report)

        total_loss = self._total(parts, losses, image.device)
        if total_loss is not None and "mano_total_loss" in losses:
            # the reference accumulates in place into the tensor ManoLoss returned, so its "mano_total_loss" IS the
            # total; reproduced as an alias
            losses["mano_total_loss"] = total_loss
        losses["total_loss"] = total_loss
        return total_loss, results, losses
