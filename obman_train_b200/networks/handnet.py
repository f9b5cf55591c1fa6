"""Mirror of mano_train/networks/handnet.py: the per-image training graph of obman_train.

Same ``HandNet(**kwargs)`` constructor, ``forward(sample, no_loss=False, return_features=False,
force_objects=False) -> (total_loss, results, losses)`` contract, attributes read from outside
(``base_net``, ``atlas_base_net``, ``atlas_branch.decoder``, ``decay_regul``, ``mano_branch.faces``,
``atlas_branch.test_faces/test_verts``) and state-dict key names as the reference
(/root/reference/mano_train/networks/handnet.py:19-392), so it drops in under traineval.py /
epochpass3d.py.  Every arithmetic stage runs in libobman_b200.so.

Not on the hot path and therefore rejected at construction: ResNet-50, the absolute / 2-D joint
branches (``absolute_lambda``, ``mano_lambda_joints2d``; dead or default-off in the reference, SURVEY.md
Appendix A.13), the residual decoder, ``mano_adapt_skeleton``, ``fc_dropout``, ``mano_use_pca=False``.
"""
from copy import deepcopy

import torch
from torch import nn

from .. import mlp, streams
from ..queries import TransQueries, BaseQueries
from .bases import resnet
from .branches.manobranch import ManoBranch, ManoLoss
from .branches.atlasbranch import AtlasBranch, AtlasLoss
from .branches.contactloss import compute_contact_loss, meshiou
from .. import functional as F_b200


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
        if int(resnet_version) == 18:
            img_feature_size = 512
            base_net = resnet.resnet18(pretrained=True)
        else:
            raise NotImplementedError("Resnet {} not supported on the B200 hot path (ResNet-18 only)".format(resnet_version))
        if absolute_lambda or mano_lambda_joints2d:
            raise NotImplementedError("absolute / 2-D joint branches are not on the hot path")
        self.adapt_atlas_decoder = adapt_atlas_decoder
        self.atlas_separate_encoder = atlas_separate_encoder
        if self.adapt_atlas_decoder:
            self.atlas_adapter = torch.nn.Linear(img_feature_size, img_feature_size)
        mano_base_neurons = [img_feature_size] + mano_neurons
        self.contact_target = contact_target
        self.contact_zones = contact_zones
        self.contact_lambda = contact_lambda
        self.contact_thresh = contact_thresh
        self.contact_mode = contact_mode
        self.collision_lambda = collision_lambda
        self.collision_thresh = collision_thresh
        self.collision_mode = collision_mode
        self.need_collisions = bool(contact_lambda or collision_lambda)
        self.base_net = base_net
        if self.atlas_separate_encoder:
            self.atlas_base_net = deepcopy(base_net)
        self.absolute_lambda = absolute_lambda
        self.mano_adapt_skeleton = mano_adapt_skeleton
        self.mano_branch = ManoBranch(ncomps=mano_comps, base_neurons=mano_base_neurons,
                                      adapt_skeleton=mano_adapt_skeleton, dropout=fc_dropout, use_trans=False,
                                      mano_root=mano_root, center_idx=mano_center_idx,
                                      use_shape=mano_use_shape, use_pca=mano_use_pca)
        self.mano_lambdas = bool(mano_lambda_verts or mano_lambda_joints3d or mano_lambda_joints2d
                                 or mano_lambda_pca)
        self.mano_loss = ManoLoss(lambda_verts=mano_lambda_verts, lambda_joints3d=mano_lambda_joints3d,
                                  lambda_shape=mano_lambda_shape, lambda_pose_reg=mano_lambda_pose_reg,
                                  lambda_pca=mano_lambda_pca)
        self.lambda_joints2d = mano_lambda_joints2d
        self.atlas_mesh = atlas_mesh
        self.atlas_branch = AtlasBranch(mode="sphere", use_residual=atlas_residual, points_nb=atlas_points_nb,
                                        predict_trans=atlas_predict_trans, predict_scale=atlas_predict_scale,
                                        inference_ico_divisions=atlas_ico_divisions,
                                        bottleneck_size=img_feature_size, use_tanh=atlas_use_tanh,
                                        out_factor=atlas_out_factor,
                                        separate_encoder=self.atlas_separate_encoder)
        self.atlas_lambda = atlas_lambda
        self.atlas_final_lambda = atlas_final_lambda
        self.atlas_trans_weight = atlas_trans_weight
        self.atlas_scale_weight = atlas_scale_weight
        self.atlas_loss = AtlasLoss(atlas_loss=atlas_loss, lambda_atlas=atlas_lambda,
                                    final_lambda_atlas=atlas_final_lambda, trans_weight=atlas_trans_weight,
                                    scale_weight=atlas_scale_weight, edge_regul_lambda=atlas_lambda_regul_edges,
                                    lambda_laplacian=atlas_lambda_laplacian,
                                    laplacian_faces=self.atlas_branch.test_faces,
                                    laplacian_verts=self.atlas_branch.test_verts)

    def decay_regul(self, gamma):
        if self.atlas_loss.edge_regul_lambda is not None:
            self.atlas_loss.edge_regul_lambda = gamma * self.atlas_loss.edge_regul_lambda
        if self.atlas_loss.lambda_laplacian is not None:
            self.atlas_loss.lambda_laplacian = gamma * self.atlas_loss.lambda_laplacian

    def forward(self, sample, no_loss=False, return_features=False, force_objects=False):
        if force_objects:
            if TransQueries.objpoints3d not in sample:
                sample[TransQueries.objpoints3d] = None
        total_loss = None
        results = {}
        losses = {}
        # the reference receives every tensor on-device from DataParallel.scatter
        for key in (TransQueries.joints3d, TransQueries.verts3d, TransQueries.objpoints3d):
            if key in sample and torch.is_tensor(sample[key]) and not sample[key].is_cuda:
                sample[key] = sample[key].cuda()
        image = sample[TransQueries.images].cuda()
        features, _ = self.base_net(image)
        if self.atlas_separate_encoder:
            atlas_infeatures, _ = self.atlas_base_net(image)
            if return_features:
                results["atlas_features"] = atlas_infeatures
        if return_features:
            results["img_features"] = features
        if ((TransQueries.joints3d in sample.keys() or TransQueries.verts3d in sample.keys()
             or (TransQueries.joints2d in sample.keys() and TransQueries.camintrs in sample.keys()))
                and BaseQueries.sides in sample.keys() and self.mano_lambdas):
            root_palm = sample["root"] == "palm"
            # The hand branch (3 small GEMMs, the MANO layer, 4 scalar losses: ~50 launch-latency-bound kernels) and the
            # object branch (the AtlasNet decoder GEMMs) only share the image features: the hand branch is issued on
            # the BRANCH auxiliary stream and joined before the first consumer of its vertices (contact loss / total);
            # autograd replays the same two lanes in the backward pass.
            streams.fork(streams.BRANCH)
            with streams.on_aux(streams.BRANCH):
                mano_results = self.mano_branch(features, sides=sample[BaseQueries.sides], root_palm=root_palm,
                                                use_stereoshape=False, side_mask=sample.get("sides_mask"))
                if not no_loss:
                    mano_total_loss, mano_losses = self.mano_loss.compute_loss(mano_results, sample)
                    if total_loss is None:
                        total_loss = mano_total_loss
                    else:
                        total_loss += mano_total_loss
                    for key, val in mano_losses.items():
                        losses[key] = val
            for key, result in mano_results.items():
                results[key] = result
            hand_lane_open = True
        else:
            hand_lane_open = False
        predict_atlas = TransQueries.objpoints3d in sample.keys() and (self.atlas_lambda or self.atlas_final_lambda)
        if not predict_atlas and hand_lane_open:
            streams.join(streams.BRANCH)
            hand_lane_open = False
        if predict_atlas:
            if self.atlas_mesh:
                if self.adapt_atlas_decoder:
                    atlas_features = mlp.linear(features, self.atlas_adapter.weight, self.atlas_adapter.bias)
                else:
                    atlas_features = features
                if self.atlas_separate_encoder:
                    atlas_results = self.atlas_branch.forward_inference(
                        atlas_features, separate_encoder_features=atlas_infeatures)
                else:
                    atlas_results = self.atlas_branch.forward_inference(atlas_features)
            else:
                atlas_results = self.atlas_branch(features)
            if hand_lane_open:
                streams.join(streams.BRANCH)
            if self.need_collisions:
                attr_loss, penetr_loss, contact_infos, contact_metrics = compute_contact_loss(
                    mano_results["verts"], self.mano_branch.faces, atlas_results["objpoints3d"],
                    self.atlas_branch.test_faces, contact_thresh=self.contact_thresh,
                    contact_mode=self.contact_mode, collision_thresh=self.collision_thresh,
                    collision_mode=self.collision_mode, contact_target=self.contact_target,
                    contact_zones=self.contact_zones)
                if not no_loss:
                    if TransQueries.verts3d in sample and TransQueries.objpoints3d in sample:
                        # GT hand->object distances for the contact-IoU metric: nearest-neighbour kernel
                        # instead of the reference's (B,778,M) matrix (handnet.py:353-357)
                        dist_h2o_gt, _, _, _ = F_b200.nearest_neighbours(
                            sample[TransQueries.verts3d], sample[TransQueries.objpoints3d], dirs=1)
                        contact_ious, contact_auc = meshiou(dist_h2o_gt, contact_infos["min_dists"])
                        contact_infos["batch_ious"] = contact_ious
                        losses["contact_auc"] = contact_auc
                    contact_loss = self.contact_lambda * attr_loss + self.collision_lambda * penetr_loss
                    total_loss += contact_loss
                    losses["penetration_loss"] = penetr_loss
                    losses["attraction_loss"] = attr_loss
                    losses["contact_loss"] = contact_loss
                    for metric_name, metric_val in contact_metrics.items():
                        losses[metric_name] = metric_val
                results["contact_info"] = contact_infos
            for key, result in atlas_results.items():
                results[key] = result
            if not no_loss:
                atlas_total_loss, atlas_losses = self.atlas_loss.compute_loss(atlas_results, sample)
                if total_loss is None:
                    total_loss = atlas_total_loss
                else:
                    total_loss += atlas_total_loss
                for key, val in atlas_losses.items():
                    losses[key] = val
        if total_loss is not None:
            losses["total_loss"] = total_loss
        else:
            losses["total_loss"] = None
        return total_loss, results, losses
