"""Mirror of mano_train/networks/branches/atlasbranch.py: AtlasNet sphere-deformation branch and its loss.

Same constructor / forward / forward_inference signatures and result keys as the reference
(/root/reference/mano_train/networks/branches/atlasbranch.py:14-150) and the same AtlasLoss.compute_loss
contract (:170-287).  The decoder and the two small MLP heads run on the tcgen05 GEMM kernels, the
Chamfer terms on the fused nearest-neighbour kernel.  The residual decoder (``use_residual``) is out of
scope: ``atlas_residual`` is never set from the CLI.  The mesh regularisers (edge lengths,
cotangent Laplacian) run as fused CUDA kernels (csrc/mesh_regul.cu); the reference's own Laplacian class is a legacy
autograd Function that fails on torch >= 1.5 (Appendix A.17), the one here computes the same loss on the device.
"""
import numpy as np
import torch
from torch import nn

from ... import functional as Fb
from ... import losshead, mlp
from ...icosphere import icosphere
from ...queries import TransQueries
from . import atlasutils


class _MLP2(nn.Sequential):
    """nn.Sequential(Linear, ReLU, Linear) whose forward runs on the tensor-core GEMM kernel."""

    def forward(self, x):
        h = mlp.linear(x, self[0].weight, self[0].bias, relu=True)
        return mlp.linear(h, self[2].weight, self[2].bias, relu=False)


class AtlasBranch(nn.Module):
    def __init__(self, use_residual=True, mode="sphere", points_nb=600, bottleneck_size=1024,
                 use_tanh=False, inference_ico_divisions=3, predict_trans=False, predict_scale=False,
                 out_factor=200, separate_encoder=False):
        super(AtlasBranch, self).__init__()
        self.mode = mode
        self.points_nb = points_nb
        self.bottleneck_size = bottleneck_size
        self.separate_encoder = separate_encoder
        self.use_residual = use_residual
        if self.use_residual:
            raise NotImplementedError("the residual AtlasNet decoder is not on the hot path "
                                      "(atlas_residual is never set by traineval.py)")
        self.decoder = atlasutils.PointGenCon(bottleneck_size=3 + self.bottleneck_size,
                                              out_factor=out_factor, use_tanh=use_tanh)
        self.predict_trans = predict_trans
        if self.predict_trans:
            self.decode_trans = _MLP2(
                torch.nn.Linear(self.bottleneck_size, int(self.bottleneck_size / 2)), torch.nn.ReLU(),
                torch.nn.Linear(int(self.bottleneck_size / 2), 3))
        self.predict_scale = predict_scale
        if self.predict_scale:
            self.decode_scale = _MLP2(
                torch.nn.Linear(self.bottleneck_size, int(self.bottleneck_size / 2)), torch.nn.ReLU(),
                torch.nn.Linear(int(self.bottleneck_size / 2), 1))
            self.decode_scale[-1].bias.data.fill_(1)
        if mode == "sphere":
            test_verts, test_faces = icosphere(subdivisions=inference_ico_divisions)
        else:
            raise ValueError("{} not in [sphere]".format(mode))
        # plain attributes like the reference (not buffers: they are not in its state dict)
        self.test_verts = torch.Tensor(np.array(test_verts).astype(np.float32))
        self.test_faces = np.array(test_faces)
        self.rand_grid = None  # optional injected (B,points_nb,3) unit-sphere sample for forward()

    def _apply(self, fn, *args, **kwargs):
        out = super(AtlasBranch, self)._apply(fn, *args, **kwargs)
        self.test_verts = fn(self.test_verts)
        return out

    def forward(self, img_features):
        """Random-points mode (atlasbranch.py:78-108): the sphere sample comes from the global torch RNG
        unless ``self.rand_grid`` has been set (parity tests inject it)."""
        if self.predict_trans:
            translations = self.decode_trans(img_features)
        if self.rand_grid is not None:
            rand_grid = self.rand_grid
        else:
            rand_grid = img_features.new_empty((img_features.size(0), 3, self.points_nb))
            rand_grid.data.normal_(0, 1)
            rand_grid = (rand_grid / torch.sqrt(torch.sum(rand_grid ** 2, dim=1, keepdim=True))).transpose(2, 1)
        verts = self.decoder.decode(img_features, rand_grid.contiguous())
        if self.predict_trans:
            objpoints3d = verts + translations.unsqueeze(1)
            results = {"objpoints3d": objpoints3d, "objtrans": translations, "objpointscentered3d": verts}
        else:
            results = {"objpoints3d": verts}
        return results

    def forward_inference(self, img_features, separate_encoder_features=None):
        """Mesh mode (atlasbranch.py:110-150): icosphere grid, optional scale / translation heads."""
        if self.predict_trans:
            translations = self.decode_trans(img_features)
        if self.predict_scale:
            scales = self.decode_scale(img_features)
        dec_features = separate_encoder_features if self.separate_encoder else img_features
        verts = self.decoder.decode(dec_features, self.test_verts)
        # per-sample similarity transform of the decoded sphere: one kernel per direction (losshead.affine_points)
        if not self.predict_scale and not self.predict_trans:
            results = {"objpoints3d": verts, "objfaces": self.test_faces}
        if self.predict_trans:
            objpoints3d = losshead.affine_points(verts, scales if self.predict_scale else None, translations)
            results = {"objpoints3d": objpoints3d, "objtrans": translations,
                       "objpointscentered3d": verts, "objfaces": self.test_faces}
        if self.predict_scale:
            results["objscale"] = scales
        return results


_faces_cache = {}


def _faces_on(faces, device, dtype=np.int64):
    """(F,3) face tensor on ``device``, cached (no host->device copy per step, CUDA-graph safe)."""
    if torch.is_tensor(faces):
        faces = faces.detach().cpu().numpy()
    arr = np.ascontiguousarray(np.asarray(faces).astype(dtype))
    key = (arr.shape, arr.dtype.str, hash(arr.tobytes()), str(device))
    if key not in _faces_cache:
        _faces_cache[key] = torch.from_numpy(arr).to(device)
    return _faces_cache[key]


_vert_faces_cache = {}


def _vert_faces_on(faces, n_verts, device):
    arr = np.ascontiguousarray(np.asarray(faces).astype(np.int32))
    key = (arr.shape, hash(arr.tobytes()), n_verts, str(device))
    if key not in _vert_faces_cache:
        from .laplacianloss import vertex_face_table
        _vert_faces_cache[key] = torch.from_numpy(vertex_face_table(n_verts, arr)).to(device)
    return _vert_faces_cache[key]


def edge_loss(edges, faces):
    """atlasbranch.py:153-167: mean absolute deviation of the squared edge lengths from their per-sample mean.
    One fused CUDA kernel pair per direction (csrc/mesh_regul.cu) instead of ~15 gather / reduce launches."""
    faces_i32 = _faces_on(faces, edges.device, np.int32)
    vert_faces = _vert_faces_on(faces, edges.shape[1], edges.device)
    return Fb.edge_loss(edges, faces_i32, vert_faces).squeeze(0)


class AtlasLoss:
    """Object-branch loss (atlasbranch.py:170-287): centred and final Chamfer terms, translation / scale regression
    against the GT centroid and extent, optional edge-length and Laplacian mesh regularisers.

    Device-side: GT statistics are one kernel (losshead.object_targets), the two regression terms one kernel
    (losshead.sq_terms), every Chamfer / regulariser term comes from its own fused kernel, and the weighted total is
    one kernel (losshead.combine) that reads the lambdas from a device vector - so ``HandNet.decay_regul``
    (handnet.py:188-196) takes effect inside a captured CUDA graph.  The lambda attributes stay plain Python numbers
    (``decay_regul`` rescales them from outside); they are mirrored into the device vector at every call."""

    _WEIGHTS = ("lambda_atlas", "final_lambda_atlas", "trans_weight", "scale_weight", "edge_regul_lambda",
                "lambda_laplacian")

    def __init__(self, lambda_atlas=1, atlas_loss="chamfer", final_lambda_atlas=1, trans_weight=0,
                 scale_weight=0, edge_regul_lambda=None, lambda_laplacian=0, laplacian_faces=None,
                 laplacian_verts=None):
        if atlas_loss != "chamfer":
            raise ValueError("Removed support for earth mover distance !")   # the reference's message (atlasbranch.py:197)
        self.atlas_loss = atlas_loss
        self.chamfer_loss = atlasutils.ChamferLoss()
        self.lambda_atlas = lambda_atlas
        self.final_lambda_atlas = final_lambda_atlas
        self.trans_weight = trans_weight
        self.scale_weight = scale_weight
        self.edge_regul_lambda = edge_regul_lambda
        self.lambda_laplacian = lambda_laplacian
        if lambda_laplacian:
            # the reference's own class is a legacy autograd Function that fails on torch >= 1.5 (SURVEY.md
            # Appendix A.17); this one computes the same loss on the device
            from .laplacianloss import LaplacianLoss
            self.laplacian_loss = LaplacianLoss(laplacian_faces, laplacian_verts)
        self._weights = losshead.LossWeights(self._WEIGHTS + ("one",))
        self._weights["one"] = 1.0
        self._scratch = losshead._Workspace()

    def sync_weights(self):
        """Mirror the lambda attributes into the device vector (a fill kernel per CHANGED value)."""
        for name in self._WEIGHTS:
            self._weights[name] = getattr(self, name)

    def _chamfer_term(self, pred_points, gt_points, weight_name):
        """Symmetric Chamfer distance as a combine term: mean_b(loss_1 + loss_2) = (sum loss_1 + sum loss_2) / B.
        A weight that is zero (the README recipe leaves --atlas_lambda at 0) still gets its
        forward value logged but is cut off from the backward pass."""
        loss_1, loss_2 = self.chamfer_loss(pred_points, gt_points)
        if not self._weights[weight_name]:
            loss_1, loss_2 = loss_1.detach(), loss_2.detach()
        return ((loss_1, loss_2), 1.0 / loss_1.shape[0], self._weights.slot[weight_name], 0)

    def compute_loss(self, preds, target):
        report = {"atlas_objpoints3d": None}
        has_points = TransQueries.objpoints3d in target
        if not ((has_points and (self.lambda_atlas or self.final_lambda_atlas))
                or (TransQueries.center3d in target and self.trans_weight)):
            return torch.zeros(1, device="cuda"), report
        self.sync_weights()
        gt = target[TransQueries.objpoints3d]
        weights = self._weights.device(gt.device)
        slot = self._weights.slot
        terms, logged = [], []   # combine terms; (report key, index of the term whose value is logged)
        mesh = None
        if has_points and "objtrans" in preds and "objpointscentered3d" in preds:
            # regression targets from the GT cloud: centroid, extent, centred copy (atlasbranch.py:211-227)
            centroid, extent, centred_gt = losshead.object_targets(gt)
            heads = [(preds["objtrans"], centroid, None, slot["trans_weight"])]
            if "objscale" in preds:
                heads.append((preds["objscale"], extent, None, slot["scale_weight"]))
            head_sum, head_vals = losshead.sq_terms(heads, weights, self._scratch)
            report["atlas_trans3d"] = head_vals[0:1]
            if "objscale" in preds:
                report["atlas_scale3d"] = head_vals[1:2]
            terms.append((head_sum, 1.0, slot["one"], 0))
            mesh = preds["objpointscentered3d"]
            logged.append(("atlas_objpoints3d", len(terms)))
            terms.append(self._chamfer_term(mesh, centred_gt, "lambda_atlas"))
            logged.append(("final_{}_loss".format(self.atlas_loss), len(terms)))
            terms.append(self._chamfer_term(preds["objpoints3d"], gt, "final_lambda_atlas"))
        elif "objpoints3d" in preds and self.lambda_atlas:
            mesh = preds["objpoints3d"]
            logged.append(("atlas_objpoints3d", len(terms)))
            terms.append(self._chamfer_term(mesh, gt, "lambda_atlas"))
        if mesh is None:
            raise RuntimeError("AtlasLoss.compute_loss: nothing to supervise (predictions lack 'objpoints3d' / "
                               "'objtrans' + 'objpointscentered3d' for the configured weights)")
        if self.edge_regul_lambda is not None and self.edge_regul_lambda > 0:
            logged.append(("atlas_edge_regul", len(terms)))
            terms.append((edge_loss(mesh, preds["objfaces"]), 1.0, slot["edge_regul_lambda"], 0))
        if self.lambda_laplacian:
            logged.append(("atlas_laplac", len(terms)))
            terms.append((self.laplacian_loss(mesh), 1.0, slot["lambda_laplacian"], 0))
        final_loss, _, values = losshead.combine(terms, weights)
        for key, k in logged:
            report[key] = values[k:k + 1]
        return final_loss, report
