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
import torch.nn.functional as torch_f

from ... import functional as Fb
from ... import mlp
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
        if self.predict_scale:
            scaled_verts = scales.unsqueeze(1) * verts
            if self.predict_trans:
                objpoints3d = scaled_verts + translations.unsqueeze(1)
        elif self.predict_trans:
            objpoints3d = verts + translations.unsqueeze(1)
        if not self.predict_scale and not self.predict_trans:
            results = {"objpoints3d": verts, "objfaces": self.test_faces}
        if self.predict_trans:
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
    def __init__(self, lambda_atlas=1, atlas_loss="chamfer", final_lambda_atlas=1, trans_weight=0,
                 scale_weight=0, edge_regul_lambda=None, lambda_laplacian=0, laplacian_faces=None,
                 laplacian_verts=None):
        self.lambda_atlas = lambda_atlas
        self.final_lambda_atlas = final_lambda_atlas
        self.trans_weight = trans_weight
        self.scale_weight = scale_weight
        self.edge_regul_lambda = edge_regul_lambda
        self.lambda_laplacian = lambda_laplacian
        if lambda_laplacian:
            # atlasbranch.py:189-192; the reference's own class is a legacy autograd Function that fails on
            # torch >= 1.5 - this one computes the same loss on the device
            from .laplacianloss import LaplacianLoss
            self.laplacian_loss = LaplacianLoss(laplacian_faces, laplacian_verts)
        self.atlas_loss = atlas_loss
        if self.atlas_loss == "chamfer":
            self.chamfer_loss = atlasutils.ChamferLoss()
        else:
            raise ValueError("Removed support for earth mover distance !")

    def compute_loss(self, preds, target):
        atlas_losses = {}
        if (TransQueries.objpoints3d in target and (self.lambda_atlas or self.final_lambda_atlas)) or (
                TransQueries.center3d in target and self.trans_weight):
            gt = target[TransQueries.objpoints3d]
            if "objtrans" in preds and TransQueries.objpoints3d in target and ("objpointscentered3d" in preds):
                obj_centroids = gt.mean(1)
                trans3d_loss = torch_f.mse_loss(preds["objtrans"], obj_centroids)
                atlas_losses["atlas_trans3d"] = trans3d_loss
                centered_objpoints3d = gt - obj_centroids.unsqueeze(1)
                if "objscale" in preds:
                    obj_scales = torch.norm(centered_objpoints3d, 2, 2).max(1)[0]
                    scale3d_loss = torch_f.mse_loss(preds["objscale"], obj_scales.unsqueeze(1))
                    atlas_losses["atlas_scale3d"] = scale3d_loss
                else:
                    scale3d_loss = 0
                loss_1, loss_2 = self.chamfer_loss(preds["objpointscentered3d"], centered_objpoints3d)
                sym_loss = torch.mean(loss_1 + loss_2)
                obj_mesh = preds["objpointscentered3d"]
                final_loss_1, final_loss_2 = self.chamfer_loss(preds["objpoints3d"], gt)
                sym_final_loss = torch.mean(final_loss_1 + final_loss_2)
                atlas_losses["final_{}_loss".format(self.atlas_loss)] = sym_final_loss
                final_loss = (self.lambda_atlas * sym_loss + self.final_lambda_atlas * sym_final_loss
                              + self.trans_weight * trans3d_loss + self.scale_weight * scale3d_loss)
            else:
                if "objpoints3d" in preds and self.lambda_atlas:
                    loss_1, loss_2 = self.chamfer_loss(preds["objpoints3d"], gt)
                    sym_loss = torch.mean((loss_1 + loss_2))
                    final_loss = self.lambda_atlas * sym_loss
                    obj_mesh = preds["objpoints3d"]
            if self.edge_regul_lambda is not None and (self.edge_regul_lambda > 0):
                edge_regul_loss = edge_loss(obj_mesh, preds["objfaces"])
                atlas_losses["atlas_edge_regul"] = edge_regul_loss
                final_loss = final_loss + self.edge_regul_lambda * edge_regul_loss
            if self.lambda_laplacian:  # atlasbranch.py:275-280
                laplacian_loss = self.laplacian_loss(obj_mesh)
                atlas_losses["atlas_laplac"] = laplacian_loss
                final_loss = final_loss + self.lambda_laplacian * laplacian_loss
        else:
            sym_loss = None
            final_loss = torch.zeros(1, device="cuda")
        atlas_losses["atlas_objpoints3d"] = sym_loss
        return final_loss, atlas_losses
