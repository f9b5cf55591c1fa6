"""Mirror of mano_train/networks/branches/atlasutils.py: ChamferLoss and the AtlasNet point decoder."""
import torch
from torch import nn

from ... import functional as F_b200


class ChamferLoss(nn.Module):
    """ChamferLoss().forward(preds, gts) -> (loss_1, loss_2), atlasutils.py:6-18.
    loss_1[b] = mean_j min_i |gt_i - pred_j|^2 ; loss_2[b] = mean_i min_j |gt_i - pred_j|^2."""

    def __init__(self):
        super(ChamferLoss, self).__init__()
        self.use_cuda = torch.cuda.is_available()

    def forward(self, preds, gts):
        return F_b200.chamfer(preds, gts)

    def batch_pairwise_dist(self, x, y):
        """Full (B,Nx,Ny) matrix for callers that ask for it (atlasutils.py:20-39)."""
        return ((x.unsqueeze(2) - y.unsqueeze(1)) ** 2).sum(-1)


class PointGenCon(nn.Module):
    """AtlasNet point decoder, atlasutils.py:42-75: Conv1d(k=1) 515->515->257->128->3 with BatchNorm1d + ReLU,
    output scaled by ``out_factor`` (tanh variant is not on the hot path: traineval.py:54 fixes use_tanh=False).

    ``forward(x)`` keeps the reference signature on the concatenated (B, 3+F, N) tensor; AtlasBranch calls
    ``decode(features, grid)`` instead, which never materialises that tensor (conv1 split)."""

    def __init__(self, bottleneck_size=2500, use_tanh=False, out_factor=200):
        if use_tanh:
            raise NotImplementedError("use_tanh=True is not on the hot path (traineval.py:54)")
        self.bottleneck_size = bottleneck_size
        self.use_tanh = use_tanh
        self.out_factor = out_factor
        super(PointGenCon, self).__init__()
        self.conv1 = torch.nn.Conv1d(self.bottleneck_size, self.bottleneck_size, 1)
        self.conv2 = torch.nn.Conv1d(self.bottleneck_size, int(self.bottleneck_size / 2), 1)
        self.conv3 = torch.nn.Conv1d(int(self.bottleneck_size / 2), int(self.bottleneck_size / 4), 1)
        self.conv4 = torch.nn.Conv1d(int(self.bottleneck_size / 4), 3, 1)
        self.bn1 = torch.nn.BatchNorm1d(self.bottleneck_size)
        self.bn2 = torch.nn.BatchNorm1d(int(self.bottleneck_size / 2))
        self.bn3 = torch.nn.BatchNorm1d(int(self.bottleneck_size / 4))

    def _params(self):
        """(conv tensors, bn tensors, momenta, training); the three BatchNorm1d layers must be in the same mode."""
        bn_mods = (self.bn1, self.bn2, self.bn3)
        modes = {bool(bn.training) for bn in bn_mods}
        if len(modes) != 1:
            raise NotImplementedError("PointGenCon: BatchNorm layers in mixed train / eval mode are not supported")
        convs = [self.conv1.weight, self.conv1.bias, self.conv2.weight, self.conv2.bias,
                 self.conv3.weight, self.conv3.bias, self.conv4.weight, self.conv4.bias]
        bns = []
        for bn in bn_mods:
            bns.extend([bn.weight, bn.bias, bn.running_mean, bn.running_var])
        momenta = [0.1 if bn.momentum is None else float(bn.momentum) for bn in bn_mods]
        return convs, bns, momenta, modes.pop()

    def decode(self, features, grid):
        """features (B,F), grid (N,3) or (B,N,3) -> (B,N,3) = out_factor * decoder(cat(grid, features)).  BatchNorm in
        eval mode (the README recipe, --freeze_batchnorm) is folded into the tensor-core GEMMs; in training mode the
        batch statistics run as separate kernels (mlp.point_decoder_train)."""
        from ... import mlp
        convs, bns, momenta, training = self._params()
        if training:
            return mlp.point_decoder_train(features, grid, float(self.out_factor), convs, bns, momenta)
        return mlp.point_decoder(features, grid, float(self.out_factor), convs, bns)

    def forward(self, x):
        """x (B, 3+F, N) with x[:, 3:] constant along N (the AtlasBranch input) -> (B,3,N)."""
        grid = x[:, :3].transpose(2, 1).contiguous()
        feats = x[:, 3:, 0].contiguous()
        return self.decode(feats, grid).transpose(2, 1)
