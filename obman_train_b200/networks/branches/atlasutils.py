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
