"""Mirror of mano_train/networks/manonetwork.py::ManoNet (the hand-only network; the reference file is
unimportable as shipped, manonetwork.py:8-9, but its ``ManoNet.forward(images, sides)`` signature is part
of the drop-in contract).  ``HandRegNet`` is out of scope (unused by traineval.py)."""
from torch import nn

from .branches.manobranch import ManoBranch


class ManoNet(nn.Module):
    def __init__(self, base_net, base_neurons=[2048, 512], ncomps=6, center_idx=9, use_shape=False,
                 use_trans=False, mano_root="misc/mano"):
        super(ManoNet, self).__init__()
        self.base_net = base_net
        self.mano_branch = ManoBranch(ncomps=ncomps, base_neurons=base_neurons, use_trans=use_trans,
                                      use_shape=use_shape, mano_root=mano_root, center_idx=center_idx)

    def forward(self, images, sides):
        features, _ = self.base_net(images)
        results = self.mano_branch(features, sides=sides)
        return results
