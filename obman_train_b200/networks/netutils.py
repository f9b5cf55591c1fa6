"""Freezing helpers with the reference's names and effects (/root/reference/mano_train/networks/netutils.py:4-19):

``freeze_batchnorm_stats(model)``  momentum = 0 on every BatchNorm layer, so running statistics stop moving even in
                                   train mode (traineval.py:91-92 pairs it with ``model.eval()`` during training);
``rec_freeze(model)``              the same plus ``requires_grad = False`` on every parameter owned by a sub-module
                                   (``--freeze_encoder`` / ``--atlas_freeze_encoder`` / ``--atlas_freeze_decoder``,
                                   traineval.py:93-102).  ``FlatAdamTrainer`` leaves such parameters out of the flat
                                   buffers, exactly like ``filter(requires_grad, model.parameters())`` does there.
"""
from torch.nn.modules.batchnorm import _BatchNorm


def freeze_batchnorm_stats(model):
    for layer in model.modules():          # modules() already walks the whole tree
        if isinstance(layer, _BatchNorm):
            layer.momentum = 0


def rec_freeze(model):
    freeze_batchnorm_stats(model)
    # parameters registered directly on `model` itself keep their flag (the reference only visits children)
    for child in model.children():
        for param in child.parameters():
            param.requires_grad = False
