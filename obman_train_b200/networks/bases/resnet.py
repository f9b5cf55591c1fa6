"""Mirror of mano_train/networks/bases/resnet.py for the hot path: ResNet-18 feature extractor.

Same module tree and state-dict keys as the reference (conv1, bn1, layer1..4.{0,1}.{conv1,bn1,conv2,bn2,
downsample.{0,1}}, fc; /root/reference/mano_train/networks/bases/resnet.py:99-152) and the same call
convention ``features, {} = net(images)`` (:184-185), but ``forward`` runs the whole encoder as one fused
node on the tcgen05 convolution kernels (obman_train_b200/encoder.py).  The nn.Conv2d / nn.BatchNorm2d
children are parameter containers only.  ResNet-34/50/101/152 are out of scope (BASELINE.json names
ResNet-18 only).
"""
import math
import warnings

import torch
import torch.nn as nn

from ... import encoder

__all__ = ["ResNet", "resnet18"]


def conv3x3(in_planes, out_planes, stride=1):
    return nn.Conv2d(in_planes, out_planes, kernel_size=3, stride=stride, padding=1, bias=False)


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super(BasicBlock, self).__init__()
        self.conv1 = conv3x3(inplanes, planes, stride)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = conv3x3(planes, planes)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample
        self.stride = stride


class ResNet(nn.Module):
    def __init__(self, block=BasicBlock, layers=(2, 2, 2, 2), num_classes=1000, features=True,
                 early_features=False, return_inter=False):
        if block is not BasicBlock or tuple(layers) != (2, 2, 2, 2):
            raise NotImplementedError("only ResNet-18 (BasicBlock, [2,2,2,2]) is on the B200 hot path")
        if early_features or return_inter or not features:
            raise NotImplementedError("only the pooled-feature output (features=True) is on the hot path")
        self.inplanes = 64
        self.features = features
        super(ResNet, self).__init__()
        self.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.layer1 = self._make_layer(block, 64, layers[0])
        self.layer2 = self._make_layer(block, 128, layers[1], stride=2)
        self.layer3 = self._make_layer(block, 256, layers[2], stride=2)
        self.layer4 = self._make_layer(block, 512, layers[3], stride=2)
        self.avgpool = nn.AvgPool2d(7, stride=1)
        self.fc = nn.Linear(512 * block.expansion, num_classes)  # constructed, checkpointed, never run
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def _make_layer(self, block, planes, blocks, stride=1):
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            downsample = nn.Sequential(
                nn.Conv2d(self.inplanes, planes * block.expansion, kernel_size=1, stride=stride, bias=False),
                nn.BatchNorm2d(planes * block.expansion))
        layers = [block(self.inplanes, planes, stride, downsample)]
        self.inplanes = planes * block.expansion
        for _ in range(1, blocks):
            layers.append(block(self.inplanes, planes))
        return nn.Sequential(*layers)

    def _unit_params(self):
        """(params, momenta, training): 5 tensors per conv+BN unit; all BatchNorm layers must be in the same mode."""
        mods = dict(self.named_modules())
        params, momenta, modes = [], [], set()
        for conv_name, bn_name, _, _, _, _ in encoder.resnet18_units():
            conv, bn = mods[conv_name], mods[bn_name]
            modes.add(bool(bn.training))
            momenta.append(0.1 if bn.momentum is None else float(bn.momentum))
            params.extend([conv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var])
        if len(modes) != 1:
            raise NotImplementedError("ResNet: BatchNorm layers in mixed train / eval mode are not supported")
        return params, momenta, modes.pop()

    def forward(self, x):
        params, momenta, training = self._unit_params()
        if training:
            # batch statistics (training without --freeze_batchnorm, epochpass3d.py:48-52)
            feats = encoder.resnet18_features_train(x, params, momenta)
        else:
            # fixed running statistics, trainable gamma / beta (--freeze_batchnorm: eval() during training)
            feats = encoder.resnet18_features(x, params)
        return feats, {}


def resnet18(pretrained=False, **kwargs):
    """ResNet-18; ``pretrained=True`` loads the torchvision ImageNet checkpoint from the local torch-hub cache
    (``resnet18-5c106cde.pth``) when it is there.  No download is attempted (no network on the B200 boxes):
    if the file is absent a warning is emitted and the random initialisation is kept."""
    model = ResNet(BasicBlock, [2, 2, 2, 2], **kwargs)
    if pretrained:
        import os
        path = os.path.join(torch.hub.get_dir(), "checkpoints", "resnet18-5c106cde.pth")
        if os.path.exists(path):
            model.load_state_dict(torch.load(path, map_location="cpu"))
        else:
            warnings.warn("ImageNet weights not found at {}; keeping random init".format(path))
    return model
