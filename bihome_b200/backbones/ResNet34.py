"""DeTone-style regression backbone (reference ``src/backbones/ResNet34.py``): torchvision ResNet-34 with a
2-channel stem and an 8-way head, emitting the 4-point offsets directly.  Same kwargs / dict protocol / names."""
import warnings

import torch
import torch.nn as nn
import torchvision.models as models


class Model(nn.Module):

    def __init__(self, **kwargs):
        super().__init__()
        self.patch_keys = kwargs['PATCH_KEYS']
        self.target_keys = kwargs['TARGET_KEYS']
        net = None
        if kwargs['PRETRAINED_RESNET']:
            try:
                net = models.resnet34(weights='DEFAULT', progress=True)
            except Exception as e:  # noqa: BLE001
                warnings.warn('bihome_b200: pretrained resnet34 weights unavailable (%s); using random init' % e)
        self.resnet34 = net if net is not None else models.resnet34(weights=None)
        self.resnet34.conv1 = nn.Conv2d(2, 64, kernel_size=(7, 7), stride=(2, 2), padding=(3, 3), bias=False)
        self.resnet34.fc = nn.Linear(512, 8, bias=True)
        self.variant = str.lower(kwargs['VARIANT']) if 'VARIANT' in kwargs else 'oneline'
        assert 'oneline' in self.variant or 'doubleline' in self.variant, 'Only OneLine or DoubleLine variant is supported'

    def single_forward(self, x):
        return self.resnet34(x).reshape(-1, 4, 2)

    def forward(self, data):
        e1, e2 = self.patch_keys
        p1, p2 = data[e1], data[e2]
        data[self.target_keys[0]] = self.single_forward(torch.cat([p1, p2], dim=1))
        if self.variant == 'doubleline':
            data[self.target_keys[1]] = self.single_forward(torch.cat([p2, p1], dim=1))
        return data

    def predict_homography(self, data):
        return self.forward(data)
