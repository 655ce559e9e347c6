"""DeTone-style regression backbone (reference ``src/backbones/ResNet34.py``): torchvision ResNet-34 with a
2-channel stem and an 8-way head, emitting the 4-point offsets directly.  Same kwargs / dict protocol / names."""
import torch
import torch.nn as nn

from .blocks import offset_regressor


class Model(nn.Module):

    def __init__(self, **kwargs):
        super().__init__()
        self.patch_keys = kwargs['PATCH_KEYS']
        self.target_keys = kwargs['TARGET_KEYS']
        self.resnet34 = offset_regressor(kwargs['PRETRAINED_RESNET'])
        self.variant = kwargs.get('VARIANT', 'oneline').lower()
        if not ('oneline' in self.variant or 'doubleline' in self.variant):
            raise AssertionError('Only OneLine or DoubleLine variant is supported')

    def single_forward(self, x):
        return self.resnet34(x).reshape(-1, 4, 2)

    def forward(self, data):
        """adds TARGET_KEYS[0] (patch 1 -> 2) and, for DoubleLine, TARGET_KEYS[1] (patch 2 -> 1) to the batch dict; the
        two directions are separate passes because BatchNorm normalises with the statistics of each call"""
        pair = [data[k] for k in self.patch_keys]
        directions = (pair, pair[::-1]) if self.variant == 'doubleline' else (pair,)
        for key, (a, b) in zip(self.target_keys, directions):
            data[key] = self.single_forward(torch.cat([a, b], dim=1))
        return data

    def predict_homography(self, data):
        return self.forward(data)
