"""Residual building blocks of the homography backbones.

Same layer graph and the same parameter names (``upper_branch.<i>`` / ``lower_branch.<i>``) as the reference's
``src/backbones/utils.py`` so that reference checkpoints load unchanged; the dense convolutions stay on cuDNN
(tensor-core work is out of scope for the custom kernels, BASELINE.json north_star).
"""
import warnings

import torch.nn as nn
import torchvision.models as models

from .. import functional as F


def _conv(cin, cout, k, stride=1, bias=False):
    return nn.Conv2d(cin, cout, kernel_size=k, padding=k // 2, stride=stride, bias=bias)


def _seq(*specs):
    """specs: ('c', cin, cout, k, stride) conv+bn | 'r' relu | ('t', cin, cout, bias) 2x2 transposed conv | ('b', c) bn"""
    layers = []
    for s in specs:
        if s == 'r':
            layers.append(nn.ReLU())
        elif s[0] == 'c':
            layers += [_conv(s[1], s[2], s[3], s[4]), nn.BatchNorm2d(s[2])]
        elif s[0] == 't':
            layers.append(nn.ConvTranspose2d(s[1], s[2], kernel_size=2, padding=0, stride=2, bias=s[3]))
        elif s[0] == 'b':
            layers.append(nn.BatchNorm2d(s[1]))
    return nn.Sequential(*layers)


class _Residual(nn.Module):
    """relu(upper_branch(x) + (lower_branch(x) or x))"""

    def __init__(self, upper, lower=None):
        super().__init__()
        self.upper_branch = upper
        if lower is not None:
            self.lower_branch = lower
        self.lower_is_identity = lower is None

    def forward(self, x):
        # a skip path that ends in its own BatchNorm (projection / up-sampling blocks): keep its un-normalised tensor and let
        # the end of the block normalise both branches, add and rectify in one stage (K7c)
        skip, skip_bn = x, None
        if not self.lower_is_identity:
            lower = list(self.lower_branch)
            if isinstance(lower[-1], nn.BatchNorm2d) and F.bnact2_enabled():
                for m in lower[:-1]:
                    skip = m(skip)
                skip_bn = lower[-1]
            else:
                skip = self.lower_branch(x)
        # the memory-bound stages between the convolutions run fused (K7b, csrc/stem.cu) on channels-last CUDA tensors in
        # training: BatchNorm2d -> ReLU inside the branch, BatchNorm2d + skip -> ReLU at its end; anything else (eval mode,
        # NCHW, CPU) takes the modules one by one
        mods = list(self.upper_branch)
        h, i, n = x, 0, len(mods)
        while i < n:
            m = mods[i]
            if isinstance(m, nn.BatchNorm2d):
                if i + 1 < n and isinstance(mods[i + 1], nn.ReLU) and F.bnact_supported(m, h):
                    h = F.bn_relu(m, h)
                    i += 2
                    continue
                if i == n - 1 and skip_bn is not None and h.shape == skip.shape and h.stride() == skip.stride() \
                        and F.bnact_supported(m, h) and F.bnact_supported(skip_bn, skip):
                    return F.bn_bn_relu(m, h, skip_bn, skip)
                if i == n - 1:
                    if skip_bn is not None:
                        skip, skip_bn = skip_bn(skip), None
                    if F.bnact_supported(m, h, skip):
                        return F.bn_relu(m, h, residual=skip)
            if isinstance(m, nn.ConvTranspose2d) and F.convt_bias_supported(m, h):
                h = F.conv_transpose_bias(m, h)      # K8: cuDNN's transposed convolution + the bias in place
            else:
                h = m(h)
            i += 1
        if skip_bn is not None:
            skip = skip_bn(skip)
        return nn.functional.relu(h + skip)


class ResNet34ConvBlock(_Residual):
    def __init__(self, input_channels, output_channels, stride):
        upper = _seq(('c', input_channels, output_channels, 3, stride), 'r', ('c', output_channels, output_channels, 3, 1))
        lower = None
        if input_channels != output_channels:
            lower = _seq(('c', input_channels, output_channels, 1, stride))
        super().__init__(upper, lower)


class ResNet34IdentityBlock(_Residual):
    def __init__(self, input_channels):
        c = input_channels
        super().__init__(_seq(('c', c, c, 3, 1), 'r', ('c', c, c, 3, 1)))


class ResNet50ConvBlock(_Residual):
    def __init__(self, input_channels, output_channels, stride):
        mid = input_channels // stride
        upper = _seq(('c', input_channels, mid, 1, stride), 'r', ('c', mid, mid, 3, 1), 'r', ('c', mid, output_channels, 1, 1))
        super().__init__(upper, _seq(('c', input_channels, output_channels, 1, stride)))


class ResNet50IdentityBlock(_Residual):
    def __init__(self, input_channels):
        c, m = input_channels, input_channels // 4
        super().__init__(_seq(('c', c, m, 1, 1), 'r', ('c', m, m, 3, 1), 'r', ('c', m, c, 1, 1)))


class ResNet50DeconvBlock(_Residual):
    """x2 up-sampling block: channels halve."""

    def __init__(self, input_channels):
        c = input_channels
        upper = _seq(('t', c, c, True), ('c', c, c, 3, 1), 'r', ('c', c, c // 2, 1, 1))
        lower = _seq(('t', c, c // 2, False), ('b', c // 2))
        super().__init__(upper, lower)


def offset_regressor(pretrained):
    """torchvision ResNet-34 turned into the 4-point regressor both the DeTone-style and the content-aware backbone use:
    a 2-channel 7x7 stem (the two patches, or the two masked feature maps) and an 8-way linear head.  ``pretrained``
    asks for the ImageNet trunk; without network access (no cached checkpoint) it degrades to random init with a warning."""
    trunk = None
    if pretrained:
        try:
            trunk = models.resnet34(weights='DEFAULT', progress=True)
        except Exception as e:  # noqa: BLE001 -- URLError / missing cache
            warnings.warn('bihome_b200: pretrained resnet34 weights unavailable (%s); using random init' % e)
    if trunk is None:
        trunk = models.resnet34(weights=None)
    trunk.conv1 = nn.Conv2d(2, 64, kernel_size=7, stride=2, padding=3, bias=False)
    trunk.fc = nn.Linear(512, 8, bias=True)
    return trunk
